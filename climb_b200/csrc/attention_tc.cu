// tcgen05 / TMEM fused attention for the ViLT hot path (ViltSelfAttention, modeling_vilt.py:355-388)
// when a whole head fits one key tile (L <= 256 keys, dh = 64): CLiMB's 40 text + 197 image tokens.
//
//   forward  (CTA = 128 query rows of one (b, h); 2 CTAs / SM):
//     TMA (3-D map over [B, L, 3*H*64]: rows past L are zero-filled, so no padding buffer exists)
//       -> Q [128 x 64], K [256 x 64], V [256 x 64] in 128B-swizzled smem
//     S = Q K^T            one 128 x 256 x 64 UMMA chain, accumulator = 256 TMEM columns
//     softmax              thread = row (TMEM lane): two passes over the row held in TMEM, additive key
//                          mask, exp2; P is written as bf16 straight into the K-major A-operand layout
//     O = P V              128 x 64 x 256 UMMA chain (V read in place as an MN-major B operand),
//                          accumulator aliases S's first 64 columns
//   backward (CTA = one (b, h); K, V, Q, dO resident; 512 TMEM columns):
//     for key tile j, query tile i:  S = Q_i K_j^T, dP = dO_i V_j^T          (two 128x128x64 chains)
//                                    P = exp2(S*c + mask - lse_i), dS = P * (dP - delta_i)   (thread = row)
//                                    dV_j += P^T dO_i, dK_j += dS^T Q_i, dQ_i += dS K_j
//     P / dS are stored once in smem and read BOTH as K-major A (for dQ) and, in place, as MN-major A
//     (for the transposed products): no transpose is ever materialised. No atomics; deterministic.
//
// Row reductions need no shuffles here: a softmax thread owns a whole row (the TMEM lane), which is the
// tcgen05 counterpart of the warp-shuffle row reductions of the general-L kernels in attention.cu.
#include "common.cuh"
#include "internal.h"

#include <cuda.h>

namespace climb {
namespace {

constexpr int kDh = 64;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kRowB = 128;              // bytes per 64-element bf16 row

__device__ __forceinline__ uint32_t swz(uint32_t row, uint32_t granule) {      // offset inside a [rows x 128 B] tile
    return row * kRowB + ((granule ^ (row & 7u)) << 4);
}

// 32 fp32 (thread = row) -> bf16 -> 4 granules of the row at column chunk c32 (32 columns) of a
// [128 rows x 64 cols]-chunked K-major tile (16 KB per 64-column chunk)
__device__ __forceinline__ void store_row_chunk_bf16(uint8_t* tile, int row, int c32, const float (&v)[32]) {
    uint8_t* chunk = tile + (c32 >> 1) * (128 * kRowB);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        uint4 u;
        u.x = pack_bf16(v[8 * g], v[8 * g + 1]);
        u.y = pack_bf16(v[8 * g + 2], v[8 * g + 3]);
        u.z = pack_bf16(v[8 * g + 4], v[8 * g + 5]);
        u.w = pack_bf16(v[8 * g + 6], v[8 * g + 7]);
        *reinterpret_cast<uint4*>(chunk + swz(row, (c32 & 1) * 4 + g)) = u;
    }
}

// a warp's 32 rows x 32 bf16 columns (thread = row, 16 packed words = 64 B) -> global rows, coalesced through a
// 2 KB swizzled staging block: two rows share one 128 B staging line
__device__ __forceinline__ void store_rows_32(uint8_t* stage, const uint32_t (&pk)[16], __nv_bfloat16* gdst,
                                              long long ld_elems, int rows_valid, int lane) {
    __syncwarp();
#pragma unroll
    for (int g = 0; g < 4; ++g)
        *reinterpret_cast<uint4*>(stage + swz(lane >> 1, (lane & 1) * 4 + g)) =
            make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int q = it * 32 + lane;
        const int rr = q >> 2, g = q & 3;
        if (rr < rows_valid) {
            const uint4 val = *reinterpret_cast<const uint4*>(stage + swz(rr >> 1, (rr & 1) * 4 + g));
            *reinterpret_cast<uint4*>(gdst + rr * ld_elems + g * 8) = val;
        }
    }
    __syncwarp();
}

// column sums of the 32 x 32 block that store_rows_32 just staged: lane l owns column l
__device__ __forceinline__ void staged_colsum_32(const uint8_t* stage, int rows_valid, int lane, float* dst) {
    float cs = 0.0f;
    for (int rr = 0; rr < rows_valid; ++rr) {
        const uint16_t h16 = *reinterpret_cast<const uint16_t*>(stage + swz(rr >> 1, (rr & 1) * 4 + (lane >> 3)) + (lane & 7) * 2);
        cs += __uint_as_float(static_cast<uint32_t>(h16) << 16);
    }
    atomicAdd(dst + lane, cs);
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
struct FwdSmem {
    static constexpr int kP = 0;                 // P [128 x 256] bf16 = 64 KB, aliases Q (16 KB) + K (32 KB)
    static constexpr int kQ = 0;
    static constexpr int kK = 16 * 1024;
    static constexpr int kV = 64 * 1024;         // 32 KB; reused as output staging after O = P V
    static constexpr int kBias = 96 * 1024;      // 256 floats
    static constexpr int kXchg = 97 * 1024;      // row max / row sum exchange between the two column halves: 2 x 256 floats
    static constexpr int kBar = 99 * 1024;
    static constexpr int kTotal = 99 * 1024 + 128 + 1024;
};

constexpr int kSoftmaxThreads = 256;             // 8 warps: warp w and w + 4 share TMEM lanes 32*(w%4).., split the columns
constexpr int kCtlWarp = 8;
constexpr int kThreads = kSoftmaxThreads + 32;

__device__ __forceinline__ void softmax_bar_sync() {      // named barrier 1: the softmax warps only
    asm volatile("bar.sync 1, %0;" ::"n"(kSoftmaxThreads) : "memory");
}

__global__ void __launch_bounds__(kThreads, 2)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                   const float* __restrict__ key_bias, __nv_bfloat16* __restrict__ ctx, float* __restrict__ lse,
                   int L, int H, float scale_log2, uint32_t drop_thresh, float inv_keep, unsigned long long seed) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    float* sBias = reinterpret_cast<float*>(sm + FwdSmem::kBias);
    uint64_t* bar_load = reinterpret_cast<uint64_t*>(sm + FwdSmem::kBar);
    uint64_t* bar_s = bar_load + 1;
    uint64_t* bar_p = bar_load + 2;
    uint64_t* bar_o = bar_load + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_load + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
    pdl_wait();

    float* sXmax = reinterpret_cast<float*>(sm + FwdSmem::kXchg);       // [2][128]
    float* sXsum = sXmax + 256;                                          // [2][128]
    if (warp == kCtlWarp) {
        if (lane == 0) {
            tma_prefetch_desc(&map_q);
            tma_prefetch_desc(&map_kv);
            mbar_init(bar_load, 1);
            mbar_init(bar_s, 1);
            mbar_init(bar_p, kSoftmaxThreads);
            mbar_init(bar_o, 1);
            fence_barrier_init();
            // loads go out before anything else so that they overlap TMEM allocation and the prologue
            mbar_arrive_expect_tx(bar_load, (128 + 256 + 256) * kRowB);
            tma_load_3d(&map_q, bar_load, sm + FwdSmem::kQ, h * kDh, q0, b);
            tma_load_3d(&map_kv, bar_load, sm + FwdSmem::kK, (H + h) * kDh, 0, b);
            tma_load_3d(&map_kv, bar_load, sm + FwdSmem::kV, (2 * H + h) * kDh, 0, b);
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 256);
        tmem_relinquish();
    } else {
        const int j = threadIdx.x;
        sBias[j] = j < L ? (key_bias ? key_bias[static_cast<long long>(b) * L + j] * kLog2e : 0.0f) : -INFINITY;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == kCtlWarp) {
        if (lane == 0) {
            mbar_wait(bar_load, 0);
            tc_fence_after();
            const uint32_t sQ = smem_u32(sm + FwdSmem::kQ), sK = smem_u32(sm + FwdSmem::kK);
            const uint32_t idesc_s = umma_instr_desc(128, 256, 0, 0);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
                umma_bf16(tmem, umma_smem_desc(sQ + kk * 32, 16, 1024), umma_smem_desc(sK + kk * 32, 16, 1024), idesc_s,
                          kk > 0 ? 1u : 0u);
            umma_commit(bar_s);
            mbar_wait(bar_p, 0);
            tc_fence_after();
            const uint32_t sP = smem_u32(sm + FwdSmem::kP), sV = smem_u32(sm + FwdSmem::kV);
            const uint32_t idesc_o = umma_instr_desc(128, 64, 0, 1);
#pragma unroll
            for (int k = 0; k < 16; ++k)
                umma_bf16(tmem, umma_smem_desc(sP + (k >> 2) * (128 * kRowB) + (k & 3) * 32, 16, 1024),
                          umma_smem_desc(sV + k * (16 * kRowB), 256 * kRowB, 1024), idesc_o, k > 0 ? 1u : 0u);
            umma_commit(bar_o);
        }
        __syncwarp();
    } else {
        const int lg = warp & 3, half = warp >> 2;          // TMEM lane group, column half
        const int row = lg * 32 + lane;
        const uint32_t t_row = tmem + (static_cast<uint32_t>(lg * 32) << 16);
        mbar_wait(bar_s, 0);
        tc_fence_after();
        float m = -INFINITY;
#pragma unroll 1
        for (int c = half * 4; c < half * 4 + 4; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(t_row + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 b4 = *reinterpret_cast<const float4*>(sBias + c * 32 + j);
                m = fmaxf(m, fmaxf(fmaxf(fmaf(__uint_as_float(r[j]), scale_log2, b4.x), fmaf(__uint_as_float(r[j + 1]), scale_log2, b4.y)),
                                   fmaxf(fmaf(__uint_as_float(r[j + 2]), scale_log2, b4.z), fmaf(__uint_as_float(r[j + 3]), scale_log2, b4.w))));
            }
        }
        sXmax[half * 128 + row] = m;
        softmax_bar_sync();
        m = fmaxf(m, sXmax[(half ^ 1) * 128 + row]);       // key 0 is always valid: m is finite
        float sum = 0.0f;
#pragma unroll 1
        for (int c = half * 4; c < half * 4 + 4; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(t_row + c * 32, r);
            tmem_ld_wait();
            float p[32];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 b4 = *reinterpret_cast<const float4*>(sBias + c * 32 + j);
                p[j] = ex2_ftz(fmaf(__uint_as_float(r[j]), scale_log2, b4.x - m));
                p[j + 1] = ex2_ftz(fmaf(__uint_as_float(r[j + 1]), scale_log2, b4.y - m));
                p[j + 2] = ex2_ftz(fmaf(__uint_as_float(r[j + 2]), scale_log2, b4.z - m));
                p[j + 3] = ex2_ftz(fmaf(__uint_as_float(r[j + 3]), scale_log2, b4.w - m));
                sum += (p[j] + p[j + 1]) + (p[j + 2] + p[j + 3]);
            }
            if (drop_thresh != 0u) {
                // dropout on the probabilities (modeling_bert.py:341-345): the normaliser keeps every term, the
                // P V product sees the kept ones scaled by 1/(1-p). Counter = (b, h, query, key / 4).
                const unsigned long long base = ((static_cast<unsigned long long>(b) * H + h) * L + (q0 + row)) * 64ull + c * 8;
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const uint4 rnd = philox4x32(seed, base + g);
                    p[4 * g] *= dropout_scale(rnd.x, drop_thresh, inv_keep);
                    p[4 * g + 1] *= dropout_scale(rnd.y, drop_thresh, inv_keep);
                    p[4 * g + 2] *= dropout_scale(rnd.z, drop_thresh, inv_keep);
                    p[4 * g + 3] *= dropout_scale(rnd.w, drop_thresh, inv_keep);
                }
            }
            store_row_chunk_bf16(sm + FwdSmem::kP, row, c, p);     // Q / K are dead: S = Q K^T has retired
        }
        sXsum[half * 128 + row] = sum;
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(bar_p);
        softmax_bar_sync();
        sum += sXsum[(half ^ 1) * 128 + row];
        mbar_wait(bar_o, 0);
        tc_fence_after();
        const float inv = 1.0f / sum;
        // O epilogue: this warp converts 32 of the row's 64 output columns
        uint32_t pk[16];
        {
            uint32_t r[32];
            tmem_ld_32x32(t_row + half * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j)
                pk[j] = pack_bf16(__uint_as_float(r[2 * j]) * inv, __uint_as_float(r[2 * j + 1]) * inv);
        }
        const int qrow0 = q0 + lg * 32;
        const int rows_valid = min(32, max(0, L - qrow0));
        store_rows_32(sm + FwdSmem::kV + warp * 2048, pk,
                      ctx + (static_cast<long long>(b) * L + qrow0) * (H * kDh) + h * kDh + half * 32,
                      static_cast<long long>(H) * kDh, rows_valid, lane);
        if (half == 0 && q0 + row < L) lse[(static_cast<long long>(b) * H + h) * L + q0 + row] = (m + log2f(sum)) * kLn2;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kCtlWarp) {
        tc_fence_after();
        tmem_dealloc(tmem, 256);
    }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
struct BwdSmem {
    static constexpr int kQ = 0;                 // [256 x 64] bf16, 32 KB each
    static constexpr int kDO = 32 * 1024;
    static constexpr int kK = 64 * 1024;
    static constexpr int kV = 96 * 1024;
    static constexpr int kP = 128 * 1024;        // [128 x 128] bf16 = two 16 KB chunks of 64 columns
    static constexpr int kDS = 160 * 1024;
    static constexpr int kLse = 192 * 1024;      // 256 floats each
    static constexpr int kDelta = 193 * 1024;
    static constexpr int kBias = 194 * 1024;
    static constexpr int kColV = 195 * 1024;     // 64 floats: column sums of dO (= of dV, see below)
    static constexpr int kBar = 195 * 1024 + 256;
    static constexpr int kStage = 196 * 1024;    // 8 x 2 KB output staging (its own region: a warp that runs ahead into
                                                 // the next block rewrites P while slower warps still stage dK / dV)
    static constexpr int kTotal = 196 * 1024 + 8 * 2048 + 1024;
};
// TMEM columns
constexpr uint32_t kTS = 0, kTdP = 128, kTdV = 256, kTdK = 320, kTdQ = 384;

// 16 elementwise warps (four per TMEM lane group, one 32-key chunk of the 128-key block each): the P / dS
// phase sits between the two MMA phases of every block and nothing else can run meanwhile, so its latency
// is the kernel's; the first 8 of them also drain the dK / dV / dQ accumulators.
constexpr int kBwdEwWarps = 16;
constexpr int kBwdEwThreads = kBwdEwWarps * 32;
constexpr int kBwdOutThreads = 256;
constexpr int kBwdCtlWarp = kBwdEwWarps;
constexpr int kBwdThreads = kBwdEwThreads + 32;

__device__ __forceinline__ void tmem_ld_32x16_a(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

// Bias gradients of the q / k / v projections = column sums of dQ / dK / dV over all tokens:
//   dV:  sum_keys dV = sum_q dO (sum_keys P) = sum_q dO     (softmax rows sum to one) -> reduced from the dO rows
//        that the delta prologue reads anyway;
//   dK:  sum_keys dK = scale * sum_q (sum_keys dS) Q = 0     (sum_keys dS = delta - delta): the key bias shifts
//        every score of a row equally -- nothing is added (the reference's value is fp32 rounding noise);
//   dQ:  no shortcut: column sums of the staged dQ tiles.
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do,
                   const float* __restrict__ key_bias, const __nv_bfloat16* __restrict__ ctx,
                   const __nv_bfloat16* __restrict__ dctx, const float* __restrict__ lse,
                   __nv_bfloat16* __restrict__ dqkv, float* __restrict__ colsum, int L, int H, float scale_log2,
                   float scale) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space
    float* sLse = reinterpret_cast<float*>(sm + BwdSmem::kLse);
    float* sDelta = reinterpret_cast<float*>(sm + BwdSmem::kDelta);
    float* sBias = reinterpret_cast<float*>(sm + BwdSmem::kBias);
    float* sColV = reinterpret_cast<float*>(sm + BwdSmem::kColV);
    uint64_t* bar_load = reinterpret_cast<uint64_t*>(sm + BwdSmem::kBar);
    uint64_t* bar_a = bar_load + 1;      // S, dP of a block are in TMEM
    uint64_t* bar_p = bar_load + 2;      // P, dS of a block are in smem (kBwdEwThreads arrivals)
    uint64_t* bar_b = bar_load + 3;      // dV / dK / dQ chains of a block have retired
    uint64_t* bar_kv = bar_load + 4;     // dV_j, dK_j have been read out of TMEM (kBwdOutThreads arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_load + 5);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.x, b = blockIdx.y;
    pdl_wait();
    // 768 CTAs on 148 SMs = 5.19 waves: once the last CTA has started, the next kernel of the stream may be
    // scheduled onto the SMs the final partial wave leaves idle (it still waits for this grid unless it was
    // declared independent: the attention-output wgrad is)
    pdl_launch_dependents();
    const long long ld = 3LL * H * kDh, ldo = static_cast<long long>(H) * kDh;
    const int n_jt = (L + 127) / 128;            // key / query tiles that contain real rows (1 or 2)

    float cv_keep[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) cv_keep[e] = 0.0f;
    if (warp == kBwdCtlWarp) {
        if (lane == 0) {
            tma_prefetch_desc(&map_qkv);
            tma_prefetch_desc(&map_do);
            mbar_init(bar_load, 1);
            mbar_init(bar_a, 1);
            mbar_init(bar_p, kBwdEwThreads);
            mbar_init(bar_b, 1);
            mbar_init(bar_kv, kBwdOutThreads);
            fence_barrier_init();
            // loads go out first: they overlap TMEM allocation and the delta prologue of the other warps
            mbar_arrive_expect_tx(bar_load, 4 * 256 * kRowB);
            tma_load_3d(&map_qkv, bar_load, sm + BwdSmem::kQ, h * kDh, 0, b);
            tma_load_3d(&map_do, bar_load, sm + BwdSmem::kDO, h * kDh, 0, b);
            tma_load_3d(&map_qkv, bar_load, sm + BwdSmem::kK, (H + h) * kDh, 0, b);
            tma_load_3d(&map_qkv, bar_load, sm + BwdSmem::kV, (2 * H + h) * kDh, 0, b);
        }
        if (lane < 16) reinterpret_cast<float4*>(sColV)[lane] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    } else {
        // per-row scalars: lse (log2 domain; +inf past L makes P = 0 there), key mask, and
        // delta = rowsum(dO * O) with coalesced loads: 8 lanes x 16 B cover one 128 B row, 4 rows per warp access
        const long long stat = (static_cast<long long>(b) * H + h) * L;
        if (threadIdx.x < 256) {
            const int r = threadIdx.x;
            sLse[r] = r < L ? lse[stat + r] * kLog2e : INFINITY;
            sBias[r] = r < L ? (key_bias ? key_bias[static_cast<long long>(b) * L + r] * kLog2e : 0.0f) : -INFINITY;
        }
        const int sub = lane & 7, rsel = lane >> 3;
        float cv[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) cv[e] = 0.0f;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int r = warp * 16 + it * 4 + rsel;
            float dl = 0.0f;
            if (r < L) {
                const long long off = (static_cast<long long>(b) * L + r) * ldo + h * kDh + sub * 8;
                const uint4 a = *reinterpret_cast<const uint4*>(ctx + off);
                const uint4 d = *reinterpret_cast<const uint4*>(dctx + off);
                const uint32_t av[4] = {a.x, a.y, a.z, a.w}, dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 x = unpack_bf16(av[e]), y = unpack_bf16(dv[e]);
                    dl = fmaf(x.x, y.x, dl);
                    dl = fmaf(x.y, y.y, dl);
                    cv[2 * e] += y.x;
                    cv[2 * e + 1] += y.y;
                }
            }
            dl += __shfl_xor_sync(0xffffffffu, dl, 1);
            dl += __shfl_xor_sync(0xffffffffu, dl, 2);
            dl += __shfl_xor_sync(0xffffffffu, dl, 4);
            if (sub == 0) sDelta[r] = dl;
        }
        if (colsum != nullptr) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                cv[e] += __shfl_xor_sync(0xffffffffu, cv[e], 8);
                cv[e] += __shfl_xor_sync(0xffffffffu, cv[e], 16);
                cv_keep[e] = cv[e];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp != kBwdCtlWarp && colsum != nullptr) {
        // v-bias gradient: warp partials (lanes 0..7 hold 8 columns each) -> CTA sums in smem -> 64 global atomics
        if (lane < 8) {
#pragma unroll
            for (int e = 0; e < 8; ++e) atomicAdd(sColV + lane * 8 + e, cv_keep[e]);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kBwdEwThreads) : "memory");
        if (threadIdx.x < 64) atomicAdd(colsum + 2 * H * kDh + h * kDh + threadIdx.x, sColV[threadIdx.x]);
    }
    const uint32_t tmem = *tmem_slot;
    const int n_blocks = n_jt * n_jt;

    if (warp == kBwdCtlWarp) {
        if (lane == 0) {
            mbar_wait(bar_load, 0);
            tc_fence_after();
            const uint32_t sQ = smem_u32(sm + BwdSmem::kQ), sDO = smem_u32(sm + BwdSmem::kDO);
            const uint32_t sK = smem_u32(sm + BwdSmem::kK), sV = smem_u32(sm + BwdSmem::kV);
            const uint32_t sP = smem_u32(sm + BwdSmem::kP), sDS = smem_u32(sm + BwdSmem::kDS);
            const uint32_t id_a = umma_instr_desc(128, 128, 0, 0);     // S / dP: both operands K-major
            const uint32_t id_t = umma_instr_desc(128, 64, 1, 1);      // dV / dK: A = P^T / dS^T in place, B MN-major
            const uint32_t id_q = umma_instr_desc(128, 64, 0, 1);      // dQ: A = dS K-major, B = K MN-major
            const uint32_t tile = 128 * kRowB;                          // 16 KB: 128 rows of a [rows x 64] tile
            // phase A of block n: S = Q_i K_j^T, dP = dO_i V_j^T
            auto phase_a = [&](int n) {
                const int j = n / n_jt, i = n - j * n_jt;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_bf16(tmem + kTS, umma_smem_desc(sQ + i * tile + kk * 32, 16, 1024),
                              umma_smem_desc(sK + j * tile + kk * 32, 16, 1024), id_a, kk > 0 ? 1u : 0u);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_bf16(tmem + kTdP, umma_smem_desc(sDO + i * tile + kk * 32, 16, 1024),
                              umma_smem_desc(sV + j * tile + kk * 32, 16, 1024), id_a, kk > 0 ? 1u : 0u);
                umma_commit(bar_a);
            };
            phase_a(0);
            for (int n = 0; n < n_blocks; ++n) {
                const int j = n / n_jt, i = n - j * n_jt;
                mbar_wait(bar_p, n & 1);            // P, dS of block n are in smem; S / dP have been read out of TMEM
                tc_fence_after();
                // the NEXT block's S / dP chains go first: the elementwise warps work on them while the three
                // accumulation chains of this block run (they only need P / dS's smem back before they store)
                if (n + 1 < n_blocks) phase_a(n + 1);
                if (i == 0 && j > 0) {                  // previous key tile's dV / dK must have been read out
                    mbar_wait(bar_kv, (j - 1) & 1);
                    tc_fence_after();
                }
                // phase B: k runs over the 128 query rows (dV, dK) or the 128 keys (dQ), 16 per UMMA
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    umma_bf16(tmem + kTdV, umma_smem_desc(sP + k * (16 * kRowB), tile, 1024),
                              umma_smem_desc(sDO + i * tile + k * (16 * kRowB), tile, 1024), id_t, (i > 0 || k > 0) ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    umma_bf16(tmem + kTdK, umma_smem_desc(sDS + k * (16 * kRowB), tile, 1024),
                              umma_smem_desc(sQ + i * tile + k * (16 * kRowB), tile, 1024), id_t, (i > 0 || k > 0) ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    umma_bf16(tmem + kTdQ + i * 64, umma_smem_desc(sDS + (k >> 2) * tile + (k & 3) * 32, 16, 1024),
                              umma_smem_desc(sK + j * tile + k * (16 * kRowB), tile, 1024), id_q, (j > 0 || k > 0) ? 1u : 0u);
                umma_commit(bar_b);
            }
        }
        __syncwarp();
    } else {
        const int lg = warp & 3, quarter = warp >> 2;       // TMEM lane group, 32-key chunk of the block
        const int row = lg * 32 + lane;
        const uint32_t t_row = tmem + (static_cast<uint32_t>(lg * 32) << 16);
        const bool out_warp = warp < 8;                     // drains accumulators: lane group lg, column half (warp >> 2)
        const int half = quarter & 1;
        uint8_t* stage = sm + BwdSmem::kStage + (warp & 7) * 2048;
        for (int n = 0; n < n_blocks; ++n) {
            const int j = n / n_jt, i = n - j * n_jt;
            mbar_wait(bar_a, n & 1);
            tc_fence_after();
            const float lse_r = sLse[i * 128 + row], dl_r = sDelta[i * 128 + row];
            const int c = quarter;
            uint32_t pp[2][8], dd[2][8];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {                    // two 16-key halves: 32 accumulator registers live
                uint32_t rs[16], rd[16];
                tmem_ld_32x16_a(t_row + kTS + c * 32 + hh * 16, rs);
                tmem_ld_32x16_a(t_row + kTdP + c * 32 + hh * 16, rd);
                float bias[16];
#pragma unroll
                for (int e = 0; e < 16; e += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(sBias + j * 128 + c * 32 + hh * 16 + e);
                    bias[e] = b4.x - lse_r; bias[e + 1] = b4.y - lse_r; bias[e + 2] = b4.z - lse_r; bias[e + 3] = b4.w - lse_r;
                }
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 16; e += 2) {
                    const float p0 = ex2_ftz(fmaf(__uint_as_float(rs[e]), scale_log2, bias[e]));
                    const float p1 = ex2_ftz(fmaf(__uint_as_float(rs[e + 1]), scale_log2, bias[e + 1]));
                    pp[hh][e >> 1] = pack_bf16(p0, p1);
                    dd[hh][e >> 1] = pack_bf16(p0 * (__uint_as_float(rd[e]) - dl_r), p1 * (__uint_as_float(rd[e + 1]) - dl_r));
                }
            }
            // everything above overlapped the previous block's accumulation chains; they must have retired before
            // P / dS are overwritten
            if (n > 0) mbar_wait(bar_b, (n - 1) & 1);
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                // 16 columns = granules (c & 1) * 4 + hh * 2 + {0, 1} of the row in 64-column chunk (c >> 1)
                uint8_t* pc = sm + BwdSmem::kP + (c >> 1) * (128 * kRowB);
                uint8_t* dc = sm + BwdSmem::kDS + (c >> 1) * (128 * kRowB);
                const int g0 = (c & 1) * 4 + hh * 2;
                *reinterpret_cast<uint4*>(pc + swz(row, g0)) = make_uint4(pp[hh][0], pp[hh][1], pp[hh][2], pp[hh][3]);
                *reinterpret_cast<uint4*>(pc + swz(row, g0 + 1)) = make_uint4(pp[hh][4], pp[hh][5], pp[hh][6], pp[hh][7]);
                *reinterpret_cast<uint4*>(dc + swz(row, g0)) = make_uint4(dd[hh][0], dd[hh][1], dd[hh][2], dd[hh][3]);
                *reinterpret_cast<uint4*>(dc + swz(row, g0 + 1)) = make_uint4(dd[hh][4], dd[hh][5], dd[hh][6], dd[hh][7]);
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(bar_p);
            if (i == n_jt - 1 && out_warp) {
                // dV_j, dK_j are complete once this block's chains retire; each warp converts 32 of the 64 columns
                mbar_wait(bar_b, n & 1);
                tc_fence_after();
                const int key0 = j * 128 + lg * 32;
                const int rows_valid = min(32, max(0, L - key0));
                __nv_bfloat16* dst = dqkv + (static_cast<long long>(b) * L + key0) * ld + h * kDh + half * 32;
#pragma unroll
                for (int which = 0; which < 2; ++which) {          // 0: dK (scaled), 1: dV
                    const uint32_t col = which == 0 ? kTdK : kTdV;
                    const float sc = which == 0 ? scale : 1.0f;
                    uint32_t r[32], pk[16];
                    tmem_ld_32x32(t_row + col + half * 32, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        pk[e] = pack_bf16(__uint_as_float(r[2 * e]) * sc, __uint_as_float(r[2 * e + 1]) * sc);
                    store_rows_32(stage, pk, dst + (which == 0 ? H * kDh : 2 * H * kDh), ld, rows_valid, lane);
                }
                tc_fence_before();
                mbar_arrive(bar_kv);
            }
        }
        // dQ tiles: complete after the last block
        if (out_warp) {
            mbar_wait(bar_b, (n_blocks - 1) & 1);
            tc_fence_after();
            for (int i = 0; i < n_jt; ++i) {
                const int q0 = i * 128 + lg * 32;
                const int rows_valid = min(32, max(0, L - q0));
                uint32_t r[32], pk[16];
                tmem_ld_32x32(t_row + kTdQ + i * 64 + half * 32, r);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 16; ++e)
                    pk[e] = pack_bf16(__uint_as_float(r[2 * e]) * scale, __uint_as_float(r[2 * e + 1]) * scale);
                store_rows_32(stage, pk, dqkv + (static_cast<long long>(b) * L + q0) * ld + h * kDh + half * 32, ld, rows_valid, lane);
                if (colsum) staged_colsum_32(stage, rows_valid, lane, colsum + h * kDh + half * 32);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kBwdCtlWarp) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

int make_map3(CUtensorMap* map, const void* ptr, int B, int L, long long row_elems, int box_rows) {
    const long long dims[3] = {row_elems, L, B};
    const long long strides[2] = {row_elems, static_cast<long long>(L) * row_elems};
    const int box[3] = {64, box_rows, 1};
    return encode_tmap_bf16(map, ptr, 3, dims, strides, box);
}

}  // namespace

int attention_tc_fwd(const void* qkv, const float* key_bias, void* ctx, float* lse, int B, int L, int H, float scale,
                     cudaStream_t stream, float p_drop, unsigned long long seed) {
    CLIMB_REQUIRE(L <= 256, "attention_tc_fwd: L=%d > 256", L);
    CLIMB_REQUIRE(p_drop >= 0.0f && p_drop < 1.0f, "attention_tc_fwd: dropout p=%f outside [0, 1)", p_drop);
    CUtensorMap mq, mkv;
    int rc = make_map3(&mq, qkv, B, L, 3LL * H * kDh, 128);
    if (rc) return rc;
    rc = make_map3(&mkv, qkv, B, L, 3LL * H * kDh, 256);
    if (rc) return rc;
    static bool attr = false;
    if (!attr) {
        CLIMB_CUDA_OK(cudaFuncSetAttribute(attn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem::kTotal));
        attr = true;
    }
    dim3 grid((L + 127) / 128, H, B);
    CLIMB_CUDA_OK(launch_pdl(attn_tc_fwd_kernel, grid, dim3(kThreads), FwdSmem::kTotal, stream, mq, mkv, key_bias,
                             static_cast<__nv_bfloat16*>(ctx), lse, L, H, scale * kLog2e,
                             p_drop > 0.0f ? dropout_threshold(p_drop) : 0u, 1.0f / (1.0f - p_drop), seed));
    CLIMB_LAUNCH_OK();
    return 0;
}

int attention_tc_bwd(const void* qkv, const float* key_bias, const void* ctx, const void* dctx, const float* lse,
                     void* dqkv, float* colsum, int B, int L, int H, float scale, cudaStream_t stream) {
    CLIMB_REQUIRE(L <= 256, "attention_tc_bwd: L=%d > 256", L);
    CUtensorMap mqkv, mdo;
    int rc = make_map3(&mqkv, qkv, B, L, 3LL * H * kDh, 256);
    if (rc) return rc;
    rc = make_map3(&mdo, dctx, B, L, static_cast<long long>(H) * kDh, 256);
    if (rc) return rc;
    static bool attr = false;
    if (!attr) {
        CLIMB_CUDA_OK(cudaFuncSetAttribute(attn_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem::kTotal));
        attr = true;
    }
    dim3 grid(H, B);
    CLIMB_CUDA_OK(launch_pdl(attn_tc_bwd_kernel, grid, dim3(kBwdThreads), BwdSmem::kTotal, stream, mqkv, mdo, key_bias,
                             static_cast<const __nv_bfloat16*>(ctx), static_cast<const __nv_bfloat16*>(dctx), lse,
                             static_cast<__nv_bfloat16*>(dqkv), colsum, L, H, scale * kLog2e, scale));
    CLIMB_LAUNCH_OK();
    return 0;
}

}  // namespace climb
