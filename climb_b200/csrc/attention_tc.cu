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

#include <algorithm>
#include <cstdlib>

// CLIMB_ATTN_TIMELINE=1 (compile time, dev only) makes CTA 0 of the persistent kernels record clock64() at the hand-over
// points of its third item; CLIMB_ATTN_TL=1 (environment) then prints the cycle stamps after the 20th launch. That timeline is
// what located the kernels' limits (attention section); the default build carries none of it.
#ifndef CLIMB_ATTN_TIMELINE
#define CLIMB_ATTN_TIMELINE 0
#endif

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

// column sums of the 32 x 32 bf16 block that store_rows_32 just staged (rows past the valid ones hold exact zeros here):
// lane = (row group of 4, 8-column granule): four independent 16 B loads, then a butterfly over the eight row groups
__device__ __forceinline__ void staged_colsum_32(const uint8_t* stage, int lane, float* dst) {
    const int g = lane & 3, rg = lane >> 2;
    float cs[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) cs[e] = 0.0f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int rr = rg * 4 + q;
        const uint4 v = *reinterpret_cast<const uint4*>(stage + swz(rr >> 1, (rr & 1) * 4 + g));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 f = unpack_bf16(w[e]);
            cs[2 * e] += f.x;
            cs[2 * e + 1] += f.y;
        }
    }
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
#pragma unroll
        for (int e = 0; e < 8; ++e) cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], o);
    }
    if (lane < 4) {
#pragma unroll
        for (int e = 0; e < 8; ++e) atomicAdd(dst + lane * 8 + e, cs[e]);
    }
    __syncwarp();
}

// One mbarrier arrival per WARP: the lanes order their own shared-memory writes (fence.proxy.async) and TMEM reads
// (tcgen05.wait::ld) first, meet at __syncwarp, and one lane arrives. Hundreds of per-thread arrivals on one barrier word
// serialise in the shared-memory atomic unit -- measured here, they were most of the kernels' time.
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}

__device__ __forceinline__ void tmem_ld_32x16_a(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

// A shared-memory descriptor split into its words: the start-address field (bits 0..13 of the low word, in 16-byte units)
// is the only thing that changes between the instructions of a chain, so an operand costs one 32-bit add
struct DescBase {
    uint32_t lo, hi;
    __device__ explicit DescBase(uint64_t d) : lo(static_cast<uint32_t>(d)), hi(static_cast<uint32_t>(d >> 32)) {}
};
__device__ __forceinline__ void umma_bf16_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
        "}\n"
        :
        : "r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}

// 2^x for x <= 0 on the FMA pipe (two values per instruction): the MUFU pipe runs 16 ex2 per clock and SM (measured,
// tools/probe/mufu_probe.cu) and is the softmax's busiest unit; the packed-FMA evaluation runs beside it.
// x = n + f, n = round(x) through the 1.5 * 2^23 trick, |f| <= 1/2; 2^f by a degree-4 minimax polynomial (relative error
// 2.7e-6, far inside the bf16 rounding of P); n goes straight into the exponent field. x is clamped at -126.
__device__ __forceinline__ void exp2_poly2(float x0, float x1, float& y0, float& y1) {
    x0 = fmaxf(x0, -126.0f);
    x1 = fmaxf(x1, -126.0f);
    const uint64_t x = pack_f32x2(x0, x1);
    const uint64_t t = fadd2(x, pack_f32x2(12582912.0f, 12582912.0f));
    const uint64_t f = fadd2(x, fadd2(pack_f32x2(12582912.0f, 12582912.0f), t ^ 0x8000000080000000ull));     // x - (t - magic)
    uint64_t pl = ffma2(pack_f32x2(0.009570099413394928f, 0.009570099413394928f), f, pack_f32x2(0.05591785907745361f, 0.05591785907745361f));
    pl = ffma2(pl, f, pack_f32x2(0.240247443318367f, 0.240247443318367f));
    pl = ffma2(pl, f, pack_f32x2(0.6931217908859253f, 0.6931217908859253f));
    pl = ffma2(pl, f, pack_f32x2(0.9999992847442627f, 0.9999992847442627f));
    float p0, p1, t0, t1;
    unpack_f32x2(pl, p0, p1);
    unpack_f32x2(t, t0, t1);
    y0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
    y1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
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
            mbar_init(bar_p, kSoftmaxThreads / 32);
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
        warp_arrive(bar_p, lane);
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
// forward, persistent (the encoder's hot call: no dropout)
//
// One CTA per SM walks over (b, h) items; both 128-row query tiles of an item are in flight at once, each owned by
// its own softmax group (8 warps: thread = row x column half) and its own MMA-issuing warp, so that one tile's
// exponentials overlap the other tile's tensor-core chains, accumulator drain and output stores:
//   producer warp   TMA: Q [256 x 64] + K [256 x 64] (single buffers, refilled as soon as both S chains of the
//                   item have retired = one item ahead of their use), V [256 x 64] in two stages
//   MMA warp g      S_g = Q_g K^T (128 x 256 x 64, TMEM columns g*256..), then O_g += P_g[:, chunk] V[chunk, :]
//                   for the four 64-key chunks as the softmax group hands them over through a two-slot ring
//                   (O_g aliases S_g's first 64 columns: chunk 0 of S has been consumed by then)
//   softmax group g row max over the row held in TMEM (two threads per row exchange through smem), exp2, P as
//                   bf16 into the ring slot in the K-major A-operand layout, 1 / sum folded into the O epilogue
// Every K / V byte is read once per item (the one-tile-per-CTA kernel above reads them twice).
// ------------------------------------------------------------------------------------------------
struct Fwd2Smem {
    static constexpr int kQ = 0;                    // [256 x 64] bf16: query tile g at g * 16 KB
    static constexpr int kK = 32 * 1024;            // [256 x 64]
    static constexpr int kV = 64 * 1024;            // 2 stages x [256 x 64]
    static constexpr int kP = 128 * 1024;           // 2 groups x 2 slots x [128 x 64] bf16; slot 0 doubles as output staging
    static constexpr int kBias = 192 * 1024;        // [2 groups][2 stages][256] floats
    static constexpr int kXchg = 196 * 1024;        // [2 groups][max, sum][2 halves][128] floats
    static constexpr int kFlag = 200 * 1024;        // [2 groups][2 stages][8] words: this 32-key chunk has a non-zero mask
    static constexpr int kBar = 200 * 1024 + 128;
    static constexpr int kTotal = 200 * 1024 + 256 + 1024;
};
constexpr int kF2GroupThreads = 256;
constexpr int kF2ProducerWarp = 16;                 // + TMEM allocation
constexpr int kF2MmaWarp0 = 17;                     // MMA warp of group g = 17 + g
constexpr int kF2Threads = 19 * 32;

template <int kId>
__device__ __forceinline__ void named_bar_sync(int id_offset, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(kId + id_offset), "r"(threads) : "memory");
}

__global__ void __launch_bounds__(kF2Threads, 1)
attn_tc_fwd2_kernel(const __grid_constant__ CUtensorMap map_qkv, const float* __restrict__ key_bias,
                    __nv_bfloat16* __restrict__ ctx, float* __restrict__ lse, int n_items, int L, int H, float scale_log2,
                    int stagger, long long* tl) {
#if CLIMB_ATTN_TIMELINE
#define TL(role, idx) do { if (tl != nullptr && blockIdx.x == 0 && it == 2) tl[(role) * 32 + (idx)] = clock64(); } while (0)
#else
#define TL(role, idx) do { } while (0)
#endif
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + Fwd2Smem::kBar);
    uint64_t* qk_full = bars;            // [1]  tx
    uint64_t* qk_free = bars + 1;        // [1]  n_groups commits
    uint64_t* v_full = bars + 2;         // [2]  tx
    uint64_t* v_free = bars + 4;         // [2]  n_groups commits
    uint64_t* s_full = bars + 6;         // [2 groups] commit
    uint64_t* s_free = bars + 8;         // [2 groups] 8 warp arrivals: O has been read out of TMEM
    uint64_t* p_full = bars + 10;        // [2 groups][2 slots] 8 warp arrivals
    uint64_t* p_free = bars + 14;        // [2 groups][2 slots] commit
    uint64_t* o_full = bars + 18;        // [2 groups] commit
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_groups = L > 128 ? 2 : 1;            // query tiles with real rows
    const int n_my = (n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    pdl_wait();
    pdl_launch_dependents();     // the next kernel's CTAs may take the SMs whose CTA here has run out of items

    if (warp == kF2ProducerWarp) {
        if (lane == 0) {
            tma_prefetch_desc(&map_qkv);
            mbar_init(qk_full, 1);
            mbar_init(qk_free, n_groups);
            for (int s = 0; s < 2; ++s) {
                mbar_init(&v_full[s], 1);
                mbar_init(&v_free[s], n_groups);
                mbar_init(&s_full[s], 1);
                mbar_init(&s_free[s], kF2GroupThreads / 32);
                mbar_init(&o_full[s], 1);
            }
            for (int s = 0; s < 4; ++s) {
                mbar_init(&p_full[s], kF2GroupThreads / 32);
                mbar_init(&p_free[s], 1);
            }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == kF2ProducerWarp) {
        if (lane == 0) {
            for (int it = 0; it < n_my; ++it) {
                const int item = blockIdx.x + it * gridDim.x;
                const int b = item / H, h = item - b * H;
                if (it > 0) mbar_wait(qk_free, (it - 1) & 1);
                mbar_arrive_expect_tx(qk_full, 2 * 256 * kRowB);
                tma_load_3d(&map_qkv, qk_full, sm + Fwd2Smem::kQ, h * kDh, 0, b);
                tma_load_3d(&map_qkv, qk_full, sm + Fwd2Smem::kK, (H + h) * kDh, 0, b);
                const int s = it & 1, k = it >> 1;
                if (k > 0) mbar_wait(&v_free[s], (k - 1) & 1);
                mbar_arrive_expect_tx(&v_full[s], 256 * kRowB);
                tma_load_3d(&map_qkv, &v_full[s], sm + Fwd2Smem::kV + s * (256 * kRowB), (2 * H + h) * kDh, 0, b);
            }
        }
        __syncwarp();
    } else if (warp >= kF2MmaWarp0) {
        const int g = warp - kF2MmaWarp0;
        if (lane == 0 && g < n_groups) {
            const DescBase dQ(umma_smem_desc(smem_u32(sm + Fwd2Smem::kQ) + g * (128 * kRowB), 16, 1024));
            const DescBase dK(umma_smem_desc(smem_u32(sm + Fwd2Smem::kK), 16, 1024));
            const DescBase dP(umma_smem_desc(smem_u32(sm + Fwd2Smem::kP) + g * (2 * 128 * kRowB), 16, 1024));
            const DescBase dV(umma_smem_desc(smem_u32(sm + Fwd2Smem::kV), 256 * kRowB, 1024));       // V read in place as an MN-major B operand
            const uint32_t t_acc = tmem + g * 256;
            const uint32_t idesc_s = umma_instr_desc(128, 256, 0, 0);
            const uint32_t idesc_o = umma_instr_desc(128, 64, 0, 1);
            // the two groups run half a tile apart (group 1 starts once group 0 has handed over its second chunk), so that
            // one group's exponentials (the MUFU pipe is the busiest unit) meet the other group's max pass, epilogue and
            // barrier round trips instead of its exponentials; nothing later couples the groups, so the offset stays
            if (g == 1 && stagger) mbar_wait(&p_full[1], 0);
            for (int it = 0; it < n_my; ++it) {
                TL(g, 0);
                mbar_wait(qk_full, it & 1);
                TL(g, 1);
                if (it > 0) mbar_wait(&s_free[g], (it - 1) & 1);
                TL(g, 2);
                tc_fence_after();
#pragma unroll
                for (uint32_t kk = 0; kk < 4; ++kk)
                    umma_bf16_lohi(t_acc, dQ.lo + kk * 2, dQ.hi, dK.lo + kk * 2, dK.hi, idesc_s, kk > 0 ? 1u : 0u);
                umma_commit(&s_full[g]);
                umma_commit(qk_free);
                TL(g, 3);
                const int s = it & 1;
                const uint32_t v_lo = dV.lo + s * ((256 * kRowB) >> 4);
                mbar_wait(&v_full[s], (it >> 1) & 1);
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    const int slot = c & 1;
                    mbar_wait(&p_full[g * 2 + slot], (2 * it + (c >> 1)) & 1);
                    TL(g, 4 + 2 * c);
                    tc_fence_after();
#pragma unroll
                    for (uint32_t kk = 0; kk < 4; ++kk)
                        umma_bf16_lohi(t_acc, dP.lo + slot * ((128 * kRowB) >> 4) + kk * 2, dP.hi, v_lo + (c * 4 + kk) * ((16 * kRowB) >> 4), dV.hi,
                                       idesc_o, (c > 0 || kk > 0) ? 1u : 0u);
                    umma_commit(&p_free[g * 2 + slot]);
                    TL(g, 5 + 2 * c);
                }
                umma_commit(&o_full[g]);
                umma_commit(&v_free[s]);
            }
        }
        __syncwarp();
    } else if ((warp >> 3) < n_groups) {
        const int g = warp >> 3, wg = warp & 7;              // group, warp inside the group
        const int lg = wg & 3, half = wg >> 2;               // TMEM lane group, 32-key half of every 64-key chunk
        const int row = lg * 32 + lane;
        const int tg = threadIdx.x & (kF2GroupThreads - 1);
        const uint32_t t_row = tmem + (static_cast<uint32_t>(lg * 32) << 16) + g * 256;
        float* sBias = reinterpret_cast<float*>(sm + Fwd2Smem::kBias) + g * 512;                 // [2 stages][256]
        float* sXmax = reinterpret_cast<float*>(sm + Fwd2Smem::kXchg) + g * 512;                 // [2][128]
        float* sXsum = sXmax + 256;
        uint8_t* ring = sm + Fwd2Smem::kP + g * (2 * 128 * kRowB);
        uint32_t* sFlag = reinterpret_cast<uint32_t*>(sm + Fwd2Smem::kFlag) + g * 16;          // [2 stages][8 key chunks]
        // additive key mask of item `it` for key tg, still in natural-log units (the scaling waits until the value is
        // stored, so that nothing stalls on the load): -inf past L
        auto load_bias = [&](int it) -> float {
            if (tg >= L) return -INFINITY;
            if (key_bias == nullptr) return 0.0f;
            const int item = blockIdx.x + it * gridDim.x;
            return __ldg(key_bias + static_cast<long long>(item / H) * L + tg);
        };
        // a 32-key chunk whose mask is all zero (the usual case away from the text padding and the tail past L) takes
        // the short path: max over the raw scores, one packed FMA per pair of scores
        auto store_bias = [&](int stage, float v) {
            sBias[stage * 256 + tg] = v * kLog2e;
            const uint32_t any = __ballot_sync(0xffffffffu, v != 0.0f);
            if (lane == 0) sFlag[stage * 8 + wg] = any;
        };
        if (n_my > 0) store_bias(0, load_bias(0));
        const uint64_t scale2 = pack_f32x2(scale_log2, scale_log2);
        for (int it = 0; it < n_my; ++it) {
            const int item = blockIdx.x + it * gridDim.x;
            const int b = item / H, h = item - b * H;
            const float bias_next = it + 1 < n_my ? load_bias(it + 1) : 0.0f;
            // the group meets here once per item: the bias row written during the previous item is visible, and every
            // warp has finished reading its output staging block (ring slot 0) before chunk 0 is written again
#define TLS(idx) do { if (wg == 0 && lane == 0) TL(2 + g, idx); } while (0)
            TLS(0);
            named_bar_sync<9>(g, kF2GroupThreads);
            TLS(1);
            const float* bias = sBias + (it & 1) * 256;
            const uint32_t* flag = sFlag + (it & 1) * 8;
            mbar_wait(&s_full[g], it & 1);
            TLS(2);
            tc_fence_after();
            // ---- pass 1: row max. The four 32-column loads alternate between two register sets, each issued before the
            //      previous one is reduced, so only the first load's latency is exposed ----
            float m = -INFINITY, m_raw = -INFINITY;
            auto row_max = [&](const uint32_t (&r)[32], int c) {
                if (flag[c * 2 + half] == 0u) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        m_raw = fmaxf(m_raw, fmaxf(fmaxf(__uint_as_float(r[j]), __uint_as_float(r[j + 1])),
                                                   fmaxf(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]))));
                } else {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(bias + c * 64 + half * 32 + j);
                        m = fmaxf(m, fmaxf(fmaxf(fmaf(__uint_as_float(r[j]), scale_log2, b4.x), fmaf(__uint_as_float(r[j + 1]), scale_log2, b4.y)),
                                           fmaxf(fmaf(__uint_as_float(r[j + 2]), scale_log2, b4.z), fmaf(__uint_as_float(r[j + 3]), scale_log2, b4.w))));
                    }
                }
            };
            uint32_t ra[16], rb[16];
            {
                uint32_t r0[32], r1[32];
                tmem_ld_32x32(t_row + half * 32, r0);
                tmem_ld_wait();
                tmem_ld_32x32(t_row + 64 + half * 32, r1);
                row_max(r0, 0);
                tmem_ld_wait();
                tmem_ld_32x32(t_row + 128 + half * 32, r0);
                row_max(r1, 1);
                tmem_ld_wait();
                tmem_ld_32x32(t_row + 192 + half * 32, r1);
                row_max(r0, 2);
                tmem_ld_wait();
                tmem_ld_32x16_a(t_row + half * 32, ra);          // pass 2's first 16 columns travel during the exchange below
                row_max(r1, 3);
            }
            m = fmaxf(m, m_raw * scale_log2);                  // scale > 0
            sXmax[half * 128 + row] = m;
            named_bar_sync<1>(g * 4 + lg, 64);
            m = fmaxf(m, sXmax[(half ^ 1) * 128 + row]);       // key 0 is always valid: m is finite
            TLS(3);
            // ---- pass 2: P = exp2(S * scale + mask - m) in 16-column steps; the next step's load is always in flight, and the
            //      next chunk's first step is issued before this chunk's store / fence / hand-over ----
            const uint64_t neg_m2 = pack_f32x2(-m, -m);
            uint64_t sum2 = 0ull;
            // 16 scores -> 16 probabilities: mask-free chunks use one packed FMA per pair and evaluate every other pair on the
            // FMA pipe instead of MUFU
            auto probs16 = [&](const uint32_t (&r)[16], float* p, bool masked, const float* bias16) {
                if (!masked) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        float x0, x1, x2, x3;
                        unpack_f32x2(ffma2(pack_u32x2(r[j], r[j + 1]), scale2, neg_m2), x0, x1);
                        unpack_f32x2(ffma2(pack_u32x2(r[j + 2], r[j + 3]), scale2, neg_m2), x2, x3);
                        p[j] = ex2_ftz(x0);
                        p[j + 1] = ex2_ftz(x1);
                        exp2_poly2(x2, x3, p[j + 2], p[j + 3]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(bias16 + j);
                        p[j] = ex2_ftz(fmaf(__uint_as_float(r[j]), scale_log2, b4.x - m));
                        p[j + 1] = ex2_ftz(fmaf(__uint_as_float(r[j + 1]), scale_log2, b4.y - m));
                        p[j + 2] = ex2_ftz(fmaf(__uint_as_float(r[j + 2]), scale_log2, b4.z - m));
                        p[j + 3] = ex2_ftz(fmaf(__uint_as_float(r[j + 3]), scale_log2, b4.w - m));
                    }
                }
            };
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                const bool masked = flag[c * 2 + half] != 0u;
                const float* bias32 = bias + c * 64 + half * 32;
                float p[32];
                tmem_ld_wait();
                tmem_ld_32x16_a(t_row + c * 64 + half * 32 + 16, rb);
                probs16(ra, p, masked, bias32);
                tmem_ld_wait();
                if (c < 3) tmem_ld_32x16_a(t_row + (c + 1) * 64 + half * 32, ra);
                probs16(rb, p + 16, masked, bias32 + 16);
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    sum2 = fadd2(sum2, fadd2(pack_f32x2(p[j], p[j + 1]), pack_f32x2(p[j + 2], p[j + 3])));
                const int slot = c & 1, use = 2 * it + (c >> 1);
                TLS(4 + 4 * c);
                if (use > 0) mbar_wait(&p_free[g * 2 + slot], (use - 1) & 1);     // the chains that read this slot have retired
                TLS(5 + 4 * c);
                store_row_chunk_bf16(ring + slot * (128 * kRowB), row, half, p);
                fence_proxy_async();
                TLS(6 + 4 * c);
                warp_arrive(&p_full[g * 2 + slot], lane);
                TLS(7 + 4 * c);
            }
            float sum, sum_hi;
            unpack_f32x2(sum2, sum, sum_hi);
            sum += sum_hi;
            sXsum[half * 128 + row] = sum;
            store_bias((it + 1) & 1, bias_next);
            named_bar_sync<1>(g * 4 + lg, 64);
            sum += sXsum[(half ^ 1) * 128 + row];
            TLS(20);
            mbar_wait(&o_full[g], it & 1);
            TLS(21);
            tc_fence_after();
            uint32_t pk[16];
            {
                uint32_t r[32];
                tmem_ld_32x32(t_row + half * 32, r);
                tmem_ld_wait();
                warp_arrive(&s_free[g], lane);         // the next item's S chain may overwrite the accumulator
                const float inv = 1.0f / sum;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    pk[j] = pack_bf16(__uint_as_float(r[2 * j]) * inv, __uint_as_float(r[2 * j + 1]) * inv);
            }
            const int qrow0 = g * 128 + lg * 32;
            const int rows_valid = min(32, max(0, L - qrow0));
            store_rows_32(ring + wg * 2048, pk, ctx + (static_cast<long long>(b) * L + qrow0) * (H * kDh) + h * kDh + half * 32,
                          static_cast<long long>(H) * kDh, rows_valid, lane);
            if (half == 0 && g * 128 + row < L) lse[(static_cast<long long>(b) * H + h) * L + g * 128 + row] = (m + log2f(sum)) * kLn2;
            TLS(22);
        }
    }
#undef TLS
#undef TL
    tc_fence_before();
    __syncthreads();
    if (warp == kF2ProducerWarp) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// ------------------------------------------------------------------------------------------------
// forward, persistent, probabilities in TENSOR MEMORY (round 2: the encoder's hot call)
//
// Same persistent structure as attn_tc_fwd2_kernel (one CTA per SM, both 128-row query tiles of a (b, h) item in flight,
// two softmax groups of 8 warps, one MMA-issuing warp per tile), but P never touches shared memory:
//   S_g = Q_g K^T          128 x 256 x 64 into TMEM columns [256 g, 256 g + 256)
//   softmax                thread = (row, key HALF): half 0 owns keys 0..127 (S columns 0..127), half 1 keys 128..255. P is
//                          written back as packed bf16 pairs (tcgen05.st) INTO THE S COLUMNS THE THREAD HAS ALREADY READ:
//                          half 0 -> columns [0, 64), half 1 -> columns [128, 192) -- always behind its own read pointer, so
//                          the two threads of a row never touch each other's unread scores
//   O_g = P_g V            sixteen 128 x 64 x 16 instructions whose A operand is read from TMEM ([a_tmem] form of
//                          tcgen05.mma: 128 lanes x 8 columns per instruction instead of a 4 KB shared-memory fetch),
//                          accumulator = columns [192, 256) of the tile's region (S keys 192..255 have been consumed by then)
// What this removes per tile and item: 64 KB of shared-memory stores of P, their proxy fences, the two-slot ring with its
// eight barrier round trips, and 4 of the 6 KB every P.V instruction fetched from shared memory.
// ------------------------------------------------------------------------------------------------
struct Fwd3Smem {
    static constexpr int kQ = 0;                    // [256 x 64] bf16: query tile g at g * 16 KB
    static constexpr int kK = 32 * 1024;            // [256 x 64]
    static constexpr int kV = 64 * 1024;            // 2 stages x [256 x 64]
    static constexpr int kStage = 128 * 1024;       // 16 warps x 2 KB output staging
    static constexpr int kBias = 160 * 1024;        // [2 groups][2 stages][256] floats
    static constexpr int kXchg = 164 * 1024;        // [2 groups][max, sum][2 halves][128] floats
    static constexpr int kFlag = 168 * 1024;        // [2 groups][2 stages][8] words: this 32-key chunk has a non-zero mask
    static constexpr int kBar = 168 * 1024 + 128;
    static constexpr int kTotal = 168 * 1024 + 256 + 1024;
};

__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        :
        : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (bf16, K-major: lane = row, one 32-bit column = two consecutive k) is
// read from tensor memory
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t"
        "}\n"
        :
        : "r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}

template <int POLY>
__global__ void __launch_bounds__(kF2Threads, 1)
attn_tc_fwd3_kernel(const __grid_constant__ CUtensorMap map_qkv, const float* __restrict__ key_bias,
                    __nv_bfloat16* __restrict__ ctx, float* __restrict__ lse, int n_items, int L, int H, float scale_log2,
                    long long* tl) {
#if CLIMB_ATTN_TIMELINE
#define TL(role, idx) do { if (tl != nullptr && blockIdx.x == 0 && it == 2) tl[(role) * 32 + (idx)] = clock64(); } while (0)
#else
#define TL(role, idx) do { } while (0)
#endif
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + Fwd3Smem::kBar);
    uint64_t* qk_full = bars;            // [1]  tx
    uint64_t* qk_free = bars + 1;        // [1]  n_groups commits
    uint64_t* v_full = bars + 2;         // [2]  tx
    uint64_t* v_free = bars + 4;         // [2]  n_groups commits
    uint64_t* s_full = bars + 6;         // [2 groups] commit
    uint64_t* s_free = bars + 8;         // [2 groups] 8 warp arrivals: O has been read out of TMEM
    uint64_t* p_full = bars + 10;        // [2 groups] 8 warp arrivals: P of the tile is in TMEM
    uint64_t* o_full = bars + 12;        // [2 groups] commit
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_groups = L > 128 ? 2 : 1;
    const int n_my = (n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    pdl_wait();
    pdl_launch_dependents();

    if (warp == kF2ProducerWarp) {
        if (lane == 0) {
            tma_prefetch_desc(&map_qkv);
            mbar_init(qk_full, 1);
            mbar_init(qk_free, n_groups);
            for (int s = 0; s < 2; ++s) {
                mbar_init(&v_full[s], 1);
                mbar_init(&v_free[s], n_groups);
                mbar_init(&s_full[s], 1);
                mbar_init(&s_free[s], kF2GroupThreads / 32);
                mbar_init(&p_full[s], kF2GroupThreads / 32);
                mbar_init(&o_full[s], 1);
            }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == kF2ProducerWarp) {
        if (lane == 0) {
            for (int it = 0; it < n_my; ++it) {
                const int item = blockIdx.x + it * gridDim.x;
                const int b = item / H, h = item - b * H;
                if (it > 0) mbar_wait(qk_free, (it - 1) & 1);
                mbar_arrive_expect_tx(qk_full, 2 * 256 * kRowB);
                tma_load_3d(&map_qkv, qk_full, sm + Fwd3Smem::kQ, h * kDh, 0, b);
                tma_load_3d(&map_qkv, qk_full, sm + Fwd3Smem::kK, (H + h) * kDh, 0, b);
                const int s = it & 1, k = it >> 1;
                if (k > 0) mbar_wait(&v_free[s], (k - 1) & 1);
                mbar_arrive_expect_tx(&v_full[s], 256 * kRowB);
                tma_load_3d(&map_qkv, &v_full[s], sm + Fwd3Smem::kV + s * (256 * kRowB), (2 * H + h) * kDh, 0, b);
            }
        }
        __syncwarp();
    } else if (warp >= kF2MmaWarp0) {
        const int g = warp - kF2MmaWarp0;
        if (lane == 0 && g < n_groups) {
            const DescBase dQ(umma_smem_desc(smem_u32(sm + Fwd3Smem::kQ) + g * (128 * kRowB), 16, 1024));
            const DescBase dK(umma_smem_desc(smem_u32(sm + Fwd3Smem::kK), 16, 1024));
            const DescBase dV(umma_smem_desc(smem_u32(sm + Fwd3Smem::kV), 256 * kRowB, 1024));       // V in place as an MN-major B operand
            const uint32_t t_acc = tmem + g * 256;
            const uint32_t idesc_s = umma_instr_desc(128, 256, 0, 0);
            const uint32_t idesc_o = umma_instr_desc(128, 64, 0, 1);
            for (int it = 0; it < n_my; ++it) {
                TL(g, 0);
                mbar_wait(qk_full, it & 1);
                TL(g, 1);
                if (it > 0) mbar_wait(&s_free[g], (it - 1) & 1);
                TL(g, 2);
                tc_fence_after();
#pragma unroll
                for (uint32_t kk = 0; kk < 4; ++kk)
                    umma_bf16_lohi(t_acc, dQ.lo + kk * 2, dQ.hi, dK.lo + kk * 2, dK.hi, idesc_s, kk > 0 ? 1u : 0u);
                umma_commit(&s_full[g]);
                umma_commit(qk_free);
                TL(g, 3);
                const int s = it & 1;
                const uint32_t v_lo = dV.lo + s * ((256 * kRowB) >> 4);
                mbar_wait(&v_full[s], (it >> 1) & 1);
                mbar_wait(&p_full[g], it & 1);
                TL(g, 4);
                tc_fence_after();
                // keys 0..127 sit in columns [0, 64) of the tile's region, keys 128..255 in [128, 192): 8 columns per instruction
#pragma unroll
                for (uint32_t ks = 0; ks < 16; ++ks)
                    umma_bf16_ts(t_acc + 192, t_acc + (ks < 8 ? ks * 8 : 128 + (ks - 8) * 8), v_lo + ks * ((16 * kRowB) >> 4), dV.hi,
                                 idesc_o, ks > 0 ? 1u : 0u);
                umma_commit(&o_full[g]);
                umma_commit(&v_free[s]);
                TL(g, 5);
            }
        }
        __syncwarp();
    } else if ((warp >> 3) < n_groups) {
        const int g = warp >> 3, wg = warp & 7;
        const int lg = wg & 3, half = wg >> 2;               // TMEM lane group, key half (keys 128 * half ..)
        const int row = lg * 32 + lane;
        const int tg = threadIdx.x & (kF2GroupThreads - 1);
        const uint32_t t_row = tmem + (static_cast<uint32_t>(lg * 32) << 16) + g * 256;
        const uint32_t t_half = t_row + half * 128;          // this thread's scores; its probabilities go to t_half + [0, 64)
        float* sBias = reinterpret_cast<float*>(sm + Fwd3Smem::kBias) + g * 512;                 // [2 stages][256]
        float* sXmax = reinterpret_cast<float*>(sm + Fwd3Smem::kXchg) + g * 512;                 // [2][128]
        float* sXsum = sXmax + 256;
        uint8_t* stage = sm + Fwd3Smem::kStage + warp * 2048;
        uint32_t* sFlag = reinterpret_cast<uint32_t*>(sm + Fwd3Smem::kFlag) + g * 16;          // [2 stages][8 x 32-key chunks]
        auto load_bias = [&](int it) -> float {
            if (tg >= L) return -INFINITY;
            if (key_bias == nullptr) return 0.0f;
            const int item = blockIdx.x + it * gridDim.x;
            return __ldg(key_bias + static_cast<long long>(item / H) * L + tg);
        };
        auto store_bias = [&](int stg, float v) {        // thread tg holds key tg: warp wg covers the 32-key chunk wg
            sBias[stg * 256 + tg] = v * kLog2e;
            const uint32_t any = __ballot_sync(0xffffffffu, v != 0.0f);
            if (lane == 0) sFlag[stg * 8 + wg] = any;
        };
        if (n_my > 0) store_bias(0, load_bias(0));
        const uint64_t scale2 = pack_f32x2(scale_log2, scale_log2);
        for (int it = 0; it < n_my; ++it) {
            const int item = blockIdx.x + it * gridDim.x;
            const int b = item / H, h = item - b * H;
            const float bias_next = it + 1 < n_my ? load_bias(it + 1) : 0.0f;
#define TLS(idx) do { if (wg == 0 && lane == 0) TL(2 + g, idx); } while (0)
            TLS(0);
            named_bar_sync<9>(g, kF2GroupThreads);       // bias row of this item visible; every warp is past its previous epilogue
            TLS(1);
            const float* bias = sBias + (it & 1) * 256 + half * 128;
            const uint32_t* flag = sFlag + (it & 1) * 8 + half * 4;
            mbar_wait(&s_full[g], it & 1);
            TLS(2);
            tc_fence_after();
            // ---- pass 1: row max over this thread's 128 keys ----
            float m = -INFINITY, m_raw = -INFINITY;
            auto row_max = [&](const uint32_t (&r)[32], int c) {
                if (flag[c] == 0u) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        m_raw = fmaxf(m_raw, fmaxf(fmaxf(__uint_as_float(r[j]), __uint_as_float(r[j + 1])),
                                                   fmaxf(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]))));
                } else {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(bias + c * 32 + j);
                        m = fmaxf(m, fmaxf(fmaxf(fmaf(__uint_as_float(r[j]), scale_log2, b4.x), fmaf(__uint_as_float(r[j + 1]), scale_log2, b4.y)),
                                           fmaxf(fmaf(__uint_as_float(r[j + 2]), scale_log2, b4.z), fmaf(__uint_as_float(r[j + 3]), scale_log2, b4.w))));
                    }
                }
            };
            uint32_t ra[16], rb[16];
            {
                uint32_t r0[32], r1[32];
                tmem_ld_32x32(t_half, r0);
                tmem_ld_wait();
                tmem_ld_32x32(t_half + 32, r1);
                row_max(r0, 0);
                tmem_ld_wait();
                tmem_ld_32x32(t_half + 64, r0);
                row_max(r1, 1);
                tmem_ld_wait();
                tmem_ld_32x32(t_half + 96, r1);
                row_max(r0, 2);
                tmem_ld_wait();
                tmem_ld_32x16_a(t_half, ra);                  // pass 2's first 16 columns travel during the exchange below
                row_max(r1, 3);
            }
            m = fmaxf(m, m_raw * scale_log2);                  // scale > 0
            sXmax[half * 128 + row] = m;
            named_bar_sync<1>(g * 4 + lg, 64);
            m = fmaxf(m, sXmax[(half ^ 1) * 128 + row]);       // key 0 is always valid: m is finite
            TLS(3);
            // ---- pass 2: P = exp2(S * scale + mask - m), 32 keys at a time, packed to bf16 and stored over the scores already read ----
            const uint64_t neg_m2 = pack_f32x2(-m, -m);
            uint64_t sum2 = 0ull;
            auto probs16 = [&](const uint32_t (&r)[16], float* p, bool masked, const float* bias16) {
                if (!masked) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        float x0, x1, x2, x3;
                        unpack_f32x2(ffma2(pack_u32x2(r[j], r[j + 1]), scale2, neg_m2), x0, x1);
                        unpack_f32x2(ffma2(pack_u32x2(r[j + 2], r[j + 3]), scale2, neg_m2), x2, x3);
                        p[j] = ex2_ftz(x0);
                        p[j + 1] = ex2_ftz(x1);
                        if constexpr (POLY != 0) {
                            exp2_poly2(x2, x3, p[j + 2], p[j + 3]);
                        } else {
                            p[j + 2] = ex2_ftz(x2);
                            p[j + 3] = ex2_ftz(x3);
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(bias16 + j);
                        p[j] = ex2_ftz(fmaf(__uint_as_float(r[j]), scale_log2, b4.x - m));
                        p[j + 1] = ex2_ftz(fmaf(__uint_as_float(r[j + 1]), scale_log2, b4.y - m));
                        p[j + 2] = ex2_ftz(fmaf(__uint_as_float(r[j + 2]), scale_log2, b4.z - m));
                        p[j + 3] = ex2_ftz(fmaf(__uint_as_float(r[j + 3]), scale_log2, b4.w - m));
                    }
                }
            };
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const bool masked = flag[c] != 0u;
                const float* bias32 = bias + c * 32;
                float p[32];
                tmem_ld_wait();
                tmem_ld_32x16_a(t_half + c * 32 + 16, rb);
                probs16(ra, p, masked, bias32);
                tmem_ld_wait();
                if (c < 3) tmem_ld_32x16_a(t_half + (c + 1) * 32, ra);
                probs16(rb, p + 16, masked, bias32 + 16);
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    sum2 = fadd2(sum2, fadd2(pack_f32x2(p[j], p[j + 1]), pack_f32x2(p[j + 2], p[j + 3])));
                    pk[j >> 1] = pack_bf16(p[j], p[j + 1]);
                    pk[(j >> 1) + 1] = pack_bf16(p[j + 2], p[j + 3]);
                }
                // keys 32 c .. 32 c + 31 of this half -> 16 packed columns at [16 c, 16 c + 16) of the half's region: scores this
                // thread read at chunk c / 2 or earlier
                tmem_st_32x16(t_half + c * 16, pk);
                TLS(4 + c);
            }
            tmem_st_wait();
            warp_arrive(&p_full[g], lane);
            TLS(8);
            float sum, sum_hi;
            unpack_f32x2(sum2, sum, sum_hi);
            sum += sum_hi;
            sXsum[half * 128 + row] = sum;
            store_bias((it + 1) & 1, bias_next);
            named_bar_sync<1>(g * 4 + lg, 64);
            sum += sXsum[(half ^ 1) * 128 + row];
            TLS(20);
            mbar_wait(&o_full[g], it & 1);
            TLS(21);
            tc_fence_after();
            uint32_t pk[16];
            {
                uint32_t r[32];
                tmem_ld_32x32(t_row + 192 + half * 32, r);
                tmem_ld_wait();
                warp_arrive(&s_free[g], lane);         // the next item's S chain may overwrite the tile's region
                const float inv = 1.0f / sum;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    pk[j] = pack_bf16(__uint_as_float(r[2 * j]) * inv, __uint_as_float(r[2 * j + 1]) * inv);
            }
            const int qrow0 = g * 128 + lg * 32;
            const int rows_valid = min(32, max(0, L - qrow0));
            store_rows_32(stage, pk, ctx + (static_cast<long long>(b) * L + qrow0) * (H * kDh) + h * kDh + half * 32,
                          static_cast<long long>(H) * kDh, rows_valid, lane);
            if (half == 0 && g * 128 + row < L) lse[(static_cast<long long>(b) * H + h) * L + g * 128 + row] = (m + log2f(sum)) * kLn2;
            TLS(22);
        }
    }
#undef TLS
#undef TL
    tc_fence_before();
    __syncthreads();
    if (warp == kF2ProducerWarp) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// ------------------------------------------------------------------------------------------------
// forward, persistent, ONLINE softmax over 64-key blocks, four query tiles in flight (round 2, default)
//
// What bounds the two-pass kernels above is not the tensor pipe but the softmax side: TMEM is read at ~128 B per clock and
// SM, the two-pass scheme reads every score twice (512 KB per item), and with two tiles in flight the chain S -> max ->
// exp -> P.V -> drain of a tile is serial, so the TMEM port, the issue slots, the MUFU pipe and the tensor pipe take turns
// instead of overlapping (their per-item times ADD UP to the measured item period). This kernel
//   * reads S once: online softmax, thread = query row (no cross-thread exchange at all), 32 keys at a time; the running
//     maximum is only raised when a chunk exceeds it by more than 2^8 (the probabilities then stay <= 256, exact in the
//     fp32 sum and harmless in bf16), in which case the row's accumulator is rescaled in TMEM -- a path taken at most a few
//     times per row;
//   * keeps FOUR 128-row tiles in flight (slots): two consecutive (b, h) items x their two query tiles; a slot owns 128 TMEM
//     columns: scores of one 64-key block in [0, 64), overwritten in place by the packed bf16 probabilities ([0, 32)), and
//     the output accumulator in [64, 128). 4 softmax warps + 1 MMA-issuing warp per slot, one TMA producer warp;
//   * P.V reads P from tensor memory ([a_tmem] operand), 32 keys (two instructions) per hand-over.
// K / V of an item are loaded once (two stages: the items of slots 0-1 and 2-3), Q per slot.
// ------------------------------------------------------------------------------------------------
struct Fwd4Smem {
    static constexpr int kK = 0;                    // 2 stages x [256 x 64] bf16
    static constexpr int kV = 64 * 1024;            // 2 stages x [256 x 64]
    static constexpr int kQ = 128 * 1024;           // 4 slots x [128 x 64]
    static constexpr int kStage = 192 * 1024;       // 16 warps x 2 KB output staging
    static constexpr int kBias = 224 * 1024;        // [2 pairs][256] bf16, log2 units
    static constexpr int kFlag = 225 * 1024;        // [2 pairs][8] words: the 32-key chunk has a non-zero mask
    static constexpr int kBar = 225 * 1024 + 64;
    static constexpr int kTotal = 225 * 1024 + 64 + 48 * 8 + 1024;
};
static_assert(Fwd4Smem::kTotal <= 227 * 1024, "attention forward shared memory budget");
constexpr int kF4SoftmaxWarps = 16;
constexpr int kF4ProducerWarp = 16;
constexpr int kF4MmaWarp0 = 17;                     // MMA warp of slot s = 17 + s
constexpr int kF4Threads = 21 * 32;
constexpr float kF4Tau = 8.0f;                      // log2 units: the running maximum may lag the true one by this much

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        :
        : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

template <int POLY>
__global__ void __launch_bounds__(kF4Threads, 1)
attn_tc_fwd4_kernel(const __grid_constant__ CUtensorMap map_kv, const __grid_constant__ CUtensorMap map_q,
                    const float* __restrict__ key_bias, __nv_bfloat16* __restrict__ ctx, float* __restrict__ lse, int n_items,
                    int L, int H, float scale_log2) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + Fwd4Smem::kBar);
    uint64_t* q_full = bars;             // [4 slots] tx
    uint64_t* q_free = bars + 4;         // [4] commit: the slot's last S chain has retired
    uint64_t* k_full = bars + 8;         // [2 stages] tx
    uint64_t* k_free = bars + 10;        // [2] n_tiles commits
    uint64_t* v_full = bars + 12;        // [2] tx
    uint64_t* v_free = bars + 14;        // [2] n_tiles commits
    uint64_t* s_full = bars + 16;        // [4] commit: S of a 64-key block is in TMEM
    uint64_t* s_free = bars + 20;        // [4] 4 warp arrivals: O has been read out, the slot may take its next tile
    uint64_t* p_half = bars + 24;        // [4] 4 warp arrivals: P of 32 keys is in TMEM
    uint64_t* pv_done = bars + 28;       // [4] commit: the P.V instructions issued so far have retired
    uint64_t* o_full = bars + 32;        // [4] commit: O of the tile is complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 36);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = L > 128 ? 2 : 1;             // query tiles (slots per item) with real rows
    const int n_kb = (L + 63) >> 6;                  // 64-key blocks with real keys
    const int n_my = (n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    pdl_wait();
    pdl_launch_dependents();

    if (warp == kF4ProducerWarp) {
        if (lane == 0) {
            tma_prefetch_desc(&map_kv);
            tma_prefetch_desc(&map_q);
            for (int i = 0; i < 4; ++i) {
                mbar_init(&q_full[i], 1);
                mbar_init(&q_free[i], 1);
                mbar_init(&s_full[i], 1);
                mbar_init(&s_free[i], 4);
                mbar_init(&p_half[i], 4);
                mbar_init(&pv_done[i], 1);
                mbar_init(&o_full[i], 1);
            }
            for (int i = 0; i < 2; ++i) {
                mbar_init(&k_full[i], 1);
                mbar_init(&k_free[i], n_tiles);
                mbar_init(&v_full[i], 1);
                mbar_init(&v_free[i], n_tiles);
            }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == kF4ProducerWarp) {
        if (lane == 0) {
            for (int it = 0; it < n_my; ++it) {
                const int item = blockIdx.x + it * gridDim.x;
                const int b = item / H, h = item - b * H;
                const int st = it & 1, u = it >> 1;                 // stage = item parity; u-th use of that stage
                if (u > 0) mbar_wait(&k_free[st], (u - 1) & 1);
                mbar_arrive_expect_tx(&k_full[st], 256 * kRowB);
                tma_load_3d(&map_kv, &k_full[st], sm + Fwd4Smem::kK + st * (256 * kRowB), (H + h) * kDh, 0, b);
                for (int g = 0; g < n_tiles; ++g) {
                    const int slot = st * 2 + g;
                    if (u > 0) mbar_wait(&q_free[slot], (u - 1) & 1);
                    mbar_arrive_expect_tx(&q_full[slot], 128 * kRowB);
                    tma_load_3d(&map_q, &q_full[slot], sm + Fwd4Smem::kQ + slot * (128 * kRowB), h * kDh, g * 128, b);
                }
                if (u > 0) mbar_wait(&v_free[st], (u - 1) & 1);
                mbar_arrive_expect_tx(&v_full[st], 256 * kRowB);
                tma_load_3d(&map_kv, &v_full[st], sm + Fwd4Smem::kV + st * (256 * kRowB), (2 * H + h) * kDh, 0, b);
            }
        }
        __syncwarp();
    } else if (warp >= kF4MmaWarp0) {
        const int slot = warp - kF4MmaWarp0;
        const int pair = slot >> 1, g = slot & 1;
        if (lane == 0 && g < n_tiles) {
            const int n_jobs = (n_my - pair + 1) >> 1;              // items pair, pair + 2, ...
            const DescBase dQ(umma_smem_desc(smem_u32(sm + Fwd4Smem::kQ) + slot * (128 * kRowB), 16, 1024));
            const DescBase dK(umma_smem_desc(smem_u32(sm + Fwd4Smem::kK) + pair * (256 * kRowB), 16, 1024));
            const DescBase dV(umma_smem_desc(smem_u32(sm + Fwd4Smem::kV) + pair * (256 * kRowB), 256 * kRowB, 1024));   // MN-major B
            const uint32_t t_slot = tmem + slot * 128;
            const uint32_t idesc_s = umma_instr_desc(128, 64, 0, 0);
            const uint32_t idesc_o = umma_instr_desc(128, 64, 0, 1);
            uint32_t cnt_p = 0, cnt_pv = 0;
            for (int u = 0; u < n_jobs; ++u) {
                mbar_wait(&k_full[pair], u & 1);
                mbar_wait(&q_full[slot], u & 1);
                if (u > 0) mbar_wait(&s_free[slot], (u - 1) & 1);
                tc_fence_after();
                for (int kb = 0; kb < n_kb; ++kb) {
                    if (kb > 0) {                                   // the previous block's P (same columns) has been consumed
                        mbar_wait(&pv_done[slot], (cnt_pv - 1) & 1);
                        tc_fence_after();
                    }
#pragma unroll
                    for (uint32_t kk = 0; kk < 4; ++kk)
                        umma_bf16_lohi(t_slot, dQ.lo + kk * 2, dQ.hi, dK.lo + kb * ((64 * kRowB) >> 4) + kk * 2, dK.hi, idesc_s,
                                       kk > 0 ? 1u : 0u);
                    umma_commit(&s_full[slot]);
                    if (kb == n_kb - 1) {
                        umma_commit(&q_free[slot]);
                        umma_commit(&k_free[pair]);
                    }
                    for (int hh = 0; hh < 2; ++hh) {
                        mbar_wait(&p_half[slot], cnt_p & 1);
                        ++cnt_p;
                        if (kb == 0 && hh == 0) mbar_wait(&v_full[pair], u & 1);
                        tc_fence_after();
#pragma unroll
                        for (uint32_t ks = 0; ks < 2; ++ks)
                            umma_bf16_ts(t_slot + 64, t_slot + hh * 16 + ks * 8,
                                         dV.lo + (kb * 4 + hh * 2 + ks) * ((16 * kRowB) >> 4), dV.hi, idesc_o,
                                         (kb > 0 || hh > 0 || ks > 0) ? 1u : 0u);
                        umma_commit(&pv_done[slot]);
                        ++cnt_pv;
                    }
                }
                umma_commit(&o_full[slot]);
                umma_commit(&v_free[pair]);
            }
        }
        __syncwarp();
    } else {
        const int slot = warp >> 2, lg = warp & 3;
        const int pair = slot >> 1, g = slot & 1;
        if (g < n_tiles) {
            const int n_jobs = (n_my - pair + 1) >> 1;
            const int row = lg * 32 + lane;
            const int tp = g * 128 + row;                           // the key whose mask this thread stages for its pair of slots
            const uint32_t t_slot = tmem + (static_cast<uint32_t>(lg * 32) << 16) + slot * 128;
            __nv_bfloat16* sBias = reinterpret_cast<__nv_bfloat16*>(sm + Fwd4Smem::kBias) + pair * 256;
            uint32_t* sFlag = reinterpret_cast<uint32_t*>(sm + Fwd4Smem::kFlag) + pair * 8;
            uint8_t* stage = sm + Fwd4Smem::kStage + warp * 2048;
            const uint64_t scale2 = pack_f32x2(scale_log2, scale_log2);
            uint32_t cnt_s = 0, cnt_c = 0;                          // S blocks waited for / 32-key chunks handed over (this slot)
            for (int u = 0; u < n_jobs; ++u) {
                const int it = pair + 2 * u;
                const int item = blockIdx.x + it * gridDim.x;
                const int b = item / H, h = item - b * H;
                // additive key mask of this item, staged by the 2 x 128 threads that work on it
                float kbias = -INFINITY;
                if (tp < L) kbias = key_bias != nullptr ? __ldg(key_bias + static_cast<long long>(b) * L + tp) : 0.0f;
                named_bar_sync<2>(pair, n_tiles * 128);             // everyone is done with the previous item's mask
                sBias[tp] = __float2bfloat16_rn(kbias * kLog2e);
                {
                    const uint32_t any = __ballot_sync(0xffffffffu, kbias != 0.0f);
                    if (lane == 0) sFlag[g * 4 + lg] = any;
                }
                named_bar_sync<2>(pair, n_tiles * 128);
                float m = -INFINITY, l = 0.0f;
                for (int c = 0; c < 2 * n_kb; ++c) {
                    const int hh = c & 1;
                    if (hh == 0) {
                        mbar_wait(&s_full[slot], cnt_s & 1);
                        ++cnt_s;
                        tc_fence_after();
                    }
                    uint32_t r[32];
                    tmem_ld_32x32(t_slot + hh * 32, r);
                    const bool masked = sFlag[c] != 0u;
                    tmem_ld_wait();
                    // x = S * scale + mask (log2 units), in place; mx = chunk maximum
                    float mx = -INFINITY;
                    if (!masked) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            mx = fmaxf(mx, fmaxf(fmaxf(__uint_as_float(r[j]), __uint_as_float(r[j + 1])),
                                                 fmaxf(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]))));
                        mx *= scale_log2;                           // scale > 0
                    } else {
                        const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(sBias + c * 32);
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            const float2 bb = __bfloat1622float2(b2[j >> 1]);
                            const float x0 = fmaf(__uint_as_float(r[j]), scale_log2, bb.x);
                            const float x1 = fmaf(__uint_as_float(r[j + 1]), scale_log2, bb.y);
                            r[j] = __float_as_uint(x0);
                            r[j + 1] = __float_as_uint(x1);
                            mx = fmaxf(mx, fmaxf(x0, x1));
                        }
                    }
                    if (c == 0) {
                        m = mx;                                     // key 0 is always valid: finite
                    } else if (__any_sync(0xffffffffu, mx > m + kF4Tau)) {
                        // rare: raise the running maximum and rescale what has been accumulated for this row
                        const float m_new = fmaxf(m, mx);
                        const float f = ex2_ftz(m - m_new);
                        mbar_wait(&pv_done[slot], (cnt_c - 1) & 1);  // every P.V issued for this slot has retired
                        tc_fence_after();
#pragma unroll 1
                        for (int q = 0; q < 2; ++q) {
                            uint32_t o[32];
                            tmem_ld_32x32(t_slot + 64 + q * 32, o);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * f);
                            tmem_st_32x32(t_slot + 64 + q * 32, o);
                        }
                        tmem_st_wait();
                        l *= f;
                        m = m_new;
                    }
                    uint32_t pk[16];
                    uint64_t sum2 = 0ull;
                    if (!masked) {
                        const uint64_t neg_m2 = pack_f32x2(-m, -m);
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float x0, x1, x2, x3, p0, p1, p2, p3;
                            unpack_f32x2(ffma2(pack_u32x2(r[j], r[j + 1]), scale2, neg_m2), x0, x1);
                            unpack_f32x2(ffma2(pack_u32x2(r[j + 2], r[j + 3]), scale2, neg_m2), x2, x3);
                            p0 = ex2_ftz(x0);
                            p1 = ex2_ftz(x1);
                            if constexpr (POLY != 0) {
                                exp2_poly2(x2, x3, p2, p3);
                            } else {
                                p2 = ex2_ftz(x2);
                                p3 = ex2_ftz(x3);
                            }
                            sum2 = fadd2(sum2, fadd2(pack_f32x2(p0, p1), pack_f32x2(p2, p3)));
                            pk[j >> 1] = pack_bf16(p0, p1);
                            pk[(j >> 1) + 1] = pack_bf16(p2, p3);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            const float p0 = ex2_ftz(__uint_as_float(r[j]) - m);
                            const float p1 = ex2_ftz(__uint_as_float(r[j + 1]) - m);
                            sum2 = fadd2(sum2, pack_f32x2(p0, p1));
                            pk[j >> 1] = pack_bf16(p0, p1);
                        }
                    }
                    float s0, s1;
                    unpack_f32x2(sum2, s0, s1);
                    l += s0 + s1;
                    // keys 32 c .. 32 c + 31 of the block -> 16 packed columns at [16 hh, 16 hh + 16): over scores already in registers
                    tmem_st_32x16(t_slot + hh * 16, pk);
                    tmem_st_wait();
                    warp_arrive(&p_half[slot], lane);
                    ++cnt_c;
                }
                // ---- epilogue: O / l -> bf16 -> staged, coalesced stores ----
                mbar_wait(&o_full[slot], u & 1);
                tc_fence_after();
                const float inv = 1.0f / l;
                const int qrow0 = g * 128 + lg * 32;
                const int rows_valid = min(32, max(0, L - qrow0));
                __nv_bfloat16* dst = ctx + (static_cast<long long>(b) * L + qrow0) * (H * kDh) + h * kDh;
#pragma unroll 1
                for (int q = 0; q < 2; ++q) {
                    uint32_t o[32], pk[16];
                    tmem_ld_32x32(t_slot + 64 + q * 32, o);
                    tmem_ld_wait();
                    if (q == 1) warp_arrive(&s_free[slot], lane);      // the slot's columns may be overwritten by its next tile
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        pk[j] = pack_bf16(__uint_as_float(o[2 * j]) * inv, __uint_as_float(o[2 * j + 1]) * inv);
                    store_rows_32(stage, pk, dst + q * 32, static_cast<long long>(H) * kDh, rows_valid, lane);
                }
                if (g * 128 + row < L) lse[(static_cast<long long>(b) * H + h) * L + g * 128 + row] = (m + log2f(l)) * kLn2;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kF4ProducerWarp) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
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
                   float scale, uint32_t drop_thresh, float inv_keep, unsigned long long seed) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space
    float* sLse = reinterpret_cast<float*>(sm + BwdSmem::kLse);
    float* sDelta = reinterpret_cast<float*>(sm + BwdSmem::kDelta);
    float* sBias = reinterpret_cast<float*>(sm + BwdSmem::kBias);
    float* sColV = reinterpret_cast<float*>(sm + BwdSmem::kColV);
    uint64_t* bar_load = reinterpret_cast<uint64_t*>(sm + BwdSmem::kBar);
    uint64_t* bar_a = bar_load + 1;      // S, dP of a block are in TMEM
    uint64_t* bar_p = bar_load + 2;      // P, dS of a block are in smem (one arrival per elementwise warp)
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
            mbar_init(bar_p, kBwdEwThreads / 32);
            mbar_init(bar_b, 1);
            mbar_init(bar_kv, kBwdOutThreads / 32);
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
                if (drop_thresh == 0u) {
#pragma unroll
                    for (int e = 0; e < 16; e += 2) {
                        const float p0 = ex2_ftz(fmaf(__uint_as_float(rs[e]), scale_log2, bias[e]));
                        const float p1 = ex2_ftz(fmaf(__uint_as_float(rs[e + 1]), scale_log2, bias[e + 1]));
                        pp[hh][e >> 1] = pack_bf16(p0, p1);
                        dd[hh][e >> 1] = pack_bf16(p0 * (__uint_as_float(rd[e]) - dl_r), p1 * (__uint_as_float(rd[e + 1]) - dl_r));
                    }
                } else {
                    // dropout on the probabilities (modeling_vilt.py:374): the forward used Pd = P * keep / (1 - p) in O = Pd V, so
                    //   dV = Pd^T dO,   dP = (dO V^T) * keep / (1 - p),   dS = P (dP - delta),  delta = rowsum(dO * O) = rowsum(Pd * dPd).
                    // The mask is regenerated from the forward's counters: (b, h, query, key / 4) -> four keys per Philox call.
                    const unsigned long long base = ((static_cast<unsigned long long>(b) * H + h) * L + (i * 128 + row)) * 64ull +
                                                    static_cast<unsigned long long>((j * 128 + c * 32 + hh * 16) >> 2);
#pragma unroll
                    for (int g4 = 0; g4 < 4; ++g4) {
                        const uint4 rnd = philox4x32(seed, base + g4);
                        const uint32_t rr[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
                        for (int e2 = 0; e2 < 4; e2 += 2) {
                            const int e = g4 * 4 + e2;
                            const float k0 = dropout_scale(rr[e2], drop_thresh, inv_keep), k1 = dropout_scale(rr[e2 + 1], drop_thresh, inv_keep);
                            const float p0 = ex2_ftz(fmaf(__uint_as_float(rs[e]), scale_log2, bias[e]));
                            const float p1 = ex2_ftz(fmaf(__uint_as_float(rs[e + 1]), scale_log2, bias[e + 1]));
                            pp[hh][e >> 1] = pack_bf16(p0 * k0, p1 * k1);
                            dd[hh][e >> 1] = pack_bf16(p0 * (__uint_as_float(rd[e]) * k0 - dl_r), p1 * (__uint_as_float(rd[e + 1]) * k1 - dl_r));
                        }
                    }
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
            warp_arrive(bar_p, lane);
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
                warp_arrive(bar_kv, lane);
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
                if (colsum) staged_colsum_32(stage, lane, colsum + h * kDh + half * 32);
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

// ------------------------------------------------------------------------------------------------
// backward, persistent (128 < L <= 256: two key tiles x two query tiles per (b, h) item)
//
// One CTA per SM walks over (b, h) items. The block loop of an item is the one of the kernel above; what changes is
// that nothing of an item's set-up or drain is exposed any more:
//   producer warp   the eight 16 KB operand tiles (K0 V0 | Q0 dO0 | Q1 dO1 | K1 V1) are refilled with the NEXT item's
//                   data as soon as the last chain that reads them has retired (K0 / V0 after block 1 of 4, ...), so the
//                   128 KB of an item stream in behind the previous item's blocks without a second set of buffers
//   16 elementwise  P / dS of the block; on the side, one sixteenth of the NEXT item's per-row scalars per block
//   warps           (delta = rowsum(dO * O) from coalesced global loads, lse, key mask, column sums of dO)
//   4 drain warps   dK_j / dV_j / dQ_i: TMEM -> bf16 -> staged, coalesced global stores (+ column sums of dQ); the
//                   elementwise warps never leave the block loop
//   MMA warp        issues the next block's S / dP chains (of the next item, at the item boundary) before the current
//                   block's three accumulation chains whenever their operands have landed
// ------------------------------------------------------------------------------------------------
struct Bwd2Smem {
    static constexpr int kQ = 0;                 // [2 tiles][128 x 64] bf16
    static constexpr int kDO = 32 * 1024;
    static constexpr int kK = 64 * 1024;
    static constexpr int kV = 96 * 1024;
    static constexpr int kP = 128 * 1024;        // [128 x 128] bf16 = two 16 KB chunks of 64 columns
    static constexpr int kDS = 160 * 1024;
    static constexpr int kLse = 192 * 1024;      // [2 stages][256] floats each
    static constexpr int kDelta = 194 * 1024;
    static constexpr int kBias = 196 * 1024;
    static constexpr int kColV = 198 * 1024;     // [64] column sums of dO, [64] column sums of dQ, over all of the CTA's items
    static constexpr int kFlag = 198 * 1024 + 512;   // [2 stages][8] words: the 32-key chunk has a non-zero mask
    static constexpr int kBar = 198 * 1024 + 640;
    static constexpr int kStage = 199 * 1024;    // 4 drain warps x 2 KB
    static constexpr int kNext = 207 * 1024;     // [O, dO][512 threads] x 16 B: the next item's rows on their way to delta (cp.async)
    static constexpr int kTotal = 223 * 1024 + 1024;
};
// Block n of an item works on key tile j = n / 2 and query tile i = 0 1 1 0 (even items) or 1 0 0 1 (odd items): the query
// tile an item ends with is the one the next item needs last, so every operand tile has at least one block of the
// previous item left to arrive in
__device__ __forceinline__ int bwd2_qtile(int it, int n) { return (((n + 1) >> 1) ^ it) & 1; }
constexpr int kB2EwThreads = 512;
constexpr int kB2MmaWarpA = 16, kB2MmaWarpB = 17, kB2DrainWarp0 = 18;     // drain warp 0 is also the TMA producer (and owns the TMEM allocation)
constexpr int kB2ProducerWarp = kB2DrainWarp0;
constexpr int kB2DrainThreads = 128;
constexpr int kB2Threads = 22 * 32;


__global__ void __launch_bounds__(kB2Threads, 1)
attn_tc_bwd2_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do,
                    const float* __restrict__ key_bias, const __nv_bfloat16* __restrict__ ctx,
                    const __nv_bfloat16* __restrict__ dctx, const float* __restrict__ lse,
                    __nv_bfloat16* __restrict__ dqkv, float* __restrict__ colsum, int n_batch, int L, int H,
                    float scale_log2, float scale, int colsum_q, long long* tl) {
#if CLIMB_ATTN_TIMELINE
#define TLB(role, idx) do { if (tl != nullptr && blockIdx.x == 0 && it == 2) tl[(role) * 64 + (idx)] = clock64(); } while (0)
#else
#define TLB(role, idx) do { } while (0)
#endif
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* sLse = reinterpret_cast<float*>(sm + Bwd2Smem::kLse);
    float* sDelta = reinterpret_cast<float*>(sm + Bwd2Smem::kDelta);
    float* sBias = reinterpret_cast<float*>(sm + Bwd2Smem::kBias);
    float* sColV = reinterpret_cast<float*>(sm + Bwd2Smem::kColV);
    uint32_t* sFlag = reinterpret_cast<uint32_t*>(sm + Bwd2Smem::kFlag);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + Bwd2Smem::kBar);
    uint64_t* full = bars;               // [4] tx: K0 V0 | Q0 dO0 | Q1 dO1 | K1 V1 of an item have landed (refilled when kv_full / dq_full say their last reader retired)
    uint64_t* bar_a = bars + 8;          // S, dP of a block are in TMEM
    uint64_t* bar_p = bars + 9;          // P, dS of a block are in smem (16 warp arrivals)
    uint64_t* bar_b = bars + 10;         // the accumulation chains of a block have retired
    uint64_t* kv_full = bars + 11;       // dK_j, dV_j complete
    uint64_t* kv_free = bars + 12;       // ... and read out of TMEM (4 warp arrivals)
    uint64_t* dq_full = bars + 13;       // [2] dQ_i complete: the query tile of blocks 1 and 2 after block 2, the other after block 3
    uint64_t* dq_free = bars + 15;       // [2] ... and read out of TMEM (4 warp arrivals)
    uint64_t* bar_s = bars + 17;         // S, dP of a block have been read out of TMEM (16 warp arrivals): the next block's chains may start
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // a CTA stays with ONE head (h = CTA index mod H) and walks over the batch: the q / v bias gradients of the head add up in
    // shared memory and leave as 128 global atomics per CTA instead of 128 per item. The host guarantees gridDim.x >= H.
    const int h = static_cast<int>(blockIdx.x) % H;
    const int b_first = static_cast<int>(blockIdx.x) / H;
    const int b_step = (static_cast<int>(gridDim.x) - h + H - 1) / H;          // CTAs that share this head
    const int n_my = b_first < n_batch ? (n_batch - b_first + b_step - 1) / b_step : 0;
    const long long ld = 3LL * H * kDh, ldo = static_cast<long long>(H) * kDh;
    constexpr uint32_t tile = 128 * kRowB;                          // 16 KB: 128 rows of a [rows x 64] tile
    pdl_wait();
    pdl_launch_dependents();

    // operand tile group t of item `it` -> smem (producer warp, one lane)
    auto issue_tiles = [&](int t, int it) {
        const int b = b_first + it * b_step;
        mbar_arrive_expect_tx(&full[t], 2 * tile);
        if (t == 0 || t == 3) {
            const int r0 = t == 0 ? 0 : 128;
            tma_load_3d(&map_qkv, &full[t], sm + Bwd2Smem::kK + (t == 0 ? 0 : tile), (H + h) * kDh, r0, b);
            tma_load_3d(&map_qkv, &full[t], sm + Bwd2Smem::kV + (t == 0 ? 0 : tile), (2 * H + h) * kDh, r0, b);
        } else {
            const int r0 = t == 1 ? 0 : 128;
            tma_load_3d(&map_qkv, &full[t], sm + Bwd2Smem::kQ + (t == 1 ? 0 : tile), h * kDh, r0, b);
            tma_load_3d(&map_do, &full[t], sm + Bwd2Smem::kDO + (t == 1 ? 0 : tile), h * kDh, r0, b);
        }
    };

    float cv_keep[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) cv_keep[e] = 0.0f;
    if (warp == kB2ProducerWarp) {
        if (lane == 0) {
            tma_prefetch_desc(&map_qkv);
            tma_prefetch_desc(&map_do);
            for (int t = 0; t < 4; ++t) {
                mbar_init(&full[t], 1);
            }
            mbar_init(bar_a, 1);
            mbar_init(bar_p, kB2EwThreads / 32);
            mbar_init(bar_s, kB2EwThreads / 32);
            mbar_init(bar_b, 1);
            mbar_init(kv_full, 1);
            mbar_init(kv_free, kB2DrainThreads / 32);
            for (int t = 0; t < 2; ++t) {
                mbar_init(&dq_full[t], 1);
                mbar_init(&dq_free[t], kB2DrainThreads / 32);
            }
            fence_barrier_init();
            for (int t = 0; t < 4; ++t) issue_tiles(t, 0);           // the first item's operands overlap the prologue
        }
        reinterpret_cast<float4*>(sColV)[lane] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    } else if (warp < 16) {
        // per-row scalars of the FIRST item (stage 0); later items get theirs from the block loop of the item before
        const int b = b_first;
        const long long stat = (static_cast<long long>(b) * H + h) * L;
        if (threadIdx.x < 256) {
            const int r = threadIdx.x;
            sLse[r] = r < L ? lse[stat + r] * kLog2e : INFINITY;
            const float kb = r < L ? (key_bias ? key_bias[static_cast<long long>(b) * L + r] : 0.0f) : -INFINITY;
            sBias[r] = kb * kLog2e;
            const uint32_t any = __ballot_sync(0xffffffffu, kb != 0.0f);
            if (lane == 0) sFlag[warp] = any;
        }
        const int sub = lane & 7, rsel = lane >> 3;
        float cv[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) cv[e] = 0.0f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int r = warp * 16 + q * 4 + rsel;
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
#pragma unroll
        for (int e = 0; e < 8; ++e) cv_keep[e] = cv[e];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == kB2MmaWarpA) {
        // S = Q_i K_j^T, dP = dO_i V_j^T of the NEXT block: issued the moment the elementwise warps have read the current
        // block's S / dP out of TMEM. The two issuing warps split the block's 32 instructions (an issue costs a thread
        // ~40 cycles of descriptor arithmetic around the tcgen05.mma; the 128 x 64 x 16 / 128 x 128 x 16 shapes here keep the
        // tensor pipe busy for 32 / 64, so a single issuer would be the kernel's speed limit).
        if (lane == 0) {
            const uint32_t id_a = umma_instr_desc(128, 128, 0, 0);     // both operands K-major
            const DescBase dQ_k(umma_smem_desc(smem_u32(sm + Bwd2Smem::kQ), 16, 1024)), dK_k(umma_smem_desc(smem_u32(sm + Bwd2Smem::kK), 16, 1024));
            const DescBase dDO_k(umma_smem_desc(smem_u32(sm + Bwd2Smem::kDO), 16, 1024)), dV_k(umma_smem_desc(smem_u32(sm + Bwd2Smem::kV), 16, 1024));
            auto phase_a = [&](int it, int n) {
                const uint32_t j = n >> 1, i = bwd2_qtile(it, n);
                mbar_wait(&full[j ? 3 : 0], it & 1);
                mbar_wait(&full[i ? 2 : 1], it & 1);
                tc_fence_after();
                const uint32_t oi = i * (tile >> 4), oj = j * (tile >> 4);
                // the two chains alternate: consecutive instructions never accumulate into the same TMEM columns
#pragma unroll
                for (uint32_t kk = 0; kk < 4; ++kk) {
                    umma_bf16_lohi(tmem + kTS, dQ_k.lo + oi + kk * 2, dQ_k.hi, dK_k.lo + oj + kk * 2, dK_k.hi, id_a, kk > 0 ? 1u : 0u);
                    umma_bf16_lohi(tmem + kTdP, dDO_k.lo + oi + kk * 2, dDO_k.hi, dV_k.lo + oj + kk * 2, dV_k.hi, id_a, kk > 0 ? 1u : 0u);
                }
                umma_commit(bar_a);
            };
            phase_a(0, 0);
            int gb = 0;
            for (int it = 0; it < n_my; ++it) {
#pragma unroll 1
                for (int n = 0; n < 4; ++n, ++gb) {
                    TLB(0, n * 8 + 0);
                    mbar_wait(bar_s, gb & 1);      // S / dP of the block have been read out of TMEM (its P / dS are still being made)
                    TLB(0, n * 8 + 1);
                    if (n == 3 && it + 1 >= n_my) break;
                    phase_a(n < 3 ? it : it + 1, n < 3 ? n + 1 : 0);
                    TLB(0, n * 8 + 2);
                }
            }
        }
        __syncwarp();
    } else if (warp == kB2MmaWarpB) {
        // dV_j += P^T dO_i, dK_j += dS^T Q_i, dQ_i += dS K_j of the current block
        if (lane == 0) {
            const uint32_t id_t = umma_instr_desc(128, 64, 1, 1);      // dV / dK: A = P^T / dS^T in place, B MN-major
            const uint32_t id_q = umma_instr_desc(128, 64, 0, 1);      // dQ: A = dS K-major, B = K MN-major
            const DescBase dP_m(umma_smem_desc(smem_u32(sm + Bwd2Smem::kP), tile, 1024)), dDS_m(umma_smem_desc(smem_u32(sm + Bwd2Smem::kDS), tile, 1024));
            const DescBase dDS_k(umma_smem_desc(smem_u32(sm + Bwd2Smem::kDS), 16, 1024));
            const DescBase dDO_m(umma_smem_desc(smem_u32(sm + Bwd2Smem::kDO), tile, 1024)), dQ_m(umma_smem_desc(smem_u32(sm + Bwd2Smem::kQ), tile, 1024));
            const DescBase dK_m(umma_smem_desc(smem_u32(sm + Bwd2Smem::kK), tile, 1024));
            constexpr uint32_t kstep = (16 * kRowB) >> 4;                 // 16 rows of a [rows x 64] tile, in descriptor units
            int gb = 0;
            for (int it = 0; it < n_my; ++it) {
#pragma unroll 1
                for (int n = 0; n < 4; ++n, ++gb) {
                    const uint32_t j = n >> 1, i = bwd2_qtile(it, n), pos = n & 1;
                    TLB(3, n * 8 + 0);
                    mbar_wait(bar_p, gb & 1);      // P, dS of the block are in smem
                    TLB(3, n * 8 + 1);
                    if (pos == 0) {                         // the previous key tile's dV / dK must have been read out
                        const int kvc = it * 2 + j;
                        if (kvc > 0) mbar_wait(kv_free, (kvc - 1) & 1);
                    }
                    if (n < 2 && it > 0) mbar_wait(&dq_free[i], (it - 1) & 1);    // first chain of the item into dQ_i
                    tc_fence_after();
                    TLB(3, n * 8 + 2);
                    const uint32_t oi = i * (tile >> 4), oj = j * (tile >> 4);
                    // k runs over the 128 query rows (dV, dK) or the 128 keys (dQ), 16 per UMMA
                    // (the three chains alternate: consecutive instructions never accumulate into the same TMEM columns)
#pragma unroll
                    for (uint32_t k = 0; k < 8; ++k) {
                        umma_bf16_lohi(tmem + kTdV, dP_m.lo + k * kstep, dP_m.hi, dDO_m.lo + oi + k * kstep, dDO_m.hi, id_t, (pos > 0 || k > 0) ? 1u : 0u);
                        umma_bf16_lohi(tmem + kTdK, dDS_m.lo + k * kstep, dDS_m.hi, dQ_m.lo + oi + k * kstep, dQ_m.hi, id_t, (pos > 0 || k > 0) ? 1u : 0u);
                        umma_bf16_lohi(tmem + kTdQ + i * 64, dDS_k.lo + (k >> 2) * (tile >> 4) + (k & 3) * 2, dDS_k.hi, dK_m.lo + oj + k * kstep, dK_m.hi,
                                       id_q, (j > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(bar_b);
                    if (pos == 1) umma_commit(kv_full);          // n == 1 also frees K0 / V0, n == 3 K1 / V1
                    if (n >= 2) umma_commit(&dq_full[i]);        // the last block of query tile i (blocks 1 + 2 or 0 + 3): dQ_i is complete,
                                                                 // and Q_i / dO_i may be refilled
                    TLB(3, n * 8 + 3);
                }
            }
        }
        __syncwarp();
    } else if (warp >= kB2DrainWarp0) {
        const int lg = warp & 3;                                 // the TMEM lane group this warp may read
        const uint32_t t_row = tmem + (static_cast<uint32_t>(lg * 32) << 16);
        uint8_t* stage = sm + Bwd2Smem::kStage + (warp - kB2DrainWarp0) * 2048;
        const bool producer = warp == kB2ProducerWarp && lane == 0;
#define TLD(idx) do { if (warp == kB2DrainWarp0 + 1 && lane == 0) TLB(2, j * 16 + (idx)); } while (0)
        for (int it = 0; it < n_my; ++it) {
            const int b = b_first + it * b_step;
            for (int j = 0; j < 2; ++j) {
                TLD(0);
                mbar_wait(kv_full, (it * 2 + j) & 1);
                TLD(1);
                tc_fence_after();
                // producer duty, in the order the tiles come free: K0 V0 after block 1 (= dK_0 / dV_0 complete), the query tile of
                // blocks 1 and 2 after block 2 (below), the rest after block 3 (= dK_1 / dV_1 complete)
                const int q_mid = bwd2_qtile(it, 1) ? 2 : 1, q_end = bwd2_qtile(it, 3) ? 2 : 1;     // tile groups of the two query tiles
                if (producer && it + 1 < n_my) {
                    if (j == 0) {
                        issue_tiles(0, it + 1);
                    } else {
                        issue_tiles(q_end, it + 1);       // block 3 has retired (kv_full of key tile 1): its query tile and K1 / V1 are free
                        issue_tiles(3, it + 1);
                    }
                }
                __syncwarp();
                const int key0 = j * 128 + lg * 32;
                const int rows_valid = min(32, max(0, L - key0));
                __nv_bfloat16* dst = dqkv + (static_cast<long long>(b) * L + key0) * ld + h * kDh;
                uint32_t pk[2][16];
#pragma unroll
                for (int which = 0; which < 2; ++which) {          // 0: dK (scaled), 1: dV
                    const uint32_t col = which == 0 ? kTdK : kTdV;
                    const float sc = which == 0 ? scale : 1.0f;
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        uint32_t r[32];
                        tmem_ld_32x32(t_row + col + half * 32, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 16; ++e)
                            pk[half][e] = pack_bf16(__uint_as_float(r[2 * e]) * sc, __uint_as_float(r[2 * e + 1]) * sc);
                    }
                    if (which == 1) warp_arrive(kv_free, lane);
                    TLD(2 + 2 * which);
#pragma unroll
                    for (int half = 0; half < 2; ++half)
                        store_rows_32(stage, pk[half], dst + (which == 0 ? H * kDh : 2 * H * kDh) + half * 32, ld, rows_valid, lane);
                    TLD(3 + 2 * which);
                }
                // then one dQ tile per pass: the query tile of blocks 1 and 2 is complete after block 2 (drained here while block 3
                // runs), the other one after block 3. The next item starts with the tile that was drained first.
                const int i = bwd2_qtile(it, j == 0 ? 1 : 3);
                mbar_wait(&dq_full[i], it & 1);
                TLD(6);
                tc_fence_after();
                if (producer && j == 0 && it + 1 < n_my) issue_tiles(q_mid, it + 1);      // block 2 has retired
                __syncwarp();
                const int q0 = i * 128 + lg * 32;
                const int q_valid = min(32, max(0, L - q0));
                uint32_t qk[2][16];
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t r[32];
                    tmem_ld_32x32(t_row + kTdQ + i * 64 + half * 32, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        qk[half][e] = pack_bf16(__uint_as_float(r[2 * e]) * scale, __uint_as_float(r[2 * e + 1]) * scale);
                }
                warp_arrive(&dq_free[i], lane);
                TLD(7);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    __nv_bfloat16* qdst = dqkv + (static_cast<long long>(b) * L + q0) * ld + h * kDh + half * 32;
                    // (staged, coalesced stores: writing the rows straight from registers -- 32 rows per instruction -- was measured
                    //  25 % slower for the whole kernel)
                    store_rows_32(stage, qk[half], qdst, ld, q_valid, lane);
                    if (colsum_q) staged_colsum_32(stage, lane, sColV + 64 + half * 32);
                    TLD(8 + half);
                }
            }
        }
    // (end of the drain role)
#undef TLD
    } else if (warp < 16) {
        const int lg = warp & 3, quarter = warp >> 2;       // TMEM lane group, 32-key chunk of the block
        const int row = lg * 32 + lane;
        const int sub = lane & 7, rsel = lane >> 3;
        const uint32_t t_row = tmem + (static_cast<uint32_t>(lg * 32) << 16);
        const uint64_t scale2 = pack_f32x2(scale_log2, scale_log2);
        const uint32_t s_next = smem_u32(sm + Bwd2Smem::kNext) + threadIdx.x * 16;
        // rows (set q = 0..3) of item `target`'s O and dO on their way to delta: through shared memory (cp.async), issued one
        // block before they are consumed, so that nothing ever waits on them
        auto prefetch_rows = [&](int target, int q) {
            const int b2 = b_first + target * b_step, h2 = h;
            const int dr = warp * 16 + q * 4 + rsel;
            const bool dvalid = dr < L;
            const long long off = dvalid ? (static_cast<long long>(b2) * L + dr) * ldo + h2 * kDh + sub * 8 : 0;
            cp_async_16(s_next, ctx + off, dvalid);
            cp_async_16(s_next + 512 * 16, dctx + off, dvalid);
            cp_async_commit();
        };
        if (n_my > 1) prefetch_rows(1, 0);
        int gb = 0;
        float n_lse = 0.0f, n_kb = 0.0f;
        for (int it = 0; it < n_my; ++it) {
            const int st = it & 1;
            const bool has_next = it + 1 < n_my;
            const int b2 = b_first + (it + 1) * b_step, h2 = h;
            const float* sLseC = sLse + st * 256;
            const float* sDeltaC = sDelta + st * 256;
            const float* sBiasC = sBias + st * 256;
            // the elementwise warps meet once per item: the scalar stage about to be rewritten (st ^ 1) was the previous item's, and
            // its last reader must be done; this also makes the stage hand-over independent of the mbarrier chain through the MMA warps
            if (it > 0) asm volatile("bar.sync 1, %0;" ::"n"(kB2EwThreads) : "memory");
#pragma unroll 1
            for (int n = 0; n < 4; ++n, ++gb) {
                const int j = n >> 1, i = bwd2_qtile(it, n);
                // ---- the NEXT item's per-row scalars (stage st ^ 1: nobody reads it before the next item's first block), one
                //      sixteenth per warp and block, in the time this warp would otherwise spend waiting for S / dP ----
                if (has_next) {
                    const int dr = warp * 16 + n * 4 + rsel;
                    cp_async_wait<0>();
                    uint4 n_o, n_do;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(n_o.x), "=r"(n_o.y), "=r"(n_o.z), "=r"(n_o.w) : "r"(s_next));
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(n_do.x), "=r"(n_do.y), "=r"(n_do.z), "=r"(n_do.w) : "r"(s_next + 512 * 16));
                    const uint32_t av[4] = {n_o.x, n_o.y, n_o.z, n_o.w}, dv[4] = {n_do.x, n_do.y, n_do.z, n_do.w};
                    float dl = 0.0f;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 x = unpack_bf16(av[e]), y = unpack_bf16(dv[e]);
                        dl = fmaf(x.x, y.x, dl);
                        dl = fmaf(x.y, y.y, dl);
                        cv_keep[2 * e] += y.x;          // column sums of dO (the v-bias gradient): per lane over ALL of the CTA's
                        cv_keep[2 * e + 1] += y.y;      // items, reduced once at the end (zero-filled rows add nothing)
                    }
                    dl += __shfl_xor_sync(0xffffffffu, dl, 1);
                    dl += __shfl_xor_sync(0xffffffffu, dl, 2);
                    dl += __shfl_xor_sync(0xffffffffu, dl, 4);
                    if (sub == 0) sDelta[(st ^ 1) * 256 + dr] = dl;
                    if (threadIdx.x < 256) {
                        if (n == 0) {
                            const int r = threadIdx.x;
                            n_lse = r < L ? __ldg(lse + (static_cast<long long>(b2) * H + h2) * L + r) : INFINITY;
                            n_kb = r < L ? (key_bias ? __ldg(key_bias + static_cast<long long>(b2) * L + r) : 0.0f) : -INFINITY;
                        } else if (n == 1) {
                            sLse[(st ^ 1) * 256 + threadIdx.x] = n_lse * kLog2e;
                            sBias[(st ^ 1) * 256 + threadIdx.x] = n_kb * kLog2e;
                            const uint32_t any = __ballot_sync(0xffffffffu, n_kb != 0.0f);
                            if (lane == 0) sFlag[(st ^ 1) * 8 + warp] = any;
                        }
                    }
                }
                {   // rows for the NEXT block's share of the scalars: a whole block of lead time
                    const int nit = n < 3 ? it : it + 1, nn = n < 3 ? n + 1 : 0;
                    if (nit + 1 < n_my) prefetch_rows(nit + 1, nn);
                }
#define TLE(idx) do { if (threadIdx.x == 0) TLB(1, n * 8 + (idx)); } while (0)
                TLE(0);
                mbar_wait(bar_a, gb & 1);
                TLE(1);
                tc_fence_after();
                const float lse_r = sLseC[i * 128 + row], dl_r = sDeltaC[i * 128 + row];
                const int c = quarter;
                const bool masked = sFlag[st * 8 + j * 4 + c] != 0u;
                uint32_t pp[2][8], dd[2][8];
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {                    // two 16-key halves: 32 accumulator registers live
                    uint32_t rs[16], rd[16];
                    tmem_ld_32x16_a(t_row + kTS + c * 32 + hh * 16, rs);
                    tmem_ld_32x16_a(t_row + kTdP + c * 32 + hh * 16, rd);
                    if (!masked) {
                        const uint64_t neg_lse2 = pack_f32x2(-lse_r, -lse_r), neg_dl2 = pack_f32x2(-dl_r, -dl_r);
                        tmem_ld_wait();
                        if (hh == 1) warp_arrive(bar_s, lane);      // this warp's share of S / dP is in registers
#pragma unroll
                        for (int e = 0; e < 16; e += 2) {
                            float x0, x1, d0, d1;
                            unpack_f32x2(ffma2(pack_u32x2(rs[e], rs[e + 1]), scale2, neg_lse2), x0, x1);
                            const float p0 = ex2_ftz(x0), p1 = ex2_ftz(x1);
                            unpack_f32x2(fmul2(pack_f32x2(p0, p1), fadd2(pack_u32x2(rd[e], rd[e + 1]), neg_dl2)), d0, d1);
                            pp[hh][e >> 1] = pack_bf16(p0, p1);
                            dd[hh][e >> 1] = pack_bf16(d0, d1);
                        }
                    } else {
                        float bias[16];
#pragma unroll
                        for (int e = 0; e < 16; e += 4) {
                            const float4 b4 = *reinterpret_cast<const float4*>(sBiasC + j * 128 + c * 32 + hh * 16 + e);
                            bias[e] = b4.x - lse_r; bias[e + 1] = b4.y - lse_r; bias[e + 2] = b4.z - lse_r; bias[e + 3] = b4.w - lse_r;
                        }
                        tmem_ld_wait();
                        if (hh == 1) warp_arrive(bar_s, lane);
#pragma unroll
                        for (int e = 0; e < 16; e += 2) {
                            const float p0 = ex2_ftz(fmaf(__uint_as_float(rs[e]), scale_log2, bias[e]));
                            const float p1 = ex2_ftz(fmaf(__uint_as_float(rs[e + 1]), scale_log2, bias[e + 1]));
                            pp[hh][e >> 1] = pack_bf16(p0, p1);
                            dd[hh][e >> 1] = pack_bf16(p0 * (__uint_as_float(rd[e]) - dl_r), p1 * (__uint_as_float(rd[e + 1]) - dl_r));
                        }
                    }
                }
                // everything above overlapped the previous block's accumulation chains; they must have retired before
                // P / dS are overwritten
                TLE(2);
                if (gb > 0) mbar_wait(bar_b, (gb - 1) & 1);
                TLE(3);
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    // 16 columns = granules (c & 1) * 4 + hh * 2 + {0, 1} of the row in 64-column chunk (c >> 1)
                    uint8_t* pc = sm + Bwd2Smem::kP + (c >> 1) * (128 * kRowB);
                    uint8_t* dc = sm + Bwd2Smem::kDS + (c >> 1) * (128 * kRowB);
                    const int g0 = (c & 1) * 4 + hh * 2;
                    *reinterpret_cast<uint4*>(pc + swz(row, g0)) = make_uint4(pp[hh][0], pp[hh][1], pp[hh][2], pp[hh][3]);
                    *reinterpret_cast<uint4*>(pc + swz(row, g0 + 1)) = make_uint4(pp[hh][4], pp[hh][5], pp[hh][6], pp[hh][7]);
                    *reinterpret_cast<uint4*>(dc + swz(row, g0)) = make_uint4(dd[hh][0], dd[hh][1], dd[hh][2], dd[hh][3]);
                    *reinterpret_cast<uint4*>(dc + swz(row, g0 + 1)) = make_uint4(dd[hh][4], dd[hh][5], dd[hh][6], dd[hh][7]);
                }
                fence_proxy_async();
                TLE(4);
                TLE(5);
                warp_arrive(bar_p, lane);
                TLE(6);
            }
        }
        if (gb > 0) mbar_wait(bar_b, (gb - 1) & 1);      // (every barrier phase is observed before the CTA retires)
        if (colsum != nullptr) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                cv_keep[e] += __shfl_xor_sync(0xffffffffu, cv_keep[e], 8);
                cv_keep[e] += __shfl_xor_sync(0xffffffffu, cv_keep[e], 16);
            }
            if (lane < 8) {
#pragma unroll
                for (int e = 0; e < 8; ++e) atomicAdd(sColV + lane * 8 + e, cv_keep[e]);
            }
        }
    }
#undef TLE
    tc_fence_before();
    __syncthreads();
    if (colsum != nullptr && n_my > 0 && threadIdx.x < (colsum_q ? 128 : 64)) {
        // [0, 64): column sums of dO = the head's v-bias gradient; [64, 128): column sums of dQ = its q-bias gradient
        const int t = threadIdx.x & 63;
        atomicAdd(colsum + (threadIdx.x < 64 ? 2 * H * kDh : 0) + h * kDh + t, sColV[threadIdx.x]);
    }
    if (warp == kB2ProducerWarp) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

#undef TLB
int make_map3(CUtensorMap* map, const void* ptr, int B, int L, long long row_elems, int box_rows) {
    const long long dims[3] = {row_elems, L, B};
    const long long strides[2] = {row_elems, static_cast<long long>(L) * row_elems};
    const int box[3] = {64, box_rows, 1};
    return encode_tmap_bf16(map, ptr, 3, dims, strides, box);
}

// CLIMB_ATTN_V1=1 keeps the one-tile-per-CTA kernels (A/B measurements only)
bool attn_v1() {
    static int v1 = -1;
    if (v1 < 0) {
        const char* e = std::getenv("CLIMB_ATTN_V1");
        v1 = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    return v1 == 1;
}

int env_flag(const char* name, int dflt) {
    const char* e = std::getenv(name);
    return e != nullptr && e[0] != 0 ? std::atoi(e) : dflt;
}

int sm_count() { return climb::num_sms(); }      // gemm_tcgen05.cu: device SM count minus the data-parallel reserve

}  // namespace

int attention_tc_fwd(const void* qkv, const float* key_bias, void* ctx, float* lse, int B, int L, int H, float scale,
                     cudaStream_t stream, float p_drop, unsigned long long seed) {
    CLIMB_REQUIRE(L <= 256, "attention_tc_fwd: L=%d > 256", L);
    CLIMB_REQUIRE(p_drop >= 0.0f && p_drop < 1.0f, "attention_tc_fwd: dropout p=%f outside [0, 1)", p_drop);
    if (p_drop == 0.0f && !attn_v1()) {
        CUtensorMap mqkv;
        int rc2 = make_map3(&mqkv, qkv, B, L, 3LL * H * kDh, 256);
        if (rc2) return rc2;
        static bool attr2 = false;
        if (!attr2) {
            CLIMB_CUDA_OK(cudaFuncSetAttribute(attn_tc_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Fwd2Smem::kTotal));
            attr2 = true;
        }
        const int n_items = B * H;
        long long* tl = nullptr;
#if CLIMB_ATTN_TIMELINE
        static long long* tl_buf = nullptr;
        static int tl_calls = 0;
        if (env_flag("CLIMB_ATTN_TL", 0) && tl_buf == nullptr) {
            cudaMalloc(&tl_buf, 4 * 32 * sizeof(long long));
            cudaMemset(tl_buf, 0, 4 * 32 * sizeof(long long));
        }
        tl = tl_buf;
#endif
        // CLIMB_ATTN_FWD = 3 (default): two-pass softmax, P in tensor memory, P.V with a TMEM A operand; 4: online softmax over
        // 64-key blocks with four tiles in flight; 2: the round-1 kernel (P through a shared-memory ring). All three measure
        // 40-42 us per launch at B = 64 ("what bounds the attention kernels"): the variants are kept for A/B runs.
        static const int fwd_variant = env_flag("CLIMB_ATTN_FWD", 3);
        const bool fwd3 = fwd_variant == 3;
        if (fwd_variant != 2 && fwd_variant != 3) {
            CUtensorMap mq128;
            rc2 = make_map3(&mq128, qkv, B, L, 3LL * H * kDh, 128);
            if (rc2) return rc2;
            // CLIMB_ATTN_EXP=0: every exponential on the MUFU pipe (default: every other pair as a degree-4 polynomial on the FMA pipe)
            static const bool poly = env_flag("CLIMB_ATTN_EXP", 1) != 0;
            static bool attr4 = false;
            if (!attr4) {
                CLIMB_CUDA_OK(cudaFuncSetAttribute(attn_tc_fwd4_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fwd4Smem::kTotal));
                CLIMB_CUDA_OK(cudaFuncSetAttribute(attn_tc_fwd4_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fwd4Smem::kTotal));
                attr4 = true;
            }
            CLIMB_CUDA_OK(launch_pdl(poly ? attn_tc_fwd4_kernel<1> : attn_tc_fwd4_kernel<0>, dim3(std::min(n_items, sm_count())),
                                     dim3(kF4Threads), Fwd4Smem::kTotal, stream, mqkv, mq128, key_bias, static_cast<__nv_bfloat16*>(ctx), lse,
                                     n_items, L, H, scale * kLog2e));
        } else if (fwd3) {
            // every exponential on the MUFU pipe by default (CLIMB_ATTN_EXP=1: every other pair as a polynomial on the FMA pipe):
            // with P in tensor memory the kernel is bound by issue slots, not by MUFU (21 % busy), and the polynomial costs
            // 6.5 issue slots per element against 1 -- measured 41.6 vs 42.7 us per launch at B = 64
            static const bool poly = env_flag("CLIMB_ATTN_EXP", 0) != 0;
            static bool attr3 = false;
            if (!attr3) {
                CLIMB_CUDA_OK(cudaFuncSetAttribute(attn_tc_fwd3_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fwd3Smem::kTotal));
                CLIMB_CUDA_OK(cudaFuncSetAttribute(attn_tc_fwd3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fwd3Smem::kTotal));
                attr3 = true;
            }
            CLIMB_CUDA_OK(launch_pdl(poly ? attn_tc_fwd3_kernel<1> : attn_tc_fwd3_kernel<0>, dim3(std::min(n_items, sm_count())),
                                     dim3(kF2Threads), Fwd3Smem::kTotal, stream, mqkv, key_bias, static_cast<__nv_bfloat16*>(ctx), lse, n_items,
                                     L, H, scale * kLog2e, tl));
        } else {
            CLIMB_CUDA_OK(launch_pdl(attn_tc_fwd2_kernel, dim3(std::min(n_items, sm_count())), dim3(kF2Threads), Fwd2Smem::kTotal, stream,
                                     mqkv, key_bias, static_cast<__nv_bfloat16*>(ctx), lse, n_items, L, H, scale * kLog2e, env_flag("CLIMB_ATTN_STAGGER", 1), tl));
        }
        CLIMB_LAUNCH_OK();
#if CLIMB_ATTN_TIMELINE
        if (tl != nullptr && ++tl_calls == 20) {
            long long h[128];
            cudaDeviceSynchronize();
            cudaMemcpy(h, tl, sizeof(h), cudaMemcpyDeviceToHost);
            long long t0 = h[0];
            for (int r = 0; r < 4; ++r) {
                printf("TL role %d:", r);
                for (int i = 0; i < 24; ++i) printf(" %lld", h[r * 32 + i] ? h[r * 32 + i] - t0 : -1);
                printf("\n");
            }
            fflush(stdout);
        }
#endif
        return 0;
    }
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
                     void* dqkv, float* colsum, int B, int L, int H, float scale, cudaStream_t stream, float p_drop,
                     unsigned long long seed) {
    CLIMB_REQUIRE(L <= 256, "attention_tc_bwd: L=%d > 256", L);
    CLIMB_REQUIRE(p_drop >= 0.0f && p_drop < 1.0f, "attention_tc_bwd: dropout p=%f outside [0, 1)", p_drop);
    if (p_drop == 0.0f && L > 128 && H <= sm_count() && !attn_v1()) {
        CUtensorMap mqkv2, mdo2;
        int rc2 = make_map3(&mqkv2, qkv, B, L, 3LL * H * kDh, 128);
        if (rc2) return rc2;
        rc2 = make_map3(&mdo2, dctx, B, L, static_cast<long long>(H) * kDh, 128);
        if (rc2) return rc2;
        static bool attr2 = false;
        if (!attr2) {
            CLIMB_CUDA_OK(cudaFuncSetAttribute(attn_tc_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Bwd2Smem::kTotal));
            attr2 = true;
        }
        // The drain warps are the backward's serial resource; transposing the dQ blocks for their column sums inside them
        // costs 13 us per launch. CLIMB_ATTN_COLSUM_Q=1 keeps the sums in the kernel.
        const int colsum_q = (colsum != nullptr && env_flag("CLIMB_ATTN_COLSUM_Q", 0)) ? 1 : 0;
        long long* tlb = nullptr;
#if CLIMB_ATTN_TIMELINE
        static long long* tlb_buf = nullptr;
        static int tl_calls = 0;
        if (env_flag("CLIMB_ATTN_TL", 0) && tlb_buf == nullptr) {
            cudaMalloc(&tlb_buf, 256 * sizeof(long long));
            cudaMemset(tlb_buf, 0, 256 * sizeof(long long));
        }
        tlb = tlb_buf;
#endif
        CLIMB_CUDA_OK(launch_pdl(attn_tc_bwd2_kernel, dim3(std::min(B * H, sm_count())), dim3(kB2Threads), Bwd2Smem::kTotal, stream,
                                 mqkv2, mdo2, key_bias, static_cast<const __nv_bfloat16*>(ctx), static_cast<const __nv_bfloat16*>(dctx),
                                 lse, static_cast<__nv_bfloat16*>(dqkv), colsum, B, L, H, scale * kLog2e, scale, colsum_q, tlb));
        CLIMB_LAUNCH_OK();
        // the q-bias gradient: a streaming pass over the dq columns just written (23 MB at the bench shape, still in L2)
        if (colsum != nullptr && !colsum_q) {
            const int rc3 = climb::colsum(dqkv, CLIMB_BF16, 3LL * H * kDh, B * L, H * kDh, colsum, stream);
            if (rc3) return rc3;
        }
#if CLIMB_ATTN_TIMELINE
        if (tlb != nullptr && ++tl_calls == 20) {
            long long h[256];
            cudaDeviceSynchronize();
            cudaMemcpy(h, tlb, sizeof(h), cudaMemcpyDeviceToHost);
            long long t0 = h[0];
            for (int r = 0; r < 4; ++r) {
                printf("TLB role %d:", r);
                for (int i = 0; i < 32; ++i) printf("%s%lld", (i % 8) ? " " : " | ", h[r * 64 + i] ? h[r * 64 + i] - t0 : -1);
                printf("\n");
            }
            fflush(stdout);
        }
#endif
        return 0;
    }
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
                             static_cast<__nv_bfloat16*>(dqkv), colsum, L, H, scale * kLog2e, scale,
                             p_drop > 0.0f ? dropout_threshold(p_drop) : 0u, 1.0f / (1.0f - p_drop), seed));
    CLIMB_LAUNCH_OK();
    return 0;
}

}  // namespace climb
