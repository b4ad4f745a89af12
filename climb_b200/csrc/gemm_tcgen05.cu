// bf16 x bf16 -> fp32 GEMM on the sm_100a tensor cores (tcgen05.mma, accumulators in TMEM),
// operands staged by TMA (cp.async.bulk.tensor, 128B swizzle) through a multi-stage mbarrier
// ring, persistent over output tiles with a double-buffered TMEM accumulator so that the
// epilogue of tile i overlaps the main loop of tile i+1.
//
//   C[M,N] = epilogue( alpha * sum_k A[m,k] * B[n,k] )
//
// Every Linear of the ViLT hot path maps onto this one kernel:
//   forward  Y = X W^T          A = X  (K-major)   B = W  (K-major)      modeling_vilt.py:356-360,409,464,482
//   dgrad    dX = dY W          A = dY (K-major)   B = W  (MN-major: the contraction runs over W's rows)
//   wgrad    dW = dY^T X        A = dY (MN-major)  B = X  (MN-major), split-K over the token dimension
// "MN-major" operands are read in place through the UMMA descriptor's major bit: no transposes
// are materialised in HBM.
//
// Warp roles (320 threads): warp 0 = TMA producer (one lane), warp 1 = TMEM owner + MMA issuer
// (one lane), warps 2..9 = epilogue (TMEM -> registers -> fused epilogue -> global).
#include "common.cuh"
#include "climb_b200.h"

#include <cstdlib>
#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched through the runtime)

namespace climb {

int colsum(const void* src, int dtype, long long ld, int rows, int cols, float* out, cudaStream_t stream);      // elementwise.cu

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;          // 64 bf16 = 128 B = one swizzle row
constexpr int kNumEpiWarps = 8;
constexpr int kNumThreads = 64 + kNumEpiWarps * 32;
constexpr int kAccStages = 2;
constexpr int kEpiHalfBytes = 4096;      // one staging block of an epilogue warp: 32 rows x 128 B

struct GemmDeviceArgs {
    int M, N, K;
    int a_mn_major, b_mn_major;
    void* C;
    long long ldc;
    int c_dtype;                  // climb_dtype
    const float* bias;            // [N] or null (added along n)
    const float* residual;        // fp32 [M, ldr] or null
    long long ldr;
    int epilogue;                 // climb_epilogue
    void* aux;                    // bf16 [M, ldaux]
    long long ldaux;
    void* c2;                     // optional bf16 copy of the final value [M, ldc2]
    long long ldc2;
    float* colsum;                // optional [N] += column sums of the (bf16) C tile
    float* colsum_a;              // pair wgrad kernel only, optional [M] += sum over k of A[k, m] (the bias gradient next to dW)
    float alpha;
    int accumulate;
    int split_k;
    int m_tiles, n_tiles, k_blocks_per_split, k_blocks_total;
    int skip_pdl_wait;            // climb_gemm_desc.independent: do not wait for the previous kernel of the stream
    int num_stages;               // smem ring depth (runtime: deeper when the epilogue needs no input prefetch)
    int scratch_bytes;            // per epilogue warp: 4096 (staging only) or 8192 (+ prefetch half)
    int full_tiles, total_vtiles; // fast kernels: tiles [0, full_tiles) are 256 columns wide, the rest come as two 128-column halves
};

template <int BLOCK_N>
struct SmemLayout {
    static constexpr int kABytes = kBlockM * kBlockK * 2;
    static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kMaxStages = 8;
    static constexpr int kBarrierBytes = 1024;
    // ring depth: as deep as the 227 KB allow next to the epilogue scratch
    static constexpr int stages_for(int scratch_per_warp) {
        int s = (227 * 1024 - 1024 - kBarrierBytes - kNumEpiWarps * scratch_per_warp) / kStageBytes;
        return s > kMaxStages ? kMaxStages : s;
    }
    static constexpr int total_for(int scratch_per_warp) {
        return stages_for(scratch_per_warp) * kStageBytes + kBarrierBytes + kNumEpiWarps * scratch_per_warp + 1024;
    }
    static_assert(stages_for(8192) >= 3, "shared memory budget");
};

// UMMA shared-memory descriptor, 128B swizzle (layout type 2), sm_100 version bit.
//   K-major : rows of 128 B, 8-row atoms 1024 B apart (SBO); LBO unused.
//   MN-major: 64-element (128 B) runs along MN, one row per k; 8 k-rows per 1024 B atom (SBO);
//             the next 64-wide MN chunk starts kBlockK*128 B later (LBO).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;     // descriptor version (Blackwell)
    d |= 2ull << 61;     // SWIZZLE_128B
    return d;
}

// un-swizzled (interleaved) layout: 8-row x 16-byte core matrices, LBO = next core matrix along K, SBO = next 8 rows
__device__ __forceinline__ uint64_t make_smem_desc_noswizzle(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;     // descriptor version (Blackwell); layout type 0 = no swizzle
    return d;
}

__device__ __forceinline__ void tmem_ld_32x1(uint32_t taddr, uint32_t& r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
}

__device__ __forceinline__ uint32_t make_instr_desc(int umma_m, int umma_n, int a_mn, int b_mn) {
    uint32_t d = 0;
    d |= 1u << 4;                       // D format = F32
    d |= 1u << 7;                       // A format = BF16
    d |= 1u << 10;                      // B format = BF16
    d |= static_cast<uint32_t>(a_mn & 1) << 15;
    d |= static_cast<uint32_t>(b_mn & 1) << 16;
    d |= static_cast<uint32_t>(umma_n >> 3) << 17;
    d |= static_cast<uint32_t>(umma_m >> 4) << 24;
    return d;
}

// One switch per 32-column chunk (not per element): keeps the unrolled epilogue compact.
__device__ __forceinline__ void act_chunk(int epi, float (&v)[32]) {
    switch (epi) {
        case CLIMB_EPI_GELU:
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_f(v[j]);
            break;
        case CLIMB_EPI_SWISH:
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = swish_f(v[j]);
            break;
        case CLIMB_EPI_RELU:
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
            break;
        case CLIMB_EPI_TANH:
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = tanhf(v[j]);
            break;
        default: break;
    }
}
// v *= act'(aux), aux given as 16 packed bf16 pairs of the same chunk
__device__ __forceinline__ void dact_chunk(int epi, float (&v)[32], const uint32_t* aux_pk) {
    switch (epi) {
        case CLIMB_EPI_DGELU:
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float2 a = unpack_bf16(aux_pk[j]);
                v[2 * j] *= dgelu_f(a.x);
                v[2 * j + 1] *= dgelu_f(a.y);
            }
            break;
        case CLIMB_EPI_DSWISH:
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float2 a = unpack_bf16(aux_pk[j]);
                v[2 * j] *= dswish_f(a.x);
                v[2 * j + 1] *= dswish_f(a.y);
            }
            break;
        case CLIMB_EPI_DRELU:
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float2 a = unpack_bf16(aux_pk[j]);
                v[2 * j] = a.x > 0.0f ? v[2 * j] : 0.0f;
                v[2 * j + 1] = a.y > 0.0f ? v[2 * j + 1] : 0.0f;
            }
            break;
        case CLIMB_EPI_MUL_AUX:
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float2 a = unpack_bf16(aux_pk[j]);
                v[2 * j] *= a.x;
                v[2 * j + 1] *= a.y;
            }
            break;
        default: break;
    }
}

// v <- gelu(v), g <- packed bf16 gelu'(v): Phi and the Gaussian density are shared between the two
__device__ __forceinline__ void gelu_and_grad_chunk(float (&v)[32], uint32_t (&g)[16]) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        float c0, e0, c1, e1;
        const float x0 = v[2 * j], x1 = v[2 * j + 1];
        phi_pdf(x0, c0, e0);
        phi_pdf(x1, c1, e1);
        v[2 * j] = x0 * c0;
        v[2 * j + 1] = x1 * c1;
        g[j] = pack_bf16(fmaf(x0 * 0.39894228040143267794f, e0, c0), fmaf(x1 * 0.39894228040143267794f, e1, c1));
    }
}

// 128B swizzle of a linear byte offset inside a warp's scratch block (16-byte granules)
__device__ __forceinline__ uint32_t swz128(uint32_t linear) {
    const uint32_t line = linear >> 7, slot = (linear >> 4) & 7u;
    return (line << 7) | ((slot ^ (line & 7u)) << 4);
}

// registers (thread = row, NCH 16-byte granules per row) -> global rows, coalesced.
// gbase points at (first row of the warp, first column of the chunk); ld_bytes = row pitch.
template <int NCH>
__device__ __forceinline__ void store_rows(uint8_t* scratch, const uint32_t* pk, uint8_t* gbase,
                                           long long ld_bytes, int rows_valid, int lane, bool atomic_f32) {
    __syncwarp();
#pragma unroll
    for (int j = 0; j < NCH; ++j)
        *reinterpret_cast<uint4*>(scratch + swz128((lane * NCH + j) * 16)) =
            make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
    __syncwarp();
#pragma unroll
    for (int it = 0; it < NCH; ++it) {
        const int q = it * 32 + lane;
        const int rr = q / NCH, j = q % NCH;
        if (rr < rows_valid) {
            const uint4 val = *reinterpret_cast<const uint4*>(scratch + swz128(q * 16));
            uint8_t* dst = gbase + rr * ld_bytes + j * 16;
            if (atomic_f32)
                atomicAdd(reinterpret_cast<float4*>(dst),
                          make_float4(__uint_as_float(val.x), __uint_as_float(val.y), __uint_as_float(val.z),
                                      __uint_as_float(val.w)));
            else
                *reinterpret_cast<uint4*>(dst) = val;
        }
    }
}

// global rows -> registers (thread = row), coalesced reads
template <int NCH>
__device__ __forceinline__ void load_rows(uint8_t* scratch, uint32_t* pk, const uint8_t* gbase,
                                          long long ld_bytes, int rows_valid, int lane) {
    __syncwarp();
#pragma unroll
    for (int it = 0; it < NCH; ++it) {
        const int q = it * 32 + lane;
        const int rr = q / NCH, j = q % NCH;
        uint4 val = make_uint4(0u, 0u, 0u, 0u);
        if (rr < rows_valid) val = *reinterpret_cast<const uint4*>(gbase + rr * ld_bytes + j * 16);
        *reinterpret_cast<uint4*>(scratch + swz128(q * 16)) = val;
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        const uint4 val = *reinterpret_cast<const uint4*>(scratch + swz128((lane * NCH + j) * 16));
        pk[4 * j] = val.x; pk[4 * j + 1] = val.y; pk[4 * j + 2] = val.z; pk[4 * j + 3] = val.w;
    }
}

// Asynchronous (cp.async, no registers) version of load_rows' first half: the rows of the NEXT chunk
// are put in flight while the current chunk is being processed; finish with prefetch_take().
template <int NCH>
__device__ __forceinline__ void prefetch_rows(uint8_t* buf, const uint8_t* gbase, long long ld_bytes,
                                              int rows_valid, int lane) {
    const uint32_t sb = smem_u32(buf);
#pragma unroll
    for (int it = 0; it < NCH; ++it) {
        const int q = it * 32 + lane;
        const int rr = q / NCH, j = q % NCH;
        const bool ok = rr < rows_valid;
        cp_async_16(sb + swz128(q * 16), gbase + (ok ? rr * ld_bytes : 0) + j * 16, ok);
    }
    cp_async_commit();
}
template <int NCH>
__device__ __forceinline__ void prefetch_take(const uint8_t* buf, uint32_t* pk, int lane) {
    cp_async_wait<0>();
    __syncwarp();
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        const uint4 val = *reinterpret_cast<const uint4*>(buf + swz128((lane * NCH + j) * 16));
        pk[4 * j] = val.x; pk[4 * j + 1] = val.y; pk[4 * j + 2] = val.z; pk[4 * j + 3] = val.w;
    }
}


// ================================ TMA producer (one thread) ================================
template <int BLOCK_N>
__device__ __forceinline__ void producer_loop(const CUtensorMap& tmap_a, const CUtensorMap& tmap_b, const GemmDeviceArgs& p,
                                              uint8_t* smem, uint64_t* full_bar, uint64_t* empty_bar, int kStages) {
    using L = SmemLayout<BLOCK_N>;
    const int tiles_mn = p.m_tiles * p.n_tiles;
    const int total_tiles = tiles_mn * p.split_k;
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int split = tile / tiles_mn;
        const int mn = tile - split * tiles_mn;
        const int m_blk = mn / p.n_tiles;
        const int n_blk = mn - m_blk * p.n_tiles;
        const int kb0 = split * p.k_blocks_per_split;
        const int kb1 = min(kb0 + p.k_blocks_per_split, p.k_blocks_total);
        for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            uint8_t* sa = smem + stage * L::kStageBytes;
            uint8_t* sb = sa + L::kABytes;
            mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
            if (!p.a_mn_major) {
                tma_load_2d(&tmap_a, &full_bar[stage], sa, kb * kBlockK, m_blk * kBlockM);
            } else {
#pragma unroll
                for (int j = 0; j < kBlockM / 64; ++j)
                    tma_load_2d(&tmap_a, &full_bar[stage], sa + j * (kBlockK * 128), m_blk * kBlockM + j * 64, kb * kBlockK);
            }
            if (!p.b_mn_major) {
                tma_load_2d(&tmap_b, &full_bar[stage], sb, kb * kBlockK, n_blk * BLOCK_N);
            } else {
#pragma unroll
                for (int j = 0; j < BLOCK_N / 64; ++j)
                    tma_load_2d(&tmap_b, &full_bar[stage], sb + j * (kBlockK * 128), n_blk * BLOCK_N + j * 64, kb * kBlockK);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
    }
}

// ================================ MMA issuer (one thread) ==================================
template <int BLOCK_N>
__device__ __forceinline__ void mma_loop(const GemmDeviceArgs& p, uint8_t* smem, uint64_t* full_bar, uint64_t* empty_bar,
                                         uint64_t* acc_full, uint64_t* acc_empty, uint32_t tmem_base, int kStages) {
    using L = SmemLayout<BLOCK_N>;
    const int tiles_mn = p.m_tiles * p.n_tiles;
    const int total_tiles = tiles_mn * p.split_k;
    const uint32_t idesc = make_instr_desc(kBlockM, BLOCK_N, p.a_mn_major, p.b_mn_major);
    // per-operand descriptor constants
    const uint32_t a_lbo = p.a_mn_major ? kBlockK * 128 : 16;
    const uint32_t b_lbo = p.b_mn_major ? kBlockK * 128 : 16;
    const uint32_t a_kstep = p.a_mn_major ? 16 * 128 : 32;   // bytes per UMMA_K = 16
    const uint32_t b_kstep = p.b_mn_major ? 16 * 128 : 32;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int split = tile / tiles_mn;
        const int kb0 = split * p.k_blocks_per_split;
        const int kb1 = min(kb0 + p.k_blocks_per_split, p.k_blocks_total);
        mbar_wait(&acc_empty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BLOCK_N);
        for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
            const uint32_t sb = sa + L::kABytes;
#pragma unroll
            for (int kk = 0; kk < kBlockK / 16; ++kk) {
                const uint64_t da = make_smem_desc(sa + kk * a_kstep, a_lbo, 1024);
                const uint64_t db = make_smem_desc(sb + kk * b_kstep, b_lbo, 1024);
                umma_bf16(d_tmem, da, db, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);       // frees the smem slot when MMAs retire
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&acc_full[acc]);              // accumulator complete -> epilogue
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1u; }
    }
}

// HAS_INPUT: the epilogue reads a tensor (derivative aux / residual); only those instantiations carry
// the input registers and the prefetch machinery.
template <int BLOCK_N, bool HAS_INPUT>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a,
                         const __grid_constant__ CUtensorMap tmap_b, const GemmDeviceArgs p) {
    using L = SmemLayout<BLOCK_N>;
    const int kStages = p.num_stages;
    constexpr uint32_t kTmemCols = kAccStages * BLOCK_N;   // 512 / 256 / 128: powers of two >= 32

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    uint8_t* bar_base = smem + kStages * L::kStageBytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
    uint64_t* empty_bar = full_bar + L::kMaxStages;
    uint64_t* acc_full = empty_bar + L::kMaxStages;
    uint64_t* acc_empty = acc_full + kAccStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kAccStages);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < kAccStages; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], kNumEpiWarps * 32);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (!p.skip_pdl_wait) pdl_wait();     // everything above overlapped the previous kernel's tail; operands / epilogue tensors come after

    const int tiles_mn = p.m_tiles * p.n_tiles;
    const int total_tiles = tiles_mn * p.split_k;

    if (warp == 0) {
        if (lane == 0) producer_loop<BLOCK_N>(tmap_a, tmap_b, p, smem, full_bar, empty_bar, kStages);
    } else if (warp == 1) {
        if (lane == 0) mma_loop<BLOCK_N>(p, smem, full_bar, empty_bar, acc_full, acc_empty, tmem_base, kStages);
    } else {
        // ================================ epilogue =========================================
        // A thread owns one accumulator row (TMEM lane) and walks it in 32-column chunks. Global
        // traffic never uses that row-per-thread mapping: every tensor the epilogue reads or writes
        // goes through a 4 KB per-warp shared-memory block (128B-swizzled) so that the warp's global
        // accesses are 16 bytes per lane over CONSECUTIVE addresses of a row (full 32 B sectors, 4-8
        // L1 wavefronts per instruction instead of 32).
        const int ew = warp - 2;                  // 0..7
        const int lane_grp = warp & 3;            // TMEM lanes [32*lane_grp, +32) are ours
        const int col_half = ew >> 2;             // two warps share a lane group: even/odd chunks
        uint8_t* scratch_base = smem + kStages * L::kStageBytes + L::kBarrierBytes + ew * p.scratch_bytes;
        int buf = 0;                              // which half of the warp's scratch the current chunk uses
        int pf_tile = -1, pf_c = -1;              // chunk whose epilogue input is in flight in the other half
        int acc = 0;
        uint32_t acc_phase = 0;
        const bool c_f32 = p.c_dtype == CLIMB_F32;
        const bool fast_c = ((p.ldc * (c_f32 ? 4 : 2)) % 16 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
        const bool fast_aux = p.aux == nullptr || (((p.ldaux * 2) % 16 == 0) && ((reinterpret_cast<uintptr_t>(p.aux) & 15) == 0));
        const bool fast_c2 = p.c2 == nullptr || (((p.ldc2 * 2) % 16 == 0) && ((reinterpret_cast<uintptr_t>(p.c2) & 15) == 0));
        const bool fast_res = p.residual == nullptr || (((p.ldr * 4) % 16 == 0) && ((reinterpret_cast<uintptr_t>(p.residual) & 15) == 0));
        const bool fast_bias = p.bias == nullptr || ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
        const bool all_fast = fast_c && fast_aux && fast_c2 && fast_res && fast_bias;
        const bool aux_in = (p.epilogue == CLIMB_EPI_DGELU || p.epilogue == CLIMB_EPI_DSWISH ||
                             p.epilogue == CLIMB_EPI_DRELU || p.epilogue == CLIMB_EPI_MUL_AUX);
        const bool save_grad = p.epilogue == CLIMB_EPI_GELU_SAVE_GRAD;
        const bool aux_out = (p.aux != nullptr) && !aux_in && !save_grad;   // pre-activation copy (bf16)
        // The one epilogue INPUT stream worth prefetching: the bf16 aux tensor of the derivative
        // epilogues (dGELU reads the saved pre-activation), else the fp32 residual.
        const int pf_kind = (!HAS_INPUT || !all_fast || p.scratch_bytes < 2 * kEpiHalfBytes) ? 0
                            : (aux_in ? 1 : (p.residual != nullptr ? 2 : 0));
        // issue the loads of chunk (t, cc) into scratch half `b`; returns false if that chunk does not
        // exist or is not on the fast path (then it will be loaded synchronously, or not at all)
        auto prefetch = [&](int t, int cc, int b) -> bool {
            if (pf_kind == 0 || t >= total_tiles) return false;
            const int sp = t / tiles_mn;
            if (sp != 0 && pf_kind == 2) return false;
            const int mn2 = t - sp * tiles_mn;
            const int mb = mn2 / p.n_tiles, nb = mn2 - mb * p.n_tiles;
            const int nn0 = nb * BLOCK_N + cc * 32;
            if (cc >= BLOCK_N / 32 || nn0 + 32 > p.N) return false;
            const int r0 = mb * kBlockM + lane_grp * 32;
            const int rv = min(32, max(0, p.M - r0));
            uint8_t* dstb = scratch_base + b * kEpiHalfBytes;
            if (pf_kind == 1)
                prefetch_rows<4>(dstb, reinterpret_cast<const uint8_t*>(p.aux) + (static_cast<long long>(r0) * p.ldaux + nn0) * 2,
                                 p.ldaux * 2, rv, lane);
            else
                prefetch_rows<8>(dstb, reinterpret_cast<const uint8_t*>(p.residual) + (static_cast<long long>(r0) * p.ldr + nn0) * 4,
                                 p.ldr * 4, rv, lane);
            return true;
        };
        if (prefetch(blockIdx.x, col_half, buf ^ 1)) { pf_tile = blockIdx.x; pf_c = col_half; }
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int split = tile / tiles_mn;
            const int mn = tile - split * tiles_mn;
            const int m_blk = mn / p.n_tiles;
            const int n_blk = mn - m_blk * p.n_tiles;
            const int row0 = m_blk * kBlockM + lane_grp * 32;      // first row of this warp
            const int row = row0 + lane;
            const bool row_ok = row < p.M;
            const int rows_valid = min(32, max(0, p.M - row0));
            const bool add_bias = p.bias != nullptr && split == 0;
            const bool add_res = p.residual != nullptr && split == 0;
            mbar_wait(&acc_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) +
                                   static_cast<uint32_t>(acc * BLOCK_N);
#pragma unroll 1
            for (int c = col_half; c < BLOCK_N / 32; c += 2) {
                const int n0 = n_blk * BLOCK_N + c * 32;
                if (n0 >= p.N) break;                      // warp-uniform
                const bool full = (n0 + 32 <= p.N);
                // epilogue input of THIS chunk: already in flight (prefetched) or fetched now
                uint32_t pin[HAS_INPUT ? 32 : 1];
                const bool have_pf = HAS_INPUT && (pf_tile == tile && pf_c == c);
                if (have_pf) buf ^= 1;            // prefetches always target the half the previous chunk did not use
                uint8_t* scratch = scratch_base + buf * kEpiHalfBytes;
                if (HAS_INPUT && full && all_fast) {
                    if (have_pf) {
                        if (pf_kind == 1) prefetch_take<4>(scratch, pin, lane);
                        else prefetch_take<8>(scratch, pin, lane);
                    } else if (aux_in) {
                        cp_async_wait<0>();
                        load_rows<4>(scratch, pin, reinterpret_cast<const uint8_t*>(p.aux) +
                                                       (static_cast<long long>(row0) * p.ldaux + n0) * 2,
                                     p.ldaux * 2, rows_valid, lane);
                    } else if (add_res) {
                        cp_async_wait<0>();
                        load_rows<8>(scratch, pin, reinterpret_cast<const uint8_t*>(p.residual) +
                                                       (static_cast<long long>(row0) * p.ldr + n0) * 4,
                                     p.ldr * 4, rows_valid, lane);
                    }
                    // next chunk of this warp: same tile two chunks on, else the first chunk of its next tile
                    int nt = tile, nc = c + 2;
                    if (nc >= BLOCK_N / 32 || n_blk * BLOCK_N + nc * 32 >= p.N) { nt = tile + gridDim.x; nc = col_half; }
                    if (prefetch(nt, nc, buf ^ 1)) { pf_tile = nt; pf_c = nc; } else { pf_tile = -1; }
                }
                float v[32];
                {
                    uint32_t r[32];
                    tmem_ld_32x32(t_row + static_cast<uint32_t>(c * 32), r);
                    tmem_ld_wait();
                    if (p.alpha == 1.0f) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
                    }
                }
                if (full && all_fast) {
                    // ------------------------------ fast path ------------------------------------
                    if (add_bias) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
                            v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
                        }
                    }
                    uint32_t pk[32];
                    if (HAS_INPUT && aux_in) {
                        dact_chunk(p.epilogue, v, pin);
                    } else if (save_grad) {
                        uint32_t gpk[16];
                        gelu_and_grad_chunk(v, gpk);
                        store_rows<4>(scratch, gpk, reinterpret_cast<uint8_t*>(p.aux) +
                                                        (static_cast<long long>(row0) * p.ldaux + n0) * 2,
                                      p.ldaux * 2, rows_valid, lane, false);
                    } else {
                        if (aux_out) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) pk[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
                            store_rows<4>(scratch, pk, reinterpret_cast<uint8_t*>(p.aux) +
                                                           (static_cast<long long>(row0) * p.ldaux + n0) * 2,
                                          p.ldaux * 2, rows_valid, lane, false);
                        }
                        act_chunk(p.epilogue, v);
                    }
                    if (HAS_INPUT && add_res) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(pin[j]);
                    }
                    if (p.c2 != nullptr) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) pk[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
                        store_rows<4>(scratch, pk, reinterpret_cast<uint8_t*>(p.c2) +
                                                       (static_cast<long long>(row0) * p.ldc2 + n0) * 2,
                                      p.ldc2 * 2, rows_valid, lane, false);
                    }
                    if (c_f32) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) pk[j] = __float_as_uint(v[j]);
                        store_rows<8>(scratch, pk, reinterpret_cast<uint8_t*>(p.C) +
                                                       (static_cast<long long>(row0) * p.ldc + n0) * 4,
                                      p.ldc * 4, rows_valid, lane, p.accumulate != 0);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) pk[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
                        store_rows<4>(scratch, pk, reinterpret_cast<uint8_t*>(p.C) +
                                                       (static_cast<long long>(row0) * p.ldc + n0) * 2,
                                      p.ldc * 2, rows_valid, lane, false);
                        if (p.colsum != nullptr) {
                            // the chunk is still staged in scratch: lane l sums column l over the warp's rows
                            __syncwarp();
                            float cs = 0.0f;
                            for (int rr = 0; rr < rows_valid; ++rr) {
                                const uint16_t h16 = *reinterpret_cast<const uint16_t*>(
                                    scratch + swz128((rr * 4 + (lane >> 3)) * 16) + (lane & 7) * 2);
                                cs += __uint_as_float(static_cast<uint32_t>(h16) << 16);
                            }
                            atomicAdd(p.colsum + n0 + lane, cs);
                        }
                    }
                    continue;
                }
                // ------------------- generic path: ragged N, unaligned leading dimensions ------------
                // (task-head logits with N = 3129 or 1, adapter widths that are not multiples of 32):
                // per-thread rows, scalar accesses, deliberately compact code.
                {
                    const int ncols = min(32, p.N - n0);
                    if (add_bias) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < ncols) v[j] += __ldg(p.bias + n0 + j);
                    }
                    if (aux_in) {
                        uint32_t apk[16];
                        const __nv_bfloat16* auxp = reinterpret_cast<const __nv_bfloat16*>(p.aux) +
                                                    static_cast<long long>(row) * p.ldaux + n0;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float lo = (row_ok && 2 * j < ncols) ? __bfloat162float(auxp[2 * j]) : 0.0f;
                            const float hi = (row_ok && 2 * j + 1 < ncols) ? __bfloat162float(auxp[2 * j + 1]) : 0.0f;
                            apk[j] = pack_bf16(lo, hi);
                        }
                        dact_chunk(p.epilogue, v, apk);
                    } else if (save_grad) {
                        uint32_t gpk[16];
                        gelu_and_grad_chunk(v, gpk);
                        if (row_ok) {
                            __nv_bfloat16* auxp = reinterpret_cast<__nv_bfloat16*>(p.aux) +
                                                  static_cast<long long>(row) * p.ldaux + n0;
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const float2 gg = unpack_bf16(gpk[j]);
                                if (2 * j < ncols) auxp[2 * j] = __float2bfloat16_rn(gg.x);
                                if (2 * j + 1 < ncols) auxp[2 * j + 1] = __float2bfloat16_rn(gg.y);
                            }
                        }
                    } else {
                        if (aux_out && row_ok) {
                            __nv_bfloat16* auxp = reinterpret_cast<__nv_bfloat16*>(p.aux) +
                                                  static_cast<long long>(row) * p.ldaux + n0;
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (j < ncols) auxp[j] = __float2bfloat16_rn(v[j]);
                        }
                        act_chunk(p.epilogue, v);
                    }
                    if (row_ok) {
                        if (add_res) {
                            const float* rp = p.residual + static_cast<long long>(row) * p.ldr + n0;
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (j < ncols) v[j] += rp[j];
                        }
                        if (p.c2 != nullptr) {
                            __nv_bfloat16* c2p = reinterpret_cast<__nv_bfloat16*>(p.c2) +
                                                 static_cast<long long>(row) * p.ldc2 + n0;
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (j < ncols) c2p[j] = __float2bfloat16_rn(v[j]);
                        }
                        if (c_f32) {
                            float* cp = reinterpret_cast<float*>(p.C) + static_cast<long long>(row) * p.ldc + n0;
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (j < ncols) {
                                    if (p.accumulate) atomicAdd(cp + j, v[j]);
                                    else cp[j] = v[j];
                                }
                        } else {
                            __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(p.C) +
                                                static_cast<long long>(row) * p.ldc + n0;
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (j < ncols) cp[j] = __float2bfloat16_rn(v[j]);
                        }
                    }
                }
            }
            // all of this thread's TMEM reads for the stage have completed (wait::ld above)
            tc_fence_before();
            mbar_arrive(&acc_empty[acc]);
            if (++acc == kAccStages) { acc = 0; acc_phase ^= 1u; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}


// ------------------------------------------------------------------------------------------
// Fast path: the hot Linears of the ViLT step (N a multiple of 256, aligned operands, no split-K).
// Same producer / MMA warps and ring as above; the epilogue is specialised per KIND at compile time
// and runs on SIXTEEN warps (four per TMEM lane group, each owning two of the tile's eight 32-column
// chunks), because with K = 768 a 128 x 256 tile leaves only ~10k cycles to drain 32k accumulators
// and the math-heavy epilogues are latency-bound on eight warps (ncu: IPC 0.4, stall_wait / long_sb).
//   FK_BF16       C(bf16) = acc + bias                               QKV, every plain dgrad
//   FK_GELU_SAVE  C(bf16) = gelu(acc + bias), aux(bf16) = gelu'(..)   FC1 forward
//   FK_MUL_AUX    C(bf16) = acc * aux(bf16)                           FC2 dgrad (du = dinter * gelu')
//   FK_RES_F32    C(f32)  = acc + bias + residual(f32)                O-proj / FC2 forward
// The unit of global I/O is 32 rows x 64 B (32 bf16 or 16 fp32 columns): written by thread = row into a
// 2 KB swizzled staging block, read back as 16 B per lane over consecutive addresses. Epilogue INPUTS
// (aux / residual) are fetched one chunk ahead straight into registers in that coalesced mapping and
// transposed through the same staging block when consumed: 64 KB of loads in flight per SM without
// spending shared memory, so all kinds keep the 4-stage operand ring.
// ------------------------------------------------------------------------------------------
enum FastKind { FK_BF16 = 0, FK_GELU_SAVE = 1, FK_MUL_AUX = 2, FK_RES_F32 = 3 };
constexpr int kFastEpiWarps = 16;
constexpr int kFastThreads = 64 + kFastEpiWarps * 32;
constexpr int kFastBlockN = 256;
constexpr int kFastStages = 4;
constexpr int kFastUnitBytes = 2048;
constexpr int kFastSmemBytes = kFastStages * SmemLayout<kFastBlockN>::kStageBytes + SmemLayout<kFastBlockN>::kBarrierBytes +
                               kFastEpiWarps * kFastUnitBytes + 1024;
static_assert(kFastSmemBytes <= 227 * 1024, "fast GEMM shared memory budget");

// tcgen05.wait::ld that also "produces" the loaded registers, so that no use of them can be scheduled above
// the wait when other work sits between the load and the wait
__device__ __forceinline__ void tmem_ld_wait_regs(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait_regs16(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}

// thread = row registers (16 words = 64 B per row) -> global, coalesced. g_lane already points at this lane's
// granule of row (lane >> 2); rows advance by 8 per iteration.
__device__ __forceinline__ void store_unit(uint8_t* stg, const uint32_t (&w)[16], uint8_t* g_lane, long long step8_bytes,
                                           int rows_valid, int lane) {
    __syncwarp();
#pragma unroll
    for (int g = 0; g < 4; ++g)
        *reinterpret_cast<uint4*>(stg + swz128((lane * 4 + g) * 16)) = make_uint4(w[4 * g], w[4 * g + 1], w[4 * g + 2], w[4 * g + 3]);
    __syncwarp();
    const int rsub = lane >> 2;
    uint4 val[4];                      // all four shared-memory reads in flight before the first global store
#pragma unroll
    for (int it = 0; it < 4; ++it) val[it] = *reinterpret_cast<const uint4*>(stg + swz128((it * 32 + lane) * 16));
#pragma unroll
    for (int it = 0; it < 4; ++it)
        if (it * 8 + rsub < rows_valid) *reinterpret_cast<uint4*>(g_lane + it * step8_bytes) = val[it];
}
// same for 32 B per row (16 bf16 columns): 8 words per thread, two granules per row, 16 rows per iteration.
// g_lane points at this lane's granule (lane & 1) of row (lane >> 1).
__device__ __forceinline__ void store_unit_half(uint8_t* stg, const uint32_t (&w)[8], uint8_t* g_lane, long long step16_bytes,
                                                int rows_valid, int lane) {
    __syncwarp();
#pragma unroll
    for (int g = 0; g < 2; ++g)
        *reinterpret_cast<uint4*>(stg + swz128((lane * 2 + g) * 16)) = make_uint4(w[4 * g], w[4 * g + 1], w[4 * g + 2], w[4 * g + 3]);
    __syncwarp();
    const int rsub = lane >> 1;
    uint4 val[2];
#pragma unroll
    for (int it = 0; it < 2; ++it) val[it] = *reinterpret_cast<const uint4*>(stg + swz128((it * 32 + lane) * 16));
#pragma unroll
    for (int it = 0; it < 2; ++it)
        if (it * 16 + rsub < rows_valid) *reinterpret_cast<uint4*>(g_lane + it * step16_bytes) = val[it];
}
// coalesced registers (4 granules per lane, rows lane>>2 + 8 it) -> thread = row registers
__device__ __forceinline__ void transpose_unit_in(uint8_t* stg, const uint4 (&pf)[4], uint32_t (&w)[16], int lane) {
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 4; ++it) *reinterpret_cast<uint4*>(stg + swz128((it * 32 + lane) * 16)) = pf[it];
    __syncwarp();
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const uint4 val = *reinterpret_cast<const uint4*>(stg + swz128((lane * 4 + g) * 16));
        w[4 * g] = val.x; w[4 * g + 1] = val.y; w[4 * g + 2] = val.z; w[4 * g + 3] = val.w;
    }
}
__device__ __forceinline__ void load_unit(uint4 (&pf)[4], const uint8_t* g_lane, long long step8_bytes, int rows_valid, int lane) {
    const int rsub = lane >> 2;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        pf[it] = make_uint4(0u, 0u, 0u, 0u);
        if (it * 8 + rsub < rows_valid) pf[it] = __ldg(reinterpret_cast<const uint4*>(g_lane + it * step8_bytes));
    }
}

// Tail tiles at half width. A persistent grid of W CTAs over T = m_tiles x n_tiles tiles runs ceil(T / W) rounds; when
// the last round has R <= W / 2 tiles (N = 768 at M = 15168: 357 tiles = 2 rounds + 61), those R tiles are issued as 2 R
// tiles of 128 columns: the last round costs ~0.55 of a full tile on twice as many SMs instead of 1.0 on R of them.
// Virtual tile v < full_tiles is tile v; v >= full_tiles is half ((v - full_tiles) & 1) of tile full_tiles + (v - full_tiles) / 2.
struct FastTile {
    int m_blk, n0, width;
};
__device__ __forceinline__ FastTile fast_tile(const GemmDeviceArgs& p, int v) {
    int t = v, half = 0, width = kFastBlockN;
    if (v >= p.full_tiles) {
        t = p.full_tiles + ((v - p.full_tiles) >> 1);
        half = (v - p.full_tiles) & 1;
        width = kFastBlockN / 2;
    }
    FastTile r;
    r.m_blk = t / p.n_tiles;
    r.n0 = (t - r.m_blk * p.n_tiles) * kFastBlockN + half * (kFastBlockN / 2);
    r.width = width;
    return r;
}

__device__ __forceinline__ void fast_producer_loop(const CUtensorMap& tmap_a, const CUtensorMap& tmap_b, const GemmDeviceArgs& p,
                                                   uint8_t* smem, uint64_t* full_bar, uint64_t* empty_bar) {
    using L = SmemLayout<kFastBlockN>;
    int stage = 0;
    uint32_t phase = 0;
    for (int v = blockIdx.x; v < p.total_vtiles; v += gridDim.x) {
        const FastTile tr = fast_tile(p, v);
        // a K-major B tile is one TMA box of 256 rows: a half tile still loads the whole box (rows past N are zero fill) and
        // the MMA reads its first 128 rows; an MN-major B tile is four boxes of 64 columns: a half tile loads two
        const uint32_t bytes = L::kABytes + ((p.b_mn_major && tr.width != kFastBlockN) ? L::kBBytes / 2 : L::kBBytes);
        for (int kb = 0; kb < p.k_blocks_total; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            uint8_t* sa = smem + stage * L::kStageBytes;
            uint8_t* sb = sa + L::kABytes;
            mbar_arrive_expect_tx(&full_bar[stage], bytes);
            if (!p.a_mn_major) {
                tma_load_2d(&tmap_a, &full_bar[stage], sa, kb * kBlockK, tr.m_blk * kBlockM);
            } else {
#pragma unroll
                for (int j = 0; j < kBlockM / 64; ++j)
                    tma_load_2d(&tmap_a, &full_bar[stage], sa + j * (kBlockK * 128), tr.m_blk * kBlockM + j * 64, kb * kBlockK);
            }
            if (!p.b_mn_major) {
                tma_load_2d(&tmap_b, &full_bar[stage], sb, kb * kBlockK, tr.n0);
            } else {
                for (int j = 0; j < tr.width / 64; ++j)
                    tma_load_2d(&tmap_b, &full_bar[stage], sb + j * (kBlockK * 128), tr.n0 + j * 64, kb * kBlockK);
            }
            if (++stage == kFastStages) { stage = 0; phase ^= 1u; }
        }
    }
}

__device__ __forceinline__ void fast_mma_loop(const GemmDeviceArgs& p, uint8_t* smem, uint64_t* full_bar, uint64_t* empty_bar,
                                              uint64_t* acc_full, uint64_t* acc_empty, uint32_t tmem_base) {
    using L = SmemLayout<kFastBlockN>;
    const uint32_t idesc_full = make_instr_desc(kBlockM, kFastBlockN, p.a_mn_major, p.b_mn_major);
    const uint32_t idesc_half = make_instr_desc(kBlockM, kFastBlockN / 2, p.a_mn_major, p.b_mn_major);
    const uint32_t a_lbo = p.a_mn_major ? kBlockK * 128 : 16;
    const uint32_t b_lbo = p.b_mn_major ? kBlockK * 128 : 16;
    const uint32_t a_kstep = p.a_mn_major ? 16 * 128 : 32;   // bytes per UMMA_K = 16
    const uint32_t b_kstep = p.b_mn_major ? 16 * 128 : 32;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int v = blockIdx.x; v < p.total_vtiles; v += gridDim.x) {
        const uint32_t idesc = v < p.full_tiles ? idesc_full : idesc_half;
        mbar_wait(&acc_empty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kFastBlockN);
        for (int kb = 0; kb < p.k_blocks_total; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
            const uint32_t sb = sa + L::kABytes;
#pragma unroll
            for (int kk = 0; kk < kBlockK / 16; ++kk) {
                const uint64_t da = make_smem_desc(sa + kk * a_kstep, a_lbo, 1024);
                const uint64_t db = make_smem_desc(sb + kk * b_kstep, b_lbo, 1024);
                umma_bf16(d_tmem, da, db, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);       // frees the smem slot when MMAs retire
            if (++stage == kFastStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&acc_full[acc]);              // accumulator complete -> epilogue
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1u; }
    }
}

// The epilogue warps of the specialised kernels (warps 2 .. 17): accumulator tile (this CTA's 128 TMEM lanes x tile width)
// -> registers -> epilogue arithmetic -> coalesced global stores. Shared by the one-CTA kernel (tile_rows = 128,
// row_off = 0, tiles blockIdx.x, + gridDim.x, ...) and the CTA-pair kernel (tile_rows = 256, row_off = 128 * cluster rank,
// tiles pair, + pairs, ...; the accumulator is handed back on the LEADER's acc_empty barrier).
template <int KIND, bool PAIR>
__device__ __forceinline__ void fast_epilogue(const GemmDeviceArgs& p, uint8_t* stg_base, uint64_t* acc_full, uint64_t* acc_empty,
                                              uint32_t acc_empty_leader, uint32_t tmem_base, int warp, int lane, int tile0,
                                              int tstride, int tile_rows, int row_off) {
    constexpr int BLOCK_N = kFastBlockN;
    constexpr bool kF32 = KIND == FK_RES_F32;
    constexpr bool kHasIn = KIND == FK_MUL_AUX || KIND == FK_RES_F32;
    constexpr int kUnits = kF32 ? 2 : 1;              // 64 B units per 32-column chunk
    constexpr int kColGroups = kFastEpiWarps / 4;
    const int ew = warp - 2;
    const int lane_grp = warp & 3;                    // TMEM lanes [32 * lane_grp, +32) are this warp's
    const int cg = ew >> 2;                           // chunks cg, cg + 4 of every tile
    uint8_t* stg = stg_base + ew * kFastUnitBytes;
    const int total_tiles = p.total_vtiles;
    const int rsub = lane >> 2, gj = lane & 3;
    // byte pitches: output rows / input rows, and 8 rows at a time for the coalesced mapping
    const long long c_pitch = p.ldc * (kF32 ? 4 : 2);
    const long long in_pitch = kF32 ? p.ldr * 4 : p.ldaux * 2;
    const uint8_t* in_base = kF32 ? reinterpret_cast<const uint8_t*>(p.residual) : reinterpret_cast<const uint8_t*>(p.aux);

    uint4 pf[kHasIn ? kUnits : 1][4];
    // fetch the input of chunk cc of tile t (coalesced mapping) into pf
    auto prefetch = [&](int t, int cc, int h) {
        const FastTile tp = fast_tile(p, t);
        const int r0 = tp.m_blk * tile_rows + row_off + lane_grp * 32;
        const int rv = min(32, max(0, p.M - r0));
        const uint8_t* g = in_base + (static_cast<long long>(r0) + rsub) * in_pitch +
                           static_cast<long long>(tp.n0 + cc * 32) * (kF32 ? 4 : 2) + gj * 16;
        load_unit(pf[h], g + h * 64, 8 * in_pitch, rv, lane);
    };
    if (kHasIn && tile0 < total_tiles) {
#pragma unroll
        for (int h = 0; h < kUnits; ++h) prefetch(tile0, cg, h);
    }

    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = tile0; tile < total_tiles; tile += tstride) {
        const FastTile tr = fast_tile(p, tile);
        const int n_chunks = tr.width / 32;
        const int row0 = tr.m_blk * tile_rows + row_off + lane_grp * 32;
        const int rows_valid = min(32, max(0, p.M - row0));
        mbar_wait(&acc_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) + static_cast<uint32_t>(acc * BLOCK_N);
#pragma unroll 1
        for (int c = cg; c < n_chunks; c += kColGroups) {
            const int n0 = tr.n0 + c * 32;
            int nt = tile, nc = c + kColGroups;           // this warp's next chunk
            if (nc >= n_chunks) { nt = tile + tstride; nc = cg; }
            uint8_t* c_lane = reinterpret_cast<uint8_t*>(p.C) + (static_cast<long long>(row0) + rsub) * c_pitch +
                              static_cast<long long>(n0) * (kF32 ? 4 : 2) + gj * 16;
            if constexpr (KIND == FK_RES_F32) {
                // two 16-column halves, each with its own in-flight residual unit
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t r[16], res[16];
                    tmem_ld_32x16(t_row + static_cast<uint32_t>(c * 32 + h * 16), r);
                    transpose_unit_in(stg, pf[h], res, lane);
                    if (nt < total_tiles) prefetch(nt, nc, h);
                    tmem_ld_wait_regs16(r);
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (p.bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + h * 16 + j));
                        r[j] = __float_as_uint(__uint_as_float(r[j]) + b4.x);
                        r[j + 1] = __float_as_uint(__uint_as_float(r[j + 1]) + b4.y);
                        r[j + 2] = __float_as_uint(__uint_as_float(r[j + 2]) + b4.z);
                        r[j + 3] = __float_as_uint(__uint_as_float(r[j + 3]) + b4.w);
                    }
                    const int hsub = lane >> 1, hg = lane & 1;
                    if (p.aux != nullptr) {          // bf16 copy of acc + bias (before the residual): adapter input
                        uint32_t hw[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) hw[j] = pack_bf16(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
                        uint8_t* a_lane = reinterpret_cast<uint8_t*>(p.aux) + (static_cast<long long>(row0) + hsub) * (p.ldaux * 2) +
                                          static_cast<long long>(n0 + h * 16) * 2 + hg * 16;
                        store_unit_half(stg, hw, a_lane, 16 * p.ldaux * 2, rows_valid, lane);
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(res[j]));
                    if (p.c2 != nullptr) {           // bf16 copy of the final value
                        uint32_t hw[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) hw[j] = pack_bf16(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
                        uint8_t* c2_lane = reinterpret_cast<uint8_t*>(p.c2) + (static_cast<long long>(row0) + hsub) * (p.ldc2 * 2) +
                                           static_cast<long long>(n0 + h * 16) * 2 + hg * 16;
                        store_unit_half(stg, hw, c2_lane, 16 * p.ldc2 * 2, rows_valid, lane);
                    }
                    store_unit(stg, r, c_lane + h * 64, 8 * c_pitch, rows_valid, lane);
                }
            } else {
                uint32_t r[32];
                tmem_ld_32x32(t_row + static_cast<uint32_t>(c * 32), r);
                uint32_t ain[16];
                if constexpr (KIND == FK_MUL_AUX) {
                    transpose_unit_in(stg, pf[0], ain, lane);
                    if (nt < total_tiles) prefetch(nt, nc, 0);
                }
                tmem_ld_wait_regs(r);
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                if (KIND != FK_MUL_AUX && p.bias != nullptr) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
                        v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
                    }
                }
                uint32_t pk[16];
                if constexpr (KIND == FK_MUL_AUX) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float2 a = unpack_bf16(ain[j]);
                        pk[j] = pack_bf16(v[2 * j] * a.x, v[2 * j + 1] * a.y);
                    }
                } else if constexpr (KIND == FK_GELU_SAVE) {
                    uint32_t gpk[16];
                    gelu_and_grad_chunk(v, gpk);
                    uint8_t* a_lane = reinterpret_cast<uint8_t*>(p.aux) + (static_cast<long long>(row0) + rsub) * (p.ldaux * 2) +
                                      static_cast<long long>(n0) * 2 + gj * 16;
                    store_unit(stg, gpk, a_lane, 8 * p.ldaux * 2, rows_valid, lane);
#pragma unroll
                    for (int j = 0; j < 16; ++j) pk[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) pk[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
                }
                store_unit(stg, pk, c_lane, 8 * c_pitch, rows_valid, lane);
            }
        }
        tc_fence_before();
        if constexpr (PAIR) mbar_arrive_cluster(acc_empty_leader + static_cast<uint32_t>(acc) * 8u);
        else mbar_arrive(&acc_empty[acc]);
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1u; }
    }

}

template <int KIND>
__global__ void __launch_bounds__(kFastThreads, 1)
gemm_fast_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const GemmDeviceArgs p) {
    constexpr int BLOCK_N = kFastBlockN;
    using L = SmemLayout<BLOCK_N>;
    constexpr int kStages = kFastStages;
    constexpr uint32_t kTmemCols = kAccStages * BLOCK_N;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    uint8_t* bar_base = smem + kStages * L::kStageBytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
    uint64_t* empty_bar = full_bar + L::kMaxStages;
    uint64_t* acc_full = empty_bar + L::kMaxStages;
    uint64_t* acc_empty = acc_full + kAccStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kAccStages);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < kAccStages; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], kFastEpiWarps * 32);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (!p.skip_pdl_wait) pdl_wait();     // everything above overlapped the previous kernel's tail; operands / epilogue tensors come after

    if (warp == 0) {
        if (lane == 0) fast_producer_loop(tmap_a, tmap_b, p, smem, full_bar, empty_bar);
    } else if (warp == 1) {
        if (lane == 0) fast_mma_loop(p, smem, full_bar, empty_bar, acc_full, acc_empty, tmem_base);
    } else {
        fast_epilogue<KIND, false>(p, smem + kStages * L::kStageBytes + L::kBarrierBytes, acc_full, acc_empty, 0u, tmem_base, warp,
                                   lane, static_cast<int>(blockIdx.x), static_cast<int>(gridDim.x), kBlockM, 0);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------
// CTA-pair variant of the specialised kernels (tcgen05 cta_group::2). A 2-CTA cluster owns a 256 x 256 tile: CTA r stages
// rows [128 r, 128 r + 128) of the A tile and columns [128 r, +128) of the B tile (32 KB per k-block instead of 48 KB: six
// ring stages), the leader issues 256 x 256 x 16 MMAs that read both CTAs' shared memory, and each CTA's TMEM receives
// the accumulator rows of its own A half -- so the epilogue is the one-CTA epilogue on rows offset by 128 r. Per FLOP every
// SM fetches 2/3 of the operand bytes of the 128 x 256 kernel from L2 and from shared memory (the
// K = 768 launches are bound by that traffic). Half-width tail tiles work as before (MMA N = 128, 64 B columns per CTA).
// Used for K-major A operands (forward Linears and dgrads) of the kinds pair_kind_enabled() selects.
// ------------------------------------------------------------------------------------------
constexpr int kPairStages = 6;
constexpr int kPairTileM = 2 * kBlockM;
constexpr int kPairABytes = kBlockM * kBlockK * 2;                 // this CTA's 128 rows of A
constexpr int kPairBBytes = (kFastBlockN / 2) * kBlockK * 2;       // this CTA's 128 columns of B
constexpr int kPairStageBytes = kPairABytes + kPairBBytes;
constexpr int kPairBarrierBytes = 1024;
constexpr int kPairSmemBytes = kPairStages * kPairStageBytes + kPairBarrierBytes + kFastEpiWarps * kFastUnitBytes + 1024;
static_assert(kPairSmemBytes <= 227 * 1024, "pair GEMM shared memory budget");

__device__ __forceinline__ void pair_producer_loop(const CUtensorMap& tmap_a, const CUtensorMap& tmap_b, const GemmDeviceArgs& p,
                                                   uint8_t* smem, uint64_t* full_bar, uint64_t* empty_bar, uint32_t rank,
                                                   int pair, int n_pairs) {
    int stage = 0;
    uint32_t phase = 0;
    for (int v = pair; v < p.total_vtiles; v += n_pairs) {
        const FastTile tr = fast_tile(p, v);
        const int half_w = tr.width / 2;                            // B columns staged by each CTA
        // a K-major B half is one TMA box of 128 rows (a half-width tile still loads the box and the MMA reads its first 64
        // rows); an MN-major B half is half_w / 64 boxes of 64 columns
        const uint32_t b_bytes = p.b_mn_major ? static_cast<uint32_t>(half_w / 64) * (kBlockK * 128) : kPairBBytes;
        for (int kb = 0; kb < p.k_blocks_total; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            uint8_t* sa = smem + stage * kPairStageBytes;
            uint8_t* sb = sa + kPairABytes;
            const uint32_t full_leader = cluster_map_shared(&full_bar[stage], 0);
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * (kPairABytes + b_bytes));   // both CTAs' bytes
            tma_load_2d_pair(&tmap_a, full_leader, sa, kb * kBlockK, tr.m_blk * kPairTileM + static_cast<int>(rank) * kBlockM);
            if (!p.b_mn_major) {
                tma_load_2d_pair(&tmap_b, full_leader, sb, kb * kBlockK, tr.n0 + static_cast<int>(rank) * half_w);
            } else {
                for (int j = 0; j < half_w / 64; ++j)
                    tma_load_2d_pair(&tmap_b, full_leader, sb + j * (kBlockK * 128), tr.n0 + static_cast<int>(rank) * half_w + j * 64,
                                     kb * kBlockK);
            }
            if (++stage == kPairStages) { stage = 0; phase ^= 1u; }
        }
    }
}

// leader CTA only
__device__ __forceinline__ void pair_mma_loop(const GemmDeviceArgs& p, uint8_t* smem, uint64_t* full_bar, uint64_t* empty_bar,
                                              uint64_t* acc_full, uint64_t* acc_empty, uint32_t tmem_base, int pair, int n_pairs) {
    const uint32_t idesc_full = make_instr_desc(kPairTileM, kFastBlockN, 0, p.b_mn_major);
    const uint32_t idesc_half = make_instr_desc(kPairTileM, kFastBlockN / 2, 0, p.b_mn_major);
    const uint32_t b_lbo = p.b_mn_major ? kBlockK * 128 : 16;
    const uint32_t b_kstep = p.b_mn_major ? 16 * 128 : 32;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int v = pair; v < p.total_vtiles; v += n_pairs) {
        const uint32_t idesc = v < p.full_tiles ? idesc_full : idesc_half;
        mbar_wait(&acc_empty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kFastBlockN);
        for (int kb = 0; kb < p.k_blocks_total; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * kPairStageBytes);
            const uint32_t sb = sa + kPairABytes;
#pragma unroll
            for (int kk = 0; kk < kBlockK / 16; ++kk) {
                const uint64_t da = make_smem_desc(sa + kk * 32, 16, 1024);
                const uint64_t db = make_smem_desc(sb + kk * b_kstep, b_lbo, 1024);
                umma_bf16_pair(d_tmem, da, db, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
            }
            umma_commit_pair(&empty_bar[stage]);      // frees the slot in BOTH CTAs when the MMAs retire
            if (++stage == kPairStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit_pair(&acc_full[acc]);             // accumulator complete -> both CTAs' epilogues
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1u; }
    }
}

template <int KIND>
__global__ void __launch_bounds__(kFastThreads, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const GemmDeviceArgs p) {
    constexpr uint32_t kTmemCols = kAccStages * kFastBlockN;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* bar_base = smem + kPairStages * kPairStageBytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
    uint64_t* empty_bar = full_bar + 8;
    uint64_t* acc_full = empty_bar + 8;
    uint64_t* acc_empty = acc_full + kAccStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kAccStages);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = static_cast<int>(blockIdx.x >> 1);
    const int n_pairs = static_cast<int>(gridDim.x >> 1);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int s = 0; s < kPairStages; ++s) {
            mbar_init(&full_bar[s], 1);        // the leader's expect_tx arrival (+ the bytes of both CTAs)
            mbar_init(&empty_bar[s], 1);       // one multicast tcgen05.commit
        }
        for (int s = 0; s < kAccStages; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], 2 * kFastEpiWarps * 32);     // the epilogue threads of both CTAs (leader's barrier is the one used)
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc_pair(tmem_slot, kTmemCols);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    __syncthreads();                           // the TMEM address in shared memory (racecheck does not take barrier.cluster as CTA-level ordering)
    cluster_sync_all();                        // barriers of BOTH CTAs are initialised before any remote arrival
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (!p.skip_pdl_wait) pdl_wait();

    if (warp == 0) {
        if (lane == 0) pair_producer_loop(tmap_a, tmap_b, p, smem, full_bar, empty_bar, rank, pair, n_pairs);
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) pair_mma_loop(p, smem, full_bar, empty_bar, acc_full, acc_empty, tmem_base, pair, n_pairs);
    } else {
        fast_epilogue<KIND, true>(p, smem + kPairStages * kPairStageBytes + kPairBarrierBytes, acc_full, acc_empty,
                                  cluster_map_shared(&acc_empty[0], 0), tmem_base, warp, lane, pair, n_pairs, kPairTileM,
                                  static_cast<int>(rank) * kBlockM);
    }

    tc_fence_before();
    cluster_sync_all();                        // both CTAs are done with the pair's tensor memory and barriers
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------
// CTA-pair weight-gradient kernel: dW[M, N] (fp32) += A^T B with BOTH operands MN-major (A = dY [tokens, M],
// B = X [tokens, N], read in place), split over the token dimension, fp32 vector atomics into dW. The wgrads are all
// main loop (119+ k-blocks per tile, a 64 KB epilogue), which is where halving the B fetch per SM pays in full.
// 256 x 256 tile per pair: CTA r stages dY columns [128 r, +128) and X columns [128 r, +128) of the tile as two
// 64-wide boxes each. Needs M % 256 == 0 and N % 256 == 0 (every Linear of ViLT / BERT-base).
// ------------------------------------------------------------------------------------------
constexpr int kWgradSmemBytes = kPairStages * kPairStageBytes + kPairBarrierBytes + kNumEpiWarps * kEpiHalfBytes + 1024;
static_assert(kWgradSmemBytes <= 227 * 1024, "pair wgrad shared memory budget");

__global__ void __launch_bounds__(kNumThreads, 1)
gemm_pair_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                       const GemmDeviceArgs p) {
    constexpr uint32_t kTmemCols = kAccStages * kFastBlockN;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* bar_base = smem + kPairStages * kPairStageBytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
    uint64_t* empty_bar = full_bar + 8;
    uint64_t* acc_full = empty_bar + 8;
    uint64_t* acc_empty = acc_full + kAccStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kAccStages);
    uint8_t* ones = bar_base + 256;                // 768 B of bf16 1.0: the B operand of the column-sum MMAs (p.colsum_a)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = static_cast<int>(blockIdx.x >> 1);
    const int n_pairs = static_cast<int>(gridDim.x >> 1);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int s = 0; s < kPairStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < kAccStages; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], 2 * kNumEpiWarps * 32);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc_pair(tmem_slot, kTmemCols);
        tmem_relinquish_pair();
        if (p.colsum_a != nullptr) {               // (same warp as the allocation: racecheck orders the two shared-memory writers)
            for (int i = lane; i < 768 / 4; i += 32) reinterpret_cast<uint32_t*>(ones)[i] = 0x3F803F80u;
            fence_proxy_async();                   // read by the tensor core (async proxy) of the leader CTA
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (!p.skip_pdl_wait) pdl_wait();

    const int tiles_mn = p.m_tiles * p.n_tiles;
    const int total_tiles = tiles_mn * p.split_k;
    constexpr int kBox = kBlockK * 128;            // one 64 (mn) x 64 (k) box
    // Bias gradient next to the weight gradient (p.colsum_a, host: only when every pair owns at most ONE tile, so the second
    // accumulator stage is free): colsum_a[m] += sum_k A[k, m] is one more MMA per k-step, D2[256, 16] += A^T[256, 16k] ONES[16k, 16]
    // into TMEM columns [256, 272) -- every column of D2 holds the sums. The tiles of one m-block share the work: the tile
    // with n-block j takes the k-blocks kb % n_tiles == j (each k-block exactly once per m-block and split), and every one
    // adds its partial sums with 128 atomics per CTA. Cost: 8 KB of operand fetch per k-step on 1 / n_tiles of the k-blocks,
    // instead of a separate streaming pass over dY (13 us for the FC1 bias at B = 64).
    constexpr uint32_t kCsCol = kFastBlockN;       // first column of the second accumulator stage

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = pair; tile < total_tiles; tile += n_pairs) {
                const int split = tile / tiles_mn;
                const int mn = tile - split * tiles_mn;
                const int m_blk = mn / p.n_tiles;
                const int n_blk = mn - m_blk * p.n_tiles;
                const int kb0 = split * p.k_blocks_per_split;
                const int kb1 = min(kb0 + p.k_blocks_per_split, p.k_blocks_total);
                const int m0 = m_blk * kPairTileM + static_cast<int>(rank) * kBlockM;
                const int n0 = n_blk * kFastBlockN + static_cast<int>(rank) * (kFastBlockN / 2);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1u);
                    uint8_t* sa = smem + stage * kPairStageBytes;
                    uint8_t* sb = sa + kPairABytes;
                    const uint32_t full_leader = cluster_map_shared(&full_bar[stage], 0);
                    if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * kPairStageBytes);
#pragma unroll
                    for (int j = 0; j < 2; ++j) tma_load_2d_pair(&tmap_a, full_leader, sa + j * kBox, m0 + j * 64, kb * kBlockK);
#pragma unroll
                    for (int j = 0; j < 2; ++j) tma_load_2d_pair(&tmap_b, full_leader, sb + j * kBox, n0 + j * 64, kb * kBlockK);
                    if (++stage == kPairStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            const uint32_t idesc = make_instr_desc(kPairTileM, kFastBlockN, 1, 1);
            const uint32_t idesc_cs = make_instr_desc(kPairTileM, 16, 1, 0);
            // ONES as a K-major, un-swizzled B operand: 8 rows (n) x 16 k per CTA = two 128-byte core matrices (LBO apart)
            const uint64_t d_ones = make_smem_desc_noswizzle(smem_u32(ones), 128, 256);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = pair; tile < total_tiles; tile += n_pairs) {
                const int split = tile / tiles_mn;
                const int n_blk = (tile - split * tiles_mn) % p.n_tiles;
                const int kb0 = split * p.k_blocks_per_split;
                const int kb1 = min(kb0 + p.k_blocks_per_split, p.k_blocks_total);
                mbar_wait(&acc_empty[acc], acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kFastBlockN);
                int cs_next = p.colsum_a != nullptr ? kb0 + (n_blk - kb0 % p.n_tiles + p.n_tiles) % p.n_tiles : kb1;
                uint32_t cs_acc = 0u;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * kPairStageBytes);
                    const uint32_t sb = sa + kPairABytes;
#pragma unroll
                    for (int kk = 0; kk < kBlockK / 16; ++kk) {
                        const uint64_t da = make_smem_desc(sa + kk * (16 * 128), kBox, 1024);
                        const uint64_t db = make_smem_desc(sb + kk * (16 * 128), kBox, 1024);
                        umma_bf16_pair(d_tmem, da, db, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
                    }
                    if (kb == cs_next) {
#pragma unroll
                        for (int kk = 0; kk < kBlockK / 16; ++kk) {
                            umma_bf16_pair(tmem_base + kCsCol, make_smem_desc(sa + kk * (16 * 128), kBox, 1024), d_ones, idesc_cs, cs_acc);
                            cs_acc = 1u;
                        }
                        cs_next += p.n_tiles;
                    }
                    umma_commit_pair(&empty_bar[stage]);
                    if (++stage == kPairStages) { stage = 0; phase ^= 1u; }
                }
                umma_commit_pair(&acc_full[acc]);
                if (++acc == kAccStages) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        const int ew = warp - 2;                  // 0..7
        const int lane_grp = warp & 3;
        const int col_half = ew >> 2;
        uint8_t* scratch = smem + kPairStages * kPairStageBytes + kPairBarrierBytes + ew * kEpiHalfBytes;
        const uint32_t acc_empty_leader = cluster_map_shared(&acc_empty[0], 0);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = pair; tile < total_tiles; tile += n_pairs) {
            const int split = tile / tiles_mn;
            const int mn = tile - split * tiles_mn;
            const int m_blk = mn / p.n_tiles;
            const int n_blk = mn - m_blk * p.n_tiles;
            const int row0 = m_blk * kPairTileM + static_cast<int>(rank) * kBlockM + lane_grp * 32;
            mbar_wait(&acc_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) + static_cast<uint32_t>(acc * kFastBlockN);
#pragma unroll 1
            for (int c = col_half; c < kFastBlockN / 32; c += 2) {
                const int n0 = n_blk * kFastBlockN + c * 32;
                uint32_t r[32];
                tmem_ld_32x32(t_row + static_cast<uint32_t>(c * 32), r);
                tmem_ld_wait();
                store_rows<8>(scratch, r, reinterpret_cast<uint8_t*>(p.C) + (static_cast<long long>(row0) * p.ldc + n0) * 4,
                              p.ldc * 4, 32, lane, true);
            }
            if (p.colsum_a != nullptr && col_half == 0) {
                const int kb0 = split * p.k_blocks_per_split;
                const int kb1 = min(kb0 + p.k_blocks_per_split, p.k_blocks_total);
                if (kb0 + (n_blk - kb0 % p.n_tiles + p.n_tiles) % p.n_tiles < kb1) {      // this tile issued column-sum MMAs
                    uint32_t v;
                    tmem_ld_32x1(tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) + kCsCol, v);
                    tmem_ld_wait();
                    atomicAdd(p.colsum_a + row0 + lane, __uint_as_float(v));
                }
            }
            tc_fence_before();
            mbar_arrive_cluster(acc_empty_leader + static_cast<uint32_t>(acc) * 8u);
            if (++acc == kAccStages) { acc = 0; acc_phase ^= 1u; }
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------
// Fused adapter bottleneck (adapters/modeling.py:160-179 as ViLT wires it, mixins/vilt.py:23-125):
//     forward   out = y + W_u act(W_d y + b_d) + b_u                 y = the site's input (bf16 copy) / residual stream (fp32)
//     backward  dy  = dout + (dout W_u * act'(pre)) W_d              (+ dpre for the two weight gradients, + its column sums)
// ONE launch per site and direction instead of two generic GEMM launches on ragged widths: a CTA owns 128 rows and runs
//     stage 1   T[128, r] = A[128, d] B1          12 k-blocks of A through a 4-stage TMA ring, accumulator = 64 TMEM columns
//     stage 2   forward: z = act(T + b_d)   backward: dpre = T * act'(pre)      thread = row; the bf16 result goes to global
//               memory (kept for the weight gradients) AND, as a K-major A operand tile, to shared memory -- the r-wide
//               intermediate never round-trips through HBM between the two projections
//     stage 3   C[128, d] (fp32 residual stream, in place) += Z[128, r] B2 (+ b_u) in six 128-column chunks through two TMEM
//               accumulators, the chunk's residual read and its result written by the epilogue warps in coalesced 64 B units
// B1 / B2 are the adapter's two weight matrices read in place: forward W_d, W_u K-major; backward W_u, W_d MN-major.
// r <= 64 (CLiMB: reduction factor 16 -> r = 48), r % 16 == 0, d % 128 == 0.
// ------------------------------------------------------------------------------------------
struct AdapterArgs {
    int M, d, r;
    int backward;               // 0 forward, 1 backward
    int act;                    // CLIMB_EPI_SWISH / CLIMB_EPI_RELU
    const float* bias1;         // forward: b_d [r]
    const float* bias2;         // forward: b_u [d]
    __nv_bfloat16* pre;         // [M, r]  forward: written (pre-activation), backward: read
    __nv_bfloat16* z;           // [M, r]  forward: written act(pre), backward: written dpre
    const float* c_in;          // fp32 [M, d] residual input
    float* c_out;               // fp32 [M, d] result (may alias c_in) or null
    __nv_bfloat16* c2;          // optional bf16 copy of the result [M, d]
    float* colsum_z;            // backward, optional [r]: += column sums of dpre (the down-projection's bias gradient)
    int n_tiles;
};
constexpr int kAdStages = 4;
constexpr int kAdStageBytes = 16384 + 8192;      // A k-block [128 x 64] + B1 k-block [64 x 64]
constexpr int kAdZ = kAdStages * kAdStageBytes;                 // Z tile [128 x 64] bf16
constexpr int kAdB2 = kAdZ + 16384;                             // 2 x 16 KB ring of B2 chunks
constexpr int kAdStg = kAdB2 + 2 * 16384;                       // 8 warps x 2 KB staging
constexpr int kAdBar = kAdStg + 8 * 2048;
constexpr int kAdSmemBytes = kAdBar + 256 + 1024;
constexpr int kAdThreads = 64 + 8 * 32;

__global__ void __launch_bounds__(kAdThreads, 1)
adapter_fused_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b1,
                     const __grid_constant__ CUtensorMap map_b2, const AdapterArgs p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kAdBar);
    uint64_t* full1 = bars;              // [4] tx
    uint64_t* empty1 = bars + 4;         // [4] commit
    uint64_t* full2 = bars + 8;          // [2] tx
    uint64_t* empty2 = bars + 10;        // [2] commit
    uint64_t* t_full = bars + 12;        // stage-1 accumulator complete
    uint64_t* t_empty = bars + 13;       // ... and read out (4 warp arrivals)
    uint64_t* z_full = bars + 14;        // Z tile is in shared memory (4 warp arrivals)
    uint64_t* acc_full = bars + 15;      // [2]
    uint64_t* acc_empty = bars + 17;     // [2] 8 warp arrivals
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kb_n = p.d / kBlockK, ch_n = p.d / 128, ks_n = p.r / 16;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b1);
        tma_prefetch_desc(&map_b2);
        for (int s = 0; s < kAdStages; ++s) { mbar_init(&full1[s], 1); mbar_init(&empty1[s], 1); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&full2[s], 1); mbar_init(&empty2[s], 1);
            mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 8);
        }
        mbar_init(t_full, 1); mbar_init(t_empty, 4); mbar_init(z_full, 4);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            int s1 = 0, s2 = 0;
            uint32_t ph1 = 0, ph2 = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                for (int kb = 0; kb < kb_n; ++kb) {
                    mbar_wait(&empty1[s1], ph1 ^ 1u);
                    uint8_t* sa = smem + s1 * kAdStageBytes;
                    mbar_arrive_expect_tx(&full1[s1], kAdStageBytes);
                    tma_load_2d(&map_a, &full1[s1], sa, kb * kBlockK, tile * kBlockM);
                    if (!p.backward) tma_load_2d(&map_b1, &full1[s1], sa + 16384, kb * kBlockK, 0);      // W_d [r, d]: rows past r are zero fill
                    else             tma_load_2d(&map_b1, &full1[s1], sa + 16384, 0, kb * kBlockK);      // W_u [d, r]: columns past r are zero fill
                    if (++s1 == kAdStages) { s1 = 0; ph1 ^= 1u; }
                }
                for (int c = 0; c < ch_n; ++c) {
                    mbar_wait(&empty2[s2], ph2 ^ 1u);
                    uint8_t* sb = smem + kAdB2 + s2 * 16384;
                    mbar_arrive_expect_tx(&full2[s2], 16384);
                    if (!p.backward) {
                        tma_load_2d(&map_b2, &full2[s2], sb, 0, c * 128);                               // W_u [d, r]: 128 rows (n) x 64 (k, zero fill past r)
                    } else {
                        tma_load_2d(&map_b2, &full2[s2], sb, c * 128, 0);                               // W_d [r, d]: 64 (k rows, zero fill past r) x 64 n
                        tma_load_2d(&map_b2, &full2[s2], sb + 8192, c * 128 + 64, 0);
                    }
                    if (++s2 == 2) { s2 = 0; ph2 ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc1 = make_instr_desc(kBlockM, 64, 0, p.backward);
            const uint32_t idesc2 = make_instr_desc(kBlockM, 128, 0, p.backward);
            const uint32_t b1_lbo = p.backward ? kBlockK * 128 : 16, b1_kstep = p.backward ? 16 * 128 : 32;
            const uint32_t b2_lbo = p.backward ? 64 * 128 : 16, b2_kstep = p.backward ? 16 * 128 : 32;
            int s1 = 0, s2 = 0, it = 0;
            uint32_t ph1 = 0, ph2 = 0, acc_n = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
                if (it > 0) mbar_wait(t_empty, (it - 1) & 1);
                tc_fence_after();
                for (int kb = 0; kb < kb_n; ++kb) {
                    mbar_wait(&full1[s1], ph1);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + s1 * kAdStageBytes), sb = sa + 16384;
#pragma unroll
                    for (int kk = 0; kk < kBlockK / 16; ++kk)
                        umma_bf16(tmem, make_smem_desc(sa + kk * 32, 16, 1024), make_smem_desc(sb + kk * b1_kstep, b1_lbo, 1024), idesc1,
                                  (kb > 0 || kk > 0) ? 1u : 0u);
                    umma_commit(&empty1[s1]);
                    if (++s1 == kAdStages) { s1 = 0; ph1 ^= 1u; }
                }
                umma_commit(t_full);
                mbar_wait(z_full, it & 1);
                tc_fence_after();
                const uint32_t sz = smem_u32(smem + kAdZ);
                for (int c = 0; c < ch_n; ++c, ++acc_n) {
                    const uint32_t ab = acc_n & 1u;
                    mbar_wait(&full2[s2], ph2);
                    if (acc_n >= 2) mbar_wait(&acc_empty[ab], ((acc_n >> 1) - 1) & 1u);
                    tc_fence_after();
                    const uint32_t sb = smem_u32(smem + kAdB2 + s2 * 16384);
                    for (int kk = 0; kk < ks_n; ++kk)
                        umma_bf16(tmem + 128 + ab * 128, make_smem_desc(sz + kk * 32, 16, 1024),
                                  make_smem_desc(sb + kk * b2_kstep, b2_lbo, 1024), idesc2, kk > 0 ? 1u : 0u);
                    umma_commit(&empty2[s2]);
                    umma_commit(&acc_full[ab]);
                    if (++s2 == 2) { s2 = 0; ph2 ^= 1u; }
                }
            }
        }
    } else {
        const int ew = warp - 2, lg = warp & 3, sub = ew >> 2;       // TMEM lane group (warp % 4), which half of a chunk's units
        uint8_t* stg = smem + kAdStg + ew * 2048;
        const uint32_t t_row = tmem + (static_cast<uint32_t>(lg * 32) << 16);
        const int rsub = lane >> 2, gj = lane & 3;
        int it = 0;
        uint32_t acc_n = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
            const int row0 = tile * kBlockM + lg * 32;
            const int rows_valid = min(32, max(0, p.M - row0));
            // the residual of the first chunk travels while stages 1 and 2 run
            const uint8_t* in_base = reinterpret_cast<const uint8_t*>(p.c_in) + (static_cast<long long>(row0) + rsub) * (p.d * 4LL) + gj * 16;
            uint8_t* out_base = p.c_out ? reinterpret_cast<uint8_t*>(p.c_out) + (static_cast<long long>(row0) + rsub) * (p.d * 4LL) + gj * 16 : nullptr;
            const long long pitch8 = 8LL * p.d * 4;
            uint4 pf[4][4];
#pragma unroll
            for (int q = 0; q < 4; ++q) load_unit(pf[q], in_base + static_cast<long long>((sub * 4 + q) * 16) * 4, pitch8, rows_valid, lane);
            // ---- stage 2 (one warp per lane group): the r-wide intermediate ----
            if (sub == 0) {
                mbar_wait(t_full, it & 1);
                tc_fence_after();
                const int row = row0 + lane;
                const bool valid = lane < rows_valid;
                uint8_t* zrow = smem + kAdZ + (lg * 32 + lane) * 128;
#pragma unroll 1
                for (int u = 0; u < 4; ++u) {                 // 16 columns at a time
                    uint32_t zp[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
                    if (u * 16 < p.r) {
                        uint32_t t[16];
                        tmem_ld_32x16(t_row + u * 16, t);
                        float v[16];
                        uint32_t prev[8];
                        if (p.backward) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) prev[j] = 0u;
                            if (valid) {
                                const uint4* src = reinterpret_cast<const uint4*>(p.pre + static_cast<long long>(row) * p.r + u * 16);
                                const uint4 a = src[0], b = src[1];
                                prev[0] = a.x; prev[1] = a.y; prev[2] = a.z; prev[3] = a.w;
                                prev[4] = b.x; prev[5] = b.y; prev[6] = b.z; prev[7] = b.w;
                            }
                        }
                        tmem_ld_wait_regs16(t);
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(t[j]);
                        if (!p.backward) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] += __ldg(p.bias1 + u * 16 + j);
                            uint32_t pp[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) pp[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
                            if (valid) {
                                uint4* dst = reinterpret_cast<uint4*>(p.pre + static_cast<long long>(row) * p.r + u * 16);
                                dst[0] = make_uint4(pp[0], pp[1], pp[2], pp[3]);
                                dst[1] = make_uint4(pp[4], pp[5], pp[6], pp[7]);
                            }
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = p.act == CLIMB_EPI_RELU ? fmaxf(v[j], 0.0f) : swish_f(v[j]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float2 a = unpack_bf16(prev[j]);
                                if (p.act == CLIMB_EPI_RELU) {
                                    v[2 * j] = a.x > 0.0f ? v[2 * j] : 0.0f;
                                    v[2 * j + 1] = a.y > 0.0f ? v[2 * j + 1] : 0.0f;
                                } else {
                                    v[2 * j] *= dswish_f(a.x);
                                    v[2 * j + 1] *= dswish_f(a.y);
                                }
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) zp[j] = valid ? pack_bf16(v[2 * j], v[2 * j + 1]) : 0u;
                        if (valid) {
                            uint4* dst = reinterpret_cast<uint4*>(p.z + static_cast<long long>(row) * p.r + u * 16);
                            dst[0] = make_uint4(zp[0], zp[1], zp[2], zp[3]);
                            dst[1] = make_uint4(zp[4], zp[5], zp[6], zp[7]);
                        }
                        if (p.backward && p.colsum_z != nullptr) {
                            // column sums of the bf16-rounded dpre over the warp's 32 rows, one atomic per column
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                float2 f = unpack_bf16(zp[j]);
#pragma unroll
                                for (int o = 16; o > 0; o >>= 1) {
                                    f.x += __shfl_xor_sync(0xffffffffu, f.x, o);
                                    f.y += __shfl_xor_sync(0xffffffffu, f.y, o);
                                }
                                if (lane == 0) {
                                    atomicAdd(p.colsum_z + u * 16 + 2 * j, f.x);
                                    atomicAdd(p.colsum_z + u * 16 + 2 * j + 1, f.y);
                                }
                            }
                        }
                    }
                    // granules 2u, 2u + 1 of the row in the K-major, 128B-swizzled Z tile (columns past r are zeros)
                    const uint32_t rr = static_cast<uint32_t>(lg * 32 + lane) & 7u;
                    *reinterpret_cast<uint4*>(zrow + (((2 * u) ^ rr) << 4)) = make_uint4(zp[0], zp[1], zp[2], zp[3]);
                    *reinterpret_cast<uint4*>(zrow + (((2 * u + 1) ^ rr) << 4)) = make_uint4(zp[4], zp[5], zp[6], zp[7]);
                }
                tc_fence_before();
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(t_empty);
                    mbar_arrive(z_full);
                }
            }
            // ---- stage 3: residual stream += Z B2, 16-column units; warp `sub` takes units sub * 4 .. sub * 4 + 3 of each chunk.
            //      The residual of a whole chunk (4 units = 8 KB per warp, 64 KB per SM) is in flight at any time: every unit's
            //      registers are refilled with the NEXT chunk's data as soon as they have been consumed ----
            for (int c = 0; c < ch_n; ++c, ++acc_n) {
                const uint32_t ab = acc_n & 1u;
                mbar_wait(&acc_full[ab], (acc_n >> 1) & 1u);
                tc_fence_after();
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int col = c * 128 + (sub * 4 + q) * 16;
                    uint32_t a[16], res[16];
                    tmem_ld_32x16(t_row + 128 + ab * 128 + (sub * 4 + q) * 16, a);
                    transpose_unit_in(stg, pf[q], res, lane);
                    if (c + 1 < ch_n) load_unit(pf[q], in_base + static_cast<long long>(col + 128) * 4, pitch8, rows_valid, lane);
                    tmem_ld_wait_regs16(a);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float v = __uint_as_float(a[j]) + __uint_as_float(res[j]);
                        if (p.bias2 != nullptr) v += __ldg(p.bias2 + col + j);
                        a[j] = __float_as_uint(v);
                    }
                    if (p.c2 != nullptr) {
                        uint32_t hw[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) hw[j] = pack_bf16(__uint_as_float(a[2 * j]), __uint_as_float(a[2 * j + 1]));
                        uint8_t* c2_lane = reinterpret_cast<uint8_t*>(p.c2) + (static_cast<long long>(row0) + (lane >> 1)) * (p.d * 2LL) +
                                           static_cast<long long>(col) * 2 + (lane & 1) * 16;
                        store_unit_half(stg, hw, c2_lane, 16LL * p.d * 2, rows_valid, lane);
                    }
                    if (out_base != nullptr) store_unit(stg, a, out_base + static_cast<long long>(col) * 4, pitch8, rows_valid, lane);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[ab]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// ------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) !=
            cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || sym == nullptr) {
        cudaGetLastError();
        return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(sym);
    return fn;
}

// 2-D bf16 tensor map: `inner` contiguous elements, `outer` rows `ld` elements apart.
int make_tmap_2d(CUtensorMap* map, const void* ptr, long long inner, long long outer, long long ld,
                 int box_inner, int box_outer) {
    EncodeTiledFn fn = get_encode_fn();
    CLIMB_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
    CLIMB_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA operand not 16-byte aligned");
    CLIMB_REQUIRE((ld * 2) % 16 == 0, "TMA operand leading dimension %lld not a multiple of 8", ld);
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(box_inner), static_cast<cuuint32_t>(box_outer)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CLIMB_REQUIRE(r == CUDA_SUCCESS,
                  "cuTensorMapEncodeTiled failed (%d) inner=%lld outer=%lld ld=%lld box=%dx%d",
                  static_cast<int>(r), inner, outer, ld, box_inner, box_outer);
    return 0;
}

}  // namespace

// SMs the persistent kernels may fill: the device's count minus a reserve the data-parallel layer sets while NCCL's
// reduction kernels are resident (climb_set_sm_reserve): a persistent grid of 148 CTAs next to k foreign CTAs runs its
// last k CTAs as a second round -- up to twice the kernel time -- while a grid of 148 - k runs one round at (148 - k) / 148 speed.
static int g_sm_reserve = 0;
int sm_reserve(int n) {
    const int old = g_sm_reserve;
    if (n >= 0) g_sm_reserve = n;
    return old;
}
int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    int avail = n - g_sm_reserve;
    avail &= ~1;                       // CTA pairs
    return avail < 8 ? 8 : avail;
}

namespace {

template <int BLOCK_N, bool HAS_INPUT>
int launch_gemm_t(const climb_gemm_desc* d, GemmDeviceArgs& a, cudaStream_t stream) {
    using L = SmemLayout<BLOCK_N>;
    CUtensorMap ta, tb;
    int rc;
    if (!d->a_mn_major) rc = make_tmap_2d(&ta, d->A, d->K, d->M, d->lda, kBlockK, kBlockM);
    else                rc = make_tmap_2d(&ta, d->A, d->M, d->K, d->lda, 64, kBlockK);
    if (rc) return rc;
    if (!d->b_mn_major) rc = make_tmap_2d(&tb, d->B, d->K, d->N, d->ldb, kBlockK, BLOCK_N);
    else                rc = make_tmap_2d(&tb, d->B, d->N, d->K, d->ldb, 64, kBlockK);
    if (rc) return rc;

    a.m_tiles = (d->M + kBlockM - 1) / kBlockM;
    a.n_tiles = (d->N + BLOCK_N - 1) / BLOCK_N;
    a.k_blocks_total = (d->K + kBlockK - 1) / kBlockK;
    int split = d->split_k;
    if (split <= 0) {
        // auto split-K (accumulating fp32 outputs only): choose the split whose tile count fills
        // whole waves of the machine best, e.g. 108 tiles x 4 = 432 = 2.92 waves instead of 0.73
        split = 1;
        const int tiles = a.m_tiles * a.n_tiles;
        if (d->accumulate && d->c_dtype == CLIMB_F32 && a.k_blocks_total >= 16) {
            // an INDEPENDENT launch runs in the previous kernel's last partial wave: size it for the ~80 % of
            // the SMs that wave leaves idle, so that none of its CTAs has to queue behind the predecessor's tail
            const int sms = a.skip_pdl_wait ? (num_sms() * 4) / 5 : num_sms();
            double best = 0.0;
            for (int s = 1; s <= 16 && a.k_blocks_total / s >= 8; ++s) {
                const long long t = 1LL * tiles * s;
                if (a.skip_pdl_wait && t > sms && s > 1) break;          // one wave on the idle SMs only
                const long long waves = (t + sms - 1) / sms;
                const double fill = static_cast<double>(t) / (waves * sms);
                if (fill > best + 0.03) { best = fill; split = s; }
            }
        }
    }
    if (split > a.k_blocks_total) split = a.k_blocks_total;
    a.k_blocks_per_split = (a.k_blocks_total + split - 1) / split;
    split = (a.k_blocks_total + a.k_blocks_per_split - 1) / a.k_blocks_per_split;   // no empty splits
    a.split_k = split;
    CLIMB_REQUIRE(split == 1 || (d->accumulate && d->c_dtype == CLIMB_F32),
                  "split_k > 1 needs accumulate=1 into an fp32 C");

    static bool attr_set = false;
    if (!attr_set) {
        CLIMB_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BLOCK_N, HAS_INPUT>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    // epilogues that read a tensor (derivative activations, residual) get a second scratch half for
    // the cp.async prefetch of the next chunk's input and pay one ring stage for it
    a.scratch_bytes = HAS_INPUT ? 2 * kEpiHalfBytes : kEpiHalfBytes;
    a.num_stages = L::stages_for(a.scratch_bytes);
    const int smem_bytes = L::total_for(a.scratch_bytes);
    const int total = a.m_tiles * a.n_tiles * a.split_k;
    const int grid = total < num_sms() ? total : num_sms();
    ProfScope prof(PROF_GEMM, 2.0 * d->M * static_cast<double>(d->N) * d->K, stream);
    if (a.skip_pdl_wait) pdl_mark_independent();
    CLIMB_CUDA_OK(launch_pdl(gemm_bf16_tcgen05_kernel<BLOCK_N, HAS_INPUT>, dim3(grid), dim3(kNumThreads), smem_bytes, stream, ta, tb, a));
    CLIMB_LAUNCH_OK();
    return 0;
}


template <int KIND>
int launch_fast(const climb_gemm_desc* d, GemmDeviceArgs& a, cudaStream_t stream) {
    CUtensorMap ta, tb;
    int rc;
    if (!d->a_mn_major) rc = make_tmap_2d(&ta, d->A, d->K, d->M, d->lda, kBlockK, kBlockM);
    else                rc = make_tmap_2d(&ta, d->A, d->M, d->K, d->lda, 64, kBlockK);
    if (rc) return rc;
    if (!d->b_mn_major) rc = make_tmap_2d(&tb, d->B, d->K, d->N, d->ldb, kBlockK, kFastBlockN);
    else                rc = make_tmap_2d(&tb, d->B, d->N, d->K, d->ldb, 64, kBlockK);
    if (rc) return rc;
    a.m_tiles = (d->M + kBlockM - 1) / kBlockM;
    a.n_tiles = d->N / kFastBlockN;
    a.k_blocks_total = (d->K + kBlockK - 1) / kBlockK;
    a.k_blocks_per_split = a.k_blocks_total;
    a.split_k = 1;
    a.num_stages = kFastStages;
    a.scratch_bytes = kFastUnitBytes;
    static bool attr_set = false;
    if (!attr_set) {
        CLIMB_CUDA_OK(cudaFuncSetAttribute(gemm_fast_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFastSmemBytes));
        attr_set = true;
    }
    const int total = a.m_tiles * a.n_tiles;
    const int grid = total < num_sms() ? total : num_sms();
    // the last round's tiles at half width when they occupy at most half of the grid (see fast_tile)
    const int rest = total % grid;
    static const bool tail_split = [] { const char* e = std::getenv("CLIMB_GEMM_TAIL_SPLIT"); return e == nullptr || e[0] != '0'; }();
    a.full_tiles = (tail_split && total > grid && rest > 0 && 2 * rest <= grid) ? total - rest : total;
    a.total_vtiles = total + (total - a.full_tiles);
    ProfScope prof(PROF_GEMM, 2.0 * d->M * static_cast<double>(d->N) * d->K, stream);
    if (a.skip_pdl_wait) pdl_mark_independent();
    CLIMB_CUDA_OK(launch_pdl(gemm_fast_kernel<KIND>, dim3(grid), dim3(kFastThreads), kFastSmemBytes, stream, ta, tb, a));
    CLIMB_LAUNCH_OK();
    return 0;
}

template <int KIND>
int launch_pair(const climb_gemm_desc* d, GemmDeviceArgs& a, cudaStream_t stream) {
    CUtensorMap ta, tb;
    int rc = make_tmap_2d(&ta, d->A, d->K, d->M, d->lda, kBlockK, kBlockM);
    if (rc) return rc;
    if (!d->b_mn_major) rc = make_tmap_2d(&tb, d->B, d->K, d->N, d->ldb, kBlockK, kFastBlockN / 2);
    else                rc = make_tmap_2d(&tb, d->B, d->N, d->K, d->ldb, 64, kBlockK);
    if (rc) return rc;
    a.m_tiles = (d->M + kPairTileM - 1) / kPairTileM;
    a.n_tiles = d->N / kFastBlockN;
    a.k_blocks_total = (d->K + kBlockK - 1) / kBlockK;
    a.k_blocks_per_split = a.k_blocks_total;
    a.split_k = 1;
    a.num_stages = kPairStages;
    a.scratch_bytes = kFastUnitBytes;
    static bool attr_set = false;
    if (!attr_set) {
        CLIMB_CUDA_OK(cudaFuncSetAttribute(gemm_pair_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemBytes));
        attr_set = true;
    }
    const int total = a.m_tiles * a.n_tiles;
    const int max_pairs = num_sms() / 2;
    const int pairs = total < max_pairs ? total : max_pairs;
    const int rest = total % pairs;
    static const bool tail_split = [] { const char* e = std::getenv("CLIMB_GEMM_TAIL_SPLIT"); return e == nullptr || e[0] != '0'; }();
    a.full_tiles = (tail_split && total > pairs && rest > 0 && 2 * rest <= pairs) ? total - rest : total;
    a.total_vtiles = total + (total - a.full_tiles);
    ProfScope prof(PROF_GEMM, 2.0 * d->M * static_cast<double>(d->N) * d->K, stream);
    if (a.skip_pdl_wait) pdl_mark_independent();
    CLIMB_CUDA_OK(launch_pdl_cluster(gemm_pair_kernel<KIND>, dim3(2 * pairs), dim3(kFastThreads), kPairSmemBytes, stream, 2u, ta, tb, a));
    CLIMB_LAUNCH_OK();
    return 0;
}

inline bool aligned16(const void* p, long long pitch_bytes) {
    return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && pitch_bytes % 16 == 0;
}

// is this the weight-gradient form the pair wgrad kernel covers?
bool pair_wgrad_ok(const climb_gemm_desc* d) {
    // independent launches (the attention-output wgrad behind the attention backward): CLIMB_WGRAD_INDEP_PAIR=1 lets them take
    // the pair kernel too (A/B switch)
    static const bool indep_pair = [] { const char* e = getenv("CLIMB_WGRAD_INDEP_PAIR"); return e && e[0] == '1'; }();
    return d->accumulate && d->c_dtype == CLIMB_F32 && d->a_mn_major && d->b_mn_major && (!d->independent || indep_pair) &&
           d->M % kPairTileM == 0 && d->N % kFastBlockN == 0 && d->K >= 16 * kBlockK && d->split_k <= 0 &&
           (d->block_n == 0 || d->block_n == kFastBlockN) && d->bias == nullptr && d->residual == nullptr && d->aux == nullptr &&
           d->c2 == nullptr && d->colsum == nullptr && d->epilogue == CLIMB_EPI_NONE && (d->alpha == 0.0f || d->alpha == 1.0f) &&
           aligned16(d->C, d->ldc * 4) && aligned16(d->A, d->lda * 2) && aligned16(d->B, d->ldb * 2);
}

int launch_pair_wgrad(const climb_gemm_desc* d, GemmDeviceArgs& a, cudaStream_t stream) {
    CUtensorMap ta, tb;
    int rc = make_tmap_2d(&ta, d->A, d->M, d->K, d->lda, 64, kBlockK);
    if (rc) return rc;
    rc = make_tmap_2d(&tb, d->B, d->N, d->K, d->ldb, 64, kBlockK);
    if (rc) return rc;
    a.m_tiles = d->M / kPairTileM;
    a.n_tiles = d->N / kFastBlockN;
    a.k_blocks_total = (d->K + kBlockK - 1) / kBlockK;
    const int pairs_max = num_sms() / 2;
    const int tiles = a.m_tiles * a.n_tiles;
    // split over the tokens so that tiles x splits fills whole rounds of the pair grid
    int split = 1;
    double best = 0.0;
    for (int sp = 1; sp <= 16 && a.k_blocks_total / sp >= 8; ++sp) {
        const long long t = 1LL * tiles * sp;
        const long long rounds = (t + pairs_max - 1) / pairs_max;
        const double fill = static_cast<double>(t) / (rounds * pairs_max);
        if (fill > best + 0.03) { best = fill; split = sp; }
    }
    a.k_blocks_per_split = (a.k_blocks_total + split - 1) / split;
    a.split_k = (a.k_blocks_total + a.k_blocks_per_split - 1) / a.k_blocks_per_split;
    a.num_stages = kPairStages;
    a.scratch_bytes = kEpiHalfBytes;
    static bool attr_set = false;
    if (!attr_set) {
        CLIMB_CUDA_OK(cudaFuncSetAttribute(gemm_pair_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgradSmemBytes));
        attr_set = true;
    }
    const int total = tiles * a.split_k;
    const int pairs = total < pairs_max ? total : pairs_max;
    // the fused bias-gradient sums live in the second accumulator stage: only when no pair owns more than one tile
    // (every FC1 / FC2 / O weight gradient of ViLT-base); otherwise they are an ordinary streaming pass over dY
    static const bool fuse_cs = [] { const char* e = getenv("CLIMB_WGRAD_COLSUM"); return !(e && e[0] == '0'); }();
    if (d->colsum_a != nullptr && !(fuse_cs && total <= pairs_max)) {
        rc = colsum(d->A, CLIMB_BF16, d->lda, d->K, d->M, d->colsum_a, stream);
        if (rc) return rc;
        a.colsum_a = nullptr;
    }
    ProfScope prof(PROF_GEMM, 2.0 * d->M * static_cast<double>(d->N) * d->K, stream);
    CLIMB_CUDA_OK(launch_pdl_cluster(gemm_pair_wgrad_kernel, dim3(2 * pairs), dim3(kNumThreads), kWgradSmemBytes, stream, 2u, ta, tb, a));
    CLIMB_LAUNCH_OK();
    return 0;
}

// which specialised kernel (if any) covers this problem; -1 = the generic kernel
int fast_kind(const climb_gemm_desc* d) {
    if (d->block_n != 0 && d->block_n != kFastBlockN) return -1;
    if (d->N % kFastBlockN != 0 || d->accumulate || d->split_k > 1 || d->colsum != nullptr) return -1;
    if (d->c2 != nullptr && !(d->c_dtype == CLIMB_F32 && aligned16(d->c2, d->ldc2 * 2))) return -1;
    if (d->alpha != 0.0f && d->alpha != 1.0f) return -1;
    const long long tiles = 1LL * ((d->M + kBlockM - 1) / kBlockM) * (d->N / kFastBlockN);
    if (tiles < num_sms()) return -1;                     // small problems: the generic kernel's narrower tiles
    if (d->bias != nullptr && (reinterpret_cast<uintptr_t>(d->bias) & 15) != 0) return -1;
    if (d->c_dtype == CLIMB_BF16) {
        if (!aligned16(d->C, d->ldc * 2) || d->residual != nullptr) return -1;
        if (d->epilogue == CLIMB_EPI_NONE && d->aux == nullptr) return FK_BF16;
        if (d->aux == nullptr || !aligned16(d->aux, d->ldaux * 2)) return -1;
        if (d->epilogue == CLIMB_EPI_GELU_SAVE_GRAD) return FK_GELU_SAVE;
        if (d->epilogue == CLIMB_EPI_MUL_AUX && d->bias == nullptr) return FK_MUL_AUX;
        return -1;
    }
    // fp32 C = acc + bias + residual, optionally with bf16 copies of the pre-residual (aux) / final (c2) value
    if (d->c_dtype == CLIMB_F32 && d->epilogue == CLIMB_EPI_NONE && d->residual != nullptr &&
        (d->aux == nullptr || aligned16(d->aux, d->ldaux * 2)) && aligned16(d->C, d->ldc * 4) && aligned16(d->residual, d->ldr * 4))
        return FK_RES_F32;
    return -1;
}

template <int BLOCK_N>
int launch_gemm(const climb_gemm_desc* d, GemmDeviceArgs& a, cudaStream_t stream) {
    const bool epi_input = d->residual != nullptr || d->epilogue == CLIMB_EPI_DGELU ||
                           d->epilogue == CLIMB_EPI_DSWISH || d->epilogue == CLIMB_EPI_DRELU ||
                           d->epilogue == CLIMB_EPI_MUL_AUX;
    return epi_input ? launch_gemm_t<BLOCK_N, true>(d, a, stream) : launch_gemm_t<BLOCK_N, false>(d, a, stream);
}

}  // namespace

// dev switch (CLIMB_GEMM_GENERIC=1): route everything through the generic kernel, for A/B timing
static const bool g_disable_fast = [] { const char* e = getenv("CLIMB_GEMM_GENERIC"); return e && e[0] == '1'; }();
// dev switch (CLIMB_GEMM_PAIR=1): the CTA-pair (cta_group::2) variants of the specialised kernels
// CTA-pair (cta_group::2) kernels. Default: the weight-gradient kernel, and the pair variants of the two specialised kinds
// that measured faster in pairs (FK_BF16 +4..6 %, FK_MUL_AUX +4 %; the epilogue-bound FK_GELU_SAVE / FK_RES_F32 did not gain).
// A/B switches: CLIMB_GEMM_PAIR=0 none of them, =1 all four kinds; CLIMB_GEMM_PAIR_WGRAD=0 the generic split-K wgrad kernel.
static int g_pair_mode = [] { const char* e = getenv("CLIMB_GEMM_PAIR"); return e == nullptr ? 2 : (e[0] == '1' ? 1 : (e[0] == '0' ? 0 : 2)); }();
static bool g_pair_wgrad = [] { const char* e = getenv("CLIMB_GEMM_PAIR_WGRAD"); return g_pair_mode != 0 && !(e && e[0] == '0'); }();
// climb_gemm_pair_mode (C ABI): 0 = one-CTA kernels only, 1 = every pair kernel, 2 = the default selection; returns the old mode
int gemm_pair_mode(int mode) {
    const int old = g_pair_mode;
    if (mode >= 0 && mode <= 2) { g_pair_mode = mode; g_pair_wgrad = mode != 0; }
    return old;
}
inline bool pair_kind_enabled(int kind) {
    if (g_pair_mode == 1) return true;
    return g_pair_mode == 2 && (kind == FK_BF16 || kind == FK_MUL_AUX);
}

int gemm_bf16(const climb_gemm_desc* d, cudaStream_t stream) {
    CLIMB_REQUIRE(d != nullptr, "null gemm descriptor");
    CLIMB_REQUIRE(d->M > 0 && d->N > 0 && d->K > 0, "gemm: empty problem M=%d N=%d K=%d", d->M, d->N, d->K);
    CLIMB_REQUIRE(d->A && d->B && d->C, "gemm: null operand");
    CLIMB_REQUIRE(d->c_dtype == CLIMB_F32 || d->c_dtype == CLIMB_BF16, "gemm: bad c_dtype");
    CLIMB_REQUIRE(!(d->accumulate && d->c_dtype != CLIMB_F32), "gemm: accumulate needs fp32 C");
    const bool aux_needed = (d->epilogue == CLIMB_EPI_DGELU || d->epilogue == CLIMB_EPI_DSWISH ||
                             d->epilogue == CLIMB_EPI_DRELU || d->epilogue == CLIMB_EPI_MUL_AUX ||
                             d->epilogue == CLIMB_EPI_GELU_SAVE_GRAD);
    if (d->colsum != nullptr) {
        CLIMB_REQUIRE(d->c_dtype == CLIMB_BF16 && d->N % 32 == 0 && (d->ldc * 2) % 16 == 0 &&
                          (reinterpret_cast<uintptr_t>(d->C) & 15) == 0 && !d->accumulate,
                      "gemm: fused colsum needs a bf16 C with N %% 32 == 0 and 16-byte aligned rows");
        // every other epilogue tensor must be on the aligned fast path too
        CLIMB_REQUIRE((d->aux == nullptr || ((d->ldaux * 2) % 16 == 0 && (reinterpret_cast<uintptr_t>(d->aux) & 15) == 0)) &&
                          (d->c2 == nullptr) && (d->residual == nullptr) &&
                          (d->bias == nullptr || (reinterpret_cast<uintptr_t>(d->bias) & 15) == 0),
                      "gemm: fused colsum needs aligned aux / bias and no residual / c2");
    }
    CLIMB_REQUIRE(!aux_needed || d->aux != nullptr, "gemm: derivative epilogue needs aux");
    CLIMB_REQUIRE(!(d->accumulate && d->c2 != nullptr), "gemm: accumulate cannot produce a bf16 copy");
    CLIMB_REQUIRE(!(d->accumulate && d->epilogue != CLIMB_EPI_NONE),
                  "gemm: accumulate cannot be combined with a non-linear epilogue");

    GemmDeviceArgs a{};
    a.M = d->M; a.N = d->N; a.K = d->K;
    a.a_mn_major = d->a_mn_major ? 1 : 0;
    a.b_mn_major = d->b_mn_major ? 1 : 0;
    a.C = d->C; a.ldc = d->ldc; a.c_dtype = d->c_dtype;
    a.bias = d->bias; a.residual = d->residual; a.ldr = d->ldr;
    a.epilogue = d->epilogue; a.aux = d->aux; a.ldaux = d->ldaux;
    a.c2 = d->c2; a.ldc2 = d->ldc2;
    a.colsum = d->colsum;
    a.colsum_a = d->colsum_a;
    a.alpha = d->alpha == 0.0f ? 1.0f : d->alpha;
    a.accumulate = d->accumulate ? 1 : 0;
    static const bool indep_ok = [] { const char* e = getenv("CLIMB_NO_INDEPENDENT"); return !(e && e[0] == '1'); }();   // dev A/B switch
    a.skip_pdl_wait = (d->independent && pdl_enabled() && indep_ok) ? 1 : 0;

    if (d->colsum_a != nullptr)
        CLIMB_REQUIRE(d->a_mn_major && d->accumulate, "gemm: colsum_a belongs to the weight-gradient form (A MN-major, accumulate)");
    if (!g_disable_fast && g_pair_wgrad && pair_wgrad_ok(d)) return launch_pair_wgrad(d, a, stream);
    if (d->colsum_a != nullptr) {       // every other kernel: the bias gradient is a streaming pass over A = dY [K rows, M columns]
        const int rc_cs = colsum(d->A, CLIMB_BF16, d->lda, d->K, d->M, d->colsum_a, stream);
        if (rc_cs) return rc_cs;
        a.colsum_a = nullptr;
    }
    if (!g_disable_fast) {
        const int kind = fast_kind(d);
        // CTA-pair kernels (cta_group::2): K-major A, at least one 256-row tile per pair of SMs
        if (kind >= 0 && pair_kind_enabled(kind) && !d->a_mn_major &&
            1LL * ((d->M + kPairTileM - 1) / kPairTileM) * (d->N / kFastBlockN) >= num_sms() / 2) {
            switch (kind) {
                case FK_BF16: return launch_pair<FK_BF16>(d, a, stream);
                case FK_GELU_SAVE: return launch_pair<FK_GELU_SAVE>(d, a, stream);
                case FK_MUL_AUX: return launch_pair<FK_MUL_AUX>(d, a, stream);
                case FK_RES_F32: return launch_pair<FK_RES_F32>(d, a, stream);
                default: break;
            }
        }
        switch (kind) {
            case FK_BF16: return launch_fast<FK_BF16>(d, a, stream);
            case FK_GELU_SAVE: return launch_fast<FK_GELU_SAVE>(d, a, stream);
            case FK_MUL_AUX: return launch_fast<FK_MUL_AUX>(d, a, stream);
            case FK_RES_F32: return launch_fast<FK_RES_F32>(d, a, stream);
            default: break;
        }
    }
    int bn = d->block_n;
    if (bn == 0) {
        // Tile heuristic. 128x256 halves the shared-memory operand traffic per FLOP of 128x128 (B300
        // microarchitecture notes: one 128x256x16 UMMA reads 12 KB per 128 cycles), so it wins unless
        // wave quantisation / N padding costs it more than ~25 % (measured: N = 768 at
        // M = 15168 runs 20-30 % faster on 357 wide tiles than on 714 narrow ones).
        const int m_tiles = (d->M + kBlockM - 1) / kBlockM;
        const bool can_split = d->accumulate && d->c_dtype == CLIMB_F32;     // split-K fills the machine
        double eff[3] = {0.0, 0.0, 0.0};
        const int cands[3] = {256, 128, 64};
        double best = 0.0;
        for (int i = 0; i < 3; ++i) {
            const int c = cands[i];
            if (c > 64 && d->N <= c / 2) continue;
            const long long n_tiles = (d->N + c - 1) / c;
            const long long tiles = 1LL * m_tiles * n_tiles;
            const long long waves = (tiles + num_sms() - 1) / num_sms();
            const double useful = static_cast<double>(d->N) / static_cast<double>(n_tiles * c);
            const double fill = can_split && tiles < num_sms() ? 1.0 : static_cast<double>(tiles) / (waves * num_sms());
            eff[i] = fill * useful;
            if (eff[i] > best) best = eff[i];
        }
        for (int i = 0; i < 3; ++i)
            if (eff[i] > 0.0 && eff[i] >= 0.75 * best) { bn = cands[i]; break; }
    }
    switch (bn) {
        case 256: return launch_gemm<256>(d, a, stream);
        case 128: return launch_gemm<128>(d, a, stream);
        case 64: return launch_gemm<64>(d, a, stream);
        default: CLIMB_REQUIRE(false, "gemm: block_n must be 0, 64, 128 or 256 (got %d)", bn);
    }
    return 0;
}

// Fused adapter bottleneck (see adapter_fused_kernel). A [M, d] bf16: the site's input (forward) or the gradient at its output
// (backward); w_down [r, d], w_up [d, r] bf16 (the shadow arena); c_in / c_out fp32 [M, d] (may alias); c2 optional bf16 copy.
bool adapter_fused_ok(int d, int r) {
    static const bool on = [] { const char* e = getenv("CLIMB_ADAPTER_FUSED"); return !(e && e[0] == '0'); }();
    return on && r > 0 && r <= 64 && r % 16 == 0 && d % 128 == 0 && d >= 128;
}

int adapter_fused(int backward, int M, int d, int r, int act, const void* A, const void* w_down, const void* w_up, const float* b_down,
                  const float* b_up, void* pre, void* z, const float* c_in, float* c_out, void* c2, float* colsum_z,
                  cudaStream_t stream) {
    CLIMB_REQUIRE(adapter_fused_ok(d, r), "adapter_fused: unsupported shape d=%d r=%d", d, r);
    CLIMB_REQUIRE(A && w_down && w_up && pre && z && c_in && M > 0, "adapter_fused: null operand");
    CLIMB_REQUIRE(act == CLIMB_EPI_SWISH || act == CLIMB_EPI_RELU, "adapter_fused: activation must be swish or relu");
    CUtensorMap ma, mb1, mb2;
    int rc = make_tmap_2d(&ma, A, d, M, d, kBlockK, kBlockM);
    if (rc) return rc;
    if (!backward) {
        rc = make_tmap_2d(&mb1, w_down, d, r, d, 64, 64);          // W_d [r, d]: K-major B1, 64 (k) x 64 rows (n, zero fill past r)
        if (rc) return rc;
        rc = make_tmap_2d(&mb2, w_up, r, d, r, 64, 128);           // W_u [d, r]: K-major B2, 64 (k, zero fill past r) x 128 rows (n)
    } else {
        rc = make_tmap_2d(&mb1, w_up, r, d, r, 64, 64);            // W_u [d, r]: MN-major B1, 64 (n, zero fill past r) x 64 k-rows
        if (rc) return rc;
        rc = make_tmap_2d(&mb2, w_down, d, r, d, 64, 64);          // W_d [r, d]: MN-major B2, 64 (n) x 64 k-rows (zero fill past r)
    }
    if (rc) return rc;
    AdapterArgs a{};
    a.M = M; a.d = d; a.r = r; a.backward = backward ? 1 : 0; a.act = act;
    a.bias1 = backward ? nullptr : b_down;
    a.bias2 = backward ? nullptr : b_up;
    a.pre = static_cast<__nv_bfloat16*>(pre); a.z = static_cast<__nv_bfloat16*>(z);
    a.c_in = c_in; a.c_out = c_out; a.c2 = static_cast<__nv_bfloat16*>(c2);
    a.colsum_z = backward ? colsum_z : nullptr;
    a.n_tiles = (M + kBlockM - 1) / kBlockM;
    CLIMB_REQUIRE(backward || b_down != nullptr, "adapter_fused: forward needs the down-projection bias");
    static bool attr_set = false;
    if (!attr_set) {
        CLIMB_CUDA_OK(cudaFuncSetAttribute(adapter_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAdSmemBytes));
        attr_set = true;
    }
    const int grid = a.n_tiles < num_sms() ? a.n_tiles : num_sms();
    CLIMB_CUDA_OK(launch_pdl(adapter_fused_kernel, dim3(grid), dim3(kAdThreads), kAdSmemBytes, stream, ma, mb1, mb2, a));
    CLIMB_LAUNCH_OK();
    return 0;
}

}  // namespace climb
