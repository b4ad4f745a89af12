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

#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched through the runtime)

namespace climb {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;          // 64 bf16 = 128 B = one swizzle row
constexpr int kNumEpiWarps = 8;
constexpr int kNumThreads = 64 + kNumEpiWarps * 32;
constexpr int kAccStages = 2;

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
    float alpha;
    int accumulate;
    int split_k;
    int m_tiles, n_tiles, k_blocks_per_split, k_blocks_total;
};

template <int BLOCK_N>
struct SmemLayout {
    static constexpr int kABytes = kBlockM * kBlockK * 2;
    static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = (BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 6 : 8);
    static constexpr int kBarrierBytes = 1024;
    static constexpr int kTotal = kStages * kStageBytes + kBarrierBytes + 1024 /*align slack*/;
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

__device__ __forceinline__ float apply_act(int epi, float v, float aux) {
    switch (epi) {
        case CLIMB_EPI_GELU: return gelu_f(v);
        case CLIMB_EPI_DGELU: return v * dgelu_f(aux);
        case CLIMB_EPI_SWISH: return swish_f(v);
        case CLIMB_EPI_DSWISH: return v * dswish_f(aux);
        case CLIMB_EPI_RELU: return fmaxf(v, 0.0f);
        case CLIMB_EPI_DRELU: return aux > 0.0f ? v : 0.0f;
        case CLIMB_EPI_TANH: return tanhf(v);
        default: return v;
    }
}

template <int BLOCK_N>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a,
                         const __grid_constant__ CUtensorMap tmap_b, const GemmDeviceArgs p) {
    using L = SmemLayout<BLOCK_N>;
    constexpr int kStages = L::kStages;
    constexpr uint32_t kTmemCols = kAccStages * BLOCK_N;   // 512 / 256 / 128: powers of two >= 32

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
    uint8_t* bar_base = smem + kStages * L::kStageBytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* acc_full = empty_bar + kStages;
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

    const int tiles_mn = p.m_tiles * p.n_tiles;
    const int total_tiles = tiles_mn * p.split_k;

    if (warp == 0) {
        // ================================ TMA producer =====================================
        if (lane == 0) {
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
                            tma_load_2d(&tmap_a, &full_bar[stage], sa + j * (kBlockK * 128),
                                        m_blk * kBlockM + j * 64, kb * kBlockK);
                    }
                    if (!p.b_mn_major) {
                        tma_load_2d(&tmap_b, &full_bar[stage], sb, kb * kBlockK, n_blk * BLOCK_N);
                    } else {
#pragma unroll
                        for (int j = 0; j < BLOCK_N / 64; ++j)
                            tma_load_2d(&tmap_b, &full_bar[stage], sb + j * (kBlockK * 128),
                                        n_blk * BLOCK_N + j * 64, kb * kBlockK);
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer =======================================
        if (lane == 0) {
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
    } else {
        // ================================ epilogue =========================================
        const int ew = warp - 2;                  // 0..7
        const int lane_grp = warp & 3;            // TMEM lanes [32*lane_grp, +32) are ours
        const int col_half = ew >> 2;             // two warps share a lane group: even/odd chunks
        int acc = 0;
        uint32_t acc_phase = 0;
        const bool vec_c = ((p.ldc * (p.c_dtype == CLIMB_F32 ? 4 : 2)) % 16 == 0) &&
                           ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
        const bool vec_aux = p.aux != nullptr && ((p.ldaux * 2) % 16 == 0) &&
                             ((reinterpret_cast<uintptr_t>(p.aux) & 15) == 0);
        const bool vec_c2 = p.c2 != nullptr && ((p.ldc2 * 2) % 16 == 0) &&
                            ((reinterpret_cast<uintptr_t>(p.c2) & 15) == 0);
        const bool vec_res = p.residual != nullptr && ((p.ldr * 4) % 16 == 0) &&
                             ((reinterpret_cast<uintptr_t>(p.residual) & 15) == 0);
        const bool aux_in = (p.epilogue == CLIMB_EPI_DGELU || p.epilogue == CLIMB_EPI_DSWISH ||
                             p.epilogue == CLIMB_EPI_DRELU);
        const bool aux_out = (p.aux != nullptr) && !aux_in;   // pre-activation copy (bf16)
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int split = tile / tiles_mn;
            const int mn = tile - split * tiles_mn;
            const int m_blk = mn / p.n_tiles;
            const int n_blk = mn - m_blk * p.n_tiles;
            const int row = m_blk * kBlockM + lane_grp * 32 + lane;
            const bool row_ok = row < p.M;
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
                uint32_t r[32];
                tmem_ld_32x32(t_row + static_cast<uint32_t>(c * 32), r);
                tmem_ld_wait();
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
                const bool full = (n0 + 32 <= p.N);
                if (add_bias) {
                    if (full) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
                            v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (n0 + j < p.N) v[j] += __ldg(p.bias + n0 + j);
                    }
                }
                if (row_ok) {
                    // ---- activation (with optional bf16 aux tensor in or out) ----
                    if (p.epilogue != CLIMB_EPI_NONE || aux_out) {
                        __nv_bfloat16* auxp = reinterpret_cast<__nv_bfloat16*>(p.aux) +
                                              static_cast<long long>(row) * p.ldaux + n0;
                        if (aux_in) {
                            if (full && vec_aux) {
#pragma unroll
                                for (int j = 0; j < 32; j += 8) {
                                    const uint4 u = *reinterpret_cast<const uint4*>(auxp + j);
                                    const float2 a0 = unpack_bf16(u.x), a1 = unpack_bf16(u.y),
                                                 a2 = unpack_bf16(u.z), a3 = unpack_bf16(u.w);
                                    v[j] = apply_act(p.epilogue, v[j], a0.x);
                                    v[j + 1] = apply_act(p.epilogue, v[j + 1], a0.y);
                                    v[j + 2] = apply_act(p.epilogue, v[j + 2], a1.x);
                                    v[j + 3] = apply_act(p.epilogue, v[j + 3], a1.y);
                                    v[j + 4] = apply_act(p.epilogue, v[j + 4], a2.x);
                                    v[j + 5] = apply_act(p.epilogue, v[j + 5], a2.y);
                                    v[j + 6] = apply_act(p.epilogue, v[j + 6], a3.x);
                                    v[j + 7] = apply_act(p.epilogue, v[j + 7], a3.y);
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j)
                                    if (n0 + j < p.N)
                                        v[j] = apply_act(p.epilogue, v[j], __bfloat162float(auxp[j]));
                            }
                        } else {
                            if (aux_out) {
                                if (full && vec_aux) {
#pragma unroll
                                    for (int j = 0; j < 32; j += 8) {
                                        uint4 u;
                                        u.x = pack_bf16(v[j], v[j + 1]);
                                        u.y = pack_bf16(v[j + 2], v[j + 3]);
                                        u.z = pack_bf16(v[j + 4], v[j + 5]);
                                        u.w = pack_bf16(v[j + 6], v[j + 7]);
                                        *reinterpret_cast<uint4*>(auxp + j) = u;
                                    }
                                } else {
#pragma unroll
                                    for (int j = 0; j < 32; ++j)
                                        if (n0 + j < p.N) auxp[j] = __float2bfloat16_rn(v[j]);
                                }
                            }
                            if (p.epilogue != CLIMB_EPI_NONE) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) v[j] = apply_act(p.epilogue, v[j], 0.0f);
                            }
                        }
                    }
                    // ---- residual add (fp32) ----
                    if (add_res) {
                        const float* rp = p.residual + static_cast<long long>(row) * p.ldr + n0;
                        if (full && vec_res) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 r4 = *reinterpret_cast<const float4*>(rp + j);
                                v[j] += r4.x; v[j + 1] += r4.y; v[j + 2] += r4.z; v[j + 3] += r4.w;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (n0 + j < p.N) v[j] += rp[j];
                        }
                    }
                    // ---- optional bf16 copy of the final value (feeds the next GEMM) ----
                    if (p.c2 != nullptr) {
                        __nv_bfloat16* c2p = reinterpret_cast<__nv_bfloat16*>(p.c2) +
                                             static_cast<long long>(row) * p.ldc2 + n0;
                        if (full && vec_c2) {
#pragma unroll
                            for (int j = 0; j < 32; j += 8) {
                                uint4 u;
                                u.x = pack_bf16(v[j], v[j + 1]);
                                u.y = pack_bf16(v[j + 2], v[j + 3]);
                                u.z = pack_bf16(v[j + 4], v[j + 5]);
                                u.w = pack_bf16(v[j + 6], v[j + 7]);
                                *reinterpret_cast<uint4*>(c2p + j) = u;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (n0 + j < p.N) c2p[j] = __float2bfloat16_rn(v[j]);
                        }
                    }
                    // ---- store ----
                    if (p.c_dtype == CLIMB_F32) {
                        float* cp = reinterpret_cast<float*>(p.C) +
                                    static_cast<long long>(row) * p.ldc + n0;
                        if (p.accumulate) {
                            if (full && vec_c) {
#pragma unroll
                                for (int j = 0; j < 32; j += 4)
                                    atomicAdd(reinterpret_cast<float4*>(cp + j),
                                              make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j)
                                    if (n0 + j < p.N) atomicAdd(cp + j, v[j]);
                            }
                        } else if (full && vec_c) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                *reinterpret_cast<float4*>(cp + j) =
                                    make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (n0 + j < p.N) cp[j] = v[j];
                        }
                    } else {
                        __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(p.C) +
                                            static_cast<long long>(row) * p.ldc + n0;
                        if (full && vec_c) {
#pragma unroll
                            for (int j = 0; j < 32; j += 8) {
                                uint4 u;
                                u.x = pack_bf16(v[j], v[j + 1]);
                                u.y = pack_bf16(v[j + 2], v[j + 3]);
                                u.z = pack_bf16(v[j + 4], v[j + 5]);
                                u.w = pack_bf16(v[j + 6], v[j + 7]);
                                *reinterpret_cast<uint4*>(cp + j) = u;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (n0 + j < p.N) cp[j] = __float2bfloat16_rn(v[j]);
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

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <int BLOCK_N>
int launch_gemm(const climb_gemm_desc* d, GemmDeviceArgs& a, cudaStream_t stream) {
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
        // auto: only worth it when the output grid cannot fill the machine and K is deep
        split = 1;
        const int tiles = a.m_tiles * a.n_tiles;
        if (d->accumulate && d->c_dtype == CLIMB_F32 && tiles < num_sms() && a.k_blocks_total >= 8) {
            split = num_sms() / tiles;
            if (split > a.k_blocks_total / 4) split = a.k_blocks_total / 4;
            if (split < 1) split = 1;
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
        CLIMB_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BLOCK_N>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
        attr_set = true;
    }
    const int total = a.m_tiles * a.n_tiles * a.split_k;
    const int grid = total < num_sms() ? total : num_sms();
    ProfScope prof(PROF_GEMM, 2.0 * d->M * static_cast<double>(d->N) * d->K, stream);
    gemm_bf16_tcgen05_kernel<BLOCK_N><<<grid, kNumThreads, L::kTotal, stream>>>(ta, tb, a);
    CLIMB_LAUNCH_OK();
    return 0;
}

}  // namespace

int gemm_bf16(const climb_gemm_desc* d, cudaStream_t stream) {
    CLIMB_REQUIRE(d != nullptr, "null gemm descriptor");
    CLIMB_REQUIRE(d->M > 0 && d->N > 0 && d->K > 0, "gemm: empty problem M=%d N=%d K=%d", d->M, d->N, d->K);
    CLIMB_REQUIRE(d->A && d->B && d->C, "gemm: null operand");
    CLIMB_REQUIRE(d->c_dtype == CLIMB_F32 || d->c_dtype == CLIMB_BF16, "gemm: bad c_dtype");
    CLIMB_REQUIRE(!(d->accumulate && d->c_dtype != CLIMB_F32), "gemm: accumulate needs fp32 C");
    const bool aux_needed = (d->epilogue == CLIMB_EPI_DGELU || d->epilogue == CLIMB_EPI_DSWISH ||
                             d->epilogue == CLIMB_EPI_DRELU);
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
    a.alpha = d->alpha == 0.0f ? 1.0f : d->alpha;
    a.accumulate = d->accumulate ? 1 : 0;

    int bn = d->block_n;
    if (bn == 0) {
        // tile heuristic: fewest idle SM-slots in the last wave, ties to the wider tile
        const int m_tiles = (d->M + kBlockM - 1) / kBlockM;
        double best = -1.0;
        const int cands[3] = {256, 128, 64};
        for (int c : cands) {
            if (c > 64 && d->N <= c / 2) continue;
            const long long tiles = 1LL * m_tiles * ((d->N + c - 1) / c);
            const long long waves = (tiles + num_sms() - 1) / num_sms();
            const double useful = static_cast<double>(d->N) / (((d->N + c - 1) / c) * c);
            double eff = static_cast<double>(tiles) / (waves * num_sms()) * useful;
            if (c == 128) eff *= 0.97;      // narrower tiles re-read A more often
            if (c == 64) eff *= 0.90;
            if (eff > best) { best = eff; bn = c; }
        }
    }
    switch (bn) {
        case 256: return launch_gemm<256>(d, a, stream);
        case 128: return launch_gemm<128>(d, a, stream);
        case 64: return launch_gemm<64>(d, a, stream);
        default: CLIMB_REQUIRE(false, "gemm: block_n must be 0, 64, 128 or 256 (got %d)", bn);
    }
    return 0;
}

}  // namespace climb
