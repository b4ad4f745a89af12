// Fused multi-head self-attention for the ViLT hot path (ViltSelfAttention.forward,
// modeling_vilt.py:355-388): ctx = softmax(Q K^T / sqrt(dh) + key_bias) V with dh = 64 and
// sequences short enough (L = 237 for 40 text + 197 image tokens) that one head's whole K and V
// live in a CTA's shared memory. FlashAttention-style: the L x L score matrix never reaches HBM;
// row maxima / sums are reduced with warp shuffles inside the 4-lane quads that own a row.
//
// forward : grid (ceil(L/64), H, B), 4 warps x 16 query rows, online softmax over 64-key blocks.
// backward: two deterministic passes, no atomics --
//   dQ pass   : same tiling as forward; recomputes P from the saved log-sum-exp, dS = P*(dP - D).
//   dK/dV pass: grid over 64-key tiles; computes S^T = K Q^T so that P^T / dS^T come out of the
//               tensor cores already in A-fragment layout for dV += P^T dO and dK += dS^T Q.
// Tensor path: mma.sync m16n8k16 bf16 (fp32 accumulate) fed by ldmatrix from XOR-swizzled smem.
#include "common.cuh"
#include "internal.h"

#include <cstdlib>

namespace climb {
namespace {

constexpr int kDh = 64;
constexpr int kRowBytes = kDh * 2;          // 128 B per (token, head) row
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ uint32_t tile_addr(uint32_t base, int row, int chunk) {
    return base + row * kRowBytes + ((chunk ^ (row & 7)) << 4);
}

// rows_total rows of 64 bf16 from global (row stride ld elements) into swizzled smem; rows >=
// rows_valid are zero-filled.
__device__ __forceinline__ void load_rows_async(uint32_t sbase, const __nv_bfloat16* g, long long ld,
                                                int rows_valid, int rows_total) {
    for (int idx = threadIdx.x; idx < rows_total * 8; idx += blockDim.x) {
        const int r = idx >> 3, c = idx & 7;
        const bool ok = r < rows_valid;
        const __nv_bfloat16* src = g + (ok ? static_cast<long long>(r) * ld : 0) + c * 8;
        cp_async_16(tile_addr(sbase, r, c), src, ok);
    }
}

// A fragments (16 rows x 64) of a row-major tile: 4 k-steps of 16
__device__ __forceinline__ void load_a_frags(uint32_t (&f)[4][4], uint32_t sbase, int row0) {
    const int l = lane_id();
    const int row = row0 + (l & 7) + ((l >> 3) & 1) * 8;
#pragma unroll
    for (int kc = 0; kc < 4; ++kc) ldmatrix_x4(f[kc], tile_addr(sbase, row, kc * 2 + (l >> 4)));
}

// acc[8][4] (16 x 64 block, columns = 64 rows of the [row][d] tile starting at n0) += A * T^T
//   i.e. acc[i][n] += sum_d A[i][d] * T[n0 + n][d]           (used for S = Q K^T, dP = dO V^T, ...)
__device__ __forceinline__ void mma_a_tT(float (&acc)[8][4], const uint32_t (&a)[4][4],
                                         uint32_t sbase, int n0) {
    const int l = lane_id();
    const int rsel = (l & 7) + ((l >> 4) & 1) * 8;
    const int csel = (l >> 3) & 1;
#pragma unroll
    for (int kc = 0; kc < 4; ++kc) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {
            uint32_t b[4];
            ldmatrix_x4(b, tile_addr(sbase, n0 + np * 16 + rsel, kc * 2 + csel));
            const uint32_t b0[2] = {b[0], b[1]}, b1[2] = {b[2], b[3]};
            mma_bf16_16816(acc[2 * np], a[kc], b0);
            mma_bf16_16816(acc[2 * np + 1], a[kc], b1);
        }
    }
}

// acc[8][4] (16 x 64 over d) += P * T      with P given as 4 A-fragments over 64 k rows of T
//   i.e. acc[i][d] += sum_k P[i][k] * T[k0 + k][d]            (used for O = P V, dQ = dS K, ...)
__device__ __forceinline__ void mma_p_t(float (&acc)[8][4], const uint32_t (&pa)[4][4],
                                        uint32_t sbase, int k0) {
    const int l = lane_id();
    const int rsel = (l & 7) + ((l >> 3) & 1) * 8;
    const int csel = l >> 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
            uint32_t b[4];
            ldmatrix_x4_trans(b, tile_addr(sbase, k0 + j * 16 + rsel, dp * 2 + csel));
            const uint32_t b0[2] = {b[0], b[1]}, b1[2] = {b[2], b[3]};
            mma_bf16_16816(acc[2 * dp], pa[j], b0);
            mma_bf16_16816(acc[2 * dp + 1], pa[j], b1);
        }
    }
}

__device__ __forceinline__ void acc_to_a_frags(uint32_t (&pa)[4][4], const float (&s)[8][4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        pa[j][0] = pack_bf16(s[2 * j][0], s[2 * j][1]);
        pa[j][1] = pack_bf16(s[2 * j][2], s[2 * j][3]);
        pa[j][2] = pack_bf16(s[2 * j + 1][0], s[2 * j + 1][1]);
        pa[j][3] = pack_bf16(s[2 * j + 1][2], s[2 * j + 1][3]);
    }
}

__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// 16 x 64 fp32 accumulator block of this warp -> bf16 rows in global memory, staged through the
// warp's own 16 rows of a swizzled smem tile so that the stores are 16-byte and coalesced.
__device__ __forceinline__ void store_block_bf16(const float (&o)[8][4], uint32_t sbase, uint8_t* sgen,
                                                 int warp_row0, __nv_bfloat16* gdst, long long ld,
                                                 int rows_valid /* rows of this warp that exist */) {
    const int l = lane_id(), g = l >> 2, t = l & 3;
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        const int r0 = warp_row0 + g, r1 = r0 + 8;
        *reinterpret_cast<uint32_t*>(sgen + (tile_addr(sbase, r0, nt) - sbase) + 4 * t) =
            pack_bf16(o[nt][0], o[nt][1]);
        *reinterpret_cast<uint32_t*>(sgen + (tile_addr(sbase, r1, nt) - sbase) + 4 * t) =
            pack_bf16(o[nt][2], o[nt][3]);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = l + 32 * i;
        const int r = idx >> 3, c = idx & 7;
        if (r < rows_valid) {
            const uint4 v =
                *reinterpret_cast<const uint4*>(sgen + (tile_addr(sbase, warp_row0 + r, c) - sbase));
            *reinterpret_cast<uint4*>(gdst + static_cast<long long>(r) * ld + c * 8) = v;
        }
    }
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ key_bias,
                __nv_bfloat16* __restrict__ ctx, float* __restrict__ lse, int L, int H, int Lpad,
                float scale_log2) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t* sQg = sm;
    uint8_t* sKg = sQg + 64 * kRowBytes;
    uint8_t* sVg = sKg + Lpad * kRowBytes;
    float* sBias = reinterpret_cast<float*>(sVg + Lpad * kRowBytes);
    const uint32_t sQ = smem_u32(sQg), sK = smem_u32(sKg), sV = smem_u32(sVg);

    const int q0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z;
    const long long ld = 3LL * H * kDh;
    const __nv_bfloat16* base = qkv + static_cast<long long>(b) * L * ld + h * kDh;
    load_rows_async(sQ, base + static_cast<long long>(q0) * ld, ld, min(64, L - q0), 64);
    load_rows_async(sK, base + H * kDh, ld, L, Lpad);
    load_rows_async(sV, base + 2 * H * kDh, ld, L, Lpad);
    cp_async_commit();
    for (int j = threadIdx.x; j < Lpad; j += blockDim.x)
        sBias[j] = j < L ? (key_bias ? key_bias[static_cast<long long>(b) * L + j] * kLog2e : 0.0f)
                         : -INFINITY;
    cp_async_wait<0>();
    __syncthreads();

    const int warp = threadIdx.x >> 5, l = lane_id(), g = l >> 2, t = l & 3;
    if (q0 + warp * 16 >= L) return;          // whole warp past the end (no later block syncs)

    uint32_t qf[4][4];
    load_a_frags(qf, sQ, warp * 16);

    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.0f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;

    const int nkb = Lpad / 64;
    for (int kb = 0; kb < nkb; ++kb) {
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.0f;
        mma_a_tT(s, qf, sK, kb * 64);
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int col = kb * 64 + nt * 8 + 2 * t;
            const float b0 = sBias[col], b1 = sBias[col + 1];
            s[nt][0] = fmaf(s[nt][0], scale_log2, b0);
            s[nt][1] = fmaf(s[nt][1], scale_log2, b1);
            s[nt][2] = fmaf(s[nt][2], scale_log2, b0);
            s[nt][3] = fmaf(s[nt][3], scale_log2, b1);
            mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
            mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
        }
        mx0 = quad_max(mx0);
        mx1 = quad_max(mx1);
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
        const float c0 = exp2f(m0 - mn0), c1 = exp2f(m1 - mn1);
        m0 = mn0; m1 = mn1;
        l0 *= c0; l1 *= c1;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            o[nt][0] *= c0; o[nt][1] *= c0; o[nt][2] *= c1; o[nt][3] *= c1;
            s[nt][0] = exp2f(s[nt][0] - mn0);
            s[nt][1] = exp2f(s[nt][1] - mn0);
            s[nt][2] = exp2f(s[nt][2] - mn1);
            s[nt][3] = exp2f(s[nt][3] - mn1);
            l0 += s[nt][0] + s[nt][1];
            l1 += s[nt][2] + s[nt][3];
        }
        uint32_t pa[4][4];
        acc_to_a_frags(pa, s);
        mma_p_t(o, pa, sV, kb * 64);
    }
    l0 = quad_sum(l0);
    l1 = quad_sum(l1);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { o[nt][0] *= i0; o[nt][1] *= i0; o[nt][2] *= i1; o[nt][3] *= i1; }

    const int row_a = q0 + warp * 16 + g, row_b = row_a + 8;
    if (t == 0) {
        float* lp = lse + (static_cast<long long>(b) * H + h) * L;
        if (row_a < L) lp[row_a] = (m0 + log2f(l0)) * kLn2;
        if (row_b < L) lp[row_b] = (m1 + log2f(l1)) * kLn2;
    }
    __nv_bfloat16* dst = ctx + (static_cast<long long>(b) * L + q0 + warp * 16) * (H * kDh) + h * kDh;
    store_block_bf16(o, sQ, sQg, warp * 16, dst, static_cast<long long>(H) * kDh,
                     min(16, L - q0 - warp * 16));
}

// ------------------------------------------------------------------------------------------
// backward prep: delta[b,h,i] = sum_d dO[i,d] * O[i,d]
// ------------------------------------------------------------------------------------------
__global__ void attn_bwd_delta_kernel(const __nv_bfloat16* __restrict__ ctx,
                                      const __nv_bfloat16* __restrict__ dctx,
                                      float* __restrict__ delta, int B, int L, int H) {
    // one 8-lane group per (b, i, h): 64 elements = 8 lanes x 16 bytes
    const long long gid = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 3;
    const int sub = threadIdx.x & 7;
    const long long total = static_cast<long long>(B) * L * H;
    float acc = 0.0f;
    const bool ok = gid < total;
    long long bi = 0;
    int h = 0;
    if (ok) {
        bi = gid / H;
        h = static_cast<int>(gid - bi * H);
        const long long off = bi * (static_cast<long long>(H) * kDh) + h * kDh + sub * 8;
        const uint4 a = *reinterpret_cast<const uint4*>(ctx + off);
        const uint4 d = *reinterpret_cast<const uint4*>(dctx + off);
        const uint32_t av[4] = {a.x, a.y, a.z, a.w}, dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 x = unpack_bf16(av[i]), y = unpack_bf16(dv[i]);
            acc = fmaf(x.x, y.x, acc);
            acc = fmaf(x.y, y.y, acc);
        }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (ok && sub == 0) {
        const long long b = bi / L;
        const int i = static_cast<int>(bi - b * L);
        delta[(b * H + h) * L + i] = acc;
    }
}

// ------------------------------------------------------------------------------------------
// backward, dQ pass
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
attn_bwd_dq_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ key_bias,
                   const __nv_bfloat16* __restrict__ dctx, const float* __restrict__ lse,
                   const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv, int L, int H,
                   int Lpad, float scale_log2, float scale) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t* sQg = sm;
    uint8_t* sDOg = sQg + 64 * kRowBytes;
    uint8_t* sKg = sDOg + 64 * kRowBytes;
    uint8_t* sVg = sKg + Lpad * kRowBytes;
    float* sBias = reinterpret_cast<float*>(sVg + Lpad * kRowBytes);
    const uint32_t sQ = smem_u32(sQg), sDO = smem_u32(sDOg), sK = smem_u32(sKg), sV = smem_u32(sVg);

    const int q0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z;
    const long long ld = 3LL * H * kDh, ldo = static_cast<long long>(H) * kDh;
    const __nv_bfloat16* base = qkv + static_cast<long long>(b) * L * ld + h * kDh;
    const int qvalid = min(64, L - q0);
    load_rows_async(sQ, base + static_cast<long long>(q0) * ld, ld, qvalid, 64);
    load_rows_async(sDO, dctx + (static_cast<long long>(b) * L + q0) * ldo + h * kDh, ldo, qvalid, 64);
    load_rows_async(sK, base + H * kDh, ld, L, Lpad);
    load_rows_async(sV, base + 2 * H * kDh, ld, L, Lpad);
    cp_async_commit();
    for (int j = threadIdx.x; j < Lpad; j += blockDim.x)
        sBias[j] = j < L ? (key_bias ? key_bias[static_cast<long long>(b) * L + j] * kLog2e : 0.0f)
                         : -INFINITY;
    cp_async_wait<0>();
    __syncthreads();

    const int warp = threadIdx.x >> 5, l = lane_id(), g = l >> 2, t = l & 3;
    if (q0 + warp * 16 >= L) return;

    uint32_t qf[4][4], dof[4][4];
    load_a_frags(qf, sQ, warp * 16);
    load_a_frags(dof, sDO, warp * 16);

    const int row_a = q0 + warp * 16 + g, row_b = row_a + 8;
    const long long stat = (static_cast<long long>(b) * H + h) * L;
    // rows past L: lse = +inf makes P = 0 there
    const float lse_a = row_a < L ? lse[stat + row_a] * kLog2e : INFINITY;
    const float lse_b = row_b < L ? lse[stat + row_b] * kLog2e : INFINITY;
    const float dl_a = row_a < L ? delta[stat + row_a] : 0.0f;
    const float dl_b = row_b < L ? delta[stat + row_b] : 0.0f;

    float dq[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.0f;

    const int nkb = Lpad / 64;
    for (int kb = 0; kb < nkb; ++kb) {
        float s[8][4], dp[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.0f;
            dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.0f;
        }
        mma_a_tT(s, qf, sK, kb * 64);
        mma_a_tT(dp, dof, sV, kb * 64);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int col = kb * 64 + nt * 8 + 2 * t;
            const float b0 = sBias[col], b1 = sBias[col + 1];
            const float p0 = exp2f(fmaf(s[nt][0], scale_log2, b0) - lse_a);
            const float p1 = exp2f(fmaf(s[nt][1], scale_log2, b1) - lse_a);
            const float p2 = exp2f(fmaf(s[nt][2], scale_log2, b0) - lse_b);
            const float p3 = exp2f(fmaf(s[nt][3], scale_log2, b1) - lse_b);
            s[nt][0] = p0 * (dp[nt][0] - dl_a);
            s[nt][1] = p1 * (dp[nt][1] - dl_a);
            s[nt][2] = p2 * (dp[nt][2] - dl_b);
            s[nt][3] = p3 * (dp[nt][3] - dl_b);
        }
        uint32_t dsa[4][4];
        acc_to_a_frags(dsa, s);
        mma_p_t(dq, dsa, sK, kb * 64);
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { dq[nt][0] *= scale; dq[nt][1] *= scale; dq[nt][2] *= scale; dq[nt][3] *= scale; }
    __nv_bfloat16* dst = dqkv + (static_cast<long long>(b) * L + q0 + warp * 16) * ld + h * kDh;
    store_block_bf16(dq, sQ, sQg, warp * 16, dst, ld, min(16, L - q0 - warp * 16));
}

// ------------------------------------------------------------------------------------------
// backward, dK/dV pass
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
attn_bwd_dkv_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ key_bias,
                    const __nv_bfloat16* __restrict__ dctx, const float* __restrict__ lse,
                    const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv, int L, int H,
                    int Lpad, float scale_log2, float scale) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t* sKg = sm;
    uint8_t* sVg = sKg + 64 * kRowBytes;
    uint8_t* sQg = sVg + 64 * kRowBytes;
    uint8_t* sDOg = sQg + Lpad * kRowBytes;
    float* sLse = reinterpret_cast<float*>(sDOg + Lpad * kRowBytes);
    float* sDelta = sLse + Lpad;
    const uint32_t sK = smem_u32(sKg), sV = smem_u32(sVg), sQ = smem_u32(sQg), sDO = smem_u32(sDOg);

    const int k0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z;
    const long long ld = 3LL * H * kDh, ldo = static_cast<long long>(H) * kDh;
    const __nv_bfloat16* base = qkv + static_cast<long long>(b) * L * ld + h * kDh;
    const int kvalid = min(64, L - k0);
    load_rows_async(sK, base + H * kDh + static_cast<long long>(k0) * ld, ld, kvalid, 64);
    load_rows_async(sV, base + 2 * H * kDh + static_cast<long long>(k0) * ld, ld, kvalid, 64);
    load_rows_async(sQ, base, ld, L, Lpad);
    load_rows_async(sDO, dctx + static_cast<long long>(b) * L * ldo + h * kDh, ldo, L, Lpad);
    cp_async_commit();
    const long long stat = (static_cast<long long>(b) * H + h) * L;
    for (int i = threadIdx.x; i < Lpad; i += blockDim.x) {
        sLse[i] = i < L ? lse[stat + i] * kLog2e : INFINITY;
        sDelta[i] = i < L ? delta[stat + i] : 0.0f;
    }
    cp_async_wait<0>();
    __syncthreads();

    const int warp = threadIdx.x >> 5, l = lane_id(), g = l >> 2, t = l & 3;
    if (k0 + warp * 16 >= L) return;

    uint32_t kf[4][4], vf[4][4];
    load_a_frags(kf, sK, warp * 16);
    load_a_frags(vf, sV, warp * 16);

    const int key_a = k0 + warp * 16 + g, key_b = key_a + 8;
    const float* kb_ptr = key_bias ? key_bias + static_cast<long long>(b) * L : nullptr;
    const float bias_a = key_a < L ? (kb_ptr ? kb_ptr[key_a] * kLog2e : 0.0f) : -INFINITY;
    const float bias_b = key_b < L ? (kb_ptr ? kb_ptr[key_b] * kLog2e : 0.0f) : -INFINITY;

    float dk[8][4], dv[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.0f;
        dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.0f;
    }

    const int nqb = Lpad / 64;
    for (int qb = 0; qb < nqb; ++qb) {
        float st[8][4], dpt[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.0f;
            dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.0f;
        }
        mma_a_tT(st, kf, sQ, qb * 64);      // S^T[key, query]
        mma_a_tT(dpt, vf, sDO, qb * 64);    // dP^T[key, query] = V dO^T
        uint32_t pta[4][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int qi = qb * 64 + nt * 8 + 2 * t;
            const float ls0 = sLse[qi], ls1 = sLse[qi + 1];
            const float d0 = sDelta[qi], d1 = sDelta[qi + 1];
            const float p0 = exp2f(fmaf(st[nt][0], scale_log2, bias_a) - ls0);
            const float p1 = exp2f(fmaf(st[nt][1], scale_log2, bias_a) - ls1);
            const float p2 = exp2f(fmaf(st[nt][2], scale_log2, bias_b) - ls0);
            const float p3 = exp2f(fmaf(st[nt][3], scale_log2, bias_b) - ls1);
            st[nt][0] = p0; st[nt][1] = p1; st[nt][2] = p2; st[nt][3] = p3;
            dpt[nt][0] = p0 * (dpt[nt][0] - d0);
            dpt[nt][1] = p1 * (dpt[nt][1] - d1);
            dpt[nt][2] = p2 * (dpt[nt][2] - d0);
            dpt[nt][3] = p3 * (dpt[nt][3] - d1);
        }
        acc_to_a_frags(pta, st);
        mma_p_t(dv, pta, sDO, qb * 64);     // dV += P^T dO
        acc_to_a_frags(pta, dpt);
        mma_p_t(dk, pta, sQ, qb * 64);      // dK += dS^T Q
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { dk[nt][0] *= scale; dk[nt][1] *= scale; dk[nt][2] *= scale; dk[nt][3] *= scale; }
    const int rows_valid = min(16, L - k0 - warp * 16);
    __nv_bfloat16* dstk = dqkv + (static_cast<long long>(b) * L + k0 + warp * 16) * ld + H * kDh + h * kDh;
    store_block_bf16(dk, sK, sKg, warp * 16, dstk, ld, rows_valid);
    store_block_bf16(dv, sV, sVg, warp * 16, dstk + H * kDh, ld, rows_valid);
}

int round_up(int a, int m) { return (a + m - 1) / m * m; }

// L <= 256 runs on the tcgen05 / TMEM kernels of attention_tc.cu; longer sequences (CLiMB's
// language-only variants with stretched text) keep the mma.sync kernels of this file.
// CLIMB_ATTN_LEGACY=1 forces the latter (A/B measurements only).
bool use_tc(int L) {
    static int legacy = -1;
    if (legacy < 0) {
        const char* e = std::getenv("CLIMB_ATTN_LEGACY");
        legacy = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    return L <= 256 && legacy == 0;
}

}  // namespace

int attention_fwd(const void* qkv, const float* key_bias, void* ctx, float* lse, int B, int L, int H,
                  float scale, cudaStream_t stream) {
    CLIMB_REQUIRE(qkv && ctx && lse, "attention_fwd: null pointer");
    CLIMB_REQUIRE(B > 0 && L > 0 && H > 0, "attention_fwd: empty problem B=%d L=%d H=%d", B, L, H);
    if (use_tc(L)) {
        ProfScope prof(PROF_ATTN_FWD, static_cast<double>(B) * (4.0 * L * H * kDh * 2 + static_cast<double>(L) * H * 4), stream);
        return attention_tc_fwd(qkv, key_bias, ctx, lse, B, L, H, scale, stream);
    }
    const int Lpad = round_up(L, 64);
    const int smem = 64 * kRowBytes + 2 * Lpad * kRowBytes + Lpad * 4;
    CLIMB_REQUIRE(smem <= 227 * 1024, "attention_fwd: L=%d does not fit one CTA's shared memory", L);
    static int smem_set = 0;
    if (smem > smem_set) {
        CLIMB_CUDA_OK(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        smem_set = smem;
    }
    dim3 grid((L + 63) / 64, H, B);
    // algorithmic bytes: Q, K, V read + O written (bf16) + LSE (fp32)  
    ProfScope prof(PROF_ATTN_FWD, static_cast<double>(B) * (4.0 * L * H * kDh * 2 + static_cast<double>(L) * H * 4), stream);
    attn_fwd_kernel<<<grid, 128, smem, stream>>>(
        static_cast<const __nv_bfloat16*>(qkv), key_bias, static_cast<__nv_bfloat16*>(ctx), lse, L, H,
        Lpad, scale * kLog2e);
    CLIMB_LAUNCH_OK();
    return 0;
}

int attention_bwd(const void* qkv, const float* key_bias, const void* ctx, const void* dctx,
                  const float* lse, float* delta, void* dqkv, float* dqkv_colsum, int B, int L, int H, float scale,
                  cudaStream_t stream) {
    CLIMB_REQUIRE(qkv && ctx && dctx && lse && delta && dqkv, "attention_bwd: null pointer");
    CLIMB_REQUIRE(B > 0 && L > 0 && H > 0, "attention_bwd: empty problem B=%d L=%d H=%d", B, L, H);
    if (use_tc(L)) {
        ProfScope prof(PROF_ATTN_BWD, static_cast<double>(B) * (8.0 * L * H * kDh * 2 + 2.0 * L * H * 4), stream);
        return attention_tc_bwd(qkv, key_bias, ctx, dctx, lse, dqkv, dqkv_colsum, B, L, H, scale, stream);
    }
    const int Lpad = round_up(L, 64);
    const int smem_dq = 2 * 64 * kRowBytes + 2 * Lpad * kRowBytes + Lpad * 4;
    const int smem_dkv = 2 * 64 * kRowBytes + 2 * Lpad * kRowBytes + 2 * Lpad * 4;
    CLIMB_REQUIRE(smem_dkv <= 227 * 1024, "attention_bwd: L=%d does not fit one CTA's shared memory", L);
    static int set_dq = 0, set_dkv = 0;
    if (smem_dq > set_dq) {
        CLIMB_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dq));
        set_dq = smem_dq;
    }
    if (smem_dkv > set_dkv) {
        CLIMB_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dkv));
        set_dkv = smem_dkv;
    }
    // algorithmic bytes: Q, K, V, O, dO read + dQ, dK, dV written (bf16) + LSE, delta (fp32)
    ProfScope prof(PROF_ATTN_BWD, static_cast<double>(B) * (8.0 * L * H * kDh * 2 + 2.0 * L * H * 4), stream);
    const long long groups = static_cast<long long>(B) * L * H;
    const int threads = 256;
    const long long blocks = (groups * 8 + threads - 1) / threads;
    attn_bwd_delta_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(ctx), static_cast<const __nv_bfloat16*>(dctx), delta, B, L, H);
    CLIMB_LAUNCH_OK();
    dim3 grid((L + 63) / 64, H, B);
    attn_bwd_dq_kernel<<<grid, 128, smem_dq, stream>>>(
        static_cast<const __nv_bfloat16*>(qkv), key_bias, static_cast<const __nv_bfloat16*>(dctx), lse,
        delta, static_cast<__nv_bfloat16*>(dqkv), L, H, Lpad, scale * kLog2e, scale);
    CLIMB_LAUNCH_OK();
    attn_bwd_dkv_kernel<<<grid, 128, smem_dkv, stream>>>(
        static_cast<const __nv_bfloat16*>(qkv), key_bias, static_cast<const __nv_bfloat16*>(dctx), lse,
        delta, static_cast<__nv_bfloat16*>(dqkv), L, H, Lpad, scale * kLog2e, scale);
    CLIMB_LAUNCH_OK();
    if (dqkv_colsum != nullptr)
        return colsum(dqkv, CLIMB_BF16, 3LL * H * kDh, B * L, 3 * H * kDh, dqkv_colsum, stream);
    return 0;
}

}  // namespace climb
