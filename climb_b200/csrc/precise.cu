// CLIMB_PREC_BF16X3: the parity gate of the encoder engine ("logits within 1e-3 rel of the reference").
//
// Same network (ViltModel.forward -> pooler_output and its backward, modeling_vilt.py:777-899), same parameter arenas,
// same entry points as engine.cu -- different arithmetic:
//   * every contraction  C = A W^T  runs on the tcgen05 GEMM kernel as THREE accumulating launches over split operands,
//       x = hi + lo,  hi = bf16(x),  lo = bf16(x - hi):    A W^T ~ Ahi Whi^T + Alo Whi^T + Ahi Wlo^T
//     (the dropped lo x lo term is 2^-16 of the product; fp32 accumulation in TMEM as always);
//   * activations stay fp32 between the contractions (LayerNorm outputs, q | k | v, attention output, GELU input / output,
//     every gradient), so no bf16 rounding is left anywhere on the path;
//   * attention (softmax(Q K^T / 8 + mask) V, modeling_vilt.py:355-388) runs in a plain fp32 kernel: thread = query row,
//     keys staged through shared memory, online softmax; its backward likewise (dQ per query row; dK, dV per key row);
//   * erf-GELU through erff / expf.
// Throughput is not the point here (about 5x the bf16 step): this is the mode the north star's tolerance is checked in,
// with the bf16 mode's measured error reported beside it (tests/test_gpu_precise.py, bench.py "precision").
#include "common.cuh"
#include "internal.h"

#include <cstring>

namespace climb {
namespace {

using bf16 = __nv_bfloat16;

inline unsigned blocks_for(long long n, int threads) { return static_cast<unsigned>((n + threads - 1) / threads); }

// ---- elementwise ------------------------------------------------------------------------------------------------
__global__ void split_kernel(const float* __restrict__ x, bf16* __restrict__ hi, bf16* __restrict__ lo, long long n) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = x[i];
    const bf16 h = __float2bfloat16_rn(v);
    if (hi != nullptr) hi[i] = h;
    if (lo != nullptr) lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// C[r, c] = bias[c] + res1[r, c] + res2[r, c]   (each term optional): the value the accumulating launches add onto
__global__ void init_rows_kernel(float* __restrict__ C, const float* __restrict__ bias, const float* __restrict__ res1,
                                 const float* __restrict__ res2, long long rows, int cols) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int c = static_cast<int>(i % cols);
    float v = bias != nullptr ? bias[c] : 0.0f;
    if (res1 != nullptr) v += res1[i];
    if (res2 != nullptr) v += res2[i];
    C[i] = v;
}

__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float dgelu_exact(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    return cdf + x * 0.39894228040143267794f * expf(-0.5f * x * x);
}
__device__ __forceinline__ float swish_exact(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float dswish_exact(float x) {
    const float s = 1.0f / (1.0f + expf(-x));
    return s * (1.0f + x * (1.0f - s));
}

// y = act(pre)
__global__ void act_fwd_kernel(const float* __restrict__ pre, float* __restrict__ y, long long n, int act) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = pre[i];
    float v = x;
    if (act == CLIMB_EPI_GELU) v = gelu_exact(x);
    else if (act == CLIMB_EPI_SWISH) v = swish_exact(x);
    else if (act == CLIMB_EPI_RELU) v = fmaxf(x, 0.0f);
    else if (act == CLIMB_EPI_TANH) v = tanhf(x);
    y[i] = v;
}

// dx = dy * act'(.)  -- GELU / SWISH / RELU take the saved pre-activation, TANH the saved output
__global__ void act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ saved, float* __restrict__ dx, long long n,
                               int act) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float s = saved[i];
    float g = 1.0f;
    if (act == CLIMB_EPI_GELU) g = dgelu_exact(s);
    else if (act == CLIMB_EPI_SWISH) g = dswish_exact(s);
    else if (act == CLIMB_EPI_RELU) g = s > 0.0f ? 1.0f : 0.0f;
    else if (act == CLIMB_EPI_TANH) g = 1.0f - s * s;
    dx[i] = dy[i] * g;
}

__global__ void residual_pixels_kernel(const float* __restrict__ px, float* __restrict__ out, long long n) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = px[i];
    out[i] = v - __bfloat162float(__float2bfloat16_rn(v));
}

// ---- fp32 attention -------------------------------------------------------------------------------------------------
constexpr int kDh = 64;
constexpr int kTile = 32;         // keys (forward / dQ) or queries (dK / dV) staged per step
constexpr int kRows = 128;        // threads per CTA = rows per CTA

// ctx[b, i, h, :] = softmax_j(q_i . k_j * scale + bias[b, j]) v_j ; lse[b, h, i] = log sum_j exp(...)
__global__ void __launch_bounds__(kRows)
attn_f32_fwd_kernel(const float* __restrict__ qkv, const float* __restrict__ key_bias, float* __restrict__ ctx,
                    float* __restrict__ lse, int L, int H, float scale) {
    __shared__ float sK[kTile][kDh];
    __shared__ float sV[kTile][kDh];
    __shared__ float sB[kTile];
    const int h = blockIdx.y, b = blockIdx.z;
    const int row = blockIdx.x * kRows + threadIdx.x;
    const bool valid = row < L;
    const long long ld = 3LL * H * kDh;
    const float* base = qkv + static_cast<long long>(b) * L * ld;
    float q[kDh], o[kDh];
#pragma unroll
    for (int e = 0; e < kDh; ++e) {
        q[e] = valid ? base[static_cast<long long>(row) * ld + h * kDh + e] * scale : 0.0f;
        o[e] = 0.0f;
    }
    float m = -INFINITY, l = 0.0f;
    for (int k0 = 0; k0 < L; k0 += kTile) {
        __syncthreads();
        for (int t = threadIdx.x; t < kTile * kDh; t += kRows) {
            const int j = t / kDh, e = t - j * kDh;
            const bool in = k0 + j < L;
            sK[j][e] = in ? base[static_cast<long long>(k0 + j) * ld + (H + h) * kDh + e] : 0.0f;
            sV[j][e] = in ? base[static_cast<long long>(k0 + j) * ld + (2 * H + h) * kDh + e] : 0.0f;
        }
        if (threadIdx.x < kTile) {
            const int j = k0 + threadIdx.x;
            sB[threadIdx.x] = j < L ? (key_bias != nullptr ? key_bias[static_cast<long long>(b) * L + j] : 0.0f) : -INFINITY;
        }
        __syncthreads();
        float s[kTile];
        float tmax = -INFINITY;
#pragma unroll
        for (int j = 0; j < kTile; ++j) {
            float acc = 0.0f;
#pragma unroll
            for (int e = 0; e < kDh; ++e) acc = fmaf(q[e], sK[j][e], acc);
            s[j] = acc + sB[j];
            tmax = fmaxf(tmax, s[j]);
        }
        const float m_new = fmaxf(m, tmax);          // key 0 of the first tile is always a real key: finite from then on
        const float corr = expf(m - m_new);          // exp(-inf) = 0 on the first tile
        l *= corr;
#pragma unroll
        for (int e = 0; e < kDh; ++e) o[e] *= corr;
#pragma unroll
        for (int j = 0; j < kTile; ++j) {
            const float p = expf(s[j] - m_new);      // -inf past L -> 0
            l += p;
#pragma unroll
            for (int e = 0; e < kDh; ++e) o[e] = fmaf(p, sV[j][e], o[e]);
        }
        m = m_new;
    }
    if (valid) {
        const float inv = 1.0f / l;
        float* dst = ctx + (static_cast<long long>(b) * L + row) * (H * kDh) + h * kDh;
#pragma unroll
        for (int e = 0; e < kDh; ++e) dst[e] = o[e] * inv;
        lse[(static_cast<long long>(b) * H + h) * L + row] = m + logf(l);
    }
}

// dQ_i = scale * sum_j dS_ij k_j,  dS_ij = P_ij (dO_i . v_j - delta_i),  delta_i = dO_i . O_i  (also written out)
__global__ void __launch_bounds__(kRows)
attn_f32_bwd_dq_kernel(const float* __restrict__ qkv, const float* __restrict__ key_bias, const float* __restrict__ ctx,
                       const float* __restrict__ dctx, const float* __restrict__ lse, float* __restrict__ delta,
                       float* __restrict__ dqkv, int L, int H, float scale) {
    __shared__ float sK[kTile][kDh];
    __shared__ float sV[kTile][kDh];
    __shared__ float sB[kTile];
    const int h = blockIdx.y, b = blockIdx.z;
    const int row = blockIdx.x * kRows + threadIdx.x;
    const bool valid = row < L;
    const long long ld = 3LL * H * kDh, ldo = static_cast<long long>(H) * kDh;
    const float* base = qkv + static_cast<long long>(b) * L * ld;
    float q[kDh], dO[kDh], dq[kDh];
    float dl = 0.0f;
#pragma unroll
    for (int e = 0; e < kDh; ++e) {
        q[e] = valid ? base[static_cast<long long>(row) * ld + h * kDh + e] * scale : 0.0f;
        dO[e] = valid ? dctx[(static_cast<long long>(b) * L + row) * ldo + h * kDh + e] : 0.0f;
        const float oe = valid ? ctx[(static_cast<long long>(b) * L + row) * ldo + h * kDh + e] : 0.0f;
        dl = fmaf(dO[e], oe, dl);
        dq[e] = 0.0f;
    }
    const float lse_i = valid ? lse[(static_cast<long long>(b) * H + h) * L + row] : 0.0f;
    if (valid) delta[(static_cast<long long>(b) * H + h) * L + row] = dl;
    for (int k0 = 0; k0 < L; k0 += kTile) {
        __syncthreads();
        for (int t = threadIdx.x; t < kTile * kDh; t += kRows) {
            const int j = t / kDh, e = t - j * kDh;
            const bool in = k0 + j < L;
            sK[j][e] = in ? base[static_cast<long long>(k0 + j) * ld + (H + h) * kDh + e] : 0.0f;
            sV[j][e] = in ? base[static_cast<long long>(k0 + j) * ld + (2 * H + h) * kDh + e] : 0.0f;
        }
        if (threadIdx.x < kTile) {
            const int j = k0 + threadIdx.x;
            sB[threadIdx.x] = j < L ? (key_bias != nullptr ? key_bias[static_cast<long long>(b) * L + j] : 0.0f) : -INFINITY;
        }
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < kTile; ++j) {
            float s = 0.0f, dp = 0.0f;
#pragma unroll
            for (int e = 0; e < kDh; ++e) {
                s = fmaf(q[e], sK[j][e], s);
                dp = fmaf(dO[e], sV[j][e], dp);
            }
            const float p = expf(s + sB[j] - lse_i);
            const float ds = p * (dp - dl) * scale;
#pragma unroll
            for (int e = 0; e < kDh; ++e) dq[e] = fmaf(ds, sK[j][e], dq[e]);
        }
    }
    if (valid) {
        float* dst = dqkv + (static_cast<long long>(b) * L + row) * ld + h * kDh;
#pragma unroll
        for (int e = 0; e < kDh; ++e) dst[e] = dq[e];
    }
}

// thread = key row j.  WHICH = 0: dV_j = sum_i P_ij dO_i ;  WHICH = 1: dK_j = scale * sum_i dS_ij q_i
template <int WHICH>
__global__ void __launch_bounds__(kRows)
attn_f32_bwd_dkv_kernel(const float* __restrict__ qkv, const float* __restrict__ key_bias, const float* __restrict__ dctx,
                        const float* __restrict__ lse, const float* __restrict__ delta, float* __restrict__ dqkv, int L, int H,
                        float scale) {
    __shared__ float sQ[kTile][kDh];
    __shared__ float sDO[kTile][kDh];
    __shared__ float sLse[kTile];
    __shared__ float sDelta[kTile];
    const int h = blockIdx.y, b = blockIdx.z;
    const int row = blockIdx.x * kRows + threadIdx.x;          // key index
    const bool valid = row < L;
    const long long ld = 3LL * H * kDh, ldo = static_cast<long long>(H) * kDh;
    const float* base = qkv + static_cast<long long>(b) * L * ld;
    float k[kDh], v[WHICH == 1 ? kDh : 1], acc[kDh];
#pragma unroll
    for (int e = 0; e < kDh; ++e) {
        k[e] = valid ? base[static_cast<long long>(row) * ld + (H + h) * kDh + e] * scale : 0.0f;
        if constexpr (WHICH == 1) v[e] = valid ? base[static_cast<long long>(row) * ld + (2 * H + h) * kDh + e] : 0.0f;
        acc[e] = 0.0f;
    }
    const float bias_j = valid ? (key_bias != nullptr ? key_bias[static_cast<long long>(b) * L + row] : 0.0f) : -INFINITY;
    for (int i0 = 0; i0 < L; i0 += kTile) {
        __syncthreads();
        for (int t = threadIdx.x; t < kTile * kDh; t += kRows) {
            const int i = t / kDh, e = t - i * kDh;
            const bool in = i0 + i < L;
            sQ[i][e] = in ? base[static_cast<long long>(i0 + i) * ld + h * kDh + e] : 0.0f;
            sDO[i][e] = in ? dctx[(static_cast<long long>(b) * L + i0 + i) * ldo + h * kDh + e] : 0.0f;
        }
        if (threadIdx.x < kTile) {
            const int i = i0 + threadIdx.x;
            sLse[threadIdx.x] = i < L ? lse[(static_cast<long long>(b) * H + h) * L + i] : INFINITY;     // +inf -> P = 0
            sDelta[threadIdx.x] = i < L ? delta[(static_cast<long long>(b) * H + h) * L + i] : 0.0f;
        }
        __syncthreads();
#pragma unroll 4
        for (int i = 0; i < kTile; ++i) {
            float s = 0.0f;
#pragma unroll
            for (int e = 0; e < kDh; ++e) s = fmaf(sQ[i][e], k[e], s);          // q_i . (scale k_j)
            const float p = expf(s + bias_j - sLse[i]);
            if constexpr (WHICH == 0) {
#pragma unroll
                for (int e = 0; e < kDh; ++e) acc[e] = fmaf(p, sDO[i][e], acc[e]);
            } else {
                float dp = 0.0f;
#pragma unroll
                for (int e = 0; e < kDh; ++e) dp = fmaf(sDO[i][e], v[e], dp);
                const float ds = p * (dp - sDelta[i]) * scale;
#pragma unroll
                for (int e = 0; e < kDh; ++e) acc[e] = fmaf(ds, sQ[i][e], acc[e]);
            }
        }
    }
    if (valid) {
        float* dst = dqkv + (static_cast<long long>(b) * L + row) * ld + (WHICH == 0 ? 2 * H + h : H + h) * kDh;
#pragma unroll
        for (int e = 0; e < kDh; ++e) dst[e] = acc[e];
    }
}

int attn_f32_fwd(const float* qkv, const float* key_bias, float* ctx, float* lse, int B, int L, int H, float scale, cudaStream_t s) {
    dim3 grid((L + kRows - 1) / kRows, H, B);
    attn_f32_fwd_kernel<<<grid, kRows, 0, s>>>(qkv, key_bias, ctx, lse, L, H, scale);
    CLIMB_LAUNCH_OK();
    return 0;
}
int attn_f32_bwd(const float* qkv, const float* key_bias, const float* ctx, const float* dctx, const float* lse, float* delta,
                 float* dqkv, int B, int L, int H, float scale, cudaStream_t s) {
    dim3 grid((L + kRows - 1) / kRows, H, B);
    attn_f32_bwd_dq_kernel<<<grid, kRows, 0, s>>>(qkv, key_bias, ctx, dctx, lse, delta, dqkv, L, H, scale);
    CLIMB_LAUNCH_OK();
    attn_f32_bwd_dkv_kernel<0><<<grid, kRows, 0, s>>>(qkv, key_bias, dctx, lse, delta, dqkv, L, H, scale);
    CLIMB_LAUNCH_OK();
    attn_f32_bwd_dkv_kernel<1><<<grid, kRows, 0, s>>>(qkv, key_bias, dctx, lse, delta, dqkv, L, H, scale);
    CLIMB_LAUNCH_OK();
    return 0;
}

// ---- split-operand contractions ----------------------------------------------------------------------------------------
#define TRY(expr)                 \
    do {                          \
        int _rc = (expr);         \
        if (_rc) return _rc;      \
    } while (0)

int split(const float* x, bf16* hi, bf16* lo, long long n, cudaStream_t s) {
    split_kernel<<<blocks_for(n, 256), 256, 0, s>>>(x, hi, lo, n);
    CLIMB_LAUNCH_OK();
    return 0;
}
int init_rows(float* C, const float* bias, const float* r1, const float* r2, long long rows, int cols, cudaStream_t s) {
    init_rows_kernel<<<blocks_for(rows * cols, 256), 256, 0, s>>>(C, bias, r1, r2, rows, cols);
    CLIMB_LAUNCH_OK();
    return 0;
}
int act_fwd(const float* pre, float* y, long long n, int act, cudaStream_t s) {
    act_fwd_kernel<<<blocks_for(n, 256), 256, 0, s>>>(pre, y, n, act);
    CLIMB_LAUNCH_OK();
    return 0;
}
int act_bwd(const float* dy, const float* saved, float* dx, long long n, int act, cudaStream_t s) {
    act_bwd_kernel<<<blocks_for(n, 256), 256, 0, s>>>(dy, saved, dx, n, act);
    CLIMB_LAUNCH_OK();
    return 0;
}

// one accumulating launch of the generic tcgen05 kernel: C (fp32) += A B^T in the given operand layouts
int gemm_acc(int M, int N, int K, const bf16* A, long long lda, int a_mn, const bf16* B, long long ldb, int b_mn, float* C,
             long long ldc, cudaStream_t s) {
    climb_gemm_desc g;
    std::memset(&g, 0, sizeof(g));
    g.M = M; g.N = N; g.K = K;
    g.A = A; g.lda = lda; g.a_mn_major = a_mn;
    g.B = B; g.ldb = ldb; g.b_mn_major = b_mn;
    g.C = C; g.ldc = ldc; g.c_dtype = CLIMB_F32;
    g.alpha = 1.0f; g.accumulate = 1; g.split_k = 0;
    return gemm_bf16(&g, s);
}

struct Ctx {
    const bf16* w_hi;       // bf16(theta)
    const bf16* w_lo;       // bf16(theta - w_hi)
    bf16 *a_hi, *a_lo;      // operand scratch: [rows, max width]
    bf16 *b_hi, *b_lo;      // second operand scratch (weight gradients: the saved activation)
    cudaStream_t s;
};

// Y[M, N] = bias + res1 + res2 + X[M, K] W[N, K]^T     (forward of nn.Linear)
int plinear(const Ctx& c, int M, int N, int K, const float* X, long long w_off, const float* bias, const float* res1,
            const float* res2, float* Y) {
    TRY(init_rows(Y, bias, res1, res2, M, N, c.s));
    TRY(split(X, c.a_hi, c.a_lo, static_cast<long long>(M) * K, c.s));
    TRY(gemm_acc(M, N, K, c.a_hi, K, 0, c.w_hi + w_off, K, 0, Y, N, c.s));
    TRY(gemm_acc(M, N, K, c.a_lo, K, 0, c.w_hi + w_off, K, 0, Y, N, c.s));
    TRY(gemm_acc(M, N, K, c.a_hi, K, 0, c.w_lo + w_off, K, 0, Y, N, c.s));
    return 0;
}
// dX[M, K] = res + dY[M, N] W[N, K]                     (W read in place as an MN-major B operand)
int pdgrad(const Ctx& c, int M, int N, int K, const float* dY, long long w_off, const float* res, float* dX) {
    TRY(init_rows(dX, nullptr, res, nullptr, M, K, c.s));
    TRY(split(dY, c.a_hi, c.a_lo, static_cast<long long>(M) * N, c.s));
    TRY(gemm_acc(M, K, N, c.a_hi, N, 0, c.w_hi + w_off, K, 1, dX, K, c.s));
    TRY(gemm_acc(M, K, N, c.a_lo, N, 0, c.w_hi + w_off, K, 1, dX, K, c.s));
    TRY(gemm_acc(M, K, N, c.a_hi, N, 0, c.w_lo + w_off, K, 1, dX, K, c.s));
    return 0;
}
// dW[N, K] += dY[M, N]^T X[M, K] ;  db[N] += column sums of dY
int pwgrad(const Ctx& c, int M, int N, int K, const float* dY, const float* X, float* dW, float* db) {
    TRY(split(dY, c.a_hi, c.a_lo, static_cast<long long>(M) * N, c.s));
    TRY(split(X, c.b_hi, c.b_lo, static_cast<long long>(M) * K, c.s));
    TRY(gemm_acc(N, K, M, c.a_hi, N, 1, c.b_hi, K, 1, dW, K, c.s));
    TRY(gemm_acc(N, K, M, c.a_lo, N, 1, c.b_hi, K, 1, dW, K, c.s));
    TRY(gemm_acc(N, K, M, c.a_hi, N, 1, c.b_lo, K, 1, dW, K, c.s));
    if (db != nullptr) TRY(colsum(dY, CLIMB_F32, N, M, N, db, c.s));
    return 0;
}

// ---- workspace -----------------------------------------------------------------------------------------------------------
struct Bump {
    uint8_t* base;
    long long off = 0;
    explicit Bump(void* b) : base(static_cast<uint8_t*>(b)) {}
    template <typename T>
    T* take(long long n) {
        off = (off + 255) & ~255LL;
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += n * static_cast<long long>(sizeof(T));
        return p;
    }
};

struct LayerAct {
    float* x_in;
    float *h1, *mean1, *rstd1;
    float *qkv, *ctx, *lse;
    float *x1, *h2, *mean2, *rstd2;
    float *u, *inter;
    float *mh_in, *mh_pre, *mh_z, *out_in, *out_pre, *out_z;
};

struct Plan {
    int B, T, Hh, Ww, hp, wp, Np, L, M, d, ff, heads, layers, Kp, r;
    const int* geom;
    float* key_bias;
    float *text_e, *text_mean, *text_rstd, *text_ln;
    float* px_lo;               // pixel residual (fp32) for the lo half of the im2col operand
    bf16 *im2col_hi, *im2col_lo;
    float *pos_table, *patch_out;
    LayerAct act[64];
    float* x_final;
    float *cls_ln, *fmean, *frstd, *pool_pre, *pooled;
    bf16 *a_hi, *a_lo;
    long long bytes;
};

int fill_plan(Plan& P, const climb_vilt_dims* dm, const climb_vilt_params* pr, const climb_vilt_batch* bt, void* base, int save) {
    CLIMB_REQUIRE(dm && pr && bt, "engine: null descriptor");
    CLIMB_REQUIRE(dm->layers > 0 && dm->layers <= 64, "engine: layers=%d out of range", dm->layers);
    CLIMB_REQUIRE(dm->hidden % 128 == 0 && dm->hidden == dm->heads * 64, "engine: hidden=%d must be heads*64 and a multiple of 128",
                  dm->hidden);
    CLIMB_REQUIRE(dm->ffn % 8 == 0, "engine: ffn=%d must be a multiple of 8", dm->ffn);
    CLIMB_REQUIRE(bt->B > 0 && bt->T > 0, "engine: empty batch (B=%d, T=%d)", bt->B, bt->T);
    CLIMB_REQUIRE(bt->H > 0 && bt->W > 0 && bt->H % dm->patch == 0 && bt->W % dm->patch == 0,
                  "engine: image %dx%d is not a multiple of the patch size %d", bt->H, bt->W, dm->patch);
    CLIMB_REQUIRE(bt->image_repeat <= 1, "engine (bf16x3): image_repeat=%d is implemented by the bf16 engine only", bt->image_repeat);
    P.B = bt->B; P.T = bt->T; P.Hh = bt->H; P.Ww = bt->W;
    P.hp = bt->H / dm->patch; P.wp = bt->W / dm->patch; P.Np = P.hp * P.wp;
    P.geom = bt->patch_geom;
    if (P.geom != nullptr) {
        CLIMB_REQUIRE(bt->n_patch_slots > 0 && bt->n_patch_slots <= P.Np, "engine: n_patch_slots=%d outside (0, %d]", bt->n_patch_slots,
                      P.Np);
        P.Np = bt->n_patch_slots;
    }
    P.L = P.T + 1 + P.Np; P.M = P.B * P.L;
    P.d = dm->hidden; P.ff = dm->ffn; P.heads = dm->heads; P.layers = dm->layers;
    P.Kp = dm->channels * dm->patch * dm->patch;
    P.r = pr->adapter_r;
    CLIMB_REQUIRE(P.r == 0 || P.r % 8 == 0, "engine: adapter width %d must be a multiple of 8", P.r);
    const long long M = P.M, d = P.d, ff = P.ff, BT = static_cast<long long>(P.B) * P.T, BN = static_cast<long long>(P.B) * P.Np;
    Bump b(base);
    P.key_bias = b.take<float>(static_cast<long long>(P.B) * P.L);
    P.text_e = b.take<float>(BT * d);
    P.text_mean = b.take<float>(BT);
    P.text_rstd = b.take<float>(BT);
    P.text_ln = b.take<float>(BT * d);
    P.px_lo = b.take<float>(static_cast<long long>(P.B) * dm->channels * P.Hh * P.Ww);
    P.im2col_hi = b.take<bf16>(BN * P.Kp);
    P.im2col_lo = b.take<bf16>(BN * P.Kp);
    P.pos_table = b.take<float>(static_cast<long long>(P.Np) * d);
    P.patch_out = b.take<float>(BN * d);
    float* xbuf[65];
    const int n_x = save ? P.layers + 1 : 2;
    for (int i = 0; i < n_x; ++i) xbuf[i] = b.take<float>(M * d);
    for (int l = 0; l < P.layers; ++l) {
        LayerAct& a = P.act[l];
        if (save || l == 0) {
            a.h1 = b.take<float>(M * d); a.mean1 = b.take<float>(M); a.rstd1 = b.take<float>(M);
            a.qkv = b.take<float>(M * 3 * d);
            a.ctx = b.take<float>(M * d);
            a.lse = b.take<float>(static_cast<long long>(P.B) * P.heads * P.L);
            a.x1 = b.take<float>(M * d);
            a.h2 = b.take<float>(M * d); a.mean2 = b.take<float>(M); a.rstd2 = b.take<float>(M);
            a.u = b.take<float>(M * ff);
            a.inter = b.take<float>(M * ff);
            if (P.r > 0) {
                a.mh_in = b.take<float>(M * d); a.mh_pre = b.take<float>(M * P.r); a.mh_z = b.take<float>(M * P.r);
                a.out_in = b.take<float>(M * d); a.out_pre = b.take<float>(M * P.r); a.out_z = b.take<float>(M * P.r);
            } else {
                a.mh_in = a.mh_pre = a.mh_z = a.out_in = a.out_pre = a.out_z = nullptr;
            }
        } else {
            a = P.act[0];
        }
        a.x_in = save ? xbuf[l] : xbuf[l & 1];
    }
    P.x_final = save ? xbuf[P.layers] : xbuf[P.layers & 1];
    P.cls_ln = b.take<float>(static_cast<long long>(P.B) * d);
    P.fmean = b.take<float>(P.B);
    P.frstd = b.take<float>(P.B);
    P.pool_pre = b.take<float>(static_cast<long long>(P.B) * d);
    P.pooled = b.take<float>(static_cast<long long>(P.B) * d);
    const long long widest = ff > 3 * d ? ff : 3 * d;
    P.a_hi = b.take<bf16>(M * widest);
    P.a_lo = b.take<bf16>(M * widest);
    P.bytes = (b.off + 255) & ~255LL;
    return 0;
}

struct BwdScratch {
    float *dxa, *dxb;           // gradient of the residual stream (ping-pong), fp32 [M, d]
    float* du;                  // [M, ff]
    float* dinter;              // [M, ff]
    float* dh;                  // [M, d]   dh2 / dctx / dh1 in turn
    float* dqkv;                // [M, 3d]
    float* delta;               // [B, H, L]
    float *dz, *dpre;           // [M, r]
    float* dmh;                 // [M, d]
    float *dy_text, *de_text;   // [B*T, d]
    float* dpatch;              // [B*Np, d]
    bf16* dpatch_h;             // bf16 view that embed_split_bwd writes (unused operand of this mode)
    float* S;                   // [2, L, d]
    float *dpool, *dpool_pre, *dcls;   // [B, d]
    bf16 *a_hi, *a_lo, *b_hi, *b_lo;
    long long bytes;
};

void fill_scratch(BwdScratch& S, const Plan& P, void* base) {
    Bump b(base);
    const long long M = P.M, d = P.d, ff = P.ff, r = P.r > 0 ? P.r : 8;
    S.dxa = b.take<float>(M * d); S.dxb = b.take<float>(M * d);
    S.du = b.take<float>(M * ff);
    S.dinter = b.take<float>(M * ff);
    S.dh = b.take<float>(M * d);
    S.dqkv = b.take<float>(M * 3 * d);
    S.delta = b.take<float>(static_cast<long long>(P.B) * P.heads * P.L);
    S.dz = b.take<float>(M * r); S.dpre = b.take<float>(M * r);
    S.dmh = b.take<float>(P.r > 0 ? M * d : 8);
    S.dy_text = b.take<float>(static_cast<long long>(P.B) * P.T * d);
    S.de_text = b.take<float>(static_cast<long long>(P.B) * P.T * d);
    S.dpatch = b.take<float>(static_cast<long long>(P.B) * P.Np * d);
    S.dpatch_h = b.take<bf16>(static_cast<long long>(P.B) * P.Np * d);
    S.S = b.take<float>(2LL * P.L * d);
    S.dpool = b.take<float>(static_cast<long long>(P.B) * d);
    S.dpool_pre = b.take<float>(static_cast<long long>(P.B) * d);
    S.dcls = b.take<float>(static_cast<long long>(P.B) * d);
    long long widest = ff > 3 * d ? ff : 3 * d;
    if (P.Kp > widest) widest = P.Kp;
    S.a_hi = b.take<bf16>(M * widest); S.a_lo = b.take<bf16>(M * widest);
    S.b_hi = b.take<bf16>(M * widest); S.b_lo = b.take<bf16>(M * widest);
    S.bytes = (b.off + 255) & ~255LL;
}

inline const float* F(const float* theta, long long off) { return off >= 0 ? theta + off : nullptr; }
inline float* G(float* grad, long long off) { return off >= 0 ? grad + off : nullptr; }

// patch rows of the residual-stream gradient, fp32: dpatch[b, p, :] = dx[b, T + 1 + p, :]
__global__ void patch_rows_kernel(const float* __restrict__ dx, float* __restrict__ dpatch, int B, int T, int Np, int d) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(B) * Np * d) return;
    const long long row = i / d;
    const int c = static_cast<int>(i - row * d);
    const int b = static_cast<int>(row / Np), p = static_cast<int>(row - static_cast<long long>(b) * Np);
    dpatch[i] = dx[(static_cast<long long>(b) * (T + 1 + Np) + T + 1 + p) * d + c];
}

}  // namespace

long long vilt_forward_workspace_bytes_precise(const climb_vilt_dims* dims, const climb_vilt_params* params,
                                               const climb_vilt_batch* batch, int save) {
    Plan P;
    if (fill_plan(P, dims, params, batch, nullptr, save)) return -1;
    return P.bytes;
}
long long vilt_backward_scratch_bytes_precise(const climb_vilt_dims* dims, const climb_vilt_params* params,
                                              const climb_vilt_batch* batch) {
    Plan P;
    if (fill_plan(P, dims, params, batch, nullptr, 1)) return -1;
    BwdScratch S;
    fill_scratch(S, P, nullptr);
    return S.bytes;
}

int split_f32_bf16x2(const float* src, void* hi, void* lo, long long n, cudaStream_t s) {
    CLIMB_REQUIRE(src && n > 0 && (hi || lo), "split_f32_bf16x2: bad arguments");
    return split(src, static_cast<bf16*>(hi), static_cast<bf16*>(lo), n, s);
}

int vilt_forward_precise(const climb_vilt_dims* dm, const climb_vilt_params* pr, const climb_vilt_batch* bt, const float* theta,
                         const void* shadow, void* workspace, long long workspace_bytes, int save, float* pooled_out,
                         cudaStream_t s) {
    Plan P;
    TRY(fill_plan(P, dm, pr, bt, workspace, save));
    CLIMB_REQUIRE(theta && shadow && workspace && pooled_out, "vilt_forward: null buffer");
    CLIMB_REQUIRE(pr->shadow_lo != nullptr, "vilt_forward (bf16x3): params.shadow_lo missing (climb_split_f32_bf16x2 of theta)");
    CLIMB_REQUIRE(workspace_bytes >= P.bytes, "vilt_forward: workspace %lld < required %lld", workspace_bytes, P.bytes);
    CLIMB_REQUIRE((bt->input_ids != nullptr) != (bt->inputs_embeds != nullptr), "vilt_forward: exactly one of input_ids / inputs_embeds");
    CLIMB_REQUIRE(bt->pixel_values != nullptr, "vilt_forward: pixel_values missing");
    CLIMB_REQUIRE(bt->image_type_idx != nullptr || (bt->image_type_idx_scalar >= 0 && bt->image_type_idx_scalar < dm->n_modality),
                  "vilt_forward: image_token_type_idx %d outside the %d-row modality table", bt->image_type_idx_scalar, dm->n_modality);
    CLIMB_REQUIRE(!(bt->training && (dm->hidden_dropout > 0.0f || dm->attn_dropout > 0.0f)),
                  "vilt_forward (bf16x3): dropout > 0 is implemented in the bf16 mode only");
    const int d = P.d, M = P.M, BT = P.B * P.T;
    Ctx c{static_cast<const bf16*>(shadow), static_cast<const bf16*>(pr->shadow_lo), P.a_hi, P.a_lo, nullptr, nullptr, s};

    // ---- embeddings (modeling_vilt.py:207-246) ----
    if (P.geom) TRY(key_bias_ragged(reinterpret_cast<const long long*>(bt->attention_mask), P.geom, P.key_bias, P.B, P.T, P.L, s, bt->patch_select));
    else if (bt->attention_mask) TRY(key_bias(reinterpret_cast<const long long*>(bt->attention_mask), P.key_bias, P.B, P.T, P.L, s));
    else CLIMB_CUDA_OK(cudaMemsetAsync(P.key_bias, 0, sizeof(float) * P.B * P.L, s));
    TRY(text_gather(reinterpret_cast<const long long*>(bt->input_ids), bt->inputs_embeds,
                    reinterpret_cast<const long long*>(bt->token_type_ids), F(theta, pr->word_emb), F(theta, pr->text_type_emb),
                    F(theta, pr->text_pos_emb), P.text_e, BT, P.T, d, s, dm->vocab_size, dm->type_vocab_size));
    TRY(layernorm_fwd(P.text_e, d, F(theta, pr->text_ln_w), F(theta, pr->text_ln_b), dm->ln_eps, nullptr, P.text_ln, P.text_mean,
                      P.text_rstd, BT, d, CLIMB_EPI_NONE, s));
    {
        const long long npx = static_cast<long long>(P.B) * dm->channels * P.Hh * P.Ww;
        residual_pixels_kernel<<<blocks_for(npx, 256), 256, 0, s>>>(bt->pixel_values, P.px_lo, npx);
        CLIMB_LAUNCH_OK();
        if (P.geom) {
            TRY(im2col_ragged(bt->pixel_values, P.geom, P.im2col_hi, P.B, dm->channels, P.Hh, P.Ww, dm->patch, P.Np, s, 1, bt->patch_select));
            TRY(im2col_ragged(P.px_lo, P.geom, P.im2col_lo, P.B, dm->channels, P.Hh, P.Ww, dm->patch, P.Np, s, 1, bt->patch_select));
        } else {
            TRY(im2col(bt->pixel_values, P.im2col_hi, P.B, dm->channels, P.Hh, P.Ww, dm->patch, s));
            TRY(im2col(P.px_lo, P.im2col_lo, P.B, dm->channels, P.Hh, P.Ww, dm->patch, s));
        }
        const int Mp = P.B * P.Np;
        TRY(init_rows(P.patch_out, F(theta, pr->patch_b), nullptr, nullptr, Mp, d, s));
        TRY(gemm_acc(Mp, d, P.Kp, P.im2col_hi, P.Kp, 0, c.w_hi + pr->patch_w, P.Kp, 0, P.patch_out, d, s));
        TRY(gemm_acc(Mp, d, P.Kp, P.im2col_lo, P.Kp, 0, c.w_hi + pr->patch_w, P.Kp, 0, P.patch_out, d, s));
        TRY(gemm_acc(Mp, d, P.Kp, P.im2col_hi, P.Kp, 0, c.w_lo + pr->patch_w, P.Kp, 0, P.patch_out, d, s));
    }
    if (P.geom) {
        TRY(embed_assemble_ragged(P.text_ln, P.patch_out, P.geom, F(theta, pr->cls_token), F(theta, pr->pos_emb), F(theta, pr->mod_emb),
                                  bt->image_type_idx, bt->image_type_idx_scalar, P.act[0].x_in, P.B, P.T, P.Np, dm->pos_grid, d, s,
                                  dm->n_modality, 0.0f, 0, 1, bt->patch_select));
    } else {
        TRY(pos_interp(F(theta, pr->pos_emb), P.pos_table, P.hp, P.wp, dm->pos_grid, d, s));
        TRY(embed_assemble(P.text_ln, P.patch_out, P.pos_table, F(theta, pr->cls_token), F(theta, pr->pos_emb), F(theta, pr->mod_emb),
                           bt->image_type_idx, bt->image_type_idx_scalar, P.act[0].x_in, P.B, P.T, P.Np, d, s, dm->n_modality));
    }

    // ---- encoder layers (modeling_vilt.py:503-525) ----
    for (int li = 0; li < P.layers; ++li) {
        const climb_vilt_layer& w = pr->layer[li];
        LayerAct& a = P.act[li];
        float* x_out = (li + 1 < P.layers) ? P.act[li + 1].x_in : P.x_final;
        const bool mh_ad = P.r > 0 && w.mh_down_w >= 0;
        const bool out_ad = P.r > 0 && w.out_down_w >= 0;
        TRY(layernorm_fwd(a.x_in, d, F(theta, w.ln1_w), F(theta, w.ln1_b), dm->ln_eps, nullptr, a.h1, a.mean1, a.rstd1, M, d,
                          CLIMB_EPI_NONE, s));
        TRY(plinear(c, M, 3 * d, d, a.h1, w.qkv_w, F(theta, w.qkv_b), nullptr, nullptr, a.qkv));
        TRY(attn_f32_fwd(a.qkv, P.key_bias, a.ctx, a.lse, P.B, P.L, P.heads, 0.125f, s));
        if (mh_ad) {
            // the adapter wraps O(ctx) + b BEFORE the layer residual (mixins/vilt.py:23-69): x1 = h + up(act(down(h))) + x_in
            TRY(plinear(c, M, d, d, a.ctx, w.o_w, F(theta, w.o_b), nullptr, nullptr, a.mh_in));
            TRY(plinear(c, M, P.r, d, a.mh_in, w.mh_down_w, F(theta, w.mh_down_b), nullptr, nullptr, a.mh_pre));
            TRY(act_fwd(a.mh_pre, a.mh_z, static_cast<long long>(M) * P.r, pr->adapter_act, s));
            TRY(plinear(c, M, d, P.r, a.mh_z, w.mh_up_w, F(theta, w.mh_up_b), a.mh_in, a.x_in, a.x1));
        } else {
            TRY(plinear(c, M, d, d, a.ctx, w.o_w, F(theta, w.o_b), a.x_in, nullptr, a.x1));
        }
        TRY(layernorm_fwd(a.x1, d, F(theta, w.ln2_w), F(theta, w.ln2_b), dm->ln_eps, nullptr, a.h2, a.mean2, a.rstd2, M, d,
                          CLIMB_EPI_NONE, s));
        TRY(plinear(c, M, P.ff, d, a.h2, w.fc1_w, F(theta, w.fc1_b), nullptr, nullptr, a.u));
        TRY(act_fwd(a.u, a.inter, static_cast<long long>(M) * P.ff, CLIMB_EPI_GELU, s));
        if (out_ad) {
            // the output adapter wraps FC2 + residual (mixins/vilt.py:79-125): x_out = y + up(act(down(y))), y = FC2(inter) + x1
            TRY(plinear(c, M, d, P.ff, a.inter, w.fc2_w, F(theta, w.fc2_b), a.x1, nullptr, a.out_in));
            TRY(plinear(c, M, P.r, d, a.out_in, w.out_down_w, F(theta, w.out_down_b), nullptr, nullptr, a.out_pre));
            TRY(act_fwd(a.out_pre, a.out_z, static_cast<long long>(M) * P.r, pr->adapter_act, s));
            TRY(plinear(c, M, d, P.r, a.out_z, w.out_up_w, F(theta, w.out_up_b), a.out_in, nullptr, x_out));
        } else {
            TRY(plinear(c, M, d, P.ff, a.inter, w.fc2_w, F(theta, w.fc2_b), a.x1, nullptr, x_out));
        }
    }

    // ---- final LayerNorm on the [CLS] rows + pooler (modeling_vilt.py:873-874, 887-899) ----
    TRY(layernorm_fwd(P.x_final, static_cast<long long>(P.L) * d, F(theta, pr->final_ln_w), F(theta, pr->final_ln_b), dm->ln_eps, nullptr,
                      P.cls_ln, P.fmean, P.frstd, P.B, d, CLIMB_EPI_NONE, s));
    TRY(plinear(c, P.B, d, d, P.cls_ln, pr->pooler_w, F(theta, pr->pooler_b), nullptr, nullptr, P.pool_pre));
    TRY(act_fwd(P.pool_pre, P.pooled, static_cast<long long>(P.B) * d, CLIMB_EPI_TANH, s));
    CLIMB_CUDA_OK(cudaMemcpyAsync(pooled_out, P.pooled, sizeof(float) * P.B * d, cudaMemcpyDeviceToDevice, s));
    return 0;
}

int vilt_backward_precise(const climb_vilt_dims* dm, const climb_vilt_params* pr, const climb_vilt_batch* bt, const float* theta,
                          const void* shadow, const void* workspace, long long workspace_bytes, void* scratch,
                          long long scratch_bytes, const float* dpooled, float* grad, int first_layer, int last_layer, int parts,
                          cudaStream_t s) {
    Plan P;
    TRY(fill_plan(P, dm, pr, bt, const_cast<void*>(workspace), 1));
    CLIMB_REQUIRE(theta && shadow && workspace && scratch && dpooled && grad, "vilt_backward: null buffer");
    CLIMB_REQUIRE(pr->shadow_lo != nullptr, "vilt_backward (bf16x3): params.shadow_lo missing");
    CLIMB_REQUIRE(workspace_bytes >= P.bytes, "vilt_backward: workspace %lld < required %lld", workspace_bytes, P.bytes);
    BwdScratch S;
    fill_scratch(S, P, scratch);
    CLIMB_REQUIRE(scratch_bytes >= S.bytes, "vilt_backward: scratch %lld < required %lld", scratch_bytes, S.bytes);
    const int d = P.d, M = P.M, ff = P.ff, r = P.r, BT = P.B * P.T;
    Ctx c{static_cast<const bf16*>(shadow), static_cast<const bf16*>(pr->shadow_lo), S.a_hi, S.a_lo, S.b_hi, S.b_lo, s};

    int lowest = P.layers;
    if (pr->embed_flags & CLIMB_TRAIN_BASE) lowest = 0;
    else
        for (int l = 0; l < P.layers; ++l)
            if (pr->layer[l].flags & (CLIMB_TRAIN_BASE | CLIMB_TRAIN_ADAPTER)) { lowest = l; break; }
    const bool tail = (pr->tail_flags & CLIMB_TRAIN_BASE) != 0;
    if (lowest == P.layers && !tail) return 0;
    CLIMB_REQUIRE(first_layer < P.layers && last_layer >= 0 && (first_layer >= last_layer || first_layer < 0),
                  "vilt_backward: bad layer range [%d, %d]", first_layer, last_layer);

    float* dx = S.dxa;      // gradient w.r.t. the current layer's output (lives in S.dxa at every layer boundary)
    float* dn = S.dxb;
    if (parts & CLIMB_BWD_TAIL) {
        // pooled = tanh(pre): dpre = dpooled (1 - pooled^2)
        TRY(act_bwd(dpooled, P.pooled, S.dpool_pre, static_cast<long long>(P.B) * d, CLIMB_EPI_TANH, s));
        if (tail) TRY(pwgrad(c, P.B, d, d, S.dpool_pre, P.cls_ln, G(grad, pr->pooler_w), G(grad, pr->pooler_b)));
        TRY(pdgrad(c, P.B, d, d, S.dpool_pre, pr->pooler_w, nullptr, S.dcls));
        if (lowest == P.layers) {
            TRY(layernorm_bwd(S.dcls, nullptr, P.x_final, static_cast<long long>(P.L) * d, F(theta, pr->final_ln_w),
                              F(theta, pr->final_ln_b), P.fmean, P.frstd, nullptr, nullptr, nullptr, G(grad, pr->final_ln_w),
                              G(grad, pr->final_ln_b), P.B, d, CLIMB_EPI_NONE, s));
            return 0;
        }
        CLIMB_CUDA_OK(cudaMemsetAsync(S.dxa, 0, sizeof(float) * M * d, s));
        TRY(layernorm_bwd(S.dcls, nullptr, P.x_final, static_cast<long long>(P.L) * d, F(theta, pr->final_ln_w), F(theta, pr->final_ln_b),
                          P.fmean, P.frstd, nullptr, S.dxa, nullptr, tail ? G(grad, pr->final_ln_w) : nullptr,
                          tail ? G(grad, pr->final_ln_b) : nullptr, P.B, d, CLIMB_EPI_NONE, s));
    }
    if (lowest == P.layers) return 0;

    for (int li = first_layer; li >= last_layer && li >= lowest; --li) {
        const climb_vilt_layer& w = pr->layer[li];
        const LayerAct& a = P.act[li];
        const bool base = (w.flags & CLIMB_TRAIN_BASE) != 0;
        const bool adp = (w.flags & CLIMB_TRAIN_ADAPTER) != 0;
        const bool mh_ad = r > 0 && w.mh_down_w >= 0;
        const bool out_ad = r > 0 && w.out_down_w >= 0;

        // ---- output adapter: out = y + up(act(down(y))) ----
        if (out_ad) {
            TRY(pdgrad(c, M, d, r, dx, w.out_up_w, nullptr, S.dz));                     // dz = dx Wu
            TRY(act_bwd(S.dz, a.out_pre, S.dpre, static_cast<long long>(M) * r, pr->adapter_act, s));
            if (adp) {
                TRY(pwgrad(c, M, d, r, dx, a.out_z, G(grad, w.out_up_w), G(grad, w.out_up_b)));
                TRY(pwgrad(c, M, r, d, S.dpre, a.out_in, G(grad, w.out_down_w), G(grad, w.out_down_b)));
            }
            TRY(pdgrad(c, M, r, d, S.dpre, w.out_down_w, dx, dn));                      // dy = dx + dpre Wd
            float* t = dx; dx = dn; dn = t;
        }
        // ---- FFN: y = FC2(GELU(FC1(LN2(x1)))) + x1 ----
        TRY(pdgrad(c, M, d, ff, dx, w.fc2_w, nullptr, S.dinter));
        TRY(act_bwd(S.dinter, a.u, S.du, static_cast<long long>(M) * ff, CLIMB_EPI_GELU, s));
        if (base) {
            TRY(pwgrad(c, M, d, ff, dx, a.inter, G(grad, w.fc2_w), G(grad, w.fc2_b)));
            TRY(pwgrad(c, M, ff, d, S.du, a.h2, G(grad, w.fc1_w), G(grad, w.fc1_b)));
        }
        TRY(pdgrad(c, M, ff, d, S.du, w.fc1_w, nullptr, S.dh));
        // dx1 = dx + LN2'(dh2)
        TRY(layernorm_bwd(S.dh, nullptr, a.x1, d, F(theta, w.ln2_w), F(theta, w.ln2_b), a.mean2, a.rstd2, dx, dn, nullptr,
                          base ? G(grad, w.ln2_w) : nullptr, base ? G(grad, w.ln2_b) : nullptr, M, d, CLIMB_EPI_NONE, s));
        // dn = dx1
        // ---- attention block: x1 = x + A, A = h (+ adapter), h = O(ctx) + b ----
        const float* dho = dn;
        if (mh_ad) {
            TRY(pdgrad(c, M, d, r, dn, w.mh_up_w, nullptr, S.dz));
            TRY(act_bwd(S.dz, a.mh_pre, S.dpre, static_cast<long long>(M) * r, pr->adapter_act, s));
            if (adp) {
                TRY(pwgrad(c, M, d, r, dn, a.mh_z, G(grad, w.mh_up_w), G(grad, w.mh_up_b)));
                TRY(pwgrad(c, M, r, d, S.dpre, a.mh_in, G(grad, w.mh_down_w), G(grad, w.mh_down_b)));
            }
            TRY(pdgrad(c, M, r, d, S.dpre, w.mh_down_w, dn, S.dmh));                    // dh = dx1 + dpre Wd
            dho = S.dmh;
        }
        TRY(pdgrad(c, M, d, d, dho, w.o_w, nullptr, S.dh));                              // dctx
        if (base) TRY(pwgrad(c, M, d, d, dho, a.ctx, G(grad, w.o_w), G(grad, w.o_b)));
        TRY(attn_f32_bwd(a.qkv, P.key_bias, a.ctx, S.dh, a.lse, S.delta, S.dqkv, P.B, P.L, P.heads, 0.125f, s));
        if (base) TRY(pwgrad(c, M, 3 * d, d, S.dqkv, a.h1, G(grad, w.qkv_w), G(grad, w.qkv_b)));
        TRY(pdgrad(c, M, 3 * d, d, S.dqkv, w.qkv_w, nullptr, S.dh));                      // dh1
        // dx_in = dx1 + LN1'(dh1), written over the old dx buffer
        TRY(layernorm_bwd(S.dh, nullptr, a.x_in, d, F(theta, w.ln1_w), F(theta, w.ln1_b), a.mean1, a.rstd1, dn, dx, nullptr,
                          base ? G(grad, w.ln1_w) : nullptr, base ? G(grad, w.ln1_b) : nullptr, M, d, CLIMB_EPI_NONE, s));
        // the layer-boundary gradient must live in S.dxa for the next (possibly separate) call
        if (dx != S.dxa) {
            CLIMB_CUDA_OK(cudaMemcpyAsync(S.dxa, dx, sizeof(float) * M * d, cudaMemcpyDeviceToDevice, s));
            dn = dx;
            dx = S.dxa;
        }
    }

    // ---- embeddings ----
    if ((parts & CLIMB_BWD_EMBED) && (pr->embed_flags & CLIMB_TRAIN_BASE)) {
        TRY(embed_split_bwd(dx, S.dy_text, S.dpatch_h, P.B, P.T, P.Np, d, s));
        patch_rows_kernel<<<blocks_for(static_cast<long long>(P.B) * P.Np * d, 256), 256, 0, s>>>(dx, S.dpatch, P.B, P.T, P.Np, d);
        CLIMB_LAUNCH_OK();
        TRY(embed_reduce_bwd(dx, bt->image_type_idx, bt->image_type_idx_scalar, S.S, G(grad, pr->cls_token), G(grad, pr->pos_emb),
                             G(grad, pr->mod_emb), G(grad, pr->patch_b), dm->n_modality, P.B, P.T, P.hp, P.wp, dm->pos_grid, d, s,
                             P.geom, P.geom ? P.Np : 0, bt->patch_select));
        TRY(layernorm_bwd(S.dy_text, nullptr, P.text_e, d, F(theta, pr->text_ln_w), F(theta, pr->text_ln_b), P.text_mean, P.text_rstd,
                          nullptr, S.de_text, nullptr, G(grad, pr->text_ln_w), G(grad, pr->text_ln_b), BT, d, CLIMB_EPI_NONE, s));
        TRY(text_scatter_bwd(S.de_text, reinterpret_cast<const long long*>(bt->input_ids),
                             reinterpret_cast<const long long*>(bt->token_type_ids), bt->input_ids ? G(grad, pr->word_emb) : nullptr,
                             G(grad, pr->text_type_emb), G(grad, pr->text_pos_emb), BT, P.T, d, s, dm->vocab_size, dm->type_vocab_size));
        // patch projection weight: dW[d, Kp] += dpatch^T im2col, with the im2col operand already split (hi / lo of the pixels)
        const int Mp = P.B * P.Np;
        TRY(split(S.dpatch, S.a_hi, S.a_lo, static_cast<long long>(Mp) * d, s));
        TRY(gemm_acc(d, P.Kp, Mp, S.a_hi, d, 1, P.im2col_hi, P.Kp, 1, G(grad, pr->patch_w), P.Kp, s));
        TRY(gemm_acc(d, P.Kp, Mp, S.a_lo, d, 1, P.im2col_hi, P.Kp, 1, G(grad, pr->patch_w), P.Kp, s));
        TRY(gemm_acc(d, P.Kp, Mp, S.a_hi, d, 1, P.im2col_lo, P.Kp, 1, G(grad, pr->patch_w), P.Kp, s));
    }
    return 0;
}

}  // namespace climb
