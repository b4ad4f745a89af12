// HBM-bound helper kernels of the ViLT hot path: parameter shadow cast (fp32 master -> bf16 tensor
// core operand), column sums (bias gradients of every Linear: autograd of modeling_vilt.py:356-360,
// 409,464,482), the additive attention-mask row (modeling_utils.py:299-311), tanh backward of the
// pooler (modeling_vilt.py:887-899), and the loss kernels of the trainers (train_vqa.py:95,157;
// train_nlvr2.py:80,133). All global accesses are 128-bit where alignment allows.
#include "common.cuh"
#include <cstdlib>
#include "climb_b200.h"

namespace climb {
namespace {

__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                     long long n) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 8;
    for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
        if (i + 8 <= n) {
            const float4 a = *reinterpret_cast<const float4*>(src + i);
            const float4 b = *reinterpret_cast<const float4*>(src + i + 4);
            uint4 o;
            o.x = pack_bf16(a.x, a.y); o.y = pack_bf16(a.z, a.w);
            o.z = pack_bf16(b.x, b.y); o.w = pack_bf16(b.z, b.w);
            *reinterpret_cast<uint4*>(dst + i) = o;
        } else {
            for (long long j = i; j < n; ++j) dst[j] = __float2bfloat16_rn(src[j]);
        }
    }
}

// out[c] += sum_r src[r, c]; 8 warps of a CTA take interleaved rows of one row-chunk, a lane owns
// VEC consecutive columns; partials meet in shared memory, then one atomic per column per CTA.
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ src, long long ld, int rows, int cols, int rows_per_cta,
              float* __restrict__ out) {
    constexpr int VEC = 16 / sizeof(T);            // 8 bf16 or 4 fp32 per 128-bit load
    __shared__ float s_part[8][32 * VEC];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c0 = (blockIdx.x * 32 + lane) * VEC;
    const int r_begin = blockIdx.y * rows_per_cta;
    const int r_end = min(rows, r_begin + rows_per_cta);
    pdl_wait();
    float acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = 0.0f;
    if (c0 < cols) {
        auto add = [&](const uint4& u) {
            if constexpr (sizeof(T) == 2) {
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = unpack_bf16(w[j]);
                    acc[2 * j] += f.x;
                    acc[2 * j + 1] += f.y;
                }
            } else {
                acc[0] += __uint_as_float(u.x); acc[1] += __uint_as_float(u.y);
                acc[2] += __uint_as_float(u.z); acc[3] += __uint_as_float(u.w);
            }
        };
        const T* base = src + c0;
        int r = r_begin + warp;
        // four independent 128-bit loads in flight per lane (the kernel is a latency-bound streaming pass)
        for (; r + 24 < r_end; r += 32) {
            const uint4 u0 = *reinterpret_cast<const uint4*>(base + static_cast<long long>(r) * ld);
            const uint4 u1 = *reinterpret_cast<const uint4*>(base + static_cast<long long>(r + 8) * ld);
            const uint4 u2 = *reinterpret_cast<const uint4*>(base + static_cast<long long>(r + 16) * ld);
            const uint4 u3 = *reinterpret_cast<const uint4*>(base + static_cast<long long>(r + 24) * ld);
            add(u0); add(u1); add(u2); add(u3);
        }
        for (; r < r_end; r += 8) add(*reinterpret_cast<const uint4*>(base + static_cast<long long>(r) * ld));
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) s_part[warp][lane * VEC + j] = acc[j];
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * VEC; i += blockDim.x) {
        const int c = blockIdx.x * 32 * VEC + i;
        if (c < cols) {
            float t = 0.0f;
#pragma unroll
            for (int w = 0; w < 8; ++w) t += s_part[w][i];
            atomicAdd(out + c, t);
        }
    }
}

// unaligned / tiny case (task-head logits with N = 3129): one thread per column
template <typename T>
__global__ void colsum_scalar_kernel(const T* __restrict__ src, long long ld, int rows, int cols,
                                     float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    float acc = 0.0f;
    for (int r = 0; r < rows; ++r) {
        if constexpr (sizeof(T) == 2) acc += __bfloat162float(src[static_cast<long long>(r) * ld + c]);
        else acc += src[static_cast<long long>(r) * ld + c];
    }
    out[c] += acc;
}

__global__ void key_bias_kernel(const long long* __restrict__ mask, float* __restrict__ out, int B,
                                int T, int L) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * L) return;
    const int b = i / L, j = i - b * L;
    // image tokens are always attended on the fixed-resolution path (pixel_mask all ones)
    out[i] = j < T ? (1.0f - static_cast<float>(mask[static_cast<long long>(b) * T + j])) * -10000.0f : 0.0f;
}

__global__ void tanh_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                __nv_bfloat16* __restrict__ out, long long n) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __float2bfloat16_rn(dy[i] * (1.0f - y[i] * y[i]));
}

// BCEWithLogits(mean) * scale and its gradient; one CTA per row, deterministic two-level sum
__global__ void __launch_bounds__(256)
bce_logits_kernel(const float* __restrict__ logits, long long ld, const float* __restrict__ target,
                  int rows, int cols, float scale, float grad_scale, float* __restrict__ row_loss,
                  float* __restrict__ dlogits, long long ldd) {
    __shared__ float s_warp[8];
    const int r = blockIdx.x;
    const float* x = logits + static_cast<long long>(r) * ld;
    const float* t = target + static_cast<long long>(r) * cols;
    float acc = 0.0f;
    const float inv = 1.0f / (static_cast<float>(rows) * static_cast<float>(cols));
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
        const float xv = x[c], tv = t[c];
        // max(x,0) - x*t + log(1 + exp(-|x|))   (the numerically stable form torch uses)
        acc += fmaxf(xv, 0.0f) - xv * tv + log1pf(expf(-fabsf(xv)));
        if (dlogits) {
            const float sig = 1.0f / (1.0f + expf(-xv));
            dlogits[static_cast<long long>(r) * ldd + c] = (sig - tv) * inv * scale * grad_scale;
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tsum = 0.0f;
        for (int w = 0; w < 8; ++w) tsum += s_warp[w];
        row_loss[r] = tsum * inv * scale;
    }
}

// CrossEntropy(mean) and its gradient; one warp per row (num_labels is 2..4 on this path)
__global__ void ce_kernel(const float* __restrict__ logits, long long ld, const long long* __restrict__ target,
                          int rows, int cols, float grad_scale, float* __restrict__ row_loss,
                          float* __restrict__ dlogits, long long ldd) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const float* x = logits + static_cast<long long>(r) * ld;
    float mx = -INFINITY;
    for (int c = lane; c < cols; c += 32) mx = fmaxf(mx, x[c]);
    mx = warp_max(mx);
    float se = 0.0f;
    for (int c = lane; c < cols; c += 32) se += expf(x[c] - mx);
    se = warp_sum(se);
    const float lse = mx + logf(se);
    const int tgt = static_cast<int>(target[r]);
    if (lane == 0) row_loss[r] = (lse - x[tgt]) / static_cast<float>(rows);
    if (dlogits) {
        for (int c = lane; c < cols; c += 32) {
            const float p = expf(x[c] - lse);
            dlogits[static_cast<long long>(r) * ldd + c] = (p - (c == tgt ? 1.0f : 0.0f)) * grad_scale / static_cast<float>(rows);
        }
    }
}

__global__ void sum_rows_kernel(const float* __restrict__ v, int n, float* __restrict__ out) {
    // single CTA, fixed order: deterministic
    __shared__ float s[256];
    float acc = 0.0f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += v[i];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = s[0];
}

}  // namespace

int cast_f32_bf16(const float* src, void* dst, long long n, cudaStream_t stream) {
    CLIMB_REQUIRE(src && dst && n > 0, "cast_f32_bf16: bad arguments");
    CLIMB_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
                  "cast_f32_bf16: buffers must be 16-byte aligned");
    long long blocks = (n / 8 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    cast_f32_bf16_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(src, static_cast<__nv_bfloat16*>(dst), n);
    CLIMB_LAUNCH_OK();
    return 0;
}

int colsum(const void* src, int dtype, long long ld, int rows, int cols, float* out, cudaStream_t stream) {
    CLIMB_REQUIRE(src && out && rows > 0 && cols > 0, "colsum: bad arguments");
    const int esz = dtype == CLIMB_F32 ? 4 : 2;
    const int vec = 16 / esz;
    const bool aligned = (cols % vec == 0) && ((ld * esz) % 16 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    if (!aligned) {
        const int threads = 128, blocks = (cols + threads - 1) / threads;
        if (dtype == CLIMB_F32)
            colsum_scalar_kernel<float><<<blocks, threads, 0, stream>>>(static_cast<const float*>(src), ld, rows, cols, out);
        else
            colsum_scalar_kernel<__nv_bfloat16><<<blocks, threads, 0, stream>>>(static_cast<const __nv_bfloat16*>(src), ld, rows, cols, out);
        CLIMB_LAUNCH_OK();
        return 0;
    }
    const int gx = (cols + 32 * vec - 1) / (32 * vec);
    int gy = (148 * 4 + gx - 1) / gx;                 // enough CTAs to fill the machine
    int rows_per_cta = (rows + gy - 1) / gy;
    if (rows_per_cta < 64) rows_per_cta = 64;
    gy = (rows + rows_per_cta - 1) / rows_per_cta;
    dim3 grid(gx, gy);
    if (dtype == CLIMB_F32)
        CLIMB_CUDA_OK(launch_pdl(colsum_kernel<float>, grid, dim3(256), 0, stream, static_cast<const float*>(src), ld, rows, cols,
                                 rows_per_cta, out));
    else
        CLIMB_CUDA_OK(launch_pdl(colsum_kernel<__nv_bfloat16>, grid, dim3(256), 0, stream, static_cast<const __nv_bfloat16*>(src), ld,
                                 rows, cols, rows_per_cta, out));
    CLIMB_LAUNCH_OK();
    return 0;
}

int key_bias(const long long* mask, float* out, int B, int T, int L, cudaStream_t stream) {
    CLIMB_REQUIRE(mask && out && B > 0 && T > 0 && L >= T, "key_bias: bad arguments");
    const int n = B * L;
    key_bias_kernel<<<(n + 255) / 256, 256, 0, stream>>>(mask, out, B, T, L);
    CLIMB_LAUNCH_OK();
    return 0;
}

int tanh_bwd(const float* dy, const float* y, void* out_bf16, long long n, cudaStream_t stream) {
    CLIMB_REQUIRE(dy && y && out_bf16 && n > 0, "tanh_bwd: bad arguments");
    tanh_bwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(dy, y, static_cast<__nv_bfloat16*>(out_bf16), n);
    CLIMB_LAUNCH_OK();
    return 0;
}

int bce_logits_loss(const float* logits, long long ld, const float* target, int rows, int cols, float scale,
                    float grad_scale, float* row_loss, float* loss, float* dlogits, long long ldd,
                    cudaStream_t stream) {
    CLIMB_REQUIRE(logits && target && row_loss && loss && rows > 0 && cols > 0, "bce_logits_loss: bad arguments");
    bce_logits_kernel<<<rows, 256, 0, stream>>>(logits, ld, target, rows, cols, scale, grad_scale, row_loss, dlogits, ldd);
    CLIMB_LAUNCH_OK();
    sum_rows_kernel<<<1, 256, 0, stream>>>(row_loss, rows, loss);
    CLIMB_LAUNCH_OK();
    return 0;
}

int cross_entropy_loss(const float* logits, long long ld, const long long* target, int rows, int cols,
                       float grad_scale, float* row_loss, float* loss, float* dlogits, long long ldd,
                       cudaStream_t stream) {
    CLIMB_REQUIRE(logits && target && row_loss && loss && rows > 0 && cols > 0, "cross_entropy_loss: bad arguments");
    ce_kernel<<<(rows + 3) / 4, 128, 0, stream>>>(logits, ld, target, rows, cols, grad_scale, row_loss, dlogits, ldd);
    CLIMB_LAUNCH_OK();
    sum_rows_kernel<<<1, 256, 0, stream>>>(row_loss, rows, loss);
    CLIMB_LAUNCH_OK();
    return 0;
}

}  // namespace climb
