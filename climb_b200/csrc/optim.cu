// Flat-arena kernels for the continual-learning algorithms and the optimizer:
//   * EWC penalty  lambda * sum F (theta - theta*)^2 and its gradient in ONE pass over device-resident
//     theta*, F (src/cl_algorithms/ewc.py:75-87 re-uploads both from the CPU every step and runs
//     ~4 ATen ops per tensor);
//   * Fisher accumulation F += g^2 on the device (ewc.py:61-64 does pow(2).cpu() per tensor per batch);
//   * AdamW over the flat parameter arena with per-segment weight decay / lr, matching
//     torch.optim.AdamW as configured by ViltContinualLearner.create_optimizer
//     (src/modeling/vilt.py:205-215: betas (0.9, 0.98), eps 1e-8, wd 1e-2 or 0).
// HBM-bound streaming kernels: 128-bit accesses, grid-stride, a few CTAs per SM.
#include "common.cuh"
#include "climb_b200.h"

namespace climb {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float block_sum(float v, float* s_warp) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.0f;
    if (threadIdx.x == 0)
        for (int w = 0; w < kThreads / 32; ++w) t += s_warp[w];
    return t;   // valid in thread 0
}

__global__ void __launch_bounds__(kThreads)
ewc_penalty_kernel(const float* __restrict__ theta, const float* __restrict__ theta_star,
                   const float* __restrict__ fisher, long long n, float* __restrict__ partials,
                   float* __restrict__ grad, float grad_coef, const float* __restrict__ grad_scale_dev) {
    __shared__ float s_warp[kThreads / 32];
    if (grad != nullptr && grad_scale_dev != nullptr) grad_coef *= *grad_scale_dev;
    float acc = 0.0f;
    const long long n4 = n / 4;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 t = reinterpret_cast<const float4*>(theta)[i];
        const float4 s = reinterpret_cast<const float4*>(theta_star)[i];
        const float4 f = reinterpret_cast<const float4*>(fisher)[i];
        const float dx = t.x - s.x, dy = t.y - s.y, dz = t.z - s.z, dw = t.w - s.w;
        acc += (f.x * dx * dx + f.y * dy * dy) + (f.z * dz * dz + f.w * dw * dw);
        if (grad) {
            float4 g = reinterpret_cast<float4*>(grad)[i];
            g.x += grad_coef * f.x * dx; g.y += grad_coef * f.y * dy;
            g.z += grad_coef * f.z * dz; g.w += grad_coef * f.w * dw;
            reinterpret_cast<float4*>(grad)[i] = g;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < n - n4 * 4) {        // tail (< 4 elements)
        const long long i = n4 * 4 + threadIdx.x;
        const float dx = theta[i] - theta_star[i];
        acc += fisher[i] * dx * dx;
        if (grad) grad[i] += grad_coef * fisher[i] * dx;
    }
    const float t = block_sum(acc, s_warp);
    if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

__global__ void finish_sum_kernel(const float* __restrict__ partials, int n, float scale, float* __restrict__ out) {
    __shared__ double s[kThreads];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += static_cast<double>(partials[i]);
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = kThreads / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = static_cast<float>(s[0] * scale);
}

__global__ void fisher_accumulate_kernel(const float* __restrict__ grad, float* __restrict__ fisher, long long n) {
    const long long n4 = n / 4;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 g = reinterpret_cast<const float4*>(grad)[i];
        float4 f = reinterpret_cast<float4*>(fisher)[i];
        f.x += g.x * g.x; f.y += g.y * g.y; f.z += g.z * g.z; f.w += g.w * g.w;
        reinterpret_cast<float4*>(fisher)[i] = f;
    }
    if (blockIdx.x == 0 && threadIdx.x < n - n4 * 4) {
        const long long i = n4 * 4 + threadIdx.x;
        fisher[i] += grad[i] * grad[i];
    }
}

__global__ void scale_kernel(float* __restrict__ x, long long n, float s) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) x[i] *= s;
}

struct GroupHyper { float lr[CLIMB_ADAMW_MAX_GROUPS]; float wd[CLIMB_ADAMW_MAX_GROUPS]; };

// One CTA per chunk of one tensor; chunks of a tensor share its param group's (lr, weight_decay).
// 28 B/param of HBM traffic (theta, grad, m, v read; theta, m, v written) + 2 B/param when the bf16
// shadow of the updated parameter is written in the same pass (saves the separate 6 B/param cast).
__global__ void __launch_bounds__(kThreads)
adamw_kernel(float* __restrict__ theta, const float* __restrict__ grad, float* __restrict__ m,
             float* __restrict__ v, __nv_bfloat16* __restrict__ shadow,
             const climb_adamw_chunk* __restrict__ chunks, const GroupHyper hp,
             float beta1, float beta2, float eps, float bc1, float bc2_sqrt) {
    const climb_adamw_chunk ch = chunks[blockIdx.x];
    const float lr = hp.lr[ch.group];
    const float decay = 1.0f - lr * hp.wd[ch.group];
    const float step = lr / bc1;
    const float inv_bc2 = 1.0f / bc2_sqrt;
    const float ob1 = 1.0f - beta1, ob2 = 1.0f - beta2;
    auto upd = [&](float& t, float g, float& mi, float& vi) {
        mi = beta1 * mi + ob1 * g;
        vi = beta2 * vi + ob2 * g * g;
        t = t * decay - step * (mi / (sqrtf(vi) * inv_bc2 + eps));
    };
    const bool vec = (ch.start & 3) == 0;
    const long long n4 = vec ? ch.length / 4 : 0;
    for (long long j = threadIdx.x; j < n4; j += blockDim.x) {
        const long long i = ch.start / 4 + j;
        float4 t = reinterpret_cast<float4*>(theta)[i];
        const float4 g = reinterpret_cast<const float4*>(grad)[i];
        float4 mi = reinterpret_cast<float4*>(m)[i];
        float4 vi = reinterpret_cast<float4*>(v)[i];
        upd(t.x, g.x, mi.x, vi.x); upd(t.y, g.y, mi.y, vi.y); upd(t.z, g.z, mi.z, vi.z); upd(t.w, g.w, mi.w, vi.w);
        reinterpret_cast<float4*>(theta)[i] = t;
        reinterpret_cast<float4*>(m)[i] = mi;
        reinterpret_cast<float4*>(v)[i] = vi;
        if (shadow) {
            uint2 o;
            o.x = pack_bf16(t.x, t.y);
            o.y = pack_bf16(t.z, t.w);
            reinterpret_cast<uint2*>(shadow)[i] = o;
        }
    }
    for (long long j = n4 * 4 + threadIdx.x; j < ch.length; j += blockDim.x) {
        const long long i = ch.start + j;
        float t = theta[i], mi = m[i], vi = v[i];
        upd(t, grad[i], mi, vi);
        theta[i] = t; m[i] = mi; v[i] = vi;
        if (shadow) shadow[i] = __float2bfloat16_rn(t);
    }
}

int grid_for(long long n) {
    long long b = (n / 4 + kThreads - 1) / kThreads;
    if (b > 148 * 8) b = 148 * 8;
    if (b < 1) b = 1;
    return static_cast<int>(b);
}

}  // namespace

int ewc_penalty(const float* theta, const float* theta_star, const float* fisher, long long n, float lambda,
                float* partials, int n_partials, float* loss, float* grad, float grad_scale,
                const float* grad_scale_dev, cudaStream_t stream) {
    CLIMB_REQUIRE(theta && theta_star && fisher && partials && loss && n > 0, "ewc_penalty: bad arguments");
    int grid = grid_for(n);
    if (grid > n_partials) grid = n_partials;
    CLIMB_REQUIRE(grid > 0, "ewc_penalty: partials buffer too small");
    ewc_penalty_kernel<<<grid, kThreads, 0, stream>>>(theta, theta_star, fisher, n, partials, grad,
                                                      2.0f * lambda * grad_scale, grad_scale_dev);
    CLIMB_LAUNCH_OK();
    finish_sum_kernel<<<1, kThreads, 0, stream>>>(partials, grid, lambda, loss);
    CLIMB_LAUNCH_OK();
    return 0;
}

int fisher_accumulate(const float* grad, float* fisher, long long n, cudaStream_t stream) {
    CLIMB_REQUIRE(grad && fisher && n > 0, "fisher_accumulate: bad arguments");
    fisher_accumulate_kernel<<<grid_for(n), kThreads, 0, stream>>>(grad, fisher, n);
    CLIMB_LAUNCH_OK();
    return 0;
}

int scale_inplace(float* x, long long n, float s, cudaStream_t stream) {
    CLIMB_REQUIRE(x && n > 0, "scale_inplace: bad arguments");
    scale_kernel<<<grid_for(n), kThreads, 0, stream>>>(x, n, s);
    CLIMB_LAUNCH_OK();
    return 0;
}

int adamw_step(float* theta, const float* grad, float* m, float* v, void* shadow_bf16, const climb_adamw_chunk* chunks_dev,
               int n_chunks, const float* group_lr, const float* group_wd, int n_groups, float beta1, float beta2,
               float eps, int step, cudaStream_t stream) {
    CLIMB_REQUIRE(theta && grad && m && v && chunks_dev && n_chunks > 0 && step > 0, "adamw_step: bad arguments");
    CLIMB_REQUIRE(group_lr && group_wd && n_groups > 0 && n_groups <= CLIMB_ADAMW_MAX_GROUPS,
                  "adamw_step: between 1 and %d param groups", CLIMB_ADAMW_MAX_GROUPS);
    GroupHyper hp;
    for (int i = 0; i < CLIMB_ADAMW_MAX_GROUPS; ++i) {
        hp.lr[i] = i < n_groups ? group_lr[i] : 0.0f;
        hp.wd[i] = i < n_groups ? group_wd[i] : 0.0f;
    }
    const float bc1 = 1.0f - powf(beta1, static_cast<float>(step));
    const float bc2 = 1.0f - powf(beta2, static_cast<float>(step));
    adamw_kernel<<<n_chunks, kThreads, 0, stream>>>(theta, grad, m, v, static_cast<__nv_bfloat16*>(shadow_bf16), chunks_dev, hp, beta1, beta2, eps, bc1, sqrtf(bc2));
    CLIMB_LAUNCH_OK();
    return 0;
}

}  // namespace climb
