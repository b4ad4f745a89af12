// Throughput probe (dev tool): MUFU.EX2 vs a packed-FMA polynomial exp2 on sm_100a, elements per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/mufu_probe tools/probe/mufu_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2_ftz(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_nf(float x) { float y; asm volatile("ex2.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// 2^x for x <= 0 on the FMA pipe: round-to-nearest split x = n + f, |f| <= 0.5, degree-3 polynomial, exponent add
__device__ __forceinline__ void exp2_poly2(float x0, float x1, float& y0, float& y1) {
    x0 = fmaxf(x0, -126.0f);
    x1 = fmaxf(x1, -126.0f);
    const uint64_t magic = pk(12582912.0f, 12582912.0f), nmagic = pk(-12582912.0f, -12582912.0f);
    const uint64_t x = pk(x0, x1);
    const uint64_t t = fadd2(x, magic);
    const uint64_t nf = fadd2(t, nmagic);
    const uint64_t f = fadd2(x, nf ^ 0x8000000080000000ull);
    uint64_t p = ffma2(pk(0.0555041f, 0.0555041f), f, pk(0.2402265f, 0.2402265f));
    p = ffma2(p, f, pk(0.6931472f, 0.6931472f));
    p = ffma2(p, f, pk(1.0f, 1.0f));
    float p0, p1, t0, t1;
    upk(p, p0, p1);
    upk(t, t0, t1);
    y0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
    y1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) probe(float* out, long long* cycles, int iters, float seed) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = -seed * (threadIdx.x + i + 1) * 1e-3f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = ex2_ftz(v[i]) - 1.5f;
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = ex2_nf(v[i]) - 1.5f;
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; i += 2) { float a, b; exp2_poly2(v[i], v[i + 1], a, b); v[i] = a - 1.5f; v[i + 1] = b - 1.5f; }
        } else if (MODE == 3) {       // half MUFU, half polynomial
#pragma unroll
            for (int i = 0; i < 4; i += 2) { float a, b; exp2_poly2(v[i], v[i + 1], a, b); v[i] = a - 1.5f; v[i + 1] = b - 1.5f; }
#pragma unroll
            for (int i = 4; i < 8; ++i) v[i] = ex2_ftz(v[i]) - 1.5f;
        } else if (MODE == 5) {       // cvt.rn.bf16x2.f32 only (4 packs of 2 per 8 values) + FADD
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                uint32_t u;
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(v[i + 1]), "f"(v[i]));
                v[i] = __uint_as_float(u << 16) - 1.5f;
                v[i + 1] = __uint_as_float(u & 0xffff0000u) - 1.5f;
            }
        } else if (MODE == 6) {       // MUFU + one cvt per pair (the softmax inner loop)
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                const float a = ex2_ftz(v[i]), b = ex2_ftz(v[i + 1]);
                uint32_t u;
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(b), "f"(a));
                v[i] = __uint_as_float(u << 16) - 1.5f;
                v[i + 1] = __uint_as_float(u & 0xffff0000u) - 1.5f;
            }
        } else if (MODE == 7) {       // MUFU + truncating pack through PRMT
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                const float a = ex2_ftz(v[i]), b = ex2_ftz(v[i + 1]);
                const uint32_t u = __byte_perm(__float_as_uint(a), __float_as_uint(b), 0x7632);
                v[i] = __uint_as_float(u << 16) - 1.5f;
                v[i + 1] = __uint_as_float(u & 0xffff0000u) - 1.5f;
            }
        } else if (MODE == 4) {       // 1/4 MUFU, 3/4 polynomial
#pragma unroll
            for (int i = 0; i < 6; i += 2) { float a, b; exp2_poly2(v[i], v[i + 1], a, b); v[i] = a - 1.5f; v[i + 1] = b - 1.5f; }
#pragma unroll
            for (int i = 6; i < 8; ++i) v[i] = ex2_ftz(v[i]) - 1.5f;
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 4096;
    const char* names[8] = {"MUFU ex2.approx.ftz (+FADD)", "MUFU ex2.approx (+FADD)", "poly deg3 packed (+FADD)", "1/2 MUFU + 1/2 poly", "1/4 MUFU + 3/4 poly", "cvt.rn.bf16x2 only (values/clk)", "MUFU + cvt.rn.bf16x2 per pair", "MUFU + PRMT pack per pair"};
    for (int mode = 0; mode < 8; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            switch (mode) {
                case 0: probe<0><<<148, 512>>>(out, cyc, iters, 1.0f); break;
                case 1: probe<1><<<148, 512>>>(out, cyc, iters, 1.0f); break;
                case 2: probe<2><<<148, 512>>>(out, cyc, iters, 1.0f); break;
                case 3: probe<3><<<148, 512>>>(out, cyc, iters, 1.0f); break;
                case 4: probe<4><<<148, 512>>>(out, cyc, iters, 1.0f); break;
                case 5: probe<5><<<148, 512>>>(out, cyc, iters, 1.0f); break;
                case 6: probe<6><<<148, 512>>>(out, cyc, iters, 1.0f); break;
                case 7: probe<7><<<148, 512>>>(out, cyc, iters, 1.0f); break;
            }
            cudaDeviceSynchronize();
        }
        long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < 148; ++i) avg += h[i];
        avg /= 148;
        printf("%-32s %.2f exp2 / clk / SM  (%.0f cycles)\n", names[mode], 512.0 * 8 * iters / avg, avg);
    }
    // accuracy of the polynomial
    return cudaGetLastError() != cudaSuccess;
}
