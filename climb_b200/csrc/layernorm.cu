// LayerNorm forward / backward for the ViLT hot path: layernorm_before / layernorm_after / final
// layernorm (modeling_vilt.py:505,517,873, eps 1e-12), the text-embedding LayerNorm (:302) and the
// task-head LayerNorm(1536)+GELU (src/modeling/vilt.py:190-197, eps 1e-5).
//
// HBM-bound: one warp owns one row, the row stays in registers (d/128 float4 per lane) between the
// statistics and the normalisation, so x is read once; all global accesses are 128-bit and
// coalesced; row statistics use warp shuffles. The backward fuses the residual-gradient add and the
// bf16 copy that feeds the next tensor-core GEMM, and reduces dgamma/dbeta per CTA before touching
// global memory with one atomic per column per CTA.
#include "common.cuh"
#include "climb_b200.h"

#include <cstdlib>

namespace climb {
namespace {

constexpr int kWarpsPerBlock = 8;
// dev A/B switch: CLIMB_LN_NO_BULK=1 keeps the register-only backward kernel
const bool g_ln_no_bulk = [] { const char* e = getenv("CLIMB_LN_NO_BULK"); return e && e[0] == '1'; }();

template <int NV>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
ln_fwd_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ gamma,
              const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ y_bf16,
              float* __restrict__ y_f32, float* __restrict__ mean_out, float* __restrict__ rstd_out,
              int rows, int act) {
    constexpr int D = NV * 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * kWarpsPerBlock + warp;
    pdl_wait();
    if (row >= rows) return;
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * ldx);
    float4 v[NV];
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i] = xr[lane + 32 * i];
        sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(sum) * (1.0f / D);
    float sq = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        sq += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
    if (lane == 0) {
        if (mean_out) mean_out[row] = mean;
        if (rstd_out) rstd_out[row] = rstd;
    }
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float4 g = __ldg(g4 + lane + 32 * i), b = __ldg(b4 + lane + 32 * i);
        float4 o;
        o.x = (v[i].x - mean) * rstd * g.x + b.x;
        o.y = (v[i].y - mean) * rstd * g.y + b.y;
        o.z = (v[i].z - mean) * rstd * g.z + b.z;
        o.w = (v[i].w - mean) * rstd * g.w + b.w;
        if (act == CLIMB_EPI_GELU) { o.x = gelu_f(o.x); o.y = gelu_f(o.y); o.z = gelu_f(o.z); o.w = gelu_f(o.w); }
        if (y_f32)
            reinterpret_cast<float4*>(y_f32 + static_cast<long long>(row) * D)[lane + 32 * i] = o;
        if (y_bf16) {
            uint2 p;
            p.x = pack_bf16(o.x, o.y);
            p.y = pack_bf16(o.z, o.w);
            reinterpret_cast<uint2*>(y_bf16 + static_cast<long long>(row) * D)[lane + 32 * i] = p;
        }
    }
}

template <int NV, bool COLSUM>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 2)
ln_bwd_kernel(const float* __restrict__ dy_f32, const __nv_bfloat16* __restrict__ dy_bf16,
              const float* __restrict__ x, long long ldx, const float* __restrict__ gamma,
              const float* __restrict__ beta, const float* __restrict__ mean_in,
              const float* __restrict__ rstd_in, const float* __restrict__ dres,
              float* __restrict__ dx_f32, __nv_bfloat16* __restrict__ dx_bf16,
              float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dx_colsum, int rows, int act) {
    constexpr int D = NV * 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
    pdl_wait();

    // Only the column accumulators stay in registers across the row: x and dy are read twice (the second
    // read hits L1/L2 -- a row is 4.5 KB) so that two CTAs fit per SM and more loads are in flight.
    // dx_colsum (COLSUM): column sums of the OUTPUT dx (after the residual-gradient add) = the bias gradient of
    // the Linear that produced this LayerNorm's input row stream (fc2 / attention-output dense): fused here
    // instead of another 23 MB pass over dx. Its per-warp accumulators live in shared memory (a third register
    // set would spill under the 128-register budget of two resident CTAs).
    __shared__ float4 s_dc[COLSUM ? kWarpsPerBlock : 1][COLSUM ? NV * 32 : 1];
    float4 dg[NV], db[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        dg[i] = db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (COLSUM) s_dc[warp][lane + 32 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    auto load_dy = [&](int row, int i, const float4& xh, const float4& g) -> float4 {
        float4 dy;
        if (dy_f32) {
            dy = reinterpret_cast<const float4*>(dy_f32 + static_cast<long long>(row) * D)[lane + 32 * i];
        } else {
            const uint2 p = reinterpret_cast<const uint2*>(dy_bf16 + static_cast<long long>(row) * D)[lane + 32 * i];
            const float2 a = unpack_bf16(p.x), b = unpack_bf16(p.y);
            dy = make_float4(a.x, a.y, b.x, b.y);
        }
        if (act == CLIMB_EPI_GELU) {
            const float4 b = __ldg(b4 + lane + 32 * i);
            dy.x *= dgelu_f(xh.x * g.x + b.x); dy.y *= dgelu_f(xh.y * g.y + b.y);
            dy.z *= dgelu_f(xh.z * g.z + b.z); dy.w *= dgelu_f(xh.w * g.w + b.w);
        }
        return dy;
    };

    for (int row = blockIdx.x * kWarpsPerBlock + warp; row < rows; row += gridDim.x * kWarpsPerBlock) {
        const float mean = mean_in[row], rstd = rstd_in[row];
        const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * ldx);
        float c1 = 0.0f, c2 = 0.0f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const float4 xv = xr[lane + 32 * i];
            const float4 g = __ldg(g4 + lane + 32 * i);
            float4 xh;
            xh.x = (xv.x - mean) * rstd; xh.y = (xv.y - mean) * rstd;
            xh.z = (xv.z - mean) * rstd; xh.w = (xv.w - mean) * rstd;
            const float4 dy = load_dy(row, i, xh, g);
            dg[i].x += dy.x * xh.x; dg[i].y += dy.y * xh.y; dg[i].z += dy.z * xh.z; dg[i].w += dy.w * xh.w;
            db[i].x += dy.x; db[i].y += dy.y; db[i].z += dy.z; db[i].w += dy.w;
            const float gx = dy.x * g.x, gy = dy.y * g.y, gz = dy.z * g.z, gw = dy.w * g.w;
            c1 += (gx + gy) + (gz + gw);
            c2 += (gx * xh.x + gy * xh.y) + (gz * xh.z + gw * xh.w);
        }
        c1 = warp_sum(c1) * (1.0f / D);
        c2 = warp_sum(c2) * (1.0f / D);
        if (dx_f32 == nullptr && dx_bf16 == nullptr) continue;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const float4 xv = xr[lane + 32 * i];
            const float4 g = __ldg(g4 + lane + 32 * i);
            float4 xh;
            xh.x = (xv.x - mean) * rstd; xh.y = (xv.y - mean) * rstd;
            xh.z = (xv.z - mean) * rstd; xh.w = (xv.w - mean) * rstd;
            const float4 dy = load_dy(row, i, xh, g);
            float4 o;
            o.x = rstd * (dy.x * g.x - c1 - xh.x * c2);
            o.y = rstd * (dy.y * g.y - c1 - xh.y * c2);
            o.z = rstd * (dy.z * g.z - c1 - xh.z * c2);
            o.w = rstd * (dy.w * g.w - c1 - xh.w * c2);
            if (dres) {
                const float4 r = reinterpret_cast<const float4*>(dres + static_cast<long long>(row) * ldx)[lane + 32 * i];
                o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
            }
            if (COLSUM) {
                float4 acc = s_dc[warp][lane + 32 * i];
                acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
                s_dc[warp][lane + 32 * i] = acc;
            }
            if (dx_f32)
                reinterpret_cast<float4*>(dx_f32 + static_cast<long long>(row) * ldx)[lane + 32 * i] = o;
            if (dx_bf16) {
                uint2 p;
                p.x = pack_bf16(o.x, o.y);
                p.y = pack_bf16(o.z, o.w);
                reinterpret_cast<uint2*>(dx_bf16 + static_cast<long long>(row) * D)[lane + 32 * i] = p;
            }
        }
    }

    if (COLSUM) {
        __syncthreads();
        const float* flat = reinterpret_cast<const float*>(&s_dc[0][0]);
        for (int c = threadIdx.x; c < D; c += kWarpsPerBlock * 32) {
            float acc = 0.0f;
#pragma unroll
            for (int w = 0; w < kWarpsPerBlock; ++w) acc += flat[w * D + c];
            atomicAdd(dx_colsum + c, acc);
        }
    }
    if (dgamma == nullptr && dbeta == nullptr) return;
    // CTA reduction of the per-warp column partials: shared-memory float atomics into one [2][D] accumulator
    // (two barriers in total), then one global atomic per column per CTA
    __shared__ float s_acc[2][D];
    for (int c = threadIdx.x; c < 2 * D; c += kWarpsPerBlock * 32) (&s_acc[0][0])[c] = 0.0f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c0 = (lane + 32 * i) * 4;
        if (dgamma) {
            atomicAdd(&s_acc[0][c0], dg[i].x); atomicAdd(&s_acc[0][c0 + 1], dg[i].y);
            atomicAdd(&s_acc[0][c0 + 2], dg[i].z); atomicAdd(&s_acc[0][c0 + 3], dg[i].w);
        }
        if (dbeta) {
            atomicAdd(&s_acc[1][c0], db[i].x); atomicAdd(&s_acc[1][c0 + 1], db[i].y);
            atomicAdd(&s_acc[1][c0 + 2], db[i].z); atomicAdd(&s_acc[1][c0 + 3], db[i].w);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += kWarpsPerBlock * 32) {
        if (dgamma) atomicAdd(dgamma + c, s_acc[0][c]);
        if (dbeta) atomicAdd(dbeta + c, s_acc[1][c]);
    }
}

// ---------------------------------------------------------------------------------------------
// Backward, streaming variant for the encoder's hot calls (dense rows, bf16 dy, residual-gradient add, both
// outputs): every warp keeps a two-deep ring of whole rows (x, dy, dres = 7.5 KB at d = 768) in shared memory,
// filled by cp.async.bulk (one lane issues three bulk copies per row against an mbarrier), so 120 KB of loads
// per SM are in flight independently of the registers; the row is read once from shared memory, the two
// reductions and the output pass run from registers. (The register-only kernel above re-reads x and dy and has
// ~3-5 KB in flight per warp: 4.1 TB/s; it remains the path for every other shape / option.)
// ---------------------------------------------------------------------------------------------
template <int NV>
struct LnBulk {
    static constexpr int D = NV * 128;
    static constexpr int kWarps = 8;
    static constexpr int kStages = 2;
    static constexpr int kRowBytes = D * 4 + D * 2 + D * 4;            // x fp32 | dy bf16 | dres fp32
    static constexpr int kSmem = kWarps * kStages * kRowBytes + kWarps * kStages * 8;
};

template <int NV>
__global__ void __launch_bounds__(LnBulk<NV>::kWarps * 32, 1)
ln_bwd_bulk_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                   const float* __restrict__ mean_in, const float* __restrict__ rstd_in, const float* __restrict__ dres,
                   float* __restrict__ dx_f32, __nv_bfloat16* __restrict__ dx_bf16, float* __restrict__ dgamma,
                   float* __restrict__ dbeta, int rows) {
    using C = LnBulk<NV>;
    constexpr int D = C::D;
    extern __shared__ __align__(128) uint8_t ln_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* ring = ln_smem + warp * (C::kStages * C::kRowBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(ln_smem + C::kWarps * C::kStages * C::kRowBytes) + warp * C::kStages;
    if (lane == 0) {
        for (int s = 0; s < C::kStages; ++s) mbar_init(&bars[s], 1);
        fence_barrier_init();
    }
    __syncwarp();
    pdl_wait();
    const int stride = gridDim.x * C::kWarps;
    const int row0 = blockIdx.x * C::kWarps + warp;
    auto issue = [&](int row, int s) {          // lane 0 only
        uint8_t* dst = ring + s * C::kRowBytes;
        mbar_arrive_expect_tx(&bars[s], C::kRowBytes);
        bulk_load_1d(dst, x + static_cast<long long>(row) * D, D * 4, &bars[s]);
        bulk_load_1d(dst + D * 4, dy + static_cast<long long>(row) * D, D * 2, &bars[s]);
        bulk_load_1d(dst + D * 6, dres + static_cast<long long>(row) * D, D * 4, &bars[s]);
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < C::kStages; ++s)
            if (row0 + s * stride < rows) issue(row0 + s * stride, s);
    }
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    float4 gam[NV], dg[NV], db[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        gam[i] = __ldg(g4 + lane + 32 * i);
        dg[i] = db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    int it = 0;
    for (int row = row0; row < rows; row += stride, ++it) {
        const int s = it % C::kStages;
        const uint32_t parity = (it / C::kStages) & 1;
        const float mean = mean_in[row], rstd = rstd_in[row];
        mbar_wait(&bars[s], parity);
        const uint8_t* buf = ring + s * C::kRowBytes;
        float4 xh[NV], dyv[NV], rs[NV];
        float c1 = 0.0f, c2 = 0.0f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const float4 xv = reinterpret_cast<const float4*>(buf)[lane + 32 * i];
            const uint2 p = reinterpret_cast<const uint2*>(buf + D * 4)[lane + 32 * i];
            rs[i] = reinterpret_cast<const float4*>(buf + D * 6)[lane + 32 * i];
            const float2 a = unpack_bf16(p.x), b = unpack_bf16(p.y);
            dyv[i] = make_float4(a.x, a.y, b.x, b.y);
            xh[i].x = (xv.x - mean) * rstd; xh[i].y = (xv.y - mean) * rstd;
            xh[i].z = (xv.z - mean) * rstd; xh[i].w = (xv.w - mean) * rstd;
            dg[i].x += dyv[i].x * xh[i].x; dg[i].y += dyv[i].y * xh[i].y; dg[i].z += dyv[i].z * xh[i].z; dg[i].w += dyv[i].w * xh[i].w;
            db[i].x += dyv[i].x; db[i].y += dyv[i].y; db[i].z += dyv[i].z; db[i].w += dyv[i].w;
            dyv[i].x *= gam[i].x; dyv[i].y *= gam[i].y; dyv[i].z *= gam[i].z; dyv[i].w *= gam[i].w;      // dy * gamma from here on
            c1 += (dyv[i].x + dyv[i].y) + (dyv[i].z + dyv[i].w);
            c2 += (dyv[i].x * xh[i].x + dyv[i].y * xh[i].y) + (dyv[i].z * xh[i].z + dyv[i].w * xh[i].w);
        }
        // the stage has been read into registers: refill it with the row two steps ahead
        __syncwarp();
        if (lane == 0 && row + C::kStages * stride < rows) issue(row + C::kStages * stride, s);
        c1 = warp_sum(c1) * (1.0f / D);
        c2 = warp_sum(c2) * (1.0f / D);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float4 o;
            o.x = rstd * (dyv[i].x - c1 - xh[i].x * c2) + rs[i].x;
            o.y = rstd * (dyv[i].y - c1 - xh[i].y * c2) + rs[i].y;
            o.z = rstd * (dyv[i].z - c1 - xh[i].z * c2) + rs[i].z;
            o.w = rstd * (dyv[i].w - c1 - xh[i].w * c2) + rs[i].w;
            reinterpret_cast<float4*>(dx_f32 + static_cast<long long>(row) * D)[lane + 32 * i] = o;
            uint2 pk;
            pk.x = pack_bf16(o.x, o.y);
            pk.y = pack_bf16(o.z, o.w);
            reinterpret_cast<uint2*>(dx_bf16 + static_cast<long long>(row) * D)[lane + 32 * i] = pk;
        }
    }
    if (dgamma == nullptr && dbeta == nullptr) return;
    // CTA reduction of the column partials through the (now idle) ring of warp 0..: [2][D] floats
    __syncthreads();
    float* s_acc = reinterpret_cast<float*>(ln_smem);
    for (int c = threadIdx.x; c < 2 * D; c += C::kWarps * 32) s_acc[c] = 0.0f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c0 = (lane + 32 * i) * 4;
        atomicAdd(&s_acc[c0], dg[i].x); atomicAdd(&s_acc[c0 + 1], dg[i].y);
        atomicAdd(&s_acc[c0 + 2], dg[i].z); atomicAdd(&s_acc[c0 + 3], dg[i].w);
        atomicAdd(&s_acc[D + c0], db[i].x); atomicAdd(&s_acc[D + c0 + 1], db[i].y);
        atomicAdd(&s_acc[D + c0 + 2], db[i].z); atomicAdd(&s_acc[D + c0 + 3], db[i].w);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += C::kWarps * 32) {
        if (dgamma) atomicAdd(dgamma + c, s_acc[c]);
        if (dbeta) atomicAdd(dbeta + c, s_acc[D + c]);
    }
}

// Forward, streaming variant for the encoder's hot calls (dense rows, bf16 output + statistics): same per-warp
// cp.async.bulk row ring as the backward (three stages of one 3 KB row, two CTAs per SM).
template <int NV>
struct LnFwdBulk {
    static constexpr int D = NV * 128;
    static constexpr int kWarps = 8;
    static constexpr int kStages = 3;
    static constexpr int kRowBytes = D * 4;
    static constexpr int kSmem = kWarps * kStages * kRowBytes + kWarps * kStages * 8;
};

template <int NV>
__global__ void __launch_bounds__(LnFwdBulk<NV>::kWarps * 32, 2)
ln_fwd_bulk_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                   __nv_bfloat16* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows) {
    using C = LnFwdBulk<NV>;
    constexpr int D = C::D;
    extern __shared__ __align__(128) uint8_t ln_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* ring = ln_smem + warp * (C::kStages * C::kRowBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(ln_smem + C::kWarps * C::kStages * C::kRowBytes) + warp * C::kStages;
    if (lane == 0) {
        for (int s = 0; s < C::kStages; ++s) mbar_init(&bars[s], 1);
        fence_barrier_init();
    }
    __syncwarp();
    pdl_wait();
    const int stride = gridDim.x * C::kWarps;
    const int row0 = blockIdx.x * C::kWarps + warp;
    auto issue = [&](int row, int s) {
        mbar_arrive_expect_tx(&bars[s], C::kRowBytes);
        bulk_load_1d(ring + s * C::kRowBytes, x + static_cast<long long>(row) * D, C::kRowBytes, &bars[s]);
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < C::kStages; ++s)
            if (row0 + s * stride < rows) issue(row0 + s * stride, s);
    }
    float4 gam[NV], bet[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        gam[i] = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
        bet[i] = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * i);
    }
    int it = 0;
    for (int row = row0; row < rows; row += stride, ++it) {
        const int s = it % C::kStages;
        mbar_wait(&bars[s], (it / C::kStages) & 1);
        const float4* buf = reinterpret_cast<const float4*>(ring + s * C::kRowBytes);
        float4 v[NV];
        float sum = 0.0f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            v[i] = buf[lane + 32 * i];
            sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
        __syncwarp();
        if (lane == 0 && row + C::kStages * stride < rows) issue(row + C::kStages * stride, s);
        const float mean = warp_sum(sum) * (1.0f / D);
        float sq = 0.0f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
            sq += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
        }
        const float rstd = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
        if (lane == 0) {
            if (mean_out) mean_out[row] = mean;
            if (rstd_out) rstd_out[row] = rstd;
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            uint2 p;
            p.x = pack_bf16(v[i].x * rstd * gam[i].x + bet[i].x, v[i].y * rstd * gam[i].y + bet[i].y);
            p.y = pack_bf16(v[i].z * rstd * gam[i].z + bet[i].z, v[i].w * rstd * gam[i].w + bet[i].w);
            reinterpret_cast<uint2*>(y + static_cast<long long>(row) * D)[lane + 32 * i] = p;
        }
    }
}

template <int NV>
int launch_fwd(const float* x, long long ldx, const float* gamma, const float* beta, float eps,
               void* y_bf16, float* y_f32, float* mean, float* rstd, int rows, int act,
               cudaStream_t stream) {
    if constexpr (NV == 6) {
        if (y_bf16 != nullptr && y_f32 == nullptr && act == CLIMB_EPI_NONE && ldx == NV * 128 && rows >= 1024 &&
            (reinterpret_cast<uintptr_t>(x) & 15) == 0 && !g_ln_no_bulk) {
            using C = LnFwdBulk<NV>;
            static bool attr = false;
            if (!attr) {
                CLIMB_CUDA_OK(cudaFuncSetAttribute(ln_fwd_bulk_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
                attr = true;
            }
            int g2 = (rows + C::kWarps - 1) / C::kWarps;
            if (g2 > 148 * 2) g2 = 148 * 2;
            CLIMB_CUDA_OK(launch_pdl(ln_fwd_bulk_kernel<NV>, dim3(g2), dim3(C::kWarps * 32), static_cast<size_t>(C::kSmem), stream, x, gamma,
                                     beta, eps, static_cast<__nv_bfloat16*>(y_bf16), mean, rstd, rows));
            CLIMB_LAUNCH_OK();
            return 0;
        }
    }
    const int grid = (rows + kWarpsPerBlock - 1) / kWarpsPerBlock;
    CLIMB_CUDA_OK(launch_pdl(ln_fwd_kernel<NV>, dim3(grid), dim3(kWarpsPerBlock * 32), 0, stream, x, ldx, gamma, beta, eps,
                             static_cast<__nv_bfloat16*>(y_bf16), y_f32, mean, rstd, rows, act));
    CLIMB_LAUNCH_OK();
    return 0;
}

template <int NV>
int launch_bwd(const float* dy_f32, const void* dy_bf16, const float* x, long long ldx,
               const float* gamma, const float* beta, const float* mean, const float* rstd,
               const float* dres, float* dx_f32, void* dx_bf16, float* dgamma, float* dbeta, float* dx_colsum, int rows,
               int act, cudaStream_t stream) {
    // the encoder's hot calls: dense fp32 rows, bf16 dy, residual-gradient add, both outputs, no fused activation
    if constexpr (NV == 6) {
        const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy_bf16) | reinterpret_cast<uintptr_t>(dres)) & 15) == 0;
        if (dy_bf16 != nullptr && dy_f32 == nullptr && dres != nullptr && dx_f32 != nullptr && dx_bf16 != nullptr && dx_colsum == nullptr &&
            act == CLIMB_EPI_NONE && ldx == NV * 128 && rows >= 1024 && aligned && !g_ln_no_bulk) {
            using C = LnBulk<NV>;
            static bool attr = false;
            if (!attr) {
                CLIMB_CUDA_OK(cudaFuncSetAttribute(ln_bwd_bulk_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
                attr = true;
            }
            int grid = (rows + C::kWarps - 1) / C::kWarps;
            if (grid > 148) grid = 148;
            CLIMB_CUDA_OK(launch_pdl(ln_bwd_bulk_kernel<NV>, dim3(grid), dim3(C::kWarps * 32), static_cast<size_t>(C::kSmem), stream,
                                     static_cast<const __nv_bfloat16*>(dy_bf16), x, gamma, mean, rstd, dres, dx_f32,
                                     static_cast<__nv_bfloat16*>(dx_bf16), dgamma, dbeta, rows));
            CLIMB_LAUNCH_OK();
            return 0;
        }
    }
    int grid = (rows + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const int cap = 148 * 4;          // two resident CTAs per SM, two rounds (measured: 45.8 us vs 50.1 us with one round of 296)
    if (grid > cap) grid = cap;
    if (dx_colsum != nullptr) {
        if constexpr (NV <= 8) {
            CLIMB_CUDA_OK(launch_pdl(ln_bwd_kernel<NV, true>, dim3(grid), dim3(kWarpsPerBlock * 32), 0, stream, dy_f32,
                                     static_cast<const __nv_bfloat16*>(dy_bf16), x, ldx, gamma, beta, mean, rstd, dres, dx_f32,
                                     static_cast<__nv_bfloat16*>(dx_bf16), dgamma, dbeta, dx_colsum, rows, act));
        } else {
            CLIMB_REQUIRE(false, "layernorm_bwd: fused dx_colsum supports d <= 1024");
        }
    } else {
        CLIMB_CUDA_OK(launch_pdl(ln_bwd_kernel<NV, false>, dim3(grid), dim3(kWarpsPerBlock * 32), 0, stream, dy_f32,
                                 static_cast<const __nv_bfloat16*>(dy_bf16), x, ldx, gamma, beta, mean, rstd, dres, dx_f32,
                                 static_cast<__nv_bfloat16*>(dx_bf16), dgamma, dbeta, dx_colsum, rows, act));
    }
    CLIMB_LAUNCH_OK();
    return 0;
}

}  // namespace

int layernorm_fwd(const float* x, long long ldx, const float* gamma, const float* beta, float eps,
                  void* y_bf16, float* y_f32, float* mean, float* rstd, int rows, int d, int act,
                  cudaStream_t stream) {
    CLIMB_REQUIRE(x && gamma && beta, "layernorm_fwd: null pointer");
    CLIMB_REQUIRE(rows > 0, "layernorm_fwd: no rows");
    CLIMB_REQUIRE(d % 128 == 0 && ldx % 4 == 0, "layernorm_fwd: d=%d / ldx=%lld must be multiples of 128 / 4", d, ldx);
    CLIMB_REQUIRE(act == CLIMB_EPI_NONE || act == CLIMB_EPI_GELU, "layernorm_fwd: act must be NONE or GELU");
#define ARGS (x, ldx, gamma, beta, eps, y_bf16, y_f32, mean, rstd, rows, act, stream)
    switch (d / 128) {
        case 1: return launch_fwd<1> ARGS;
        case 2: return launch_fwd<2> ARGS;
        case 3: return launch_fwd<3> ARGS;
        case 4: return launch_fwd<4> ARGS;
        case 6: return launch_fwd<6> ARGS;
        case 8: return launch_fwd<8> ARGS;
        case 12: return launch_fwd<12> ARGS;
        default: break;
    }
#undef ARGS
    CLIMB_REQUIRE(false, "layernorm_fwd: unsupported width d=%d (128,256,384,512,768,1024,1536)", d);
    return -1;
}

int layernorm_bwd(const float* dy_f32, const void* dy_bf16, const float* x, long long ldx,
                  const float* gamma, const float* beta, const float* mean, const float* rstd,
                  const float* dres, float* dx_f32, void* dx_bf16, float* dgamma, float* dbeta,
                  int rows, int d, int act, cudaStream_t stream, float* dx_colsum) {
    CLIMB_REQUIRE(dx_colsum == nullptr || dx_f32 != nullptr || dx_bf16 != nullptr, "layernorm_bwd: dx_colsum needs a dx output");
    CLIMB_REQUIRE((dy_f32 != nullptr) != (dy_bf16 != nullptr), "layernorm_bwd: exactly one of dy_f32 / dy_bf16");
    CLIMB_REQUIRE(x && gamma && beta && mean && rstd, "layernorm_bwd: null pointer");
    CLIMB_REQUIRE(rows > 0, "layernorm_bwd: no rows");
    CLIMB_REQUIRE(d % 128 == 0 && ldx % 4 == 0, "layernorm_bwd: d=%d / ldx=%lld must be multiples of 128 / 4", d, ldx);
    CLIMB_REQUIRE(act == CLIMB_EPI_NONE || act == CLIMB_EPI_GELU, "layernorm_bwd: act must be NONE or GELU");
#define ARGS (dy_f32, dy_bf16, x, ldx, gamma, beta, mean, rstd, dres, dx_f32, dx_bf16, dgamma, dbeta, dx_colsum, rows, act, stream)
    switch (d / 128) {
        case 1: return launch_bwd<1> ARGS;
        case 2: return launch_bwd<2> ARGS;
        case 3: return launch_bwd<3> ARGS;
        case 4: return launch_bwd<4> ARGS;
        case 6: return launch_bwd<6> ARGS;
        case 8: return launch_bwd<8> ARGS;
        case 12: return launch_bwd<12> ARGS;
        default: break;
    }
#undef ARGS
    CLIMB_REQUIRE(false, "layernorm_bwd: unsupported width d=%d (128,256,384,512,768,1024,1536)", d);
    return -1;
}

}  // namespace climb
