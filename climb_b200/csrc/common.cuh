// Shared device helpers for the climb_b200 sm_100a kernels: PTX wrappers for
// mbarrier / TMA / tcgen05, warp reductions, bf16 packing and the activation
// functions used by the GEMM epilogues.
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>

namespace climb {

#ifndef CLIMB_SPIN_GUARD
#define CLIMB_SPIN_GUARD 1   // trap instead of hanging forever on a lost mbarrier arrive
#endif

// ---------------------------------------------------------------------------------------------
// Error plumbing: every C-ABI entry returns 0 on success or a negative climb_status; the last
// message is kept per thread and read with climb_last_error().
// ---------------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);

// launch accounting (climb_launch_count) and the optional per-category device-time profiler
// (climb_profile_begin / climb_profile_end): a Scope brackets one launcher with two events on the
// launch stream when profiling is on and costs one branch when it is off.
extern unsigned long long g_launch_count;
enum ProfCategory { PROF_GEMM = 0, PROF_ATTN_FWD = 1, PROF_ATTN_BWD = 2, PROF_OTHER = 3, PROF_NUM = 4 };
struct ProfScope {
    int idx;
    cudaStream_t stream;
    ProfScope(int category, double work, cudaStream_t s);
    ~ProfScope();
};

#define CLIMB_CUDA_OK(expr)                                                                  \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            ::climb::set_last_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),  \
                                    __FILE__, __LINE__);                                     \
            return -2;                                                                       \
        }                                                                                    \
    } while (0)

#define CLIMB_REQUIRE(cond, ...)                                                             \
    do {                                                                                     \
        if (!(cond)) {                                                                       \
            ::climb::set_last_error(__VA_ARGS__);                                            \
            return -1;                                                                       \
        }                                                                                    \
    } while (0)

#define CLIMB_LAUNCH_OK()                                                                    \
    do {                                                                                     \
        ++::climb::g_launch_count;                                                           \
        cudaError_t _e = cudaGetLastError();                                                 \
        if (_e != cudaSuccess) {                                                             \
            ::climb::set_last_error("kernel launch failed: %s (%s:%d)",                      \
                                    cudaGetErrorString(_e), __FILE__, __LINE__);             \
            return -2;                                                                       \
        }                                                                                    \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Small device utilities
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(v);
}

// erf-GELU of ViLT (activations.py:37-56 -> nn.functional.gelu) and its derivative.
// Phi(x) = 0.5 (1 + erf(x / sqrt 2)) is evaluated with Abramowitz-Stegun 7.1.26
//   erfc(a) = (a1 t + a2 t^2 + a3 t^3 + a4 t^4 + a5 t^5) exp(-a^2),  t = 1 / (1 + p a),  |error| <= 1.5e-7
// on a = |x| / sqrt 2, using erfc directly for x < 0 so that the tail suffers no cancellation.
// The GEMM epilogues that apply it are issue-bound, so every instruction counts. exp(-a^2) is shared
// with the Gaussian density of the derivative.
__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// packed fp32 pairs (FFMA2 / FADD2 / FMUL2 of sm_100): one issue slot for two lanes of work
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ uint64_t pack_u32x2(uint32_t lo, uint32_t hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// 19 instructions per (gelu, gelu') pair: the single-instruction MUFU forms (no denormal fix-up code),
// 0.5 and 1/sqrt 2 folded into the constants, exp(-x^2/2) = 2^(-(x sqrt(log2 e / 2))^2).
__device__ __forceinline__ void phi_pdf(float x, float& cdf, float& pdf_times_sqrt2pi) {
    const float t = rcp_ftz(fmaf(0.3275911f * 0.70710678118654752440f, fabsf(x), 1.0f));
    float poly = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
    poly = fmaf(poly, t, 0.5f * 1.421413741f);
    poly = fmaf(poly, t, 0.5f * -0.284496736f);
    poly = fmaf(poly, t, 0.5f * 0.254829592f);
    const float s = x * 0.84932180028801904272f;
    const float e = ex2_ftz(-(s * s));                 // = exp(-x^2 / 2)
    const float half_erfc = (poly * t) * e;            // 0.5 * erfc(|x| / sqrt 2)
    cdf = x >= 0.0f ? 1.0f - half_erfc : half_erfc;
    pdf_times_sqrt2pi = e;
}
__device__ __forceinline__ float gelu_f(float x) {
    float cdf, e;
    phi_pdf(x, cdf, e);
    return x * cdf;
}
__device__ __forceinline__ float dgelu_f(float x) {
    float cdf, e;
    phi_pdf(x, cdf, e);
    return fmaf(x * 0.39894228040143267794f, e, cdf);
}
// swish / SiLU (Houlsby adapters, activations.py:154-168) and derivative
__device__ __forceinline__ float swish_f(float x) { return x / (1.0f + __expf(-x)); }
__device__ __forceinline__ float dswish_f(float x) {
    const float s = 1.0f / (1.0f + __expf(-x));
    return s * (1.0f + x * (1.0f - s));
}

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch: a kernel launched through launch_pdl may be scheduled while the
// previous kernel of the stream is still draining (its CTAs occupy the SMs that have gone idle in the
// predecessor's last wave and run their prologue); it must execute pdl_wait() before it reads or
// writes any global memory the predecessor touches. pdl_wait() is a no-op for ordinary launches.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();      // capi.cu: CLIMB_PDL=0 switches it off
// INDEPENDENT launches: a kernel that touches nothing the previous launch of its stream reads or writes may skip its
// pdl_wait() and run concurrently with that launch (in its last partial wave, or co-resident with its CTAs when the
// resources allow). The caller marks it with pdl_mark_independent() right before launch_pdl() and passes the kernel a
// flag that makes it skip the wait. Because such a kernel's completion no longer implies its predecessor's, the first
// ORDINARY launch after it is issued without the programmatic attribute = a full stream barrier.
void pdl_mark_independent();
bool pdl_take_independent();
void pdl_fence_next();
bool pdl_take_fence();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    const bool independent = pdl_take_independent();
    cfg.numAttrs = (pdl_enabled() && (independent || !pdl_take_fence())) ? 1 : 0;
    if (independent) pdl_fence_next();
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// The same for a kernel launched as thread-block clusters of `cluster_x` CTAs (CTA pairs for tcgen05 cta_group::2).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                      unsigned cluster_x, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster_x;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    const bool independent = pdl_take_independent();
    cfg.numAttrs = (pdl_enabled() && (independent || !pdl_take_fence())) ? 2 : 1;
    if (independent) pdl_fence_next();
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------------------------
// Counter-based random numbers (Philox4x32-10): dropout masks are a pure function of (seed, element
// index), so a backward pass -- or a test -- can regenerate them without storing anything.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32(unsigned long long seed, unsigned long long counter) {
    uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
    uint4 c = make_uint4(static_cast<uint32_t>(counter), static_cast<uint32_t>(counter >> 32), 0x636c696du, 0x62323030u);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}
// keep-scale of one element: 0 if dropped, 1/(1-p) if kept; r = 32 random bits, thresh = p * 2^32
__device__ __forceinline__ float dropout_scale(uint32_t r, uint32_t thresh, float inv_keep) {
    return r >= thresh ? inv_keep : 0.0f;
}
__host__ __device__ __forceinline__ uint32_t dropout_threshold(float p) {
    const double t = static_cast<double>(p) * 4294967296.0;
    return t >= 4294967295.0 ? 0xFFFFFFFFu : static_cast<uint32_t>(t);
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#if CLIMB_SPIN_GUARD
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = 0;
    for (uint32_t spins = 1; !mbar_try_wait(bar, parity); ++spins) {
        if ((spins & 255u) != 0) continue;              // keep the polling loop to two instructions
        if (t0 == 0) t0 = clock64();
        if (clock64() - t0 > 4000000000LL) {   // ~2 s at 2 GHz: a lost arrive, not a slow one
            printf("climb_b200: mbarrier wait timed out (block %d thread %d parity %u)\n",
                   blockIdx.x, threadIdx.x, parity);
            __trap();
        }
    }
#else
    while (!mbar_try_wait(bar, parity)) {}
#endif
}

// ---------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) 2-D tile load into shared memory, completion on an mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* desc) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const void* desc, uint64_t* bar, void* smem_dst,
                                            int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        :
        : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)),
          "r"(c0), "r"(c1)
        : "memory");
}

// 1-D bulk copy global -> shared (no tensor map): `bytes` a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :
                 : "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void tma_load_3d(const void* desc, uint64_t* bar, void* smem_dst,
                                            int32_t c0, int32_t c1, int32_t c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];"
        :
        : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)),
          "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------
// UMMA shared-memory descriptor, 128B swizzle (layout type 2), sm_100 version bit.
//   K-major : rows of 128 B, 8-row atoms 1024 B apart (SBO); LBO unused.
//   MN-major: 64-element (128 B) runs along MN, one row per k; 8 k-rows per 1024 B atom (SBO);
//             the next 64-wide MN chunk starts LBO bytes later.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;     // descriptor version (Blackwell)
    d |= 2ull << 61;     // SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor: bf16 x bf16 -> fp32, M x N tile, per-operand major bits
__device__ __forceinline__ uint32_t umma_instr_desc(int umma_m, int umma_n, int a_mn, int b_mn) {
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
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate. One thread issues.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        :
        : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on an mbarrier once all tcgen05 ops previously issued by this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::
                     "r"(smem_u32(bar))
                 : "memory");
}
// ---------------------------------------------------------------------------------------------
// CTA pairs (tcgen05 cta_group::2): the two CTAs of a 2-CTA cluster share one MMA. Each CTA stages its own
// 128 rows of A and its own HALF of the B tile; the leader (cluster rank 0) issues tcgen05.mma.cta_group::2 with
// M = 256 and each CTA's TMEM receives the accumulator rows of its own A half. Protocol as in CUTLASS's sm100 2-SM
// kernels (cute/arch/copy_sm100_tma.hpp, cutlass/arch/barrier.h, cutlass/pipeline/sm100_pipeline.hpp): TMA loads of
// BOTH CTAs complete on the LEADER's full barrier (one expect_tx for the bytes of both), tcgen05.commit multicasts the
// "slot free" / "accumulator ready" arrivals to both CTAs, epilogue warps release the accumulator on the leader's barrier.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (an address in THIS CTA's shared memory) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t cluster_map_shared(const void* p, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA tile load into THIS CTA's shared memory whose completion bytes are counted on an mbarrier given by its
// shared::cluster address (the leader CTA's barrier)
__device__ __forceinline__ void tma_load_2d_pair(const void* desc, uint32_t bar_cluster_addr, void* smem_dst,
                                                 int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        :
        : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar_cluster_addr),
          "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        :
        : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the mbarrier at this offset in BOTH CTAs of the pair once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::
                     "r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
                 : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives lane (base_lane + t).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
          "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// Legacy tensor path (mma.sync m16n8k16 bf16) + ldmatrix + cp.async: used by the attention
// kernels, whose tiles (L <= 256 keys, dh = 64) live entirely in one CTA's shared memory.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4],
                                               const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 "
        "{%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void cp_async_16(uint32_t smem_addr, const void* gptr, bool valid) {
    const int sz = valid ? 16 : 0;   // src-size 0 => zero-fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr), "l"(gptr),
                 "r"(sz)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

}  // namespace climb
