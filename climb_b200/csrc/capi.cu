// extern "C" surface of libclimb_b200.so (see include/climb_b200.h). Thin: argument plumbing and
// the per-thread error string only; the kernels live in the sibling translation units.
#include "common.cuh"
#include <cstdlib>
#include "internal.h"

#include <cstdarg>
#include <cstring>
#include <vector>

namespace climb {

static thread_local char g_last_error[1024] = "";

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

unsigned long long g_launch_count = 0;

// ---- sticky device-side error flags (climb_error_flags) -------------------------------------------
static unsigned int* g_err_host = nullptr;
static unsigned int* g_err_dev = nullptr;
unsigned int* device_error_word() {
    if (g_err_host == nullptr) {
        unsigned int* h = nullptr;
        if (cudaHostAlloc(reinterpret_cast<void**>(&h), sizeof(unsigned int), cudaHostAllocMapped) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        *h = 0u;
        unsigned int* d = nullptr;
        if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&d), h, 0) != cudaSuccess) {
            cudaGetLastError();
            cudaFreeHost(h);
            return nullptr;
        }
        g_err_host = h;
        g_err_dev = d;
    }
    return g_err_dev;
}

// ---- profiler ----------------------------------------------------------------------------------
namespace {
struct ProfRecord { cudaEvent_t a, b; int category; double work; };
bool g_prof_on = false;
std::vector<ProfRecord> g_prof;
}  // namespace

ProfScope::ProfScope(int category, double work, cudaStream_t s) : idx(-1), stream(s) {
    if (!g_prof_on) return;
    ProfRecord r;
    r.category = category;
    r.work = work;
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    cudaEventRecord(r.a, s);
    g_prof.push_back(r);
    idx = static_cast<int>(g_prof.size()) - 1;
}
ProfScope::~ProfScope() {
    if (idx >= 0) cudaEventRecord(g_prof[idx].b, stream);
}

}  // namespace climb

using namespace climb;

static inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

namespace climb {
bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("CLIMB_PDL"); return !(e && e[0] == '0'); }();
    return on;
}
static bool g_pdl_fence = false, g_pdl_independent = false;
void pdl_mark_independent() { g_pdl_independent = pdl_enabled(); }
bool pdl_take_independent() {
    const bool f = g_pdl_independent;
    g_pdl_independent = false;
    return f;
}
void pdl_fence_next() { g_pdl_fence = true; }
bool pdl_take_fence() {
    const bool f = g_pdl_fence;
    g_pdl_fence = false;
    return f;
}
}  // namespace climb

extern "C" {

const char* climb_last_error(void) { return g_last_error; }
int climb_version(void) { return 100; }
uint64_t climb_launch_count(void) { return g_launch_count; }
int climb_gemm_pair_mode(int mode) { return climb::gemm_pair_mode(mode); }
int climb_set_sm_reserve(int n) { return climb::sm_reserve(n); }
uint32_t climb_error_flags(void) {
    if (g_err_host == nullptr) return 0u;
    return __atomic_exchange_n(g_err_host, 0u, __ATOMIC_ACQ_REL);
}

int climb_profile_begin(void) {
    for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof.clear();
    g_prof_on = true;
    return 0;
}
int climb_profile_end(double* ms, double* work, int64_t* launches, int n_categories) {
    g_prof_on = false;
    CLIMB_REQUIRE(ms && work && launches && n_categories >= PROF_NUM, "climb_profile_end: need %d categories", PROF_NUM);
    CLIMB_CUDA_OK(cudaDeviceSynchronize());
    for (int i = 0; i < n_categories; ++i) { ms[i] = 0.0; work[i] = 0.0; launches[i] = 0; }
    for (auto& r : g_prof) {
        float t = 0.0f;
        if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
            ms[r.category] += t;
            work[r.category] += r.work;
            launches[r.category] += 1;
        }
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    g_prof.clear();
    return 0;
}

int climb_gemm_bf16(const climb_gemm_desc* desc, void* stream) { return gemm_bf16(desc, S(stream)); }

int climb_adapter_fused(int backward, int M, int d, int r, int act, const void* a_bf16, const void* w_down_bf16, const void* w_up_bf16,
                        const float* b_down, const float* b_up, void* pre_bf16, void* z_bf16, const float* c_in, float* c_out,
                        void* c2_bf16, float* colsum_z, void* stream) {
    return adapter_fused(backward, M, d, r, act, a_bf16, w_down_bf16, w_up_bf16, b_down, b_up, pre_bf16, z_bf16, c_in, c_out, c2_bf16,
                         colsum_z, S(stream));
}

int climb_attention_fwd(const void* qkv, const float* key_bias, void* ctx, float* lse, int B, int L,
                        int H, float scale, void* stream) {
    return attention_fwd(qkv, key_bias, ctx, lse, B, L, H, scale, S(stream));
}
int climb_attention_fwd_dropout(const void* qkv, const float* key_bias, void* ctx, float* lse, int B, int L, int H,
                                float scale, float p, uint64_t seed, void* stream) {
    CLIMB_REQUIRE(qkv && ctx && lse && B > 0 && L > 0 && H > 0, "attention_fwd_dropout: bad arguments");
    return attention_tc_fwd(qkv, key_bias, ctx, lse, B, L, H, scale, S(stream), p, seed);
}
int climb_attention_bwd(const void* qkv, const float* key_bias, const void* ctx, const void* dctx,
                        const float* lse, float* delta, void* dqkv, float* dqkv_colsum, int B, int L, int H,
                        float scale, void* stream) {
    return attention_bwd(qkv, key_bias, ctx, dctx, lse, delta, dqkv, dqkv_colsum, B, L, H, scale, S(stream));
}

int climb_layernorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps,
                        void* y_bf16, float* y_f32, float* mean, float* rstd, int rows, int d, int act,
                        void* stream) {
    return layernorm_fwd(x, ldx, gamma, beta, eps, y_bf16, y_f32, mean, rstd, rows, d, act, S(stream));
}
int climb_layernorm_bwd(const float* dy_f32, const void* dy_bf16, const float* x, int64_t ldx,
                        const float* gamma, const float* beta, const float* mean, const float* rstd,
                        const float* dres, float* dx_f32, void* dx_bf16, float* dgamma, float* dbeta,
                        int rows, int d, int act, void* stream) {
    return layernorm_bwd(dy_f32, dy_bf16, x, ldx, gamma, beta, mean, rstd, dres, dx_f32, dx_bf16,
                         dgamma, dbeta, rows, d, act, S(stream));
}
int climb_layernorm_bwd_colsum(const float* dy_f32, const void* dy_bf16, const float* x, int64_t ldx,
                               const float* gamma, const float* beta, const float* mean, const float* rstd,
                               const float* dres, float* dx_f32, void* dx_bf16, float* dgamma, float* dbeta,
                               float* dx_colsum, int rows, int d, int act, void* stream) {
    return layernorm_bwd(dy_f32, dy_bf16, x, ldx, gamma, beta, mean, rstd, dres, dx_f32, dx_bf16, dgamma, dbeta, rows, d,
                         act, S(stream), dx_colsum);
}


int climb_cast_f32_bf16(const float* src, void* dst_bf16, int64_t n, void* stream) {
    return cast_f32_bf16(src, dst_bf16, n, S(stream));
}
int climb_split_f32_bf16x2(const float* src, void* hi_bf16, void* lo_bf16, int64_t n, void* stream) {
    return split_f32_bf16x2(src, hi_bf16, lo_bf16, n, S(stream));
}
int climb_colsum(const void* src, int dtype, int64_t ld, int rows, int cols, float* out, void* stream) {
    return colsum(src, dtype, ld, rows, cols, out, S(stream));
}
int climb_bce_logits_loss(const float* logits, int64_t ld, const float* target, int rows, int cols, float scale,
                          float grad_scale, float* row_loss, float* loss, float* dlogits, int64_t ldd, void* stream) {
    return bce_logits_loss(logits, ld, target, rows, cols, scale, grad_scale, row_loss, loss, dlogits, ldd, S(stream));
}
int climb_cross_entropy_loss(const float* logits, int64_t ld, const int64_t* target, int rows, int cols,
                             float grad_scale, float* row_loss, float* loss, float* dlogits, int64_t ldd,
                             void* stream) {
    return cross_entropy_loss(logits, ld, reinterpret_cast<const long long*>(target), rows, cols, grad_scale, row_loss,
                              loss, dlogits, ldd, S(stream));
}
int climb_ewc_penalty(const float* theta, const float* theta_star, const float* fisher, int64_t n, float lambda,
                      float* partials, int n_partials, float* loss, float* grad, float grad_scale,
                      const float* grad_scale_dev, void* stream) {
    return ewc_penalty(theta, theta_star, fisher, n, lambda, partials, n_partials, loss, grad, grad_scale,
                       grad_scale_dev, S(stream));
}
int climb_fisher_accumulate(const float* grad, float* fisher, int64_t n, void* stream) {
    return fisher_accumulate(grad, fisher, n, S(stream));
}
int climb_scale_inplace(float* x, int64_t n, float s, void* stream) { return scale_inplace(x, n, s, S(stream)); }
int climb_adamw_step(float* theta, const float* grad, float* exp_avg, float* exp_avg_sq, void* shadow_bf16,
                     const climb_adamw_chunk* chunks_dev, int n_chunks, const float* group_lr_host,
                     const float* group_wd_host, int n_groups, float beta1, float beta2, float eps, int step,
                     void* stream) {
    return adamw_step(theta, grad, exp_avg, exp_avg_sq, shadow_bf16, chunks_dev, n_chunks, group_lr_host, group_wd_host, n_groups,
                      beta1, beta2, eps, step, S(stream));
}
int64_t climb_vilt_forward_workspace_bytes(const climb_vilt_dims* dims, const climb_vilt_params* params,
                                           const climb_vilt_batch* batch, int save_for_backward) {
    return vilt_forward_workspace_bytes(dims, params, batch, save_for_backward);
}
int64_t climb_vilt_backward_scratch_bytes(const climb_vilt_dims* dims, const climb_vilt_params* params,
                                          const climb_vilt_batch* batch) {
    return vilt_backward_scratch_bytes(dims, params, batch);
}
int climb_vilt_forward(const climb_vilt_dims* dims, const climb_vilt_params* params, const climb_vilt_batch* batch,
                       const float* theta, const void* shadow, void* workspace, int64_t workspace_bytes,
                       int save_for_backward, float* pooled_out, void* stream) {
    return vilt_forward(dims, params, batch, theta, shadow, workspace, workspace_bytes, save_for_backward, pooled_out,
                        S(stream));
}
int climb_vilt_backward(const climb_vilt_dims* dims, const climb_vilt_params* params, const climb_vilt_batch* batch,
                        const float* theta, const void* shadow, const void* workspace, int64_t workspace_bytes,
                        void* scratch, int64_t scratch_bytes, const float* dpooled, float* grad, int first_layer,
                        int last_layer, int parts, void* stream) {
    return vilt_backward(dims, params, batch, theta, shadow, workspace, workspace_bytes, scratch, scratch_bytes, dpooled,
                         grad, first_layer, last_layer, parts, S(stream));
}

int64_t climb_bert_forward_workspace_bytes(const climb_bert_dims* dims, const climb_bert_batch* batch) {
    return bert_forward_workspace_bytes(dims, batch);
}
int climb_bert_forward(const climb_bert_dims* dims, const climb_bert_params* params, const climb_bert_batch* batch,
                       const float* theta, const void* shadow, void* workspace, int64_t workspace_bytes,
                       float hidden_dropout, float attn_dropout, uint64_t seed, float* last_hidden_state, void* stream) {
    return bert_forward(dims, params, batch, theta, shadow, workspace, workspace_bytes, hidden_dropout, attn_dropout, seed,
                        last_hidden_state, S(stream));
}
uint64_t climb_dropout_site_seed(uint64_t base, int layer, int site) { return dropout_site_seed(base, layer, site); }
int climb_dropout_keep_mask(float* out, int64_t n, float p, uint64_t seed, void* stream) {
    return dropout_mask_f32(nullptr, out, n, p, seed, S(stream));
}
int climb_attention_dropout_keep_mask(float* out, int B, int H, int L, float p, uint64_t seed, void* stream) {
    return attn_dropout_mask(out, B, H, L, p, seed, S(stream));
}
int climb_dropout_add(const float* x, const float* res, float* y, int64_t n, float p, uint64_t seed, void* stream) {
    return dropout_add(x, res, y, n, p, seed, S(stream));
}

int climb_image_preprocess(const uint8_t* src, uint8_t* tmp, const climb_image_desc* descs, const int32_t* tables, int B,
                           int64_t max_tmp_pixels, float* pixel_values, int64_t* pixel_mask, int Hp, int Wp, const float* mean,
                           const float* std, void* stream) {
    return image_preprocess(src, tmp, descs, tables, B, max_tmp_pixels, pixel_values, reinterpret_cast<long long*>(pixel_mask), Hp, Wp,
                            mean, std, S(stream));
}

}  // extern "C"
