// extern "C" surface of libclimb_b200.so (see include/climb_b200.h). Thin: argument plumbing and
// the per-thread error string only; the kernels live in the sibling translation units.
#include "common.cuh"
#include "internal.h"

#include <cstdarg>
#include <cstring>

namespace climb {

static thread_local char g_last_error[1024] = "";

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

}  // namespace climb

using namespace climb;

static inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

extern "C" {

const char* climb_last_error(void) { return g_last_error; }
int climb_version(void) { return 100; }

int climb_gemm_bf16(const climb_gemm_desc* desc, void* stream) { return gemm_bf16(desc, S(stream)); }

int climb_attention_fwd(const void* qkv, const float* key_bias, void* ctx, float* lse, int B, int L,
                        int H, float scale, void* stream) {
    return attention_fwd(qkv, key_bias, ctx, lse, B, L, H, scale, S(stream));
}
int climb_attention_bwd(const void* qkv, const float* key_bias, const void* ctx, const void* dctx,
                        const float* lse, float* delta, void* dqkv, int B, int L, int H, float scale,
                        void* stream) {
    return attention_bwd(qkv, key_bias, ctx, dctx, lse, delta, dqkv, B, L, H, scale, S(stream));
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

}  // extern "C"
