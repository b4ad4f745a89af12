/* climb_b200 -- C ABI of the B200-native ViLT hot path (libclimb_b200.so).
 *
 * The reference (GLAMOR-USC/CLiMB) has no FFI on this path: its boundary is a Python registry plus
 * a duck-typed nn.Module (src/modeling/__init__.py:4-12, src/modeling/vilt.py:111-124,205-239).
 * This header is the native surface that the Python mirror (climb_b200/modeling) binds with ctypes;
 * every entry point names the reference code whose arithmetic it replaces.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in _host; no torch types cross the ABI
 *   - `stream` is a cudaStream_t passed as void*; every call is stream-ordered and never syncs
 *   - return value: 0 = ok, <0 = error (climb_last_error() returns the message of the last
 *     failure on the calling thread)
 *   - "rows" are tokens: a batch of B sequences of L = T text + Li image tokens is a row-major
 *     [B*L, d] matrix; bf16 tensors feed the tensor cores, fp32 carries the residual stream,
 *     statistics, parameters and gradients.
 */
#ifndef CLIMB_B200_H
#define CLIMB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { CLIMB_BF16 = 0, CLIMB_F32 = 1 } climb_dtype;

typedef enum {
    CLIMB_EPI_NONE = 0,
    CLIMB_EPI_GELU = 1,   /* erf GELU, ViltIntermediate (modeling_vilt.py:461-466); aux <- pre-activation */
    CLIMB_EPI_DGELU = 2,  /* C = acc * gelu'(aux) */
    CLIMB_EPI_SWISH = 3,  /* Houlsby adapter non-linearity (adapters/modeling.py:120-201) */
    CLIMB_EPI_DSWISH = 4,
    CLIMB_EPI_RELU = 5,   /* Pfeiffer adapter non-linearity */
    CLIMB_EPI_DRELU = 6,
    CLIMB_EPI_TANH = 7    /* ViltPooler (modeling_vilt.py:887-899) */
} climb_epilogue;

const char* climb_last_error(void);
int climb_version(void);

/* ---------------------------------------------------------------------------------------------
 * GEMM: C[M,N] = epilogue(alpha * A[M,K] * B[N,K]^T + bias[n]) + residual[m,n]
 * Replaces every nn.Linear / Conv2d-as-GEMM of the path (modeling_vilt.py:309-328,356-360,
 * 407-414,461-466,480-487,887-899; src/modeling/vilt.py:190-202) and their autograd backward.
 *   a_mn_major = 0: A[m*lda + k]   (activations in forward/dgrad)
 *   a_mn_major = 1: A[k*lda + m]   (dY read in place as dY^T for wgrad)
 *   b_mn_major likewise for B[n*ldb + k] / B[k*ldb + n].
 *   accumulate = 1: C += result with fp32 atomics (gradient accumulation, split-K).
 *   split_k = 0 picks a split automatically (only when accumulate = 1), block_n = 0 picks a tile.
 *   aux: bf16 [M, ldaux]; written with the pre-activation (acc*alpha+bias, before the activation
 *        and the residual) for GELU/SWISH/RELU/TANH/NONE when non-null, read by the D* epilogues.
 *   c2 : optional bf16 [M, ldc2] copy of the final value, i.e. the operand of the next GEMM.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int M, N, K;
    const void* A; int64_t lda; int a_mn_major;
    const void* B; int64_t ldb; int b_mn_major;
    void* C; int64_t ldc; int c_dtype;
    const float* bias;
    const float* residual; int64_t ldr;
    int epilogue;
    void* aux; int64_t ldaux;
    void* c2; int64_t ldc2;   /* optional bf16 copy of the final C (after residual) */
    float alpha;          /* 0 is read as 1 */
    int accumulate;
    int split_k;
    int block_n;
} climb_gemm_desc;

int climb_gemm_bf16(const climb_gemm_desc* desc, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused multi-head self-attention, head dim 64 (ViltSelfAttention, modeling_vilt.py:355-388):
 *   ctx = softmax(Q K^T / 8 + key_bias) V, key_bias = (1 - mask) * -10000 (modeling_utils.py:299-311)
 * qkv  bf16 [B, L, 3*H*64] (q | k | v, head-major inside each third)
 * ctx  bf16 [B, L, H*64];  lse fp32 [B, H, L] (natural-log-sum-exp, saved for backward)
 * backward: dqkv bf16 [B, L, 3*H*64] from dctx bf16 [B, L, H*64]; delta fp32 [B, H, L] scratch.
 * ------------------------------------------------------------------------------------------- */
int climb_attention_fwd(const void* qkv, const float* key_bias, void* ctx, float* lse,
                        int B, int L, int H, float scale, void* stream);
int climb_attention_bwd(const void* qkv, const float* key_bias, const void* ctx, const void* dctx,
                        const float* lse, float* delta, void* dqkv,
                        int B, int L, int H, float scale, void* stream);

/* ---------------------------------------------------------------------------------------------
 * LayerNorm over the last dim (layernorm_before/after/final, text LayerNorm, head LayerNorm:
 * modeling_vilt.py:302,505,517,873; src/modeling/vilt.py:192).
 *   fwd: y = (x - mean) * rstd * gamma + beta; y_bf16 and/or y_f32 may be null; mean/rstd saved.
 *        act = CLIMB_EPI_GELU applies erf-GELU after the affine (task head LN->GELU).
 *   bwd: dx_f32 = LN'(dy) (+ dres when non-null); dx_bf16 optional copy; dgamma/dbeta += (atomics).
 *        dy is fp32 (dy_f32) or bf16 (dy_bf16) -- exactly one non-null. For act=GELU the saved
 *        fp32 pre-activation is recomputed from x, mean, rstd.
 * x / dres / dx_f32 rows are ldx elements apart (ldx = d for a dense matrix; ldx = L*d selects the
 * [CLS] row of every sequence for the final LayerNorm, whose other rows CLiMB never reads:
 * src/modeling/vilt.py:123-124); y, dy and dx_bf16 rows are dense. d in {128..1536} step 128.
 * ------------------------------------------------------------------------------------------- */
int climb_layernorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps,
                        void* y_bf16, float* y_f32, float* mean, float* rstd,
                        int rows, int d, int act, void* stream);
int climb_layernorm_bwd(const float* dy_f32, const void* dy_bf16, const float* x, int64_t ldx,
                        const float* gamma, const float* beta, const float* mean, const float* rstd,
                        const float* dres, float* dx_f32, void* dx_bf16,
                        float* dgamma, float* dbeta, int rows, int d, int act, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CLIMB_B200_H */
