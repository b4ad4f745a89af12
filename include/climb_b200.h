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
    CLIMB_EPI_TANH = 7,   /* ViltPooler (modeling_vilt.py:887-899) */
    CLIMB_EPI_GELU_SAVE_GRAD = 8, /* C = gelu(pre), aux <- gelu'(pre): the backward then needs no erf at all */
    CLIMB_EPI_MUL_AUX = 9         /* C = acc * aux (backward of GELU_SAVE_GRAD) */
} climb_epilogue;

const char* climb_last_error(void);
int climb_version(void);

/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
uint64_t climb_launch_count(void);

/* Which tcgen05 GEMM kernels climb_gemm_bf16 may choose (measurement / test switch; the default is the product path):
 * 0 = one-CTA kernels only, 1 = every CTA-pair (cta_group::2) kernel, 2 = default selection (pair weight-gradient kernel +
 * the pair variants of the plain and multiply-by-aux epilogues). Any other value only queries. Returns the previous mode.
 * The environment variable CLIMB_GEMM_PAIR (0 / 1) sets the initial mode. */
int climb_gemm_pair_mode(int mode);

/* SMs left free by the persistent kernels (GEMMs, attention): their grids are sized to (SM count - n). The data-parallel
 * layer (climb_b200/distributed.py) sets n = the number of CTAs NCCL's reduction kernels occupy while a gradient all-reduce is
 * in flight and 0 afterwards, so that a persistent grid never has to run its last CTAs as a second round behind them.
 * n >= 0 sets, n < 0 only queries; returns the previous value. */
int climb_set_sm_reserve(int n);

/* Device-time profiler used by bench.py's roofline: between begin and end every GEMM / attention
 * launcher is bracketed by two CUDA events ON ITS LAUNCH STREAM. climb_profile_end synchronises and
 * returns per category (0 = tcgen05 GEMM, 1 = attention fwd, 2 = attention bwd [3 kernels], 3 = unused)
 * the summed device milliseconds, the summed algorithmic work (GEMM: FLOPs = 2 M N K; attention:
 * algorithmic HBM bytes of SURVEY.md 8d) and the number of launches. Arrays of >= 4 entries. */
int climb_profile_begin(void);
int climb_profile_end(double* ms, double* work, int64_t* launches, int n_categories);

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
    float* colsum;            /* optional [N]: colsum[n] += sum_m C[m, n] (bias gradient of the layer that
                                 produced this dY), bf16 C, N % 32 == 0 and 16-byte aligned rows only */
    float alpha;          /* 0 is read as 1 */
    int accumulate;
    int split_k;
    int block_n;
    /* Programmatic dependent launch: every kernel of this library waits (griddepcontrol.wait) for the previous kernel
     * of its stream before touching global memory. independent = 1 declares that this GEMM reads / writes nothing the
     * PREVIOUS launch on the stream produces or consumes, so its CTAs may start on the SMs that launch's last wave
     * leaves idle; the launch AFTER it is then issued as a full stream barrier. The engine uses it for the
     * attention-output wgrad, which follows the attention backward (768 CTAs = 5.19 waves) without depending on it. */
    int independent;
    /* Weight-gradient form only (a_mn_major = 1, accumulate = 1: A = dY [K tokens, M], C = dW): optional [M],
     * colsum_a[m] += sum_k A[k, m] -- the bias gradient that belongs to dW (nn.Linear's backward produces both from the
     * same dY). Inside the CTA-pair weight-gradient kernel these are one extra tcgen05.mma per k-step against a tile of
     * ones (no second pass over dY); on every other kernel a streaming column-sum launch issued by climb_gemm_bf16. */
    float* colsum_a;
} climb_gemm_desc;

int climb_gemm_bf16(const climb_gemm_desc* desc, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused adapter bottleneck (Adapter.forward as ViLT wires it: adapters/modeling.py:160-179, mixins/vilt.py:23-125), one launch
 * per site and direction; the r-wide intermediate never round-trips through HBM between the two projections.
 *   backward = 0:  pre = A W_d^T + b_d ;  z = act(pre) ;  c_out = c_in + z W_u^T + b_u                 (A = the site's input)
 *   backward = 1:  z   = (A W_u) * act'(pre)  (= dpre) ;  c_out = c_in + z W_d ;  colsum_z += column sums of z
 *                  (A = gradient at the site's output; dpre and A then feed the two weight gradients)
 * A bf16 [M, d]; W_d bf16 [r, d]; W_u bf16 [d, r]; pre, z bf16 [M, r]; c_in / c_out fp32 [M, d] (may alias; c_out may be NULL);
 * c2 optional bf16 copy of the result [M, d]. act = CLIMB_EPI_SWISH or CLIMB_EPI_RELU. r <= 64, r % 16 == 0, d % 128 == 0
 * (CLiMB's reduction factor 16 on ViLT-base: d = 768, r = 48); other widths run as two climb_gemm_bf16 launches in the engine.
 * ------------------------------------------------------------------------------------------- */
int climb_adapter_fused(int backward, int M, int d, int r, int act, const void* a_bf16, const void* w_down_bf16, const void* w_up_bf16,
                        const float* b_down, const float* b_up, void* pre_bf16, void* z_bf16, const float* c_in, float* c_out,
                        void* c2_bf16, float* colsum_z, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused multi-head self-attention, head dim 64 (ViltSelfAttention, modeling_vilt.py:355-388):
 *   ctx = softmax(Q K^T / 8 + key_bias) V, key_bias = (1 - mask) * -10000 (modeling_utils.py:299-311)
 * qkv  bf16 [B, L, 3*H*64] (q | k | v, head-major inside each third)
 * ctx  bf16 [B, L, H*64];  lse fp32 [B, H, L] (natural-log-sum-exp, saved for backward)
 * backward: dqkv bf16 [B, L, 3*H*64] from dctx bf16 [B, L, H*64]; delta fp32 [B, H, L] scratch;
 *           dqkv_colsum (nullable, fp32 [3*H*64]) += column sums of dqkv = the q/k/v bias gradients.
 * L <= 256 runs on tcgen05 / TMEM kernels (attention_tc.cu), longer sequences on mma.sync kernels.
 * ------------------------------------------------------------------------------------------- */
int climb_attention_fwd(const void* qkv, const float* key_bias, void* ctx, float* lse,
                        int B, int L, int H, float scale, void* stream);
/* forward with dropout on the attention probabilities (BertSelfAttention, modeling_bert.py:341-345):
 * ctx = dropout(softmax(..), p) V; L <= 256. The mask is a pure function of (seed, b, h, query, key). */
int climb_attention_fwd_dropout(const void* qkv, const float* key_bias, void* ctx, float* lse,
                                int B, int L, int H, float scale, float p, uint64_t seed, void* stream);
int climb_attention_bwd(const void* qkv, const float* key_bias, const void* ctx, const void* dctx,
                        const float* lse, float* delta, void* dqkv, float* dqkv_colsum,
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
/* same, plus dx_colsum[d] += column sums of the output dx (nullable): the bias gradient of the Linear whose
 * output row stream this LayerNorm normalised (fc2 / attention-output dense), fused into the pass */
int climb_layernorm_bwd_colsum(const float* dy_f32, const void* dy_bf16, const float* x, int64_t ldx,
                               const float* gamma, const float* beta, const float* mean, const float* rstd,
                               const float* dres, float* dx_f32, void* dx_bf16,
                               float* dgamma, float* dbeta, float* dx_colsum, int rows, int d, int act, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Streaming helpers
 * ------------------------------------------------------------------------------------------- */
/* bf16 shadow of the fp32 master parameters (tensor-core operands); n elements, 16-byte aligned */
int climb_cast_f32_bf16(const float* src, void* dst_bf16, int64_t n, void* stream);
/* split operands of CLIMB_PREC_BF16X3: hi = bf16(x), lo = bf16(x - float(hi)); either output may be NULL */
int climb_split_f32_bf16x2(const float* src, void* hi_bf16, void* lo_bf16, int64_t n, void* stream);

/* Sticky device-side error flags (kernels clamp a bad index instead of faulting and raise a bit here; the word lives in
 * host-mapped pinned memory, so reading it needs no synchronisation -- a kernel's bit becomes visible once that kernel has
 * run). Returns the bits raised since the last call and clears them. */
#define CLIMB_ERR_TOKEN_ID 1        /* input_ids outside [0, vocab_size)              (nn.Embedding: IndexError) */
#define CLIMB_ERR_TOKEN_TYPE 2      /* token_type_ids outside [0, type_vocab_size) */
#define CLIMB_ERR_MODALITY 4        /* image_token_type_idx outside the modality table */
uint32_t climb_error_flags(void);
/* out[c] += sum_r src[r*ld + c]  (bias gradients); dtype = CLIMB_BF16 / CLIMB_F32 */
int climb_colsum(const void* src, int dtype, int64_t ld, int rows, int cols, float* out, void* stream);

/* Trainer losses with fused gradient (train_vqa.py:95,157: BCEWithLogits(mean) * num_labels;
 * train_nlvr2.py:80,133 / train_snli_ve.py / train_vcr.py:83,135: CrossEntropy(mean)).
 * row_loss: [rows] scratch; loss: 1 float; dlogits (nullable) = grad_scale * dloss/dlogits. */
int climb_bce_logits_loss(const float* logits, int64_t ld, const float* target, int rows, int cols,
                          float scale, float grad_scale, float* row_loss, float* loss,
                          float* dlogits, int64_t ldd, void* stream);
int climb_cross_entropy_loss(const float* logits, int64_t ld, const int64_t* target, int rows, int cols,
                             float grad_scale, float* row_loss, float* loss,
                             float* dlogits, int64_t ldd, void* stream);

/* ---------------------------------------------------------------------------------------------
 * EWC (src/cl_algorithms/ewc.py) on device-resident flat arenas of n floats
 *   climb_ewc_penalty : *loss = lambda * sum F (theta - theta*)^2 ; if grad != NULL also
 *                       grad += grad_scale * (*grad_scale_dev) * 2 lambda F (theta - theta*)  (ewc.py:75-87)
 *                       (grad_scale_dev: optional device scalar = the upstream gradient, read without a sync)
 *                       partials: scratch of n_partials floats (>= 1184 is always enough)
 *   climb_fisher_accumulate : fisher += grad^2                                    (ewc.py:61-64)
 *   climb_scale_inplace     : x *= s   (the division by the sample count, ewc.py:70-71)
 * ------------------------------------------------------------------------------------------- */
int climb_ewc_penalty(const float* theta, const float* theta_star, const float* fisher, int64_t n,
                      float lambda, float* partials, int n_partials, float* loss,
                      float* grad, float grad_scale, const float* grad_scale_dev, void* stream);
int climb_fisher_accumulate(const float* grad, float* fisher, int64_t n, void* stream);
int climb_scale_inplace(float* x, int64_t n, float s, void* stream);

/* ---------------------------------------------------------------------------------------------
 * AdamW over the flat arena (torch.optim.AdamW semantics; src/modeling/vilt.py:205-215).
 * chunks: DEVICE array; every chunk lies inside one parameter tensor and names that tensor's param
 * group (the reference has two: decayed / not decayed); the groups' current lr / weight decay are
 * HOST arrays passed by value each step, so an lr scheduler never forces a table rebuild.
 * ------------------------------------------------------------------------------------------- */
#define CLIMB_ADAMW_MAX_GROUPS 8
typedef struct {
    int64_t start;
    int32_t length;
    int32_t group;       /* index into group_lr / group_wd */
} climb_adamw_chunk;
int climb_adamw_step(float* theta, const float* grad, float* exp_avg, float* exp_avg_sq,
                     void* shadow_bf16 /* nullable: bf16 copy of the updated theta, same offsets */,
                     const climb_adamw_chunk* chunks_dev, int n_chunks,
                     const float* group_lr_host, const float* group_wd_host, int n_groups,
                     float beta1, float beta2, float eps, int step, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Whole-encoder engine: ViltModel.forward -> pooler_output and its backward
 * (modeling_vilt.py:777-899 with ViltEmbeddings :92-328, 12 x ViltLayer :503-525, final LayerNorm,
 * ViltPooler; adapters per adapters/mixins/vilt.py:23-125). One C call launches every kernel of
 * the pass on `stream`.
 *
 * Parameters live in ONE flat fp32 arena `theta`; `shadow` is its bf16 copy (climb_cast_f32_bf16)
 * and `grad` the fp32 gradient arena, all three with the same element offsets (below, -1 = absent).
 * query/key/value weights (and biases) must be contiguous in that order: qkv_w points at query.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int hidden, layers, heads, ffn;
    int patch, channels, pos_grid;      /* pos_grid = image_size / patch_size of the position table */
    int n_modality;                     /* rows of token_type_embeddings (2, or 3 with NLVR2) */
    float ln_eps;
    int vocab_size, type_vocab_size;    /* rows of the word / text token-type tables: ids are range-checked on the device
                                           (nn.Embedding raises IndexError; here: clamped + climb_error_flags()); 0 = unchecked */
    int precision;                      /* CLIMB_PREC_BF16 (throughput mode) or CLIMB_PREC_BF16X3 (parity gate, below) */
    float hidden_dropout, attn_dropout; /* ViltConfig.hidden_dropout_prob / attention_probs_dropout_prob (modeling_vilt.py:303,
                                           374,410,482), applied only when climb_vilt_batch.training != 0; defaults 0.0 */
} climb_vilt_dims;

/* Arithmetic of the engine.
 *   CLIMB_PREC_BF16   : bf16 tensor-core operands, fp32 accumulation / residual stream / statistics. The throughput mode;
 *                       2e-3 .. 6e-3 relative error on pooled / logits against the fp32 reference.
 *   CLIMB_PREC_BF16X3 : the parity gate of the north star ("logits within 1e-3 rel"): every contraction runs as three
 *                       accumulating tcgen05 launches over split operands  x = hi + lo (bf16 each):  A.W ~ Ahi.Whi + Alo.Whi
 *                       + Ahi.Wlo  (error ~2^-16 per product instead of 2^-8), activations stay fp32 between the
 *                       contractions and attention runs in an fp32 kernel. ~5x slower; same entry points, same gradients. */
#define CLIMB_PREC_BF16 0
#define CLIMB_PREC_BF16X3 1

typedef struct {
    int64_t qkv_w, qkv_b, o_w, o_b, fc1_w, fc1_b, fc2_w, fc2_b;
    int64_t ln1_w, ln1_b, ln2_w, ln2_b;
    int64_t mh_down_w, mh_down_b, mh_up_w, mh_up_b;        /* attention.output.adapters.<task> */
    int64_t out_down_w, out_down_b, out_up_w, out_up_b;    /* output.adapters.<task> */
    int32_t flags;                                         /* CLIMB_TRAIN_* bits, backward only */
    int32_t pad_;
} climb_vilt_layer;

#define CLIMB_TRAIN_BASE 1      /* base weights of this layer / of the embeddings need gradients */
#define CLIMB_TRAIN_ADAPTER 2   /* the active adapter of this layer needs gradients */

typedef struct {
    int64_t cls_token, pos_emb, word_emb, text_pos_emb, text_type_emb, text_ln_w, text_ln_b;
    int64_t patch_w, patch_b, mod_emb, final_ln_w, final_ln_b, pooler_w, pooler_b;
    const climb_vilt_layer* layer;      /* HOST array [dims.layers] */
    int adapter_r;                      /* bottleneck width of the active adapter, 0 = none */
    int adapter_act;                    /* CLIMB_EPI_SWISH (houlsby) or CLIMB_EPI_RELU (pfeiffer) */
    int32_t embed_flags;                /* CLIMB_TRAIN_BASE if the embeddings need gradients */
    int32_t tail_flags;                 /* CLIMB_TRAIN_BASE if final LayerNorm + pooler need gradients */
    const void* shadow_lo;              /* CLIMB_PREC_BF16X3 only: bf16(theta - float(shadow)), same offsets (climb_split_f32_bf16x2) */
} climb_vilt_params;

typedef struct {
    int B, T, H, W;                     /* sequences, text tokens, image height / width (pixels) */
    const int64_t* input_ids;           /* [B, T] or NULL when inputs_embeds is given */
    const float* inputs_embeds;         /* [B, T, hidden] (ViLT-BERT) or NULL */
    const int64_t* token_type_ids;      /* [B, T] or NULL (= zeros) */
    const int64_t* attention_mask;      /* [B, T] or NULL (= ones) */
    const float* pixel_values;          /* [B, C, H, W] */
    const int32_t* image_type_idx;      /* [B] or NULL -> image_type_idx_scalar for every sequence */
    int image_type_idx_scalar;
    /* Variable resolution (images padded to a common H x W, ViltEmbeddings.visual_embed modeling_vilt.py:121-205):
     * patch_geom [B, 2] int32 = valid patch rows / columns (h_b, w_b) of every image = what the reference derives from
     * pixel_mask (:126-129), n_patch_slots = patch rows per sequence (>= max_b h_b * w_b; the reference uses exactly the
     * maximum, :163-170). Slots past h_b * w_b are padding: zero pixels, no position embedding, masked as attention keys.
     * NULL / 0 = fixed resolution: every image fills the whole (H / patch) x (W / patch) grid. */
    const int32_t* patch_geom;
    int n_patch_slots;
    int training;                       /* != 0: dims.hidden_dropout / attn_dropout are live (nn.Module.train()) */
    uint64_t dropout_seed;              /* Philox key of this forward; the backward regenerates the masks from it */
    /* > 1: pixel_values holds B / image_repeat images, image i belongs to the image_repeat consecutive sequences
     * i * image_repeat ... (VCR: one image, four answer choices, src/modeling/vilt.py:334-347 feeds the same pixels four
     * times): im2col + patch projection run once per image, the embedding assembly broadcasts the rows, the backward sums
     * the patch gradients of the sequences that share an image. patch_geom stays per sequence. 0 / 1 = one image per sequence. */
    int image_repeat;
    /* config.max_image_length > 0 (modeling_vilt.py:163-189: an image with more valid patches than the cap keeps a random
     * subset, drawn with torch.multinomial): patch_select [B, n_patch_slots] int32 = raster index (in the image's own
     * h_b x w_b grid) of the patch in every sequence slot, -1 = padding slot. The HOST draws the subset with the same torch
     * calls in the same order as the reference (climb_b200/modeling/vilt_model.py), so a seeded run keeps the same patches.
     * Needs patch_geom; NULL = slots 0 .. h_b w_b - 1 hold all valid patches in raster order. */
    const int32_t* patch_select;
} climb_vilt_batch;

/* bytes of the activation workspace a forward needs (save_for_backward = 1 keeps every layer's
 * activations; 0 recycles one layer's worth) and of the scratch a backward needs */
int64_t climb_vilt_forward_workspace_bytes(const climb_vilt_dims* dims, const climb_vilt_params* params,
                                           const climb_vilt_batch* batch, int save_for_backward);
int64_t climb_vilt_backward_scratch_bytes(const climb_vilt_dims* dims, const climb_vilt_params* params,
                                          const climb_vilt_batch* batch);

int climb_vilt_forward(const climb_vilt_dims* dims, const climb_vilt_params* params,
                       const climb_vilt_batch* batch, const float* theta, const void* shadow,
                       void* workspace, int64_t workspace_bytes, int save_for_backward,
                       float* pooled_out /* [B, hidden] */, void* stream);

/* grad arena += d(pooled . dpooled)/d(theta) for every parameter whose flag asks for it.
 * `workspace` is the one the matching forward (save_for_backward = 1) filled.
 * The pass may be issued in several calls over the same scratch, top of the network first, so that the
 * data-parallel gradient all-reduce of the layers already done overlaps the rest of the backward:
 *   parts & CLIMB_BWD_TAIL  : pooler + final LayerNorm (must be in the first call)
 *   layers first_layer .. last_layer (descending, inclusive)
 *   parts & CLIMB_BWD_EMBED : embeddings (must be in the last call)
 * One call with first_layer = layers-1, last_layer = 0, parts = TAIL|EMBED does everything. */
#define CLIMB_BWD_TAIL 1
#define CLIMB_BWD_EMBED 2
int climb_vilt_backward(const climb_vilt_dims* dims, const climb_vilt_params* params,
                        const climb_vilt_batch* batch, const float* theta, const void* shadow,
                        const void* workspace, int64_t workspace_bytes,
                        void* scratch, int64_t scratch_bytes,
                        const float* dpooled /* [B, hidden] */, float* grad,
                        int first_layer, int last_layer, int parts, void* stream);


/* ---------------------------------------------------------------------------------------------
 * Frozen BERT text encoder of ViLT-BERT (src/modeling/viltbert.py:115-120 get_bert_outputs:
 * BertModel(...).last_hidden_state under torch.no_grad(), fed to ViltModel as inputs_embeds :135-151).
 * Post-LN encoder of adapter-transformers' modeling_bert.py: BertEmbeddings (:171-228), BertSelfAttention
 * (:231-360), BertSelfOutput (:362-376), BertIntermediate (:430-442), BertOutput (:445-459).
 * Forward only (the reference never differentiates through it). Offsets index the BERT parameter
 * arena (theta fp32 / shadow bf16); q, k, v weights (and biases) of a layer are adjacent.
 * hidden_dropout / attn_dropout > 0 reproduce the reference's train-mode behaviour (its BertModel keeps
 * hidden_dropout_prob = attention_probs_dropout_prob = 0.1 active inside no_grad when the learner is in
 * train mode) with a counter-based generator keyed by `seed`; 0 = eval mode = deterministic.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int hidden, layers, heads, ffn;
    float ln_eps;
} climb_bert_dims;

typedef struct {
    int64_t qkv_w, qkv_b, o_w, o_b, attn_ln_w, attn_ln_b, fc1_w, fc1_b, fc2_w, fc2_b, out_ln_w, out_ln_b;
} climb_bert_layer;

typedef struct {
    int64_t word_emb, pos_emb, type_emb, emb_ln_w, emb_ln_b;
    const climb_bert_layer* layer;      /* HOST array [dims.layers] */
} climb_bert_params;

typedef struct {
    int B, T;
    const int64_t* input_ids;           /* [B, T] */
    const int64_t* token_type_ids;      /* [B, T] or NULL (= zeros) */
    const int64_t* attention_mask;      /* [B, T] or NULL (= ones) */
} climb_bert_batch;

int64_t climb_bert_forward_workspace_bytes(const climb_bert_dims* dims, const climb_bert_batch* batch);
int climb_bert_forward(const climb_bert_dims* dims, const climb_bert_params* params, const climb_bert_batch* batch,
                       const float* theta, const void* shadow, void* workspace, int64_t workspace_bytes,
                       float hidden_dropout, float attn_dropout, uint64_t seed,
                       float* last_hidden_state /* [B, T, hidden] */, void* stream);

/* y = dropout(x, p) (inverted scaling 1/(1-p)) + res (res may be NULL); x, res, y fp32 [n], y may alias x.
 * Element i of stream `seed` is kept iff philox(seed, i) >= p: the mask is a pure function of (seed, i). */
int climb_dropout_add(const float* x, const float* res, float* y, int64_t n, float p, uint64_t seed, void* stream);

/* Dropout inside the ViLT encoder (dims.hidden_dropout / attn_dropout > 0 with batch.training != 0; modeling_vilt.py:201,303,374,
 * 410,482). Every mask is a pure function of (site seed, element index); the backward regenerates it. These entry points expose
 * the masks so that a test can hand the SAME masks to the CPU oracle:
 *   climb_dropout_site_seed(batch.dropout_seed, layer, site): layer -1 / site 1 = embeddings ([B, L, hidden] rows after assembly,
 *       before the modality-type rows are added); layer l / site 0 = attention probabilities, 1 = self-output dense, 2 = output dense
 *   climb_dropout_keep_mask          : out[i] = 0 or 1 / (1 - p) for the n elements of a hidden-state site (n % 4 == 0)
 *   climb_attention_dropout_keep_mask: out[b, h, q, k] likewise for the probabilities, fp32 [B, H, L, L] */
uint64_t climb_dropout_site_seed(uint64_t base_seed, int layer, int site);
int climb_dropout_keep_mask(float* out, int64_t n, float p, uint64_t seed, void* stream);
int climb_attention_dropout_keep_mask(float* out, int B, int H, int L, float p, uint64_t seed, void* stream);

/* ---- image side of the input pipeline ------------------------------------------------------
 * Replaces ViltFeatureExtractor.__call__ (adapter-transformers/src/transformers/models/vilt/feature_extraction_vilt.py:
 * 253-292, called from ViltEncoderWrapper.process_inputs, src/modeling/vilt.py:83-96): Pillow BICUBIC resize of uint8 RGB
 * images, float32 (x / 255 - mean) / std, zero padding to the batch maximum and the pixel mask, bit-exact with the CPU path.
 * All pointers are DEVICE pointers except mean / std (host, 3 floats each). `tables` holds, per image and axis, the bounds
 * [out, 2] (first input index, tap count) and int32 coefficients [out, ksize] (2^-22 units) of Pillow's precompute_coeffs /
 * normalize_coeffs_8bpc; the offsets in the descriptors index it in ints. tmp receives the horizontally resampled images. */
typedef struct {
    int64_t src_off;                    /* bytes into src: image [in_h, in_w, 3] uint8 */
    int64_t tmp_off;                    /* bytes into tmp: [in_h, out_w, 3] uint8 */
    int32_t in_h, in_w, out_h, out_w;
    int32_t ksize_h, ksize_v;
    int64_t bounds_h_off, coef_h_off, bounds_v_off, coef_v_off;
} climb_image_desc;

int climb_image_preprocess(const uint8_t* src, uint8_t* tmp, const climb_image_desc* descs, const int32_t* tables, int B,
                           int64_t max_tmp_pixels /* max over images of in_h * out_w */,
                           float* pixel_values /* [B, 3, Hp, Wp] */, int64_t* pixel_mask /* [B, Hp, Wp] */, int Hp, int Wp,
                           const float* mean, const float* std, void* stream);

/* ---- text side of the input pipeline -------------------------------------------------------
 * Replaces the tokenizer call of ViltProcessor.__call__ (processing_vilt.py:72-89, from ViltEncoderWrapper.process_inputs,
 * src/modeling/vilt.py:83-96): BertTokenizerFast(text, padding=True, truncation=True, max_length=...) = special-token split,
 * BertNormalizer, BertPreTokenizer, WordPiece, [CLS] ... [SEP], truncation, padding. HOST code (no device work): rows are
 * written into caller memory, typically the pinned staging buffer of the batch's host-to-device copy.
 * vocab: the bytes of a BERT vocab.txt (one token per line, id = line number). lowercase = do_lower_case (also strips accents,
 * as BERT does when strip_accents is unset). Returns NULL (see climb_last_error) if [UNK] [CLS] [SEP] [PAD] are missing. */
typedef struct climb_wordpiece climb_wordpiece;
climb_wordpiece* climb_wordpiece_create(const char* vocab, int64_t vocab_bytes, int lowercase, int handle_chinese_chars);
void climb_wordpiece_destroy(climb_wordpiece* tokenizer);
/* texts: n UTF-8 strings back to back, text i = bytes [offsets[i], offsets[i + 1]). Outputs are [n, max_length] int64, every
 * element written: ids padded with [PAD], mask 1 on real tokens, token types 0. *longest = the longest row incl. [CLS] / [SEP]
 * (padding=True keeps columns [0, longest)). Texts are split over up to n_threads host threads. */
int climb_wordpiece_encode(const climb_wordpiece* tokenizer, const char* texts, const int64_t* offsets, int n, int max_length,
                           int64_t* input_ids, int64_t* attention_mask, int64_t* token_type_ids, int* longest, int n_threads);

#ifdef __cplusplus
}
#endif
#endif /* CLIMB_B200_H */
