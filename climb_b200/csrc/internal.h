// Internal (C++) declarations of the kernels' host launchers; capi.cu wraps these in extern "C".
#pragma once
#include <cuda_runtime.h>
#include "climb_b200.h"

namespace climb {

// capi.cu: device pointer of the sticky error word (host-mapped pinned memory), nullptr if it cannot be allocated
unsigned int* device_error_word();

int gemm_bf16(const climb_gemm_desc* d, cudaStream_t stream);
int gemm_pair_mode(int mode);
int num_sms();                  // SMs the persistent kernels may fill (device count minus sm_reserve)
int sm_reserve(int n);          // n >= 0 sets the reserve, n < 0 queries; returns the previous value
// fused adapter bottleneck (gemm_tcgen05.cu): down -> activation -> up -> residual in one launch, forward and backward
bool adapter_fused_ok(int d, int r);
int adapter_fused(int backward, int M, int d, int r, int act, const void* A, const void* w_down, const void* w_up, const float* b_down,
                  const float* b_up, void* pre, void* z, const float* c_in, float* c_out, void* c2, float* colsum_z,
                  cudaStream_t stream);

int attention_fwd(const void* qkv, const float* key_bias, void* ctx, float* lse, int B, int L, int H,
                  float scale, cudaStream_t stream);
int attention_bwd(const void* qkv, const float* key_bias, const void* ctx, const void* dctx,
                  const float* lse, float* delta, void* dqkv, float* dqkv_colsum, int B, int L, int H, float scale,
                  cudaStream_t stream);

int layernorm_fwd(const float* x, long long ldx, const float* gamma, const float* beta, float eps,
                  void* y_bf16, float* y_f32, float* mean, float* rstd, int rows, int d, int act,
                  cudaStream_t stream);
int layernorm_bwd(const float* dy_f32, const void* dy_bf16, const float* x, long long ldx,
                  const float* gamma, const float* beta, const float* mean, const float* rstd,
                  const float* dres, float* dx_f32, void* dx_bf16, float* dgamma, float* dbeta,
                  int rows, int d, int act, cudaStream_t stream, float* dx_colsum = nullptr);


// tma_host.cu
int encode_tmap_bf16(void* map_out, const void* ptr, int rank, const long long* dims, const long long* strides_elems,
                     const int* box);

// attention_tc.cu (tcgen05 / TMEM attention for L <= 256; attention.cu keeps the general-L kernels)
int attention_tc_fwd(const void* qkv, const float* key_bias, void* ctx, float* lse, int B, int L, int H, float scale,
                     cudaStream_t stream, float p_drop = 0.0f, unsigned long long seed = 0);
int attention_tc_bwd(const void* qkv, const float* key_bias, const void* ctx, const void* dctx, const float* lse,
                     void* dqkv, float* colsum, int B, int L, int H, float scale, cudaStream_t stream, float p_drop = 0.0f,
                     unsigned long long seed = 0);

// elementwise.cu
int cast_f32_bf16(const float* src, void* dst, long long n, cudaStream_t stream);
int colsum(const void* src, int dtype, long long ld, int rows, int cols, float* out, cudaStream_t stream);
int key_bias(const long long* mask, float* out, int B, int T, int L, cudaStream_t stream);
int tanh_bwd(const float* dy, const float* y, void* out_bf16, long long n, cudaStream_t stream);
int bce_logits_loss(const float* logits, long long ld, const float* target, int rows, int cols, float scale,
                    float grad_scale, float* row_loss, float* loss, float* dlogits, long long ldd,
                    cudaStream_t stream);
int cross_entropy_loss(const float* logits, long long ld, const long long* target, int rows, int cols,
                       float grad_scale, float* row_loss, float* loss, float* dlogits, long long ldd,
                       cudaStream_t stream);

// embed.cu
int text_gather(const long long* ids, const float* inputs_embeds, const long long* tt, const float* word,
                const float* type_emb, const float* pos, float* e, int rows, int T, int d, cudaStream_t stream,
                int vocab = 0, int n_types = 0);      // > 0: ids are range-checked (clamped + climb_error_flags)
int im2col(const float* px, void* out, int B, int C, int H, int W, int P, cudaStream_t stream);
int pos_interp(const float* pos_emb, float* table, int hp, int wp, int G, int d, cudaStream_t stream);
int embed_assemble(const float* text_ln, const float* patch, const float* table, const float* cls,
                   const float* pos_emb, const float* mod, const int* type_idx, int type_idx_scalar, float* x,
                   int B, int T, int Np, int d, cudaStream_t stream, int n_mod = 0, float p_drop = 0.0f,
                   unsigned long long seed = 0, int rep = 1);
// rep > 1 (climb_vilt_batch.image_repeat): `patch` / `dpatch` hold B / rep images, each shared by rep consecutive sequences
int embed_split_bwd(const float* dx, float* dy_text, void* dpatch, int B, int T, int Np, int d, cudaStream_t stream, int rep = 1);
int embed_reduce_bwd(const float* dx, const int* type_idx, int type_idx_scalar, float* S, float* d_cls,
                     float* d_pos, float* d_mod, float* d_patch_bias, int n_mod, int B, int T, int hp, int wp,
                     int G, int d, cudaStream_t stream, const int* geom = nullptr, int ragged_np = 0, const int* sel = nullptr);
// variable-resolution (padded images): geom = [B, 2] valid patch rows / cols per image, Np patch slots per sequence
// sel (optional, [B, Np] int32): raster index of the patch in every slot or -1 (climb_vilt_batch.patch_select)
int im2col_ragged(const float* px, const int* geom, void* out, int B, int C, int H, int W, int P, int Np, cudaStream_t stream,
                  int rep = 1, const int* sel = nullptr);
int embed_assemble_ragged(const float* text_ln, const float* patch, const int* geom, const float* cls, const float* pos_emb,
                          const float* mod, const int* type_idx, int type_idx_scalar, float* x, int B, int T, int Np, int G,
                          int d, cudaStream_t stream, int n_mod = 0, float p_drop = 0.0f, unsigned long long seed = 0, int rep = 1,
                          const int* sel = nullptr);
int key_bias_ragged(const long long* mask, const int* geom, float* out, int B, int T, int L, cudaStream_t stream, const int* sel = nullptr);
int text_scatter_bwd(const float* de, const long long* ids, const long long* tt, float* d_word, float* d_type,
                     float* d_pos, int rows, int T, int d, cudaStream_t stream, int vocab = 0, int n_types = 0);

// optim.cu
int ewc_penalty(const float* theta, const float* theta_star, const float* fisher, long long n, float lambda,
                float* partials, int n_partials, float* loss, float* grad, float grad_scale,
                const float* grad_scale_dev, cudaStream_t stream);
int fisher_accumulate(const float* grad, float* fisher, long long n, cudaStream_t stream);
int scale_inplace(float* x, long long n, float s, cudaStream_t stream);
int adamw_step(float* theta, const float* grad, float* m, float* v, void* shadow_bf16, const climb_adamw_chunk* chunks_dev,
               int n_chunks, const float* group_lr, const float* group_wd, int n_groups, float beta1, float beta2,
               float eps, int step, cudaStream_t stream);

// bert_engine.cu
long long bert_forward_workspace_bytes(const climb_bert_dims* dims, const climb_bert_batch* batch);
int bert_forward(const climb_bert_dims* dm, const climb_bert_params* pr, const climb_bert_batch* bt, const float* theta,
                 const void* shadow, void* workspace, long long workspace_bytes, float hidden_dropout, float attn_dropout,
                 unsigned long long seed, float* out, cudaStream_t s);
int dropout_add(const float* x, const float* res, float* y, long long n, float p, unsigned long long seed, cudaStream_t s);

// image_pre.cu
int image_preprocess(const uint8_t* src, uint8_t* tmp, const climb_image_desc* descs_dev, const int* tables_dev, int B,
                     long long max_tmp_pixels, float* pixel_values, long long* pixel_mask, int Hp, int Wp, const float* mean,
                     const float* stdv, cudaStream_t stream);

// engine.cu
long long vilt_forward_workspace_bytes(const climb_vilt_dims* dims, const climb_vilt_params* params,
                                       const climb_vilt_batch* batch, int save);
long long vilt_backward_scratch_bytes(const climb_vilt_dims* dims, const climb_vilt_params* params,
                                      const climb_vilt_batch* batch);
int vilt_forward(const climb_vilt_dims* dm, const climb_vilt_params* pr, const climb_vilt_batch* bt,
                 const float* theta, const void* shadow, void* workspace, long long workspace_bytes, int save,
                 float* pooled_out, cudaStream_t s);
int vilt_backward(const climb_vilt_dims* dm, const climb_vilt_params* pr, const climb_vilt_batch* bt,
                  const float* theta, const void* shadow, const void* workspace, long long workspace_bytes,
                  void* scratch, long long scratch_bytes, const float* dpooled, float* grad, int first_layer,
                  int last_layer, int parts, cudaStream_t s);

// dropout.cu (ViLT encoder dropout p > 0: hidden-state sites, mask export for the tests)
unsigned long long dropout_site_seed(unsigned long long base, int layer, int site);
int dropout_res(const float* x, const float* res, float* y, void* pre_bf16, void* post_bf16, long long n, float p,
                unsigned long long seed, cudaStream_t s);
int dropout_mask_bf16(const void* src, void* dst, long long n, float p, unsigned long long seed, cudaStream_t s);
int dropout_mask_f32(const float* src, float* dst, long long n, float p, unsigned long long seed, cudaStream_t s);
int attn_dropout_mask(float* out, int B, int H, int L, float p, unsigned long long seed, cudaStream_t s);

// precise.cu (CLIMB_PREC_BF16X3: split-operand contractions, fp32 activations and attention)
long long vilt_forward_workspace_bytes_precise(const climb_vilt_dims* dims, const climb_vilt_params* params,
                                               const climb_vilt_batch* batch, int save);
long long vilt_backward_scratch_bytes_precise(const climb_vilt_dims* dims, const climb_vilt_params* params,
                                              const climb_vilt_batch* batch);
int vilt_forward_precise(const climb_vilt_dims* dm, const climb_vilt_params* pr, const climb_vilt_batch* bt, const float* theta,
                         const void* shadow, void* workspace, long long workspace_bytes, int save, float* pooled_out,
                         cudaStream_t s);
int vilt_backward_precise(const climb_vilt_dims* dm, const climb_vilt_params* pr, const climb_vilt_batch* bt, const float* theta,
                          const void* shadow, const void* workspace, long long workspace_bytes, void* scratch,
                          long long scratch_bytes, const float* dpooled, float* grad, int first_layer, int last_layer, int parts,
                          cudaStream_t s);
int split_f32_bf16x2(const float* src, void* hi, void* lo, long long n, cudaStream_t s);

}  // namespace climb
