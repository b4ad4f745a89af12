// Internal (C++) declarations of the kernels' host launchers; capi.cu wraps these in extern "C".
#pragma once
#include <cuda_runtime.h>
#include "climb_b200.h"

namespace climb {

int gemm_bf16(const climb_gemm_desc* d, cudaStream_t stream);

int attention_fwd(const void* qkv, const float* key_bias, void* ctx, float* lse, int B, int L, int H,
                  float scale, cudaStream_t stream);
int attention_bwd(const void* qkv, const float* key_bias, const void* ctx, const void* dctx,
                  const float* lse, float* delta, void* dqkv, int B, int L, int H, float scale,
                  cudaStream_t stream);

int layernorm_fwd(const float* x, long long ldx, const float* gamma, const float* beta, float eps,
                  void* y_bf16, float* y_f32, float* mean, float* rstd, int rows, int d, int act,
                  cudaStream_t stream);
int layernorm_bwd(const float* dy_f32, const void* dy_bf16, const float* x, long long ldx,
                  const float* gamma, const float* beta, const float* mean, const float* rstd,
                  const float* dres, float* dx_f32, void* dx_bf16, float* dgamma, float* dbeta,
                  int rows, int d, int act, cudaStream_t stream);

}  // namespace climb
