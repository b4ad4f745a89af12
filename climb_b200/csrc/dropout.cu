// Dropout inside the ViLT encoder (config.hidden_dropout_prob / attention_probs_dropout_prob > 0 in train mode):
//   modeling_vilt.py:303   TextEmbeddings: dropout(LayerNorm(..))                      } one mask over the assembled
//   modeling_vilt.py:201   visual_embed:   dropout(patches + position embeddings)      } [B, L, d] rows, before the
//                                                                                         modality-type rows are added
//   modeling_vilt.py:374   ViltSelfAttention: dropout(softmax(..))   -> attention_tc.cu (Philox inside the kernels)
//   modeling_vilt.py:410   ViltSelfOutput: dropout(dense(ctx))       -> dropout_res below, before adapter / residual
//   modeling_vilt.py:482   ViltOutput:     dropout(dense(inter)) + x -> dropout_res below
// Every mask is a pure function of (seed of the site, element index): element 4 i + k is kept iff word k of
// philox4x32(seed, i) >= p * 2^32, kept values are scaled by 1 / (1 - p). The backward regenerates the masks; nothing is
// stored. ViltConfig and every CLiMB script leave both probabilities at 0.0, so none of this runs on the benchmark path.
#include "common.cuh"
#include "internal.h"

namespace climb {
namespace {

inline unsigned blocks_for(long long n, int threads) { return static_cast<unsigned>((n + threads - 1) / threads); }

__device__ __forceinline__ float4 keep4(unsigned long long seed, long long i, uint32_t thresh, float inv_keep) {
    const uint4 r = philox4x32(seed, static_cast<unsigned long long>(i));
    return make_float4(dropout_scale(r.x, thresh, inv_keep), dropout_scale(r.y, thresh, inv_keep),
                       dropout_scale(r.z, thresh, inv_keep), dropout_scale(r.w, thresh, inv_keep));
}

// y = dropout(x) + res ; pre = bf16(dropout(x)) ; post = bf16(y)      (res / pre / post optional)
__global__ void dropout_res_kernel(const float* __restrict__ x, const float* __restrict__ res, float* __restrict__ y,
                                   __nv_bfloat16* __restrict__ pre, __nv_bfloat16* __restrict__ post, long long n4, uint32_t thresh,
                                   float inv_keep, unsigned long long seed) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 v = reinterpret_cast<const float4*>(x)[i];
    const float4 k = keep4(seed, i, thresh, inv_keep);
    v.x *= k.x; v.y *= k.y; v.z *= k.z; v.w *= k.w;
    if (pre != nullptr) reinterpret_cast<uint2*>(pre)[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    if (res != nullptr) {
        const float4 q = reinterpret_cast<const float4*>(res)[i];
        v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
    }
    reinterpret_cast<float4*>(y)[i] = v;
    if (post != nullptr) reinterpret_cast<uint2*>(post)[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
}

// dst = src * keep   (gradient through a dropout site), bf16 -> bf16
__global__ void dropout_mask_bf16_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n4,
                                         uint32_t thresh, float inv_keep, unsigned long long seed) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const uint2 u = reinterpret_cast<const uint2*>(src)[i];
    const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
    const float4 k = keep4(seed, i, thresh, inv_keep);
    reinterpret_cast<uint2*>(dst)[i] = make_uint2(pack_bf16(a.x * k.x, a.y * k.y), pack_bf16(b.x * k.z, b.y * k.w));
}
// fp32 -> fp32 (and / or the keep factors themselves when src == nullptr: what the tests hand to the oracle)
__global__ void dropout_mask_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n4, uint32_t thresh,
                                        float inv_keep, unsigned long long seed) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 k = keep4(seed, i, thresh, inv_keep);
    if (src != nullptr) {
        const float4 v = reinterpret_cast<const float4*>(src)[i];
        k.x *= v.x; k.y *= v.y; k.z *= v.z; k.w *= v.w;
    }
    reinterpret_cast<float4*>(dst)[i] = k;
}

// keep factors of the attention-probability dropout as the kernels generate them: counter ((b H + h) L + q) * 64 + key / 4
__global__ void attn_dropout_mask_kernel(float* __restrict__ out, int B, int H, int L, uint32_t thresh, float inv_keep,
                                         unsigned long long seed) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;      // (b, h, q, key / 4)
    const int k4 = (L + 3) / 4;
    if (i >= static_cast<long long>(B) * H * L * k4) return;
    const int kq = static_cast<int>(i % k4);
    const long long row = i / k4;                                                            // (b H + h) L + q
    const uint4 r = philox4x32(seed, static_cast<unsigned long long>(row) * 64ull + kq);
    const float k[4] = {dropout_scale(r.x, thresh, inv_keep), dropout_scale(r.y, thresh, inv_keep),
                        dropout_scale(r.z, thresh, inv_keep), dropout_scale(r.w, thresh, inv_keep)};
    for (int e = 0; e < 4; ++e)
        if (kq * 4 + e < L) out[row * L + kq * 4 + e] = k[e];
}

}  // namespace

unsigned long long dropout_site_seed(unsigned long long base, int layer, int site) {
    // distinct Philox keys per dropout site: layer -1 = embeddings; site 0 = attention probabilities, 1 = self-output, 2 = output
    return base + 0x9E3779B97F4A7C15ull * static_cast<unsigned long long>((layer + 1) * 4 + site + 1);
}

int dropout_res(const float* x, const float* res, float* y, void* pre_bf16, void* post_bf16, long long n, float p,
                unsigned long long seed, cudaStream_t s) {
    CLIMB_REQUIRE(x && y && n > 0 && n % 4 == 0, "dropout_res: bad arguments (n=%lld must be a positive multiple of 4)", n);
    CLIMB_REQUIRE(p >= 0.0f && p < 1.0f, "dropout_res: p=%f outside [0, 1)", p);
    dropout_res_kernel<<<blocks_for(n / 4, 256), 256, 0, s>>>(x, res, y, static_cast<__nv_bfloat16*>(pre_bf16),
                                                           static_cast<__nv_bfloat16*>(post_bf16), n / 4,
                                                           p > 0.0f ? dropout_threshold(p) : 0u, 1.0f / (1.0f - p), seed);
    CLIMB_LAUNCH_OK();
    return 0;
}
int dropout_mask_bf16(const void* src, void* dst, long long n, float p, unsigned long long seed, cudaStream_t s) {
    CLIMB_REQUIRE(src && dst && n > 0 && n % 4 == 0, "dropout_mask_bf16: bad arguments");
    dropout_mask_bf16_kernel<<<blocks_for(n / 4, 256), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(src), static_cast<__nv_bfloat16*>(dst),
                                                                 n / 4, p > 0.0f ? dropout_threshold(p) : 0u, 1.0f / (1.0f - p), seed);
    CLIMB_LAUNCH_OK();
    return 0;
}
int dropout_mask_f32(const float* src, float* dst, long long n, float p, unsigned long long seed, cudaStream_t s) {
    CLIMB_REQUIRE(dst && n > 0 && n % 4 == 0, "dropout_mask_f32: bad arguments");
    dropout_mask_f32_kernel<<<blocks_for(n / 4, 256), 256, 0, s>>>(src, dst, n / 4, p > 0.0f ? dropout_threshold(p) : 0u,
                                                                1.0f / (1.0f - p), seed);
    CLIMB_LAUNCH_OK();
    return 0;
}
int attn_dropout_mask(float* out, int B, int H, int L, float p, unsigned long long seed, cudaStream_t s) {
    CLIMB_REQUIRE(out && B > 0 && H > 0 && L > 0, "attn_dropout_mask: bad arguments");
    const long long n = static_cast<long long>(B) * H * L * ((L + 3) / 4);
    attn_dropout_mask_kernel<<<blocks_for(n, 256), 256, 0, s>>>(out, B, H, L, p > 0.0f ? dropout_threshold(p) : 0u, 1.0f / (1.0f - p), seed);
    CLIMB_LAUNCH_OK();
    return 0;
}

}  // namespace climb
