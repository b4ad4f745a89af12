// Frozen BERT text encoder of ViLT-BERT: BertModel(...).last_hidden_state under no_grad
// (src/modeling/viltbert.py:115-120), fed to ViltModel as inputs_embeds (:135-151). One C call runs
// adapter-transformers' modeling_bert.py forward as a fixed sequence of this directory's kernels:
//   BertEmbeddings (:171-228)   word[ids] + type[tt] + pos[0:T] -> LayerNorm -> dropout
//   12 x BertLayer (:462-545)   POST-LN: x = LN(x + drop(O(attn(x))));  x = LN(x + drop(FC2(gelu(FC1(x)))))
// Forward only: the reference never differentiates through it, so nothing is saved. Every Linear is the
// tcgen05 GEMM, attention the tcgen05 kernel (T <= 256), LayerNorm writes the fp32 stream and the bf16
// GEMM operand in one pass.
//
// Dropout (hidden_dropout_prob / attention_probs_dropout_prob = 0.1 in BertConfig) is live in the
// reference whenever the learner is in train mode, no_grad or not. With p = 0 (eval) the residual adds are
// fused into the GEMM epilogues; with p > 0 the dense output goes through dropout_add first.
#include "common.cuh"
#include "internal.h"

#include <cstring>

namespace climb {
namespace {

using bf16 = __nv_bfloat16;

__global__ void dropout_add_kernel(const float* __restrict__ x, const float* __restrict__ res, float* __restrict__ y,
                                   long long n4, uint32_t thresh, float inv_keep, unsigned long long seed) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 v = reinterpret_cast<const float4*>(x)[i];
    const uint4 r = philox4x32(seed, static_cast<unsigned long long>(i));
    v.x *= dropout_scale(r.x, thresh, inv_keep);
    v.y *= dropout_scale(r.y, thresh, inv_keep);
    v.z *= dropout_scale(r.z, thresh, inv_keep);
    v.w *= dropout_scale(r.w, thresh, inv_keep);
    if (res != nullptr) {
        const float4 q = reinterpret_cast<const float4*>(res)[i];
        v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
    }
    reinterpret_cast<float4*>(y)[i] = v;
}

struct Plan {
    int B, T, M, d, ff, heads, layers;
    float* key_bias;    // [B, T]
    float* e;           // [M, d] embedding sum / pre-LN sums
    float* x;           // [M, d] fp32 stream (LN output)
    bf16* xb;           // bf16 copy = GEMM operand
    float* x1;          // [M, d] after the attention block
    bf16* x1b;
    bf16* qkv;          // [M, 3d]
    bf16* ctx;          // [M, d]
    float* lse;         // [B, heads, T]
    bf16* inter;        // [M, ff]
    long long bytes;
};

struct Bump {
    uint8_t* base;
    long long off = 0;
    explicit Bump(void* b) : base(static_cast<uint8_t*>(b)) {}
    template <typename T>
    T* take(long long n) {
        off = (off + 255) & ~255LL;
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += n * static_cast<long long>(sizeof(T));
        return p;
    }
};

int fill_plan(Plan& P, const climb_bert_dims* dm, const climb_bert_batch* bt, void* base) {
    CLIMB_REQUIRE(dm && bt, "bert: null descriptor");
    CLIMB_REQUIRE(dm->layers > 0 && dm->hidden % 128 == 0 && dm->hidden == dm->heads * 64 && dm->ffn % 8 == 0,
                  "bert: hidden=%d must be heads*64 and a multiple of 128, ffn=%d a multiple of 8", dm->hidden, dm->ffn);
    CLIMB_REQUIRE(bt->B > 0 && bt->T > 0, "bert: empty batch (B=%d, T=%d)", bt->B, bt->T);
    P.B = bt->B; P.T = bt->T; P.M = bt->B * bt->T;
    P.d = dm->hidden; P.ff = dm->ffn; P.heads = dm->heads; P.layers = dm->layers;
    const long long M = P.M, d = P.d;
    Bump b(base);
    P.key_bias = b.take<float>(M);
    P.e = b.take<float>(M * d);
    P.x = b.take<float>(M * d);
    P.xb = b.take<bf16>(M * d);
    P.x1 = b.take<float>(M * d);
    P.x1b = b.take<bf16>(M * d);
    P.qkv = b.take<bf16>(M * 3 * d);
    P.ctx = b.take<bf16>(M * d);
    P.lse = b.take<float>(static_cast<long long>(P.B) * P.heads * P.T);
    P.inter = b.take<bf16>(M * P.ff);
    P.bytes = (b.off + 255) & ~255LL;
    return 0;
}

// C = epi(A W^T + bias) (+ residual)
int linear(int M, int N, int K, const bf16* A, const bf16* W, const float* bias, void* C, int c_dtype, int epi,
           const float* residual, cudaStream_t s) {
    climb_gemm_desc g;
    std::memset(&g, 0, sizeof(g));
    g.M = M; g.N = N; g.K = K;
    g.A = A; g.lda = K;
    g.B = W; g.ldb = K;
    g.C = C; g.ldc = N; g.c_dtype = c_dtype;
    g.bias = bias; g.residual = residual; g.ldr = N;
    g.epilogue = epi;
    g.alpha = 1.0f;
    return gemm_bf16(&g, s);
}

#define TRY(expr)                 \
    do {                          \
        int _rc = (expr);         \
        if (_rc) return _rc;      \
    } while (0)

}  // namespace

int dropout_add(const float* x, const float* res, float* y, long long n, float p, unsigned long long seed, cudaStream_t s) {
    CLIMB_REQUIRE(x && y && n > 0 && n % 4 == 0, "dropout_add: bad arguments (n=%lld must be a positive multiple of 4)", n);
    CLIMB_REQUIRE(p >= 0.0f && p < 1.0f, "dropout_add: p=%f outside [0, 1)", p);
    const long long n4 = n / 4;
    dropout_add_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, s>>>(x, res, y, n4, p > 0.0f ? dropout_threshold(p) : 0u,
                                                                           1.0f / (1.0f - p), seed);
    CLIMB_LAUNCH_OK();
    return 0;
}

long long bert_forward_workspace_bytes(const climb_bert_dims* dims, const climb_bert_batch* batch) {
    Plan P;
    if (fill_plan(P, dims, batch, nullptr)) return -1;
    return P.bytes;
}

int bert_forward(const climb_bert_dims* dm, const climb_bert_params* pr, const climb_bert_batch* bt, const float* theta,
                 const void* shadow, void* workspace, long long workspace_bytes, float p_hid, float p_attn,
                 unsigned long long seed, float* out, cudaStream_t s) {
    Plan P;
    TRY(fill_plan(P, dm, bt, workspace));
    CLIMB_REQUIRE(pr && pr->layer && theta && shadow && workspace && out && bt->input_ids, "bert_forward: null buffer");
    CLIMB_REQUIRE(workspace_bytes >= P.bytes, "bert_forward: workspace %lld < required %lld", workspace_bytes, P.bytes);
    CLIMB_REQUIRE(P.T <= 256 || p_attn == 0.0f, "bert_forward: attention dropout needs T <= 256 (T=%d)", P.T);
    CLIMB_REQUIRE(p_hid >= 0.0f && p_hid < 1.0f && p_attn >= 0.0f && p_attn < 1.0f, "bert_forward: dropout outside [0, 1)");
    const int M = P.M, d = P.d;
    const long long Md = static_cast<long long>(M) * d;
    auto F = [&](long long off) { return theta + off; };
    auto H = [&](long long off) { return static_cast<const bf16*>(shadow) + off; };
    unsigned long long stream_id = seed * 0x9E3779B97F4A7C15ull + 1;     // one Philox stream per dropout site
    auto next_seed = [&]() { stream_id += 0xD1B54A32D192ED03ull; return stream_id; };

    if (bt->attention_mask) TRY(key_bias(reinterpret_cast<const long long*>(bt->attention_mask), P.key_bias, P.B, P.T, P.T, s));
    else CLIMB_CUDA_OK(cudaMemsetAsync(P.key_bias, 0, sizeof(float) * M, s));
    // ---- BertEmbeddings (modeling_bert.py:194-228) ----
    TRY(text_gather(reinterpret_cast<const long long*>(bt->input_ids), nullptr, reinterpret_cast<const long long*>(bt->token_type_ids),
                    F(pr->word_emb), F(pr->type_emb), F(pr->pos_emb), P.e, M, P.T, d, s));
    if (p_hid > 0.0f) {
        TRY(layernorm_fwd(P.e, d, F(pr->emb_ln_w), F(pr->emb_ln_b), dm->ln_eps, nullptr, P.x, nullptr, nullptr, M, d, CLIMB_EPI_NONE, s));
        TRY(dropout_add(P.x, nullptr, P.x, Md, p_hid, next_seed(), s));
        TRY(cast_f32_bf16(P.x, P.xb, Md, s));
    } else {
        TRY(layernorm_fwd(P.e, d, F(pr->emb_ln_w), F(pr->emb_ln_b), dm->ln_eps, P.xb, P.x, nullptr, nullptr, M, d, CLIMB_EPI_NONE, s));
    }
    // ---- BertLayer x N ----
    for (int li = 0; li < P.layers; ++li) {
        const climb_bert_layer& w = pr->layer[li];
        const bool last = li + 1 == P.layers;
        TRY(linear(M, 3 * d, d, P.xb, H(w.qkv_w), F(w.qkv_b), P.qkv, CLIMB_BF16, CLIMB_EPI_NONE, nullptr, s));
        if (p_attn > 0.0f) TRY(attention_tc_fwd(P.qkv, P.key_bias, P.ctx, P.lse, P.B, P.T, P.heads, 0.125f, s, p_attn, next_seed()));
        else TRY(attention_fwd(P.qkv, P.key_bias, P.ctx, P.lse, P.B, P.T, P.heads, 0.125f, s));
        // BertSelfOutput (:372-376): LayerNorm(dropout(dense(ctx)) + x)
        if (p_hid > 0.0f) {
            TRY(linear(M, d, d, P.ctx, H(w.o_w), F(w.o_b), P.e, CLIMB_F32, CLIMB_EPI_NONE, nullptr, s));
            TRY(dropout_add(P.e, P.x, P.e, Md, p_hid, next_seed(), s));
        } else {
            TRY(linear(M, d, d, P.ctx, H(w.o_w), F(w.o_b), P.e, CLIMB_F32, CLIMB_EPI_NONE, P.x, s));
        }
        TRY(layernorm_fwd(P.e, d, F(w.attn_ln_w), F(w.attn_ln_b), dm->ln_eps, P.x1b, P.x1, nullptr, nullptr, M, d, CLIMB_EPI_NONE, s));
        // BertIntermediate (:439-442) + BertOutput (:455-459)
        TRY(linear(M, P.ff, d, P.x1b, H(w.fc1_w), F(w.fc1_b), P.inter, CLIMB_BF16, CLIMB_EPI_GELU, nullptr, s));
        if (p_hid > 0.0f) {
            TRY(linear(M, d, P.ff, P.inter, H(w.fc2_w), F(w.fc2_b), P.e, CLIMB_F32, CLIMB_EPI_NONE, nullptr, s));
            TRY(dropout_add(P.e, P.x1, P.e, Md, p_hid, next_seed(), s));
        } else {
            TRY(linear(M, d, P.ff, P.inter, H(w.fc2_w), F(w.fc2_b), P.e, CLIMB_F32, CLIMB_EPI_NONE, P.x1, s));
        }
        TRY(layernorm_fwd(P.e, d, F(w.out_ln_w), F(w.out_ln_b), dm->ln_eps, last ? nullptr : P.xb, last ? out : P.x, nullptr,
                          nullptr, M, d, CLIMB_EPI_NONE, s));
    }
    return 0;
}

}  // namespace climb
