// Whole-encoder engine: one C call runs ViltModel.forward -> pooler_output (or its backward) as a
// fixed sequence of the sm_100a kernels in this directory on one stream. Mirrors
// modeling_vilt.py:777-899: ViltEmbeddings (:92-328) -> 12 x ViltLayer (:503-525, pre-LN) -> final
// LayerNorm -> ViltPooler, with the single active bottleneck adapter of adapters/mixins/vilt.py.
//
// Data layout in HBM (B sequences of L = T + 1 + Np tokens, M = B*L rows, d = hidden):
//   residual stream x_l          fp32 [M, d]    one per layer boundary (saved for LayerNorm backward)
//   LN outputs h1/h2, ctx        bf16 [M, d]    tensor-core A operands, re-read by the wgrads
//   qkv                          bf16 [M, 3d]   q|k|v; the attention kernels index heads in place
//   u (pre-GELU), inter          bf16 [M, ff]
//   parameters                   fp32 flat arena + bf16 shadow + fp32 gradient arena, same offsets
// Every GEMM is the tcgen05 kernel of gemm_tcgen05.cu; dgrad / wgrad read W and dY in place through
// MN-major UMMA descriptors, so no transposed copies exist anywhere.
#include "common.cuh"
#include "internal.h"

#include <cstring>

namespace climb {
namespace {

using bf16 = __nv_bfloat16;

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

struct LayerAct {
    float* x_in;        // [M, d]  input of the layer (= previous layer's output)
    bf16* h1;           // LN1(x_in)
    float *mean1, *rstd1;
    bf16* qkv;
    bf16* ctx;
    float* lse;
    float* x1;          // after the attention block
    bf16* h2;
    float *mean2, *rstd2;
    bf16* u;            // gelu'(pre-activation), saved by the FC1 epilogue for the backward
    bf16* inter;        // GELU(pre-activation)
    // adapters
    bf16* mh_in;        // O-proj output (adapter input), [M, d]
    bf16* mh_pre;       // [M, r] pre-activation
    bf16* mh_z;         // [M, r]
    bf16* out_in;       // FC2 + residual (adapter input), [M, d]
    bf16* out_pre;
    bf16* out_z;
};

struct Plan {
    int B, T, Hh, Ww, hp, wp, Np, L, M, d, ff, heads, layers, Kp, r;
    int rep, Bi;        // climb_vilt_batch.image_repeat: Bi = B / rep images, each shared by rep consecutive sequences
    const int* geom;    // [B, 2] valid patch rows / cols per image (padded batches) or null
    // embeddings
    float* key_bias;    // [B, L]
    float* text_e;      // [B*T, d] pre-LN sum
    float *text_mean, *text_rstd;
    float* text_ln;     // [B*T, d]
    bf16* im2col;       // [B*Np, Kp]
    float* pos_table;   // [Np, d]
    float* patch_out;   // [B*Np, d]
    LayerAct act[64];
    float* x_final;     // output of the last layer
    bf16* cls_ln;       // [B, d]
    float *fmean, *frstd;
    float* pooled;      // [B, d]
    float* drop_tmp;    // [M, d] dense output before dropout (hidden dropout p > 0 in train mode only)
    bool drop_hidden;
    long long bytes;
};

int fill_plan(Plan& P, const climb_vilt_dims* dm, const climb_vilt_params* pr, const climb_vilt_batch* bt,
              void* base, int save) {
    CLIMB_REQUIRE(dm && pr && bt, "engine: null descriptor");
    CLIMB_REQUIRE(dm->layers > 0 && dm->layers <= 64, "engine: layers=%d out of range", dm->layers);
    CLIMB_REQUIRE(dm->hidden % 128 == 0 && dm->hidden == dm->heads * 64,
                  "engine: hidden=%d must be heads*64 and a multiple of 128", dm->hidden);
    CLIMB_REQUIRE(dm->ffn % 8 == 0, "engine: ffn=%d must be a multiple of 8", dm->ffn);
    CLIMB_REQUIRE(bt->B > 0 && bt->T > 0, "engine: empty batch (B=%d, T=%d)", bt->B, bt->T);
    CLIMB_REQUIRE(bt->H > 0 && bt->W > 0 && bt->H % dm->patch == 0 && bt->W % dm->patch == 0,
                  "engine: image %dx%d is not a multiple of the patch size %d (fixed-resolution path)",
                  bt->H, bt->W, dm->patch);
    P.B = bt->B; P.T = bt->T; P.Hh = bt->H; P.Ww = bt->W;
    P.rep = bt->image_repeat > 1 ? bt->image_repeat : 1;
    CLIMB_REQUIRE(P.B % P.rep == 0, "engine: image_repeat=%d does not divide the %d sequences", P.rep, P.B);
    P.Bi = P.B / P.rep;
    P.hp = bt->H / dm->patch; P.wp = bt->W / dm->patch; P.Np = P.hp * P.wp;
    P.geom = bt->patch_geom;
    if (P.geom != nullptr) {
        CLIMB_REQUIRE(bt->n_patch_slots > 0 && bt->n_patch_slots <= P.Np,
                      "engine: n_patch_slots=%d outside (0, %d] for the padded %d x %d patch grid", bt->n_patch_slots, P.Np, P.hp, P.wp);
        P.Np = bt->n_patch_slots;
    }
    CLIMB_REQUIRE(bt->patch_select == nullptr || (P.geom != nullptr && P.rep == 1),
                  "engine: patch_select needs patch_geom and one image per sequence (image_repeat <= 1)");
    P.L = P.T + 1 + P.Np; P.M = P.B * P.L;
    P.d = dm->hidden; P.ff = dm->ffn; P.heads = dm->heads; P.layers = dm->layers;
    P.Kp = dm->channels * dm->patch * dm->patch;
    P.r = pr->adapter_r;
    CLIMB_REQUIRE(P.r == 0 || P.r % 8 == 0, "engine: adapter width %d must be a multiple of 8", P.r);
    const long long M = P.M, d = P.d, ff = P.ff, BT = static_cast<long long>(P.B) * P.T;
    Bump b(base);
    P.key_bias = b.take<float>(static_cast<long long>(P.B) * P.L);
    P.text_e = b.take<float>(BT * d);
    P.text_mean = b.take<float>(BT);
    P.text_rstd = b.take<float>(BT);
    P.text_ln = b.take<float>(BT * d);
    P.im2col = b.take<bf16>(static_cast<long long>(P.B) * P.Np * P.Kp);
    P.pos_table = b.take<float>(static_cast<long long>(P.Np) * d);
    P.patch_out = b.take<float>(static_cast<long long>(P.B) * P.Np * d);
    // residual stream: one buffer per layer boundary when saving for backward, else ping-pong
    float* xbuf[65];
    const int n_x = save ? P.layers + 1 : 2;
    for (int i = 0; i < n_x; ++i) xbuf[i] = b.take<float>(M * d);
    for (int l = 0; l < P.layers; ++l) {
        LayerAct& a = P.act[l];
        if (save || l == 0) {
            a.h1 = b.take<bf16>(M * d);
            a.mean1 = b.take<float>(M); a.rstd1 = b.take<float>(M);
            a.qkv = b.take<bf16>(M * 3 * d);
            a.ctx = b.take<bf16>(M * d);
            a.lse = b.take<float>(static_cast<long long>(P.B) * P.heads * P.L);
            a.x1 = b.take<float>(M * d);
            a.h2 = b.take<bf16>(M * d);
            a.mean2 = b.take<float>(M); a.rstd2 = b.take<float>(M);
            a.u = b.take<bf16>(M * ff);
            a.inter = b.take<bf16>(M * ff);
            if (P.r > 0) {
                a.mh_in = b.take<bf16>(M * d); a.mh_pre = b.take<bf16>(M * P.r); a.mh_z = b.take<bf16>(M * P.r);
                a.out_in = b.take<bf16>(M * d); a.out_pre = b.take<bf16>(M * P.r); a.out_z = b.take<bf16>(M * P.r);
            } else {
                a.mh_in = a.mh_pre = a.mh_z = a.out_in = a.out_pre = a.out_z = nullptr;
            }
        } else {
            a = P.act[0];        // inference: recycle layer 0's activation buffers
        }
        a.x_in = save ? xbuf[l] : xbuf[l & 1];
    }
    P.x_final = save ? xbuf[P.layers] : xbuf[P.layers & 1];
    P.cls_ln = b.take<bf16>(static_cast<long long>(P.B) * d);
    P.fmean = b.take<float>(P.B);
    P.frstd = b.take<float>(P.B);
    P.pooled = b.take<float>(static_cast<long long>(P.B) * d);
    P.drop_hidden = bt->training && dm->hidden_dropout > 0.0f;
    P.drop_tmp = P.drop_hidden ? b.take<float>(M * d) : nullptr;
    P.bytes = (b.off + 255) & ~255LL;
    return 0;
}

// ---- GEMM conveniences -----------------------------------------------------------------------
struct Lin {
    int M, N, K;
    const bf16* A; long long lda;
    const bf16* W;                 // [N, K] row-major
    const float* bias = nullptr;
    void* C = nullptr; int c_dtype = CLIMB_BF16; long long ldc = 0;
    int epi = CLIMB_EPI_NONE;
    bf16* aux = nullptr; long long ldaux = 0;
    const float* residual = nullptr; long long ldr = 0;
    bf16* c2 = nullptr; long long ldc2 = 0;
};

int run_linear(const Lin& l, cudaStream_t s) {
    climb_gemm_desc g;
    std::memset(&g, 0, sizeof(g));
    g.M = l.M; g.N = l.N; g.K = l.K;
    g.A = l.A; g.lda = l.lda; g.a_mn_major = 0;
    g.B = l.W; g.ldb = l.K; g.b_mn_major = 0;
    g.C = l.C; g.ldc = l.ldc ? l.ldc : l.N; g.c_dtype = l.c_dtype;
    g.bias = l.bias; g.residual = l.residual; g.ldr = l.ldr ? l.ldr : l.N;
    g.epilogue = l.epi; g.aux = l.aux; g.ldaux = l.ldaux ? l.ldaux : l.N;
    g.c2 = l.c2; g.ldc2 = l.ldc2 ? l.ldc2 : l.N;
    g.alpha = 1.0f;
    return gemm_bf16(&g, s);
}

// dX[M, K] = epi(dY[M, N] * W[N, K]) (+ residual); W read in place (MN-major B operand)
int run_dgrad(int M, int N, int K, const bf16* dY, const bf16* W, void* dX, int c_dtype, int epi, const bf16* aux,
              long long ldaux, const float* residual, bf16* c2, cudaStream_t s, float* colsum_out = nullptr) {
    climb_gemm_desc g;
    std::memset(&g, 0, sizeof(g));
    g.M = M; g.N = K; g.K = N;
    g.A = dY; g.lda = N; g.a_mn_major = 0;
    g.B = W; g.ldb = K; g.b_mn_major = 1;
    g.C = dX; g.ldc = K; g.c_dtype = c_dtype;
    g.epilogue = epi; g.aux = const_cast<bf16*>(aux); g.ldaux = ldaux;
    g.residual = residual; g.ldr = K;
    g.c2 = c2; g.ldc2 = K;
    g.colsum = colsum_out;
    g.alpha = 1.0f;
    return gemm_bf16(&g, s);
}

// dW[N, K] += dY[M, N]^T * X[M, K]; both operands read in place (MN-major), split over tokens
int run_wgrad(int M, int N, int K, const bf16* dY, long long lddy, const bf16* X, long long ldx, float* dW,
              cudaStream_t s, int independent = 0, float* dbias = nullptr) {
    climb_gemm_desc g;
    std::memset(&g, 0, sizeof(g));
    g.M = N; g.N = K; g.K = M;
    g.A = dY; g.lda = lddy; g.a_mn_major = 1;
    g.B = X; g.ldb = ldx; g.b_mn_major = 1;
    g.C = dW; g.ldc = K; g.c_dtype = CLIMB_F32;
    g.alpha = 1.0f; g.accumulate = 1; g.split_k = 0;
    g.independent = independent;
    g.colsum_a = dbias;           // the Linear's bias gradient = column sums of dY, summed inside the weight-gradient kernel
    return gemm_bf16(&g, s);
}

#define TRY(expr)                 \
    do {                          \
        int _rc = (expr);         \
        if (_rc) return _rc;      \
    } while (0)

inline const float* F(const float* theta, long long off) { return off >= 0 ? theta + off : nullptr; }
inline const bf16* H(const void* shadow, long long off) {
    return off >= 0 ? static_cast<const bf16*>(shadow) + off : nullptr;
}
inline float* G(float* grad, long long off) { return off >= 0 ? grad + off : nullptr; }

}  // namespace

long long vilt_forward_workspace_bytes(const climb_vilt_dims* dims, const climb_vilt_params* params,
                                       const climb_vilt_batch* batch, int save) {
    if (dims && dims->precision == CLIMB_PREC_BF16X3) return vilt_forward_workspace_bytes_precise(dims, params, batch, save);
    Plan P;
    if (fill_plan(P, dims, params, batch, nullptr, save)) return -1;
    return P.bytes;
}

struct BwdScratch {
    float *dxa, *dxb;           // ping-pong gradient of the residual stream, fp32 [M, d]
    bf16 *dxa_h, *dxb_h;        // bf16 copies (tensor-core operands)
    bf16* du;                   // [M, ff]
    bf16* dh;                   // [M, d]  (dh2 / dctx / dh1 in turn)
    bf16* dqkv;                 // [M, 3d]
    float* delta;               // [B, H, L]
    bf16* dz;                   // [M, r]
    bf16* dmh;                  // [M, d] gradient at the O-proj output when the mh adapter is on
    float* dy_text;             // [B*T, d]
    float* de_text;             // [B*T, d]
    bf16* dpatch;               // [B*Np, d]
    float* S;                   // [2, L, d]
    bf16* dpool;                // [B, d]
    bf16* dcls;                 // [B, d]
    bf16* dm_h;                 // [M, d]  gradient behind a hidden-dropout site (p > 0 only)
    float* dxm;                 // [M, d]  embedding gradient behind the embedding dropout (p > 0 only)
    long long bytes;
};

static void fill_scratch(BwdScratch& S, const Plan& P, void* base) {
    Bump b(base);
    const long long M = P.M, d = P.d;
    S.dxa = b.take<float>(M * d); S.dxb = b.take<float>(M * d);
    S.dxa_h = b.take<bf16>(M * d); S.dxb_h = b.take<bf16>(M * d);
    S.du = b.take<bf16>(M * P.ff);
    S.dh = b.take<bf16>(M * d);
    S.dqkv = b.take<bf16>(M * 3 * d);
    S.delta = b.take<float>(static_cast<long long>(P.B) * P.heads * P.L);
    S.dz = b.take<bf16>(M * (P.r > 0 ? P.r : 8));
    S.dmh = b.take<bf16>(P.r > 0 ? M * d : 8);
    S.dy_text = b.take<float>(static_cast<long long>(P.B) * P.T * d);
    S.de_text = b.take<float>(static_cast<long long>(P.B) * P.T * d);
    S.dpatch = b.take<bf16>(static_cast<long long>(P.B) * P.Np * d);
    S.S = b.take<float>(2LL * P.L * d);
    S.dpool = b.take<bf16>(static_cast<long long>(P.B) * d);
    S.dcls = b.take<bf16>(static_cast<long long>(P.B) * d);
    S.dm_h = P.drop_hidden ? b.take<bf16>(M * d) : nullptr;
    S.dxm = P.drop_hidden ? b.take<float>(M * d) : nullptr;
    S.bytes = (b.off + 255) & ~255LL;
}

long long vilt_backward_scratch_bytes(const climb_vilt_dims* dims, const climb_vilt_params* params,
                                      const climb_vilt_batch* batch) {
    if (dims && dims->precision == CLIMB_PREC_BF16X3) return vilt_backward_scratch_bytes_precise(dims, params, batch);
    Plan P;
    if (fill_plan(P, dims, params, batch, nullptr, 1)) return -1;
    BwdScratch S;
    fill_scratch(S, P, nullptr);
    return S.bytes;
}

int vilt_forward(const climb_vilt_dims* dm, const climb_vilt_params* pr, const climb_vilt_batch* bt,
                 const float* theta, const void* shadow, void* workspace, long long workspace_bytes, int save,
                 float* pooled_out, cudaStream_t s) {
    if (dm && dm->precision == CLIMB_PREC_BF16X3)
        return vilt_forward_precise(dm, pr, bt, theta, shadow, workspace, workspace_bytes, save, pooled_out, s);
    Plan P;
    TRY(fill_plan(P, dm, pr, bt, workspace, save));
    CLIMB_REQUIRE(theta && shadow && workspace && pooled_out, "vilt_forward: null buffer");
    CLIMB_REQUIRE(workspace_bytes >= P.bytes, "vilt_forward: workspace %lld < required %lld", workspace_bytes, P.bytes);
    CLIMB_REQUIRE((bt->input_ids != nullptr) != (bt->inputs_embeds != nullptr),
                  "vilt_forward: exactly one of input_ids / inputs_embeds");
    CLIMB_REQUIRE(bt->pixel_values != nullptr, "vilt_forward: pixel_values missing");
    CLIMB_REQUIRE(bt->image_type_idx != nullptr ||
                      (bt->image_type_idx_scalar >= 0 && bt->image_type_idx_scalar < dm->n_modality),
                  "vilt_forward: image_token_type_idx %d outside the %d-row modality table",
                  bt->image_type_idx_scalar, dm->n_modality);
    const int d = P.d, M = P.M, BT = P.B * P.T;
    // dropout (train mode, ViltConfig.hidden_dropout_prob / attention_probs_dropout_prob > 0: csrc/dropout.cu)
    const float p_h = bt->training ? dm->hidden_dropout : 0.0f, p_a = bt->training ? dm->attn_dropout : 0.0f;
    CLIMB_REQUIRE(p_h >= 0.0f && p_h < 1.0f && p_a >= 0.0f && p_a < 1.0f, "vilt_forward: dropout probabilities must lie in [0, 1)");
    CLIMB_REQUIRE(p_a == 0.0f || P.L <= 256, "vilt_forward: dropout on the attention probabilities is implemented for sequences of "
                  "up to 256 tokens (L = %d)", P.L);
    const unsigned long long dseed = bt->dropout_seed;

    // ---- embeddings (modeling_vilt.py:207-246) ----
    if (P.geom) TRY(key_bias_ragged(reinterpret_cast<const long long*>(bt->attention_mask), P.geom, P.key_bias, P.B, P.T, P.L, s, bt->patch_select));
    else if (bt->attention_mask) TRY(key_bias(reinterpret_cast<const long long*>(bt->attention_mask), P.key_bias, P.B, P.T, P.L, s));
    else CLIMB_CUDA_OK(cudaMemsetAsync(P.key_bias, 0, sizeof(float) * P.B * P.L, s));
    TRY(text_gather(reinterpret_cast<const long long*>(bt->input_ids), bt->inputs_embeds,
                    reinterpret_cast<const long long*>(bt->token_type_ids), F(theta, pr->word_emb),
                    F(theta, pr->text_type_emb), F(theta, pr->text_pos_emb), P.text_e, BT, P.T, d, s, dm->vocab_size,
                    dm->type_vocab_size));
    TRY(layernorm_fwd(P.text_e, d, F(theta, pr->text_ln_w), F(theta, pr->text_ln_b), dm->ln_eps, nullptr, P.text_ln,
                      P.text_mean, P.text_rstd, BT, d, CLIMB_EPI_NONE, s));
    // image_repeat > 1 (VCR: four answer choices per image, vilt.py:334-347): the patch projection runs once per image
    if (P.geom) TRY(im2col_ragged(bt->pixel_values, P.geom, P.im2col, P.Bi, dm->channels, P.Hh, P.Ww, dm->patch, P.Np, s, P.rep, bt->patch_select));
    else TRY(im2col(bt->pixel_values, P.im2col, P.Bi, dm->channels, P.Hh, P.Ww, dm->patch, s));
    {
        Lin l{P.Bi * P.Np, d, P.Kp, P.im2col, P.Kp, H(shadow, pr->patch_w)};
        l.bias = F(theta, pr->patch_b); l.C = P.patch_out; l.c_dtype = CLIMB_F32;
        TRY(run_linear(l, s));
    }
    if (P.geom) {
        TRY(embed_assemble_ragged(P.text_ln, P.patch_out, P.geom, F(theta, pr->cls_token), F(theta, pr->pos_emb),
                                  F(theta, pr->mod_emb), bt->image_type_idx, bt->image_type_idx_scalar, P.act[0].x_in, P.B,
                                  P.T, P.Np, dm->pos_grid, d, s, dm->n_modality, p_h, dropout_site_seed(dseed, -1, 1), P.rep, bt->patch_select));
    } else {
        TRY(pos_interp(F(theta, pr->pos_emb), P.pos_table, P.hp, P.wp, dm->pos_grid, d, s));
        TRY(embed_assemble(P.text_ln, P.patch_out, P.pos_table, F(theta, pr->cls_token), F(theta, pr->pos_emb),
                           F(theta, pr->mod_emb), bt->image_type_idx, bt->image_type_idx_scalar, P.act[0].x_in, P.B, P.T,
                           P.Np, d, s, dm->n_modality, p_h, dropout_site_seed(dseed, -1, 1), P.rep));
    }

    // ---- encoder layers (modeling_vilt.py:503-525) ----
    const float scale = 0.125f;   // 1 / sqrt(64)
    const bool fused_ad = P.r > 0 && adapter_fused_ok(d, P.r);      // one-launch bottleneck (gemm_tcgen05.cu: adapter_fused_kernel)
    for (int li = 0; li < P.layers; ++li) {
        const climb_vilt_layer& w = pr->layer[li];
        LayerAct& a = P.act[li];
        float* x_out = (li + 1 < P.layers) ? P.act[li + 1].x_in : P.x_final;
        const bool mh_ad = P.r > 0 && w.mh_down_w >= 0;
        const bool out_ad = P.r > 0 && w.out_down_w >= 0;
        TRY(layernorm_fwd(a.x_in, d, F(theta, w.ln1_w), F(theta, w.ln1_b), dm->ln_eps, a.h1, nullptr, a.mean1, a.rstd1,
                          M, d, CLIMB_EPI_NONE, s));
        {
            Lin l{M, 3 * d, d, a.h1, d, H(shadow, w.qkv_w)};
            l.bias = F(theta, w.qkv_b); l.C = a.qkv;
            TRY(run_linear(l, s));
        }
        if (p_a > 0.0f) TRY(attention_tc_fwd(a.qkv, P.key_bias, a.ctx, a.lse, P.B, P.L, P.heads, scale, s, p_a, dropout_site_seed(dseed, li, 0)));
        else TRY(attention_fwd(a.qkv, P.key_bias, a.ctx, a.lse, P.B, P.L, P.heads, scale, s));
        if (p_h > 0.0f) {
            // ViltSelfOutput: dropout(dense(ctx)) before the adapter / residual (modeling_vilt.py:407-414)
            Lin l{M, d, d, a.ctx, d, H(shadow, w.o_w)};
            l.bias = F(theta, w.o_b); l.C = P.drop_tmp; l.c_dtype = CLIMB_F32;
            TRY(run_linear(l, s));
            TRY(dropout_res(P.drop_tmp, a.x_in, a.x1, mh_ad ? a.mh_in : nullptr, nullptr, static_cast<long long>(M) * d, p_h,
                            dropout_site_seed(dseed, li, 1), s));
        } else {
            Lin l{M, d, d, a.ctx, d, H(shadow, w.o_w)};
            l.bias = F(theta, w.o_b); l.C = a.x1; l.c_dtype = CLIMB_F32; l.residual = a.x_in;
            if (mh_ad) l.aux = a.mh_in;                       // adapter sees O(ctx)+b before the residual
            TRY(run_linear(l, s));
        }
        if (mh_ad && fused_ad) {
            // x1 (= x_in + h) += W_u act(W_d h + b_d) + b_u in ONE launch: the r-wide intermediate stays on the SM
            TRY(adapter_fused(0, M, d, P.r, pr->adapter_act, a.mh_in, H(shadow, w.mh_down_w), H(shadow, w.mh_up_w), F(theta, w.mh_down_b),
                              F(theta, w.mh_up_b), a.mh_pre, a.mh_z, a.x1, a.x1, nullptr, nullptr, s));
        } else if (mh_ad) {
            Lin dn{M, P.r, d, a.mh_in, d, H(shadow, w.mh_down_w)};
            dn.bias = F(theta, w.mh_down_b); dn.C = a.mh_z; dn.epi = pr->adapter_act; dn.aux = a.mh_pre;
            TRY(run_linear(dn, s));
            Lin up{M, d, P.r, a.mh_z, P.r, H(shadow, w.mh_up_w)};
            up.bias = F(theta, w.mh_up_b); up.C = a.x1; up.c_dtype = CLIMB_F32; up.residual = a.x1;
            TRY(run_linear(up, s));
        }
        TRY(layernorm_fwd(a.x1, d, F(theta, w.ln2_w), F(theta, w.ln2_b), dm->ln_eps, a.h2, nullptr, a.mean2, a.rstd2, M,
                          d, CLIMB_EPI_NONE, s));
        {
            Lin l{M, P.ff, d, a.h2, d, H(shadow, w.fc1_w)};
            l.bias = F(theta, w.fc1_b); l.C = a.inter; l.epi = CLIMB_EPI_GELU_SAVE_GRAD; l.aux = a.u;   // u <- gelu'(pre)
            TRY(run_linear(l, s));
        }
        if (p_h > 0.0f) {
            // ViltOutput: dropout(dense(inter)) + residual (modeling_vilt.py:480-484)
            Lin l{M, d, P.ff, a.inter, P.ff, H(shadow, w.fc2_w)};
            l.bias = F(theta, w.fc2_b); l.C = P.drop_tmp; l.c_dtype = CLIMB_F32;
            TRY(run_linear(l, s));
            TRY(dropout_res(P.drop_tmp, a.x1, x_out, nullptr, out_ad ? a.out_in : nullptr, static_cast<long long>(M) * d, p_h,
                            dropout_site_seed(dseed, li, 2), s));
        } else {
            Lin l{M, d, P.ff, a.inter, P.ff, H(shadow, w.fc2_w)};
            l.bias = F(theta, w.fc2_b); l.C = x_out; l.c_dtype = CLIMB_F32; l.residual = a.x1;
            if (out_ad) l.c2 = a.out_in;                      // adapter sees FC2 + residual
            TRY(run_linear(l, s));
        }
        if (out_ad && fused_ad) {
            TRY(adapter_fused(0, M, d, P.r, pr->adapter_act, a.out_in, H(shadow, w.out_down_w), H(shadow, w.out_up_w), F(theta, w.out_down_b),
                              F(theta, w.out_up_b), a.out_pre, a.out_z, x_out, x_out, nullptr, nullptr, s));
        } else if (out_ad) {
            Lin dn{M, P.r, d, a.out_in, d, H(shadow, w.out_down_w)};
            dn.bias = F(theta, w.out_down_b); dn.C = a.out_z; dn.epi = pr->adapter_act; dn.aux = a.out_pre;
            TRY(run_linear(dn, s));
            Lin up{M, d, P.r, a.out_z, P.r, H(shadow, w.out_up_w)};
            up.bias = F(theta, w.out_up_b); up.C = x_out; up.c_dtype = CLIMB_F32; up.residual = x_out;
            TRY(run_linear(up, s));
        }
    }

    // ---- final LayerNorm on the [CLS] rows + pooler (modeling_vilt.py:873-874, 887-899) ----
    TRY(layernorm_fwd(P.x_final, static_cast<long long>(P.L) * d, F(theta, pr->final_ln_w), F(theta, pr->final_ln_b),
                      dm->ln_eps, P.cls_ln, nullptr, P.fmean, P.frstd, P.B, d, CLIMB_EPI_NONE, s));
    {
        Lin l{P.B, d, d, P.cls_ln, d, H(shadow, pr->pooler_w)};
        l.bias = F(theta, pr->pooler_b); l.C = P.pooled; l.c_dtype = CLIMB_F32; l.epi = CLIMB_EPI_TANH;
        TRY(run_linear(l, s));
    }
    CLIMB_CUDA_OK(cudaMemcpyAsync(pooled_out, P.pooled, sizeof(float) * P.B * d, cudaMemcpyDeviceToDevice, s));
    return 0;
}

int vilt_backward(const climb_vilt_dims* dm, const climb_vilt_params* pr, const climb_vilt_batch* bt,
                  const float* theta, const void* shadow, const void* workspace, long long workspace_bytes,
                  void* scratch, long long scratch_bytes, const float* dpooled, float* grad, int first_layer,
                  int last_layer, int parts, cudaStream_t s) {
    if (dm && dm->precision == CLIMB_PREC_BF16X3)
        return vilt_backward_precise(dm, pr, bt, theta, shadow, workspace, workspace_bytes, scratch, scratch_bytes, dpooled, grad,
                                     first_layer, last_layer, parts, s);
    Plan P;
    TRY(fill_plan(P, dm, pr, bt, const_cast<void*>(workspace), 1));
    CLIMB_REQUIRE(theta && shadow && workspace && scratch && dpooled && grad, "vilt_backward: null buffer");
    CLIMB_REQUIRE(workspace_bytes >= P.bytes, "vilt_backward: workspace %lld < required %lld", workspace_bytes, P.bytes);
    BwdScratch S;
    fill_scratch(S, P, scratch);
    CLIMB_REQUIRE(scratch_bytes >= S.bytes, "vilt_backward: scratch %lld < required %lld", scratch_bytes, S.bytes);
    const int d = P.d, M = P.M, ff = P.ff, r = P.r, BT = P.B * P.T;
    const int dact = pr->adapter_act == CLIMB_EPI_RELU ? CLIMB_EPI_DRELU : CLIMB_EPI_DSWISH;
    const float p_h = bt->training ? dm->hidden_dropout : 0.0f, p_a = bt->training ? dm->attn_dropout : 0.0f;
    const unsigned long long dseed = bt->dropout_seed;
    const bool fused_ad = r > 0 && adapter_fused_ok(d, r);

    // lowest layer that still needs a gradient (everything below is skipped)
    int lowest = P.layers;
    if (pr->embed_flags & CLIMB_TRAIN_BASE) lowest = 0;
    else
        for (int l = 0; l < P.layers; ++l)
            if (pr->layer[l].flags & (CLIMB_TRAIN_BASE | CLIMB_TRAIN_ADAPTER)) { lowest = l; break; }
    const bool tail = (pr->tail_flags & CLIMB_TRAIN_BASE) != 0;
    if (lowest == P.layers && !tail) return 0;      // nothing inside the encoder is trainable
    CLIMB_REQUIRE(first_layer < P.layers && last_layer >= 0 && (first_layer >= last_layer || first_layer < 0),
                  "vilt_backward: bad layer range [%d, %d]", first_layer, last_layer);

    // The gradient of the residual stream lives in S.dxa / S.dxa_h at every layer boundary, so the pass
    // can be issued in several calls (top layers first) with gradient all-reduces launched in between.
    float* dx = S.dxa;  bf16* dx_h = S.dxa_h;      // gradient w.r.t. the current layer's output
    float* dn = S.dxb;  bf16* dn_h = S.dxb_h;      // scratch for the next one
    if (parts & CLIMB_BWD_TAIL) {
    // ---- pooler + final LayerNorm ----
    TRY(tanh_bwd(dpooled, P.pooled, S.dpool, static_cast<long long>(P.B) * d, s));
    if (tail) {
        TRY(run_wgrad(P.B, d, d, S.dpool, d, P.cls_ln, d, G(grad, pr->pooler_w), s));
        TRY(colsum(S.dpool, CLIMB_BF16, d, P.B, d, G(grad, pr->pooler_b), s));
    }
    if (lowest == P.layers) {
        // only the tail trains: still need d(final LN params)
        TRY(run_dgrad(P.B, d, d, S.dpool, H(shadow, pr->pooler_w), S.dcls, CLIMB_BF16, CLIMB_EPI_NONE, nullptr, 0, nullptr, nullptr, s));
        TRY(layernorm_bwd(nullptr, S.dcls, P.x_final, static_cast<long long>(P.L) * d, F(theta, pr->final_ln_w),
                          F(theta, pr->final_ln_b), P.fmean, P.frstd, nullptr, nullptr, nullptr,
                          G(grad, pr->final_ln_w), G(grad, pr->final_ln_b), P.B, d, CLIMB_EPI_NONE, s));
        return 0;
    }
    TRY(run_dgrad(P.B, d, d, S.dpool, H(shadow, pr->pooler_w), S.dcls, CLIMB_BF16, CLIMB_EPI_NONE, nullptr, 0, nullptr, nullptr, s));
    CLIMB_CUDA_OK(cudaMemsetAsync(S.dxa, 0, sizeof(float) * M * d, s));
    TRY(layernorm_bwd(nullptr, S.dcls, P.x_final, static_cast<long long>(P.L) * d, F(theta, pr->final_ln_w),
                      F(theta, pr->final_ln_b), P.fmean, P.frstd, nullptr, S.dxa, nullptr,
                      tail ? G(grad, pr->final_ln_w) : nullptr, tail ? G(grad, pr->final_ln_b) : nullptr, P.B, d,
                      CLIMB_EPI_NONE, s));
    TRY(cast_f32_bf16(S.dxa, S.dxa_h, static_cast<long long>(M) * d, s));
    }   // CLIMB_BWD_TAIL
    if (lowest == P.layers) return 0;

    for (int li = first_layer; li >= last_layer && li >= lowest; --li) {
        const climb_vilt_layer& w = pr->layer[li];
        const LayerAct& a = P.act[li];
        const bool base = (w.flags & CLIMB_TRAIN_BASE) != 0;
        const bool adp = (w.flags & CLIMB_TRAIN_ADAPTER) != 0;
        const bool mh_ad = r > 0 && w.mh_down_w >= 0;
        const bool out_ad = r > 0 && w.out_down_w >= 0;

        // ---- output adapter: out = y + up(act(down(y))) ----
        if (out_ad && fused_ad) {
            // the up-projection's gradients need dx_h as it is NOW (the fused launch refreshes it in place)
            if (adp) {
                TRY(run_wgrad(M, d, r, dx_h, d, a.out_z, r, G(grad, w.out_up_w), s));
                TRY(colsum(dx_h, CLIMB_BF16, d, M, d, G(grad, w.out_up_b), s));
            }
            // dpre = (dx W_u) * act'(pre) -> S.dz (+ its column sums = the down-bias gradient); dx += dpre W_d in place, dx_h refreshed
            TRY(adapter_fused(1, M, d, r, pr->adapter_act, dx_h, H(shadow, w.out_down_w), H(shadow, w.out_up_w), nullptr, nullptr,
                              a.out_pre, S.dz, dx, dx, dx_h, adp ? G(grad, w.out_down_b) : nullptr, s));
            if (adp) TRY(run_wgrad(M, r, d, S.dz, r, a.out_in, d, G(grad, w.out_down_w), s));
        } else if (out_ad) {
            TRY(run_dgrad(M, d, r, dx_h, H(shadow, w.out_up_w), S.dz, CLIMB_BF16, dact, a.out_pre, r, nullptr, nullptr, s));
            if (adp) {
                TRY(run_wgrad(M, d, r, dx_h, d, a.out_z, r, G(grad, w.out_up_w), s));
                TRY(colsum(dx_h, CLIMB_BF16, d, M, d, G(grad, w.out_up_b), s));
                TRY(run_wgrad(M, r, d, S.dz, r, a.out_in, d, G(grad, w.out_down_w), s));
                TRY(colsum(S.dz, CLIMB_BF16, r, M, r, G(grad, w.out_down_b), s));
            }
            // dy = dx + dz Wd  (in place over dx, bf16 copy refreshed)
            TRY(run_dgrad(M, r, d, S.dz, H(shadow, w.out_down_w), dx, CLIMB_F32, CLIMB_EPI_NONE, nullptr, 0, dx, dx_h, s));
        }
        // ---- FFN: y = FC2(GELU(FC1(LN2(x1)))) + x1 ----
        // du = (dx W2) * gelu'(pre). The bias gradients are column sums = streaming passes over dx / du. (Fused into
        // the K = 768 epilogue the FC1 sums cost 57 us per launch, a third of that kernel's instructions; issued as
        // INDEPENDENT background launches beside the dgrad GEMMs they measured 2 % slower on the whole step than as
        // ordinary 8-13 us passes: the co-resident CTAs take issue slots and L2 bandwidth from an epilogue-bound GEMM.)
        const bf16* dy_h = dx_h;               // gradient at the FC2 output: behind the hidden dropout when it is on
        if (p_h > 0.0f) {
            TRY(dropout_mask_bf16(dx_h, S.dm_h, static_cast<long long>(M) * d, p_h, dropout_site_seed(dseed, li, 2), s));
            dy_h = S.dm_h;
        }
        TRY(run_dgrad(M, d, ff, dy_h, H(shadow, w.fc2_w), S.du, CLIMB_BF16, CLIMB_EPI_MUL_AUX, a.u, ff, nullptr, nullptr, s,
                      nullptr));
        if (base) {
            // the two bias gradients (column sums of dy / du) ride inside the weight-gradient kernels: one more MMA per
            // k-step against a tile of ones instead of two streaming passes (gemm_pair_wgrad_kernel, colsum_a)
            TRY(run_wgrad(M, d, ff, dy_h, d, a.inter, ff, G(grad, w.fc2_w), s, 0, G(grad, w.fc2_b)));
            TRY(run_wgrad(M, ff, d, S.du, ff, a.h2, d, G(grad, w.fc1_w), s, 0, G(grad, w.fc1_b)));
        }
        TRY(run_dgrad(M, ff, d, S.du, H(shadow, w.fc1_w), S.dh, CLIMB_BF16, CLIMB_EPI_NONE, nullptr, 0, nullptr, nullptr, s));
        // dx1 = dx + LN2'(dh2)
        // (layernorm_bwd can also emit the column sums of its output = the o_b / fc2_b gradients; measured on B200 the
        // fused variant costs 15 us more per launch than the 8 us streaming colsum it replaces, so it is not used here)
        TRY(layernorm_bwd(nullptr, S.dh, a.x1, d, F(theta, w.ln2_w), F(theta, w.ln2_b), a.mean2, a.rstd2, dx, dn, dn_h,
                          base ? G(grad, w.ln2_w) : nullptr, base ? G(grad, w.ln2_b) : nullptr, M, d, CLIMB_EPI_NONE, s));
        // now dn / dn_h = dx1
        // ---- attention block: x1 = x + A, A = h (+ adapter), h = O(ctx) + b ----
        const bf16* dho = dn_h;            // gradient at the O-proj output
        if (mh_ad && fused_ad) {
            if (adp) {
                TRY(run_wgrad(M, d, r, dn_h, d, a.mh_z, r, G(grad, w.mh_up_w), s));
                TRY(colsum(dn_h, CLIMB_BF16, d, M, d, G(grad, w.mh_up_b), s));
            }
            // gradient at the O-proj output: dmh (bf16) = dn + ((dn W_u) * act'(pre)) W_d; dn itself (the residual path) is untouched
            TRY(adapter_fused(1, M, d, r, pr->adapter_act, dn_h, H(shadow, w.mh_down_w), H(shadow, w.mh_up_w), nullptr, nullptr,
                              a.mh_pre, S.dz, dn, nullptr, S.dmh, adp ? G(grad, w.mh_down_b) : nullptr, s));
            if (adp) TRY(run_wgrad(M, r, d, S.dz, r, a.mh_in, d, G(grad, w.mh_down_w), s));
            dho = S.dmh;
        } else if (mh_ad) {
            TRY(run_dgrad(M, d, r, dn_h, H(shadow, w.mh_up_w), S.dz, CLIMB_BF16, dact, a.mh_pre, r, nullptr, nullptr, s));
            if (adp) {
                TRY(run_wgrad(M, d, r, dn_h, d, a.mh_z, r, G(grad, w.mh_up_w), s));
                TRY(colsum(dn_h, CLIMB_BF16, d, M, d, G(grad, w.mh_up_b), s));
                TRY(run_wgrad(M, r, d, S.dz, r, a.mh_in, d, G(grad, w.mh_down_w), s));
                TRY(colsum(S.dz, CLIMB_BF16, r, M, r, G(grad, w.mh_down_b), s));
            }
            TRY(run_dgrad(M, r, d, S.dz, H(shadow, w.mh_down_w), S.dmh, CLIMB_BF16, CLIMB_EPI_NONE, nullptr, 0, dn, nullptr, s));
            dho = S.dmh;
        }
        if (p_h > 0.0f) {                       // behind ViltSelfOutput's dropout
            TRY(dropout_mask_bf16(dho, S.dm_h, static_cast<long long>(M) * d, p_h, dropout_site_seed(dseed, li, 1), s));
            dho = S.dm_h;
        }
        TRY(run_dgrad(M, d, d, dho, H(shadow, w.o_w), S.dh, CLIMB_BF16, CLIMB_EPI_NONE, nullptr, 0, nullptr, nullptr, s));   // dctx
        if (p_a > 0.0f) {
            // dropout on the probabilities: mask regenerated inside the kernel; its analytic v-bias shortcut (rows of P sum to
            // one) does not hold for the dropped probabilities, so the q | k | v bias gradients are plain column sums of dqkv
            TRY(attention_tc_bwd(a.qkv, P.key_bias, a.ctx, S.dh, a.lse, S.dqkv, nullptr, P.B, P.L, P.heads, 0.125f, s, p_a,
                                 dropout_site_seed(dseed, li, 0)));
            if (base) TRY(colsum(S.dqkv, CLIMB_BF16, 3 * d, M, 3 * d, G(grad, w.qkv_b), s));
        } else {
        // the q/k/v bias gradient (column sums of dqkv) comes out of the attention backward's epilogue
        TRY(attention_bwd(a.qkv, P.key_bias, a.ctx, S.dh, a.lse, S.delta, S.dqkv, base ? G(grad, w.qkv_b) : nullptr, P.B,
                          P.L, P.heads, 0.125f, s));
        }
        if (base) {
            // dW_o = dho^T ctx depends on nothing the attention backward writes: issued right behind it as an
            // INDEPENDENT launch, its CTAs fill the SMs that kernel's last partial wave leaves idle
            TRY(run_wgrad(M, d, d, dho, d, a.ctx, d, G(grad, w.o_w), s, /*independent=*/1, G(grad, w.o_b)));
        }
        if (base) TRY(run_wgrad(M, 3 * d, d, S.dqkv, 3 * d, a.h1, d, G(grad, w.qkv_w), s));
        TRY(run_dgrad(M, 3 * d, d, S.dqkv, H(shadow, w.qkv_w), S.dh, CLIMB_BF16, CLIMB_EPI_NONE, nullptr, 0, nullptr, nullptr, s));  // dh1
        // dx_in = dx1 + LN1'(dh1)   (written over the old dx buffers)
        TRY(layernorm_bwd(nullptr, S.dh, a.x_in, d, F(theta, w.ln1_w), F(theta, w.ln1_b), a.mean1, a.rstd1, dn, dx, dx_h,
                          base ? G(grad, w.ln1_w) : nullptr, base ? G(grad, w.ln1_b) : nullptr, M, d, CLIMB_EPI_NONE, s));
        // dx / dx_h now hold the gradient w.r.t. this layer's input = next iteration's output grad
    }

    // ---- embeddings ----
    if ((parts & CLIMB_BWD_EMBED) && (pr->embed_flags & CLIMB_TRAIN_BASE)) {
        const float* dxe = dx;                  // gradient of the content rows: behind the embedding dropout when it is on
        if (p_h > 0.0f) {
            TRY(dropout_mask_f32(dx, S.dxm, static_cast<long long>(M) * d, p_h, dropout_site_seed(dseed, -1, 1), s));
            dxe = S.dxm;
            // the modality-type rows are added AFTER the dropout: their gradient is the unmasked one
            TRY(embed_reduce_bwd(dx, bt->image_type_idx, bt->image_type_idx_scalar, S.S, nullptr, nullptr, G(grad, pr->mod_emb),
                                 nullptr, dm->n_modality, P.B, P.T, P.hp, P.wp, dm->pos_grid, d, s, P.geom, P.geom ? P.Np : 0, bt->patch_select));
        }
        TRY(embed_split_bwd(dxe, S.dy_text, S.dpatch, P.B, P.T, P.Np, d, s, P.rep));
        TRY(embed_reduce_bwd(dxe, bt->image_type_idx, bt->image_type_idx_scalar, S.S, G(grad, pr->cls_token),
                             G(grad, pr->pos_emb), p_h > 0.0f ? nullptr : G(grad, pr->mod_emb), G(grad, pr->patch_b), dm->n_modality,
                             P.B, P.T, P.hp, P.wp, dm->pos_grid, d, s, P.geom, P.geom ? P.Np : 0, bt->patch_select));
        TRY(layernorm_bwd(S.dy_text, nullptr, P.text_e, d, F(theta, pr->text_ln_w), F(theta, pr->text_ln_b), P.text_mean,
                          P.text_rstd, nullptr, S.de_text, nullptr, G(grad, pr->text_ln_w), G(grad, pr->text_ln_b), BT, d,
                          CLIMB_EPI_NONE, s));
        TRY(text_scatter_bwd(S.de_text, reinterpret_cast<const long long*>(bt->input_ids),
                             reinterpret_cast<const long long*>(bt->token_type_ids),
                             bt->input_ids ? G(grad, pr->word_emb) : nullptr, G(grad, pr->text_type_emb),
                             G(grad, pr->text_pos_emb), BT, P.T, d, s, dm->vocab_size, dm->type_vocab_size));
        TRY(run_wgrad(P.Bi * P.Np, d, P.Kp, S.dpatch, d, P.im2col, P.Kp, G(grad, pr->patch_w), s));
    }
    return 0;
}

}  // namespace climb
