// Embedding kernels of the ViLT hot path (ViltEmbeddings / TextEmbeddings / PatchEmbeddings,
// modeling_vilt.py:92-328) on the fixed-resolution path (one H x W per batch, pixel_mask all ones):
//   text : word[ids] (or inputs_embeds) + segment[tt] + position[t]            (:292-301; LN runs in layernorm.cu)
//   image: im2col of the 32x32/stride-32 conv (:309-328) so that it becomes one tcgen05 GEMM,
//          bilinear (align_corners) resize of the position table (:130-147), [cls] + pos[0] (:195-200)
//   both : + modality-type rows and concatenation text || image (:231-246), written straight into the
//          [B, L, d] fp32 residual stream.
// Backward kernels reduce the gradient of that stream onto the embedding tables. All HBM-bound:
// float4 / uint4 accesses, one pass over the data.
#include "common.cuh"
#include "climb_b200.h"
#include "internal.h"

namespace climb {
namespace {

// ids outside their table are clamped to row 0 and reported through the sticky error word (climb_error_flags): nn.Embedding
// raises IndexError in the reference; a silent out-of-bounds read (or, in the backward, a corrupting write) must not happen
__device__ __forceinline__ long long checked_index(long long v, int n, unsigned int* err, unsigned int bit) {
    if (n > 0 && (v < 0 || v >= n)) {
        if (err != nullptr) atomicOr_system(err, bit);
        return 0;
    }
    return v;
}

__global__ void text_gather_kernel(const long long* __restrict__ ids, const float* __restrict__ inputs_embeds,
                                   const long long* __restrict__ tt, const float* __restrict__ word,
                                   const float* __restrict__ type_emb, const float* __restrict__ pos,
                                   float* __restrict__ e, int rows, int T, int d4, int vocab, int n_types, unsigned int* err) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(rows) * d4) return;
    const int r = static_cast<int>(i / d4), c = static_cast<int>(i - static_cast<long long>(r) * d4);
    const int t = r % T;
    const float4 w = inputs_embeds ? reinterpret_cast<const float4*>(inputs_embeds)[i]
                                   : reinterpret_cast<const float4*>(word)[checked_index(ids[r], vocab, err, CLIMB_ERR_TOKEN_ID) * d4 + c];
    const long long ty = tt ? checked_index(tt[r], n_types, err, CLIMB_ERR_TOKEN_TYPE) : 0;
    const float4 s = reinterpret_cast<const float4*>(type_emb)[ty * d4 + c];
    const float4 p = reinterpret_cast<const float4*>(pos)[static_cast<long long>(t) * d4 + c];
    reinterpret_cast<float4*>(e)[i] = make_float4(w.x + s.x + p.x, w.y + s.y + p.y, w.z + s.z + p.z, w.w + s.w + p.w);
}

// pixel fp32 [B, C, H, W] -> bf16 [B*hp*wp, C*P*P], k = (c, ky, kx) as in the conv weight [d, C, P, P]
__global__ void im2col_kernel(const float* __restrict__ px, __nv_bfloat16* __restrict__ out, int B, int C,
                              int H, int W, int P, int hp, int wp) {
    const int K = C * P * P;
    const int k8 = K / 8;
    const long long total = static_cast<long long>(B) * hp * wp * k8;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long m = i / k8;
    const int k0 = static_cast<int>(i - m * k8) * 8;
    const int c = k0 / (P * P), rem = k0 - c * P * P, ky = rem / P, kx = rem - ky * P;
    const int b = static_cast<int>(m / (hp * wp)), pr = static_cast<int>(m - static_cast<long long>(b) * hp * wp);
    const int py = pr / wp, pxi = pr - py * wp;
    const float* src = px + ((static_cast<long long>(b) * C + c) * H + (py * P + ky)) * W + pxi * P + kx;
    const float4 a = *reinterpret_cast<const float4*>(src);
    const float4 b4 = *reinterpret_cast<const float4*>(src + 4);
    uint4 o;
    o.x = pack_bf16(a.x, a.y); o.y = pack_bf16(a.z, a.w);
    o.z = pack_bf16(b4.x, b4.y); o.w = pack_bf16(b4.z, b4.w);
    *reinterpret_cast<uint4*>(out + m * K + k0) = o;
}

struct Taps { int i00, i01, i10, i11; float w00, w01, w10, w11; };

// torch's bilinear / align_corners=True source index (UpSample.h area_pixel_compute_source_index)
__device__ __forceinline__ Taps bilinear_taps(int py, int pxi, int hp, int wp, int G) {
    const float sy = hp > 1 ? static_cast<float>(G - 1) / static_cast<float>(hp - 1) : 0.0f;
    const float sx = wp > 1 ? static_cast<float>(G - 1) / static_cast<float>(wp - 1) : 0.0f;
    const float fy = sy * py, fx = sx * pxi;
    const int y0 = min(static_cast<int>(fy), G - 1), x0 = min(static_cast<int>(fx), G - 1);
    const int y1 = min(y0 + 1, G - 1), x1 = min(x0 + 1, G - 1);
    const float ly = fy - y0, lx = fx - x0;
    Taps t;
    t.i00 = y0 * G + x0; t.i01 = y0 * G + x1; t.i10 = y1 * G + x0; t.i11 = y1 * G + x1;
    t.w00 = (1.0f - ly) * (1.0f - lx); t.w01 = (1.0f - ly) * lx; t.w10 = ly * (1.0f - lx); t.w11 = ly * lx;
    return t;
}

// pos_emb [1 + G*G, d] -> table [hp*wp, d]
__global__ void pos_interp_kernel(const float* __restrict__ pos_emb, float* __restrict__ table, int hp, int wp,
                                  int G, int d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hp * wp * d) return;
    const int p = i / d, c = i - p * d;
    const Taps t = bilinear_taps(p / wp, p % wp, hp, wp, G);
    const float* g = pos_emb + d;      // skip row 0 (the [cls] position)
    table[i] = t.w00 * g[t.i00 * d + c] + t.w01 * g[t.i01 * d + c] + t.w10 * g[t.i10 * d + c] + t.w11 * g[t.i11 * d + c];
}

// x[b, l, :] for l < T: text_ln[b*T+l] + mod[0]; l == T: cls + pos[0] + mod[idx_b]; l > T: patch + table + mod[idx_b]
__global__ void embed_assemble_kernel(const float* __restrict__ text_ln, const float* __restrict__ patch,
                                      const float* __restrict__ table, const float* __restrict__ cls,
                                      const float* __restrict__ pos_emb, const float* __restrict__ mod,
                                      const int* __restrict__ type_idx, int type_idx_scalar,
                                      float* __restrict__ x, int B, int T, int Np, int d4, int n_mod, unsigned int* err,
                                      uint32_t thresh, float inv_keep, unsigned long long seed, int rep) {
    const int L = T + 1 + Np;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(B) * L * d4) return;
    const long long row = i / d4;
    const int c = static_cast<int>(i - row * d4);
    const int b = static_cast<int>(row / L), l = static_cast<int>(row - static_cast<long long>(b) * L);
    float4 v, m;
    if (l < T) {
        v = reinterpret_cast<const float4*>(text_ln)[(static_cast<long long>(b) * T + l) * d4 + c];
        m = reinterpret_cast<const float4*>(mod)[c];
    } else {
        const int idx = static_cast<int>(checked_index(type_idx ? type_idx[b] : type_idx_scalar, n_mod, err, CLIMB_ERR_MODALITY));
        m = reinterpret_cast<const float4*>(mod)[static_cast<long long>(idx) * d4 + c];
        if (l == T) {
            const float4 a = reinterpret_cast<const float4*>(cls)[c];
            const float4 p0 = reinterpret_cast<const float4*>(pos_emb)[c];
            v = make_float4(a.x + p0.x, a.y + p0.y, a.z + p0.z, a.w + p0.w);
        } else {
            const int p = l - T - 1;
            const float4 a = reinterpret_cast<const float4*>(patch)[(static_cast<long long>(b / rep) * Np + p) * d4 + c];   // image b / rep
            const float4 t = reinterpret_cast<const float4*>(table)[static_cast<long long>(p) * d4 + c];
            v = make_float4(a.x + t.x, a.y + t.y, a.z + t.z, a.w + t.w);
        }
    }
    if (thresh != 0u) {       // embedding dropout (modeling_vilt.py:201, :303): on the content rows, before the modality-type rows are added
        const uint4 rnd = philox4x32(seed, static_cast<unsigned long long>(i));
        v.x *= dropout_scale(rnd.x, thresh, inv_keep);
        v.y *= dropout_scale(rnd.y, thresh, inv_keep);
        v.z *= dropout_scale(rnd.z, thresh, inv_keep);
        v.w *= dropout_scale(rnd.w, thresh, inv_keep);
    }
    reinterpret_cast<float4*>(x)[i] = make_float4(v.x + m.x, v.y + m.y, v.z + m.z, v.w + m.w);
}

// ---- variable resolution (ViltEmbeddings.visual_embed with padded images, modeling_vilt.py:121-205) ----------
// geom[b] = (h_b, w_b): the valid patch rectangle of image b (top-left of the padded hp x wp grid). Image b's
// valid patches occupy slots s = py * w_b + px < h_b * w_b of its Np slots in raster order; the remaining slots
// are padding: zero pixels, no position embedding, masked out as attention keys. (The reference fills them with
// randomly chosen masked patches and permutes the valid ones: every output CLiMB consumes is invariant to both.)
// Which patch of image b's own h_b x w_b grid sits in sequence slot `slot` (raster index, -1 = padding slot).
// sel == nullptr: slots 0 .. h_b w_b - 1 hold the valid patches in raster order. sel [B, Np] (config.max_image_length > 0,
// modeling_vilt.py:163-189): the host drew which valid patches each image keeps, exactly as the reference does.
__device__ __forceinline__ int slot_patch(const int* __restrict__ sel, int b, int Np, int slot, int hb, int wb) {
    if (sel != nullptr) return sel[static_cast<long long>(b) * Np + slot];
    return slot < hb * wb ? slot : -1;
}

__global__ void im2col_ragged_kernel(const float* __restrict__ px, const int* __restrict__ geom,
                                     __nv_bfloat16* __restrict__ out, int B, int C, int H, int W, int P, int Np, int rep,
                                     const int* __restrict__ sel) {
    const int K = C * P * P;
    const int k8 = K / 8;
    const long long total = static_cast<long long>(B) * Np * k8;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long m = i / k8;
    const int k0 = static_cast<int>(i - m * k8) * 8;
    const int b = static_cast<int>(m / Np), slot = static_cast<int>(m - static_cast<long long>(b) * Np);
    const int hb = geom[2 * b * rep], wb = geom[2 * b * rep + 1];       // geom is per sequence; image b serves sequences b * rep ..
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    const int pj = slot_patch(sel, b * rep, Np, slot, hb, wb);
    if (pj >= 0) {
        const int py = pj / wb, pxi = pj - py * wb;
        const int c = k0 / (P * P), rem = k0 - c * P * P, ky = rem / P, kx = rem - ky * P;
        const float* src = px + ((static_cast<long long>(b) * C + c) * H + (py * P + ky)) * W + pxi * P + kx;
        const float4 a = *reinterpret_cast<const float4*>(src);
        const float4 e = *reinterpret_cast<const float4*>(src + 4);
        o.x = pack_bf16(a.x, a.y); o.y = pack_bf16(a.z, a.w); o.z = pack_bf16(e.x, e.y); o.w = pack_bf16(e.z, e.w);
    }
    reinterpret_cast<uint4*>(out)[i] = o;
}

__global__ void embed_assemble_ragged_kernel(const float* __restrict__ text_ln, const float* __restrict__ patch,
                                             const int* __restrict__ geom, const float* __restrict__ cls,
                                             const float* __restrict__ pos_emb, const float* __restrict__ mod,
                                             const int* __restrict__ type_idx, int type_idx_scalar,
                                             float* __restrict__ x, int B, int T, int Np, int G, int d4, int n_mod, unsigned int* err,
                                             uint32_t thresh, float inv_keep, unsigned long long seed, int rep,
                                             const int* __restrict__ sel) {
    const int L = T + 1 + Np;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(B) * L * d4) return;
    const long long row = i / d4;
    const int c = static_cast<int>(i - row * d4);
    const int b = static_cast<int>(row / L), l = static_cast<int>(row - static_cast<long long>(b) * L);
    float4 v, m;
    if (l < T) {
        v = reinterpret_cast<const float4*>(text_ln)[(static_cast<long long>(b) * T + l) * d4 + c];
        m = reinterpret_cast<const float4*>(mod)[c];
    } else {
        const int idx = static_cast<int>(checked_index(type_idx ? type_idx[b] : type_idx_scalar, n_mod, err, CLIMB_ERR_MODALITY));
        m = reinterpret_cast<const float4*>(mod)[static_cast<long long>(idx) * d4 + c];
        if (l == T) {
            const float4 a = reinterpret_cast<const float4*>(cls)[c];
            const float4 p0 = reinterpret_cast<const float4*>(pos_emb)[c];
            v = make_float4(a.x + p0.x, a.y + p0.y, a.z + p0.z, a.w + p0.w);
        } else {
            const int slot = l - T - 1;
            v = reinterpret_cast<const float4*>(patch)[(static_cast<long long>(b / rep) * Np + slot) * d4 + c];
            const int hb = geom[2 * b], wb = geom[2 * b + 1];
            const int pj = slot_patch(sel, b, Np, slot, hb, wb);
            if (pj >= 0) {
                const Taps t = bilinear_taps(pj / wb, pj % wb, hb, wb, G);
                const float4* g = reinterpret_cast<const float4*>(pos_emb) + d4;      // skip row 0 (the [cls] position)
                const float4 q00 = g[static_cast<long long>(t.i00) * d4 + c], q01 = g[static_cast<long long>(t.i01) * d4 + c];
                const float4 q10 = g[static_cast<long long>(t.i10) * d4 + c], q11 = g[static_cast<long long>(t.i11) * d4 + c];
                v.x += t.w00 * q00.x + t.w01 * q01.x + t.w10 * q10.x + t.w11 * q11.x;
                v.y += t.w00 * q00.y + t.w01 * q01.y + t.w10 * q10.y + t.w11 * q11.y;
                v.z += t.w00 * q00.z + t.w01 * q01.z + t.w10 * q10.z + t.w11 * q11.z;
                v.w += t.w00 * q00.w + t.w01 * q01.w + t.w10 * q10.w + t.w11 * q11.w;
            }
        }
    }
    if (thresh != 0u) {       // embedding dropout (modeling_vilt.py:201, :303): on the content rows, before the modality-type rows are added
        const uint4 rnd = philox4x32(seed, static_cast<unsigned long long>(i));
        v.x *= dropout_scale(rnd.x, thresh, inv_keep);
        v.y *= dropout_scale(rnd.y, thresh, inv_keep);
        v.z *= dropout_scale(rnd.z, thresh, inv_keep);
        v.w *= dropout_scale(rnd.w, thresh, inv_keep);
    }
    reinterpret_cast<float4*>(x)[i] = make_float4(v.x + m.x, v.y + m.y, v.z + m.z, v.w + m.w);
}

// d_pos[1 + g, :] += sum over images / valid slots of the transposed bilinear taps (per-image grids: no batch pre-sum)
__global__ void pos_scatter_ragged_bwd_kernel(const float* __restrict__ dx, const int* __restrict__ geom,
                                              float* __restrict__ d_pos, int B, int T, int Np, int G, int d,
                                              const int* __restrict__ sel) {
    const int L = T + 1 + Np;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(B) * Np * d) return;
    const long long m = i / d;
    const int c = static_cast<int>(i - m * d);
    const int b = static_cast<int>(m / Np), slot = static_cast<int>(m - static_cast<long long>(b) * Np);
    const int hb = geom[2 * b], wb = geom[2 * b + 1];
    const int pj = slot_patch(sel, b, Np, slot, hb, wb);
    if (pj < 0) return;
    const float g = dx[(static_cast<long long>(b) * L + T + 1 + slot) * d + c];
    const Taps t = bilinear_taps(pj / wb, pj % wb, hb, wb, G);
    float* dp = d_pos + d;
    if (t.w00 != 0.0f) atomicAdd(dp + static_cast<long long>(t.i00) * d + c, t.w00 * g);
    if (t.w01 != 0.0f) atomicAdd(dp + static_cast<long long>(t.i01) * d + c, t.w01 * g);
    if (t.w10 != 0.0f) atomicAdd(dp + static_cast<long long>(t.i10) * d + c, t.w10 * g);
    if (t.w11 != 0.0f) atomicAdd(dp + static_cast<long long>(t.i11) * d + c, t.w11 * g);
}

// key_bias for padded images: text mask | 0 for [cls] | 0 for valid slots, -10000 for padding slots
__global__ void key_bias_ragged_kernel(const long long* __restrict__ mask, const int* __restrict__ geom,
                                       float* __restrict__ out, int B, int T, int L, const int* __restrict__ sel) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * L) return;
    const int b = i / L, j = i - b * L;
    float v = 0.0f;
    if (j < T) v = mask ? (1.0f - static_cast<float>(mask[static_cast<long long>(b) * T + j])) * -10000.0f : 0.0f;
    else if (j > T) v = slot_patch(sel, b, L - T - 1, j - T - 1, geom[2 * b], geom[2 * b + 1]) >= 0 ? 0.0f : -10000.0f;
    out[i] = v;
}

// dx [B, L, d] -> dy_text fp32 [B*T, d] (input of the text LayerNorm backward) and
//                 dpatch bf16 [B*Np, d] (A operand of the patch-projection wgrad)
__global__ void embed_split_bwd_kernel(const float* __restrict__ dx, float* __restrict__ dy_text,
                                       __nv_bfloat16* __restrict__ dpatch, int B, int T, int Np, int d4) {
    const int L = T + 1 + Np;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(B) * L * d4) return;
    const long long row = i / d4;
    const int c = static_cast<int>(i - row * d4);
    const int b = static_cast<int>(row / L), l = static_cast<int>(row - static_cast<long long>(b) * L);
    if (l == T) return;
    const float4 v = reinterpret_cast<const float4*>(dx)[i];
    if (l < T) {
        if (dy_text) reinterpret_cast<float4*>(dy_text)[(static_cast<long long>(b) * T + l) * d4 + c] = v;
    } else if (dpatch) {
        uint2 o;
        o.x = pack_bf16(v.x, v.y);
        o.y = pack_bf16(v.z, v.w);
        reinterpret_cast<uint2*>(dpatch)[(static_cast<long long>(b) * Np + (l - T - 1)) * d4 + c] = o;
    }
}

// dpatch[i, p, :] = sum over the `rep` sequences that share image i of dx[i * rep + j, T + 1 + p, :]  (VCR: one image, four
// answer choices -- vilt.py:334-347 -- the patch projection runs once per image and its gradient collects all four)
__global__ void embed_patch_sum_bwd_kernel(const float* __restrict__ dx, __nv_bfloat16* __restrict__ dpatch, int Bi, int rep,
                                           int T, int Np, int d4) {
    const int L = T + 1 + Np;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(Bi) * Np * d4) return;
    const long long row = i / d4;
    const int c = static_cast<int>(i - row * d4);
    const int img = static_cast<int>(row / Np), p = static_cast<int>(row - static_cast<long long>(img) * Np);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < rep; ++j) {
        const float4 v = reinterpret_cast<const float4*>(dx)[((static_cast<long long>(img) * rep + j) * L + T + 1 + p) * d4 + c];
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    uint2 o;
    o.x = pack_bf16(a.x, a.y);
    o.y = pack_bf16(a.z, a.w);
    reinterpret_cast<uint2*>(dpatch)[i] = o;
}

// S[k][l, c] = sum over the sequences b whose image type index is k+1 of dx[b, l, c]  (k = 0, 1).
// Text rows (l < T) all go to S[0]. One thread per (l, float4 column), sequential over b: fixed
// summation order, coalesced across the warp.
__global__ void embed_reduce_bwd_kernel(const float* __restrict__ dx, const int* __restrict__ type_idx,
                                        int type_idx_scalar, float* __restrict__ S, int B, int T, int L, int d4) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L * d4) return;
    const int l = i / d4;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    for (int b = 0; b < B; ++b) {
        const float4 v = reinterpret_cast<const float4*>(dx)[static_cast<long long>(b) * L * d4 + i];
        const int idx = (l < T) ? 1 : (type_idx ? type_idx[b] : type_idx_scalar);
        if (idx == 2) { a1.x += v.x; a1.y += v.y; a1.z += v.z; a1.w += v.w; }
        else          { a0.x += v.x; a0.y += v.y; a0.z += v.z; a0.w += v.w; }
    }
    reinterpret_cast<float4*>(S)[i] = a0;
    reinterpret_cast<float4*>(S)[static_cast<long long>(L) * d4 + i] = a1;
}

// Folds S onto cls_token, position_embeddings (row 0 + the transposed bilinear taps), the
// modality-type table and the patch-projection bias. grid.x covers the columns, grid.y slices the
// L rows; every CTA reduces its slice and finishes with one atomic per column and target.
__global__ void embed_finalize_bwd_kernel(const float* __restrict__ S, float* __restrict__ d_cls,
                                          float* __restrict__ d_pos, float* __restrict__ d_mod,
                                          float* __restrict__ d_patch_bias, int n_mod, int T, int hp, int wp,
                                          int G, int d, int ragged_np) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d) return;
    const int Np = ragged_np > 0 ? ragged_np : hp * wp, L = T + 1 + Np;
    const int per = (L + gridDim.y - 1) / gridDim.y;
    const int l0 = blockIdx.y * per, l1 = min(L, l0 + per);
    const float* S0 = S;
    const float* S1 = S + static_cast<long long>(L) * d;
    float m0 = 0.0f, m1 = 0.0f, m2 = 0.0f, pb = 0.0f;
    for (int l = l0; l < l1; ++l) {
        const float a = S0[l * d + c], b = S1[l * d + c];
        if (l < T) { m0 += a; continue; }
        m1 += a;
        m2 += b;
        const float g = a + b;
        if (l == T) {
            if (d_cls) atomicAdd(d_cls + c, g);
            if (d_pos) atomicAdd(d_pos + c, g);
            continue;
        }
        pb += g;
        if (d_pos && ragged_np == 0) {       // per-image grids: pos_scatter_ragged_bwd_kernel does the taps
            const int p = l - T - 1;
            const Taps t = bilinear_taps(p / wp, p % wp, hp, wp, G);
            float* dp = d_pos + d;
            atomicAdd(dp + t.i00 * d + c, t.w00 * g);
            atomicAdd(dp + t.i01 * d + c, t.w01 * g);
            atomicAdd(dp + t.i10 * d + c, t.w10 * g);
            atomicAdd(dp + t.i11 * d + c, t.w11 * g);
        }
    }
    if (d_mod) {
        if (m0 != 0.0f) atomicAdd(d_mod + c, m0);
        if (l1 > T) {
            atomicAdd(d_mod + d + c, m1);
            if (n_mod > 2) atomicAdd(d_mod + 2 * d + c, m2);
        }
    }
    if (d_patch_bias && l1 > T + 1) atomicAdd(d_patch_bias + c, pb);
}

// de [B*T, d] -> word / segment / position embedding gradients (dense tables, atomics on collisions)
__global__ void text_scatter_bwd_kernel(const float* __restrict__ de, const long long* __restrict__ ids,
                                        const long long* __restrict__ tt, float* __restrict__ d_word,
                                        float* __restrict__ d_type, float* __restrict__ d_pos, int rows, int T, int d,
                                        int vocab, int n_types) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(rows) * d) return;
    const int r = static_cast<int>(i / d), c = static_cast<int>(i - static_cast<long long>(r) * d);
    const float g = de[i];
    // (the forward already reported bad ids; here they only must not write outside the tables)
    if (d_word && ids) atomicAdd(d_word + checked_index(ids[r], vocab, nullptr, 0u) * d + c, g);
    if (d_type) atomicAdd(d_type + (tt ? checked_index(tt[r], n_types, nullptr, 0u) : 0) * d + c, g);
    if (d_pos) atomicAdd(d_pos + static_cast<long long>(r % T) * d + c, g);
}

inline unsigned blocks_for(long long n, int threads) { return static_cast<unsigned>((n + threads - 1) / threads); }

}  // namespace

int text_gather(const long long* ids, const float* inputs_embeds, const long long* tt, const float* word,
                const float* type_emb, const float* pos, float* e, int rows, int T, int d, cudaStream_t stream, int vocab,
                int n_types) {
    CLIMB_REQUIRE((ids != nullptr) != (inputs_embeds != nullptr), "text_gather: exactly one of input_ids / inputs_embeds");
    CLIMB_REQUIRE(type_emb && pos && e && rows > 0 && d % 4 == 0, "text_gather: bad arguments");
    CLIMB_REQUIRE(ids == nullptr || word != nullptr, "text_gather: word table missing");
    text_gather_kernel<<<blocks_for(static_cast<long long>(rows) * (d / 4), 256), 256, 0, stream>>>(
        ids, inputs_embeds, tt, word, type_emb, pos, e, rows, T, d / 4, vocab, n_types, device_error_word());
    CLIMB_LAUNCH_OK();
    return 0;
}

int im2col(const float* px, void* out, int B, int C, int H, int W, int P, cudaStream_t stream) {
    CLIMB_REQUIRE(px && out && B > 0, "im2col: bad arguments");
    CLIMB_REQUIRE(P % 8 == 0 && H % P == 0 && W % P == 0,
                  "im2col: fixed-resolution path needs H, W multiples of the patch size (%d x %d, P=%d)", H, W, P);
    CLIMB_REQUIRE((reinterpret_cast<uintptr_t>(px) & 15) == 0 && W % 4 == 0, "im2col: pixel rows must be 16-byte aligned");
    const int hp = H / P, wp = W / P;
    const long long total = static_cast<long long>(B) * hp * wp * (C * P * P / 8);
    im2col_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(px, static_cast<__nv_bfloat16*>(out), B, C, H, W, P, hp, wp);
    CLIMB_LAUNCH_OK();
    return 0;
}

int pos_interp(const float* pos_emb, float* table, int hp, int wp, int G, int d, cudaStream_t stream) {
    CLIMB_REQUIRE(pos_emb && table && hp > 0 && wp > 0 && G > 0, "pos_interp: bad arguments");
    pos_interp_kernel<<<blocks_for(static_cast<long long>(hp) * wp * d, 256), 256, 0, stream>>>(pos_emb, table, hp, wp, G, d);
    CLIMB_LAUNCH_OK();
    return 0;
}

int embed_assemble(const float* text_ln, const float* patch, const float* table, const float* cls,
                   const float* pos_emb, const float* mod, const int* type_idx, int type_idx_scalar, float* x,
                   int B, int T, int Np, int d, cudaStream_t stream, int n_mod, float p_drop, unsigned long long seed, int rep) {
    CLIMB_REQUIRE(text_ln && patch && table && cls && pos_emb && mod && x && d % 4 == 0 && rep >= 1, "embed_assemble: bad arguments");
    const long long total = static_cast<long long>(B) * (T + 1 + Np) * (d / 4);
    embed_assemble_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(text_ln, patch, table, cls, pos_emb, mod,
                                                                     type_idx, type_idx_scalar, x, B, T, Np, d / 4, n_mod,
                                                                     device_error_word(), p_drop > 0.0f ? dropout_threshold(p_drop) : 0u,
                                                                     1.0f / (1.0f - p_drop), seed, rep);
    CLIMB_LAUNCH_OK();
    return 0;
}

int embed_split_bwd(const float* dx, float* dy_text, void* dpatch, int B, int T, int Np, int d, cudaStream_t stream, int rep) {
    CLIMB_REQUIRE(dx && d % 4 == 0 && rep >= 1 && B % rep == 0, "embed_split_bwd: bad arguments");
    const long long total = static_cast<long long>(B) * (T + 1 + Np) * (d / 4);
    embed_split_bwd_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(dx, dy_text, rep > 1 ? nullptr : static_cast<__nv_bfloat16*>(dpatch), B, T, Np, d / 4);
    CLIMB_LAUNCH_OK();
    if (rep > 1 && dpatch != nullptr) {
        const long long tp = static_cast<long long>(B / rep) * Np * (d / 4);
        embed_patch_sum_bwd_kernel<<<blocks_for(tp, 256), 256, 0, stream>>>(dx, static_cast<__nv_bfloat16*>(dpatch), B / rep, rep, T, Np, d / 4);
        CLIMB_LAUNCH_OK();
    }
    return 0;
}

int im2col_ragged(const float* px, const int* geom, void* out, int B, int C, int H, int W, int P, int Np, cudaStream_t stream, int rep,
                  const int* sel) {
    CLIMB_REQUIRE(px && geom && out && B > 0 && Np > 0, "im2col_ragged: bad arguments");
    CLIMB_REQUIRE(P % 8 == 0 && H % P == 0 && W % P == 0, "im2col_ragged: H, W must be multiples of the patch size (%d x %d, P=%d)", H, W, P);
    CLIMB_REQUIRE((reinterpret_cast<uintptr_t>(px) & 15) == 0 && W % 4 == 0, "im2col_ragged: pixel rows must be 16-byte aligned");
    const long long total = static_cast<long long>(B) * Np * (C * P * P / 8);
    im2col_ragged_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(px, geom, static_cast<__nv_bfloat16*>(out), B, C, H, W, P, Np, rep, sel);
    CLIMB_LAUNCH_OK();
    return 0;
}

int embed_assemble_ragged(const float* text_ln, const float* patch, const int* geom, const float* cls, const float* pos_emb,
                          const float* mod, const int* type_idx, int type_idx_scalar, float* x, int B, int T, int Np, int G,
                          int d, cudaStream_t stream, int n_mod, float p_drop, unsigned long long seed, int rep, const int* sel) {
    CLIMB_REQUIRE(text_ln && patch && geom && cls && pos_emb && mod && x && d % 4 == 0 && rep >= 1, "embed_assemble_ragged: bad arguments");
    const long long total = static_cast<long long>(B) * (T + 1 + Np) * (d / 4);
    embed_assemble_ragged_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(text_ln, patch, geom, cls, pos_emb, mod, type_idx,
                                                                            type_idx_scalar, x, B, T, Np, G, d / 4, n_mod, device_error_word(),
                                                                            p_drop > 0.0f ? dropout_threshold(p_drop) : 0u,
                                                                            1.0f / (1.0f - p_drop), seed, rep, sel);
    CLIMB_LAUNCH_OK();
    return 0;
}

int key_bias_ragged(const long long* mask, const int* geom, float* out, int B, int T, int L, cudaStream_t stream, const int* sel) {
    CLIMB_REQUIRE(geom && out && B > 0 && T > 0 && L > T, "key_bias_ragged: bad arguments");
    key_bias_ragged_kernel<<<(B * L + 255) / 256, 256, 0, stream>>>(mask, geom, out, B, T, L, sel);
    CLIMB_LAUNCH_OK();
    return 0;
}

int embed_reduce_bwd(const float* dx, const int* type_idx, int type_idx_scalar, float* S, float* d_cls,
                     float* d_pos, float* d_mod, float* d_patch_bias, int n_mod, int B, int T, int hp, int wp,
                     int G, int d, cudaStream_t stream, const int* geom, int ragged_np, const int* sel) {
    CLIMB_REQUIRE(dx && S && d % 4 == 0, "embed_reduce_bwd: bad arguments");
    if (geom == nullptr) ragged_np = 0;
    const int L = T + 1 + (ragged_np > 0 ? ragged_np : hp * wp);
    embed_reduce_bwd_kernel<<<blocks_for(static_cast<long long>(L) * (d / 4), 128), 128, 0, stream>>>(
        dx, type_idx, type_idx_scalar, S, B, T, L, d / 4);
    CLIMB_LAUNCH_OK();
    embed_finalize_bwd_kernel<<<dim3(blocks_for(d, 128), 32), 128, 0, stream>>>(S, d_cls, d_pos, d_mod, d_patch_bias, n_mod, T, hp, wp, G, d,
                                                                               ragged_np);
    CLIMB_LAUNCH_OK();
    if (ragged_np > 0 && d_pos != nullptr) {
        pos_scatter_ragged_bwd_kernel<<<blocks_for(static_cast<long long>(B) * ragged_np * d, 256), 256, 0, stream>>>(
            dx, geom, d_pos, B, T, ragged_np, G, d, sel);
        CLIMB_LAUNCH_OK();
    }
    return 0;
}

int text_scatter_bwd(const float* de, const long long* ids, const long long* tt, float* d_word, float* d_type,
                     float* d_pos, int rows, int T, int d, cudaStream_t stream, int vocab, int n_types) {
    CLIMB_REQUIRE(de && rows > 0, "text_scatter_bwd: bad arguments");
    text_scatter_bwd_kernel<<<blocks_for(static_cast<long long>(rows) * d, 256), 256, 0, stream>>>(
        de, ids, tt, d_word, d_type, d_pos, rows, T, d, vocab, n_types);
    CLIMB_LAUNCH_OK();
    return 0;
}

}  // namespace climb
