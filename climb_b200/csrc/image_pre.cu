// GPU input pipeline for the image side of ViltEncoderWrapper.process_inputs (src/modeling/vilt.py:83-96 ->
// ViltFeatureExtractor.__call__, feature_extraction_vilt.py:253-292): Pillow's 8-bit bicubic resize, the float32
// normalisation, zero padding to the batch maximum and the pixel mask, for a whole batch in two launches.
//
// Bit-exact with the reference's CPU path. Pillow's resampler (src/libImaging/Resample.c) is integer
// arithmetic: per output pixel a window of up to ksize input pixels times int32 coefficients in 2^-22 units, + 2^21, >> 22,
// clipped to [0, 255]; the horizontal pass runs first and its uint8 result feeds the vertical pass. The coefficient tables
// depend only on (input size, output size) and are computed on the host exactly as precompute_coeffs / normalize_coeffs_8bpc
// do (climb_b200/image_processing.py); the kernels apply them. An axis whose size does not change gets the identity table,
// which reproduces Pillow skipping that pass.
//
//   image_resample_h_kernel   src u8 [in_h, in_w, 3] -> tmp u8 [in_h, out_w, 3]          (thread = output pixel, 3 channels)
//   image_resample_v_kernel   tmp -> pixel_values f32 [B, 3, Hp, Wp] (x / 255 - mean) / std, zeros outside the image,
//                             pixel_mask i64 [B, Hp, Wp]                                  (thread = pixel of the padded canvas)
// Both are byte-sized integer work bound by L2 / HBM traffic of a few MB per batch; nothing here goes near the tensor cores.
#include "common.cuh"
#include "internal.h"

namespace climb {
namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;        // Resample.c PRECISION_BITS

__device__ __forceinline__ int clip8(int v) {
    v >>= kPrecisionBits;
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

__global__ void __launch_bounds__(256)
image_resample_h_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ tmp, const climb_image_desc* __restrict__ descs,
                        const int* __restrict__ tables) {
    const climb_image_desc d = descs[blockIdx.y];
    const long long n = static_cast<long long>(d.in_h) * d.out_w;
    const int* bounds = tables + d.bounds_h_off;
    const int* coef = tables + d.coef_h_off;
    const uint8_t* s = src + d.src_off;
    uint8_t* t = tmp + d.tmp_off;
    for (long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; p < n; p += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int y = static_cast<int>(p / d.out_w), xx = static_cast<int>(p - static_cast<long long>(y) * d.out_w);
        const int xmin = bounds[2 * xx], cnt = bounds[2 * xx + 1];
        const int* k = coef + static_cast<long long>(xx) * d.ksize_h;
        const uint8_t* row = s + (static_cast<long long>(y) * d.in_w + xmin) * 3;
        int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
        for (int x = 0; x < cnt; ++x) {
            const int w = k[x];
            a0 += row[3 * x] * w;
            a1 += row[3 * x + 1] * w;
            a2 += row[3 * x + 2] * w;
        }
        uint8_t* o = t + p * 3;
        o[0] = static_cast<uint8_t>(clip8(a0));
        o[1] = static_cast<uint8_t>(clip8(a1));
        o[2] = static_cast<uint8_t>(clip8(a2));
    }
}

__global__ void __launch_bounds__(256)
image_resample_v_kernel(const uint8_t* __restrict__ tmp, const climb_image_desc* __restrict__ descs, const int* __restrict__ tables,
                        float* __restrict__ pixel_values, long long* __restrict__ pixel_mask, int Hp, int Wp, float mean0, float mean1,
                        float mean2, float std0, float std1, float std2) {
    const int b = blockIdx.y;
    const climb_image_desc d = descs[b];
    const long long n = static_cast<long long>(Hp) * Wp;
    const int* bounds = tables + d.bounds_v_off;
    const int* coef = tables + d.coef_v_off;
    const uint8_t* t = tmp + d.tmp_off;
    float* pv = pixel_values + static_cast<long long>(b) * 3 * n;
    long long* pm = pixel_mask + static_cast<long long>(b) * n;
    for (long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; p < n; p += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int yy = static_cast<int>(p / Wp), x = static_cast<int>(p - static_cast<long long>(yy) * Wp);
        if (yy >= d.out_h || x >= d.out_w) {            // padding up to the batch maximum (feature_extraction_vilt.py:270-281)
            pv[p] = 0.0f;
            pv[n + p] = 0.0f;
            pv[2 * n + p] = 0.0f;
            pm[p] = 0;
            continue;
        }
        const int ymin = bounds[2 * yy], cnt = bounds[2 * yy + 1];
        const int* k = coef + static_cast<long long>(yy) * d.ksize_v;
        const uint8_t* col = t + (static_cast<long long>(ymin) * d.out_w + x) * 3;
        const long long pitch = static_cast<long long>(d.out_w) * 3;
        int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
        for (int y = 0; y < cnt; ++y) {
            const int w = k[y];
            a0 += col[y * pitch] * w;
            a1 += col[y * pitch + 1] * w;
            a2 += col[y * pitch + 2] * w;
        }
        // to_numpy_array: x.astype(float32) / 255.0; normalize: (x - mean) / std, all in float32 (IEEE division: no fast math)
        pv[p] = __fdiv_rn(__fdiv_rn(static_cast<float>(clip8(a0)), 255.0f) - mean0, std0);
        pv[n + p] = __fdiv_rn(__fdiv_rn(static_cast<float>(clip8(a1)), 255.0f) - mean1, std1);
        pv[2 * n + p] = __fdiv_rn(__fdiv_rn(static_cast<float>(clip8(a2)), 255.0f) - mean2, std2);
        pm[p] = 1;
    }
}

}  // namespace

int image_preprocess(const uint8_t* src, uint8_t* tmp, const climb_image_desc* descs_dev, const int* tables_dev, int B,
                     long long max_tmp_pixels, float* pixel_values, long long* pixel_mask, int Hp, int Wp, const float* mean,
                     const float* stdv, cudaStream_t stream) {
    CLIMB_REQUIRE(src && tmp && descs_dev && tables_dev && pixel_values && pixel_mask && mean && stdv, "image_preprocess: null pointer");
    CLIMB_REQUIRE(B > 0 && Hp > 0 && Wp > 0 && max_tmp_pixels > 0, "image_preprocess: empty batch (B=%d, canvas %d x %d)", B, Hp, Wp);
    CLIMB_REQUIRE(stdv[0] != 0.0f && stdv[1] != 0.0f && stdv[2] != 0.0f, "image_preprocess: zero std");
    const int threads = 256;
    const long long per_img_h = (max_tmp_pixels + threads - 1) / threads;
    dim3 grid_h(static_cast<unsigned>(per_img_h < 4096 ? per_img_h : 4096), B);
    image_resample_h_kernel<<<grid_h, threads, 0, stream>>>(src, tmp, descs_dev, tables_dev);
    CLIMB_LAUNCH_OK();
    const long long per_img_v = (static_cast<long long>(Hp) * Wp + threads - 1) / threads;
    dim3 grid_v(static_cast<unsigned>(per_img_v < 4096 ? per_img_v : 4096), B);
    image_resample_v_kernel<<<grid_v, threads, 0, stream>>>(tmp, descs_dev, tables_dev, pixel_values, pixel_mask, Hp, Wp, mean[0], mean[1],
                                                            mean[2], stdv[0], stdv[1], stdv[2]);
    CLIMB_LAUNCH_OK();
    return 0;
}

}  // namespace climb
