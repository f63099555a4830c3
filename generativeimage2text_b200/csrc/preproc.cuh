// Test-time image transform on the GPU: decoded RGB bytes -> CLIP-normalised fp32 planes, bit-identical to the
// reference's PIL / torchvision pipeline (reference inference.py:111-132 `get_image_transform`:
// Resize(BICUBIC) -> CenterCrop -> ToTensor -> Normalize, and :29-64 `MinMaxResizeForTest`).
//
// What Pillow's `Image.resize(size, BICUBIC)` computes for 8-bit images (libImaging/Resample.c), and therefore what
// these kernels compute: a separable resampling, horizontal pass first, where output pixel o of an axis takes the
// window [lo, lo+cnt) of the input axis with weights of the Keys cubic (a = -0.5) stretched by max(in/out, 1)
// ("antialias"), normalised to sum 1 in double, rounded to 22-bit fixed point; the pass accumulates
// 2^21 + sum(pixel * weight) in int32, shifts right by 22 and clamps to [0, 255]. The intermediate image between the two
// passes is uint8 -- the rounding in the middle is part of the result, so the passes cannot be merged algebraically.
//
// Device work is integer / byte, HBM bound: pass 1 reads each needed source row once and writes `out_w` bytes x 3 per
// row; pass 2 reads the intermediate (L2 resident) and writes 12 bytes per output pixel. Only the rows / columns the
// crop window needs are produced. The weight tables are built on the host (double precision, no FMA contraction --
// exactly the arithmetic of Resample.c `precompute_coeffs` + `normalize_coeffs_8bpc`), cached per (in, out) size pair
// and shipped with the per-image descriptors in one H2D copy per call.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <map>
#include <utility>
#include <vector>

namespace gitb200 {

constexpr int kPrecisionBits = 32 - 8 - 2;   // Resample.c PRECISION_BITS

// ---- host: weight tables ---------------------------------------------------------------------------------
struct AxisCoeffs {
  int ksize = 0;
  std::vector<int32_t> bounds;   // [out][2]: first tap, tap count
  std::vector<int32_t> kk;       // [out][ksize] fixed point
};

// The volatile stores keep every intermediate rounded to double exactly where the C source of Resample.c rounds it
// (a generic x86-64 build of Pillow has no FMA contraction; this translation unit is compiled by nvcc's host compiler
// whose flags we do not want to depend on).
static inline double pre_bicubic(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) {
    volatile double t = (a + 2.0) * x;
    volatile double u = t - (a + 3.0);
    volatile double v = u * x;
    volatile double w = v * x;
    return w + 1;
  }
  if (x < 2.0) {
    volatile double t = (x - 5) * x;
    volatile double u = (t + 8) * x;
    volatile double v = u - 4;
    return v * a;
  }
  return 0.0;
}

static inline void build_axis_coeffs(int in_size, int out_size, AxisCoeffs* c) {
  if (in_size == out_size) {   // ImagingResample skips the pass: identity window (pixel * 2^22 + 2^21) >> 22 == pixel
    c->ksize = 1;
    c->bounds.resize(static_cast<size_t>(out_size) * 2);
    c->kk.assign(out_size, 1 << kPrecisionBits);
    for (int i = 0; i < out_size; ++i) { c->bounds[2 * i] = i; c->bounds[2 * i + 1] = 1; }
    return;
  }
  const float in0 = 0.0f, in1 = static_cast<float>(in_size);   // the box is passed as C floats
  const double scale = static_cast<double>(in1 - in0) / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 2.0 * filterscale;
  const int ksize = static_cast<int>(std::ceil(support)) * 2 + 1;
  c->ksize = ksize;
  c->bounds.assign(static_cast<size_t>(out_size) * 2, 0);
  c->kk.assign(static_cast<size_t>(out_size) * ksize, 0);
  const double ss = 1.0 / filterscale;
  std::vector<double> w(ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    volatile double prod = (xx + 0.5) * scale;
    const double center = in0 + prod;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    volatile double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      volatile double arg0 = x + xmin - center + 0.5;
      volatile double arg = arg0 * ss;
      w[x] = pre_bicubic(arg);
      ww = ww + w[x];
    }
    for (int x = 0; x < xmax; ++x) {
      volatile double v = (ww != 0.0) ? w[x] / ww : w[x];
      volatile double s = v * (1 << kPrecisionBits);
      c->kk[static_cast<size_t>(xx) * ksize + x] = v < 0 ? static_cast<int>(-0.5 + s) : static_cast<int>(0.5 + s);
    }
    c->bounds[2 * xx] = xmin;
    c->bounds[2 * xx + 1] = xmax;
  }
}

// ---- device ------------------------------------------------------------------------------------------------
struct PreImage {            // one per image, built on the host
  long long src_off;         // bytes into the packed source buffer (RGB, HWC)
  long long tmp_off;         // bytes into the intermediate buffer: [rows_needed][out_w][3]
  long long dst_off;         // elements into the fp32 output: [3][out_h][out_w]
  int src_h, src_w;
  int out_h, out_w;
  int row0, rows;            // source rows the vertical windows of the crop touch: [row0, row0 + rows)
  int kh, kv;                // taps per output column / row
  int hb_off, hk_off;        // int32 offsets into the table buffer: horizontal bounds [out_w][2], weights [out_w][kh]
  int vb_off, vk_off;        // vertical bounds [out_h][2] (first tap relative to row0), weights [out_h][kv]
};

// Pass 1: tmp[r][x][c] = clip8((2^21 + sum_k src[row0 + r][lo_x + k][c] * w_x[k]) >> 22)
__global__ void __launch_bounds__(256)
pre_horizontal_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ tmp, const PreImage* __restrict__ imgs,
                      const int32_t* __restrict__ tab) {
  const PreImage im = imgs[blockIdx.y];
  const long long total = static_cast<long long>(im.rows) * im.out_w;
  const uint8_t* s0 = src + im.src_off;
  uint8_t* t0 = tmp + im.tmp_off;
  const int32_t* hb = tab + im.hb_off;
  const int32_t* hk = tab + im.hk_off;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / im.out_w);
    const int x = static_cast<int>(i - static_cast<long long>(r) * im.out_w);
    const int lo = __ldg(hb + 2 * x), cnt = __ldg(hb + 2 * x + 1);
    const int32_t* k = hk + static_cast<long long>(x) * im.kh;
    const uint8_t* p = s0 + (static_cast<long long>(im.row0 + r) * im.src_w + lo) * 3;
    int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
    for (int j = 0; j < cnt; ++j) {
      const int w = __ldg(k + j);
      a0 += static_cast<int>(p[3 * j]) * w;
      a1 += static_cast<int>(p[3 * j + 1]) * w;
      a2 += static_cast<int>(p[3 * j + 2]) * w;
    }
    uint8_t* o = t0 + i * 3;
    o[0] = static_cast<uint8_t>(min(max(a0 >> kPrecisionBits, 0), 255));
    o[1] = static_cast<uint8_t>(min(max(a1 >> kPrecisionBits, 0), 255));
    o[2] = static_cast<uint8_t>(min(max(a2 >> kPrecisionBits, 0), 255));
  }
}

// Pass 2 + ToTensor + Normalize: out[c][y][x] = ((clip8(...) / 255) - mean[c]) / std[c], each step rounded to fp32
// (IEEE division, like the CPU tensors of the reference pipeline).
__global__ void __launch_bounds__(256)
pre_vertical_norm_kernel(const uint8_t* __restrict__ tmp, float* __restrict__ out, const PreImage* __restrict__ imgs,
                         const int32_t* __restrict__ tab, float m0, float m1, float m2, float s0, float s1, float s2) {
  const PreImage im = imgs[blockIdx.y];
  const long long plane = static_cast<long long>(im.out_h) * im.out_w;
  const uint8_t* t0 = tmp + im.tmp_off;
  float* o = out + im.dst_off;
  const int32_t* vb = tab + im.vb_off;
  const int32_t* vk = tab + im.vk_off;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < plane;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int y = static_cast<int>(i / im.out_w);
    const int x = static_cast<int>(i - static_cast<long long>(y) * im.out_w);
    const int lo = __ldg(vb + 2 * y), cnt = __ldg(vb + 2 * y + 1);
    const int32_t* k = vk + static_cast<long long>(y) * im.kv;
    const uint8_t* p = t0 + (static_cast<long long>(lo) * im.out_w + x) * 3;
    const long long pitch = static_cast<long long>(im.out_w) * 3;
    int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
    for (int j = 0; j < cnt; ++j) {
      const int w = __ldg(k + j);
      a0 += static_cast<int>(p[0]) * w;
      a1 += static_cast<int>(p[1]) * w;
      a2 += static_cast<int>(p[2]) * w;
      p += pitch;
    }
    const float v0 = static_cast<float>(min(max(a0 >> kPrecisionBits, 0), 255));
    const float v1 = static_cast<float>(min(max(a1 >> kPrecisionBits, 0), 255));
    const float v2 = static_cast<float>(min(max(a2 >> kPrecisionBits, 0), 255));
    o[i] = __fdiv_rn(__fsub_rn(__fdiv_rn(v0, 255.0f), m0), s0);
    o[plane + i] = __fdiv_rn(__fsub_rn(__fdiv_rn(v1, 255.0f), m1), s1);
    o[2 * plane + i] = __fdiv_rn(__fsub_rn(__fdiv_rn(v2, 255.0f), m2), s2);
  }
}


}  // namespace gitb200
