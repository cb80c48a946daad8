// On-device image preprocessing: the reference's `model.processor` (timm eval transform on PIL images,
// /root/reference/models/global_eval_models.py:333-334, used by /root/reference/datasets/global_eval_torch_ds.py:20-21)
// for batches of decoded uint8 HWC images: bicubic resize of the shorter side (Pillow's 8-bit algorithm: separable,
// antialiased support, Q22 fixed-point coefficients, uint8 rounding after each pass, horizontal pass first) ->
// center crop -> /255 -> (x - mean) / std -> NCHW.  Bit-exact against Pillow + torchvision (tests/test_preprocess*.py).
// Only the cropped window is computed.
#include <math.h>

#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace semdiff {

constexpr int PRECISION_BITS = 32 - 8 - 2;

static double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

// Pillow: precompute_coeffs + normalize_coeffs_8bpc (same double arithmetic, same truncations)
int resize_ksize(int in_size, int out_size) {
  double filterscale = (double)in_size / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  return (int)ceil(2.0 * filterscale) * 2 + 1;
}
void resize_coeffs(int in_size, int out_size, int32_t* bounds, int32_t* coeffs) {
  const double scale = (double)in_size / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 2.0 * filterscale, ss = 1.0 / filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  std::vector<double> k(ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      k[x] = bicubic_filter((x + xmin - center + 0.5) * ss);
      ww += k[x];
    }
    for (int x = 0; x < ksize; ++x) {
      double v = x < xmax ? (ww != 0.0 ? k[x] / ww : k[x]) : 0.0;
      coeffs[(int64_t)xx * ksize + x] = v < 0 ? (int32_t)(-0.5 + v * (1 << PRECISION_BITS)) : (int32_t)(0.5 + v * (1 << PRECISION_BITS));
    }
    bounds[xx * 2 + 0] = xmin;
    bounds[xx * 2 + 1] = xmax;
  }
}

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= PRECISION_BITS;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// horizontal pass: src [n, Hs, Ws, 3] -> tmp [n, Hs, crop_w, 3] for resized columns [x0, x0 + crop_w)
__global__ void __launch_bounds__(256) resize_h_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ tmp, int Hs, int Ws,
                                                       int crop_w, int x0, const int32_t* __restrict__ bounds,
                                                       const int32_t* __restrict__ coeffs, int ksize, int row_lo, int rows) {
  const int img = blockIdx.y;
  const int total = rows * crop_w;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int y = row_lo + i / crop_w, xo = i % crop_w;
    const int xx = x0 + xo;
    const int xmin = bounds[xx * 2], cnt = bounds[xx * 2 + 1];
    const int32_t* k = coeffs + (int64_t)xx * ksize;
    const uint8_t* p = src + (((int64_t)img * Hs + y) * Ws + xmin) * 3;
    int a0 = 1 << (PRECISION_BITS - 1), a1 = a0, a2 = a0;
    for (int x = 0; x < cnt; ++x) {
      const int c = __ldg(k + x);
      a0 += p[x * 3 + 0] * c; a1 += p[x * 3 + 1] * c; a2 += p[x * 3 + 2] * c;
    }
    uint8_t* o = tmp + (((int64_t)img * Hs + y) * crop_w + xo) * 3;
    o[0] = clip8(a0); o[1] = clip8(a1); o[2] = clip8(a2);
  }
}

// vertical pass + crop + ToTensor + Normalize: tmp [n, Hs, crop_w, 3] -> out NCHW [n, 3, crop_h, crop_w]
template <typename TOut>
__global__ void __launch_bounds__(256) resize_v_kernel(const uint8_t* __restrict__ tmp, TOut* __restrict__ out, int Hs, int crop_h,
                                                       int crop_w, int y0, const int32_t* __restrict__ bounds,
                                                       const int32_t* __restrict__ coeffs, int ksize, float m0, float m1, float m2,
                                                       float s0, float s1, float s2) {
  const int img = blockIdx.y;
  const int total = crop_h * crop_w;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int yo = i / crop_w, xo = i % crop_w;
    const int yy = y0 + yo;
    const int ymin = bounds[yy * 2], cnt = bounds[yy * 2 + 1];
    const int32_t* k = coeffs + (int64_t)yy * ksize;
    const uint8_t* p = tmp + (((int64_t)img * Hs + ymin) * crop_w + xo) * 3;
    int a0 = 1 << (PRECISION_BITS - 1), a1 = a0, a2 = a0;
    for (int y = 0; y < cnt; ++y) {
      const int c = __ldg(k + y);
      const uint8_t* q = p + (int64_t)y * crop_w * 3;
      a0 += q[0] * c; a1 += q[1] * c; a2 += q[2] * c;
    }
    // ToTensor: uint8 -> float / 255; Normalize: (x - mean) / std   (IEEE division, same operation order as torchvision)
    const float v0 = ((float)clip8(a0) / 255.0f - m0) / s0;
    const float v1 = ((float)clip8(a1) / 255.0f - m1) / s1;
    const float v2 = ((float)clip8(a2) / 255.0f - m2) / s2;
    TOut* o = out + (int64_t)img * 3 * total + i;
    o[0] = Elem<TOut>::from_f(v0); o[total] = Elem<TOut>::from_f(v1); o[2 * total] = Elem<TOut>::from_f(v2);
  }
}

int launch_preprocess(const uint8_t* src, int n, int Hs, int Ws, int Hr, int Wr, int top, int left, int crop_h, int crop_w,
                      const int32_t* bounds_x, const int32_t* coeffs_x, int ksize_x, const int32_t* bounds_y,
                      const int32_t* coeffs_y, int ksize_y, const float* mean, const float* stdv, uint8_t* tmp, void* out,
                      int out_precision, cudaStream_t st) {
  if (n <= 0 || Hs <= 0 || Ws <= 0 || top < 0 || left < 0 || top + crop_h > Hr || left + crop_w > Wr || n > 65535) {
    set_error("preprocess: bad geometry (crop %dx%d at (%d,%d) of %dx%d)", crop_h, crop_w, top, left, Hr, Wr);
    return SEMDIFF_ERR_ARG;
  }
  const int total_h = Hs * crop_w, total_v = crop_h * crop_w;
  dim3 gh((unsigned)std::min((total_h + 255) / 256, 256), (unsigned)n), gv((unsigned)std::min((total_v + 255) / 256, 256), (unsigned)n);
  resize_h_kernel<<<gh, 256, 0, st>>>(src, tmp, Hs, Ws, crop_w, left, bounds_x, coeffs_x, ksize_x, 0, Hs);
  SEMDIFF_CUDA_OK(cudaGetLastError());
  switch (out_precision) {
    case SEMDIFF_FP32:
      resize_v_kernel<float><<<gv, 256, 0, st>>>(tmp, (float*)out, Hs, crop_h, crop_w, top, bounds_y, coeffs_y, ksize_y, mean[0], mean[1], mean[2], stdv[0], stdv[1], stdv[2]);
      break;
    case SEMDIFF_BF16:
      resize_v_kernel<__nv_bfloat16><<<gv, 256, 0, st>>>(tmp, (__nv_bfloat16*)out, Hs, crop_h, crop_w, top, bounds_y, coeffs_y, ksize_y, mean[0], mean[1], mean[2], stdv[0], stdv[1], stdv[2]);
      break;
    case SEMDIFF_FP16:
      resize_v_kernel<__half><<<gv, 256, 0, st>>>(tmp, (__half*)out, Hs, crop_h, crop_w, top, bounds_y, coeffs_y, ksize_y, mean[0], mean[1], mean[2], stdv[0], stdv[1], stdv[2]);
      break;
    default: set_error("preprocess: bad output precision %d", out_precision); return SEMDIFF_ERR_ARG;
  }
  SEMDIFF_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace semdiff

extern "C" {
#pragma GCC visibility push(default)
int32_t semdiff_resize_ksize(int32_t in_size, int32_t out_size) {
  if (in_size <= 0 || out_size <= 0) return SEMDIFF_ERR_ARG;
  return semdiff::resize_ksize(in_size, out_size);
}
int semdiff_resize_coeffs(int32_t in_size, int32_t out_size, int32_t* bounds, int32_t* coeffs) {
  if (in_size <= 0 || out_size <= 0 || bounds == nullptr || coeffs == nullptr) { semdiff::set_error("resize_coeffs: bad arguments"); return SEMDIFF_ERR_ARG; }
  semdiff::resize_coeffs(in_size, out_size, bounds, coeffs);
  return 0;
}
int semdiff_preprocess_u8(const uint8_t* src, int32_t n, int32_t Hs, int32_t Ws, int32_t Hr, int32_t Wr, int32_t top, int32_t left,
                          int32_t crop_h, int32_t crop_w, const int32_t* bounds_x, const int32_t* coeffs_x, int32_t ksize_x,
                          const int32_t* bounds_y, const int32_t* coeffs_y, int32_t ksize_y, const float* mean, const float* stdv,
                          uint8_t* tmp, void* out, int32_t out_precision, semdiff_stream_t st) {
  if (src == nullptr || tmp == nullptr || out == nullptr || mean == nullptr || stdv == nullptr) { semdiff::set_error("preprocess: null argument"); return SEMDIFF_ERR_ARG; }
  return semdiff::launch_preprocess(src, n, Hs, Ws, Hr, Wr, top, left, crop_h, crop_w, bounds_x, coeffs_x, ksize_x, bounds_y,
                                    coeffs_y, ksize_y, mean, stdv, tmp, out, out_precision, reinterpret_cast<cudaStream_t>(st));
}
#pragma GCC visibility pop
}
