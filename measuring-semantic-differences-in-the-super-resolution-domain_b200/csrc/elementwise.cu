// Bandwidth-bound helpers of the trunk: image packing, max/avg pooling.  NHWC, 128-bit accesses.
#include <algorithm>
#include <type_traits>

#include "common.cuh"
#include "kernels.h"

namespace semdiff {

// two consecutive image elements (fp32 / bf16 / fp16 input planes) as floats
template <typename TIn> __device__ __forceinline__ void load2(const TIn* p, float& a, float& b) {
  if constexpr (sizeof(TIn) == 4) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(p));
    a = v.x; b = v.y;
  } else {
    const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(p));
    const TIn* h = reinterpret_cast<const TIn*>(&u);
    a = Elem<TIn>::to_f(h[0]); b = Elem<TIn>::to_f(h[1]);
  }
}

// fp32 NCHW [n,3,H,W] (gt, sr) -> NHWC [2n,H,W,8], channels 3..7 = 0.  One thread per pixel: the three
// plane reads are coalesced across the warp, the write is one 16 B (32 B for fp32) vector per thread.
template <typename T, typename TIn>
__global__ void __launch_bounds__(256) pack_kernel(const TIn* __restrict__ gt, const TIn* __restrict__ sr,
                                                   int n_pairs, int img0, int n_imgs, int hw, T* __restrict__ out) {
  pdl_trigger();
  pdl_wait();  // inputs are the previous kernel's output
  const int64_t total = (int64_t)n_imgs * hw;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int img = img0 + (int)(i / hw);
    const int pix = (int)(i % hw);
    const TIn* src = (img < n_pairs ? gt + (int64_t)img * 3 * hw : sr + (int64_t)(img - n_pairs) * 3 * hw) + pix;
    float f[8] = {Elem<TIn>::to_f(src[0]), Elem<TIn>::to_f(src[hw]), Elem<TIn>::to_f(src[2 * hw]), 0.f, 0.f, 0.f, 0.f, 0.f};
    if constexpr (sizeof(T) == 2) {
      *reinterpret_cast<uint4*>(out + i * 8) = pack8<T>(f);
    } else {
      float4* o = reinterpret_cast<float4*>(out + i * 8);
      o[0] = make_float4(f[0], f[1], f[2], 0.f);
      o[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// Stem layout (SEMDIFF_INPUT_S2D_ROW4) for 7x7 stride-2 pad-3 stems: the conv becomes a 4x1 stride-1 conv over
//   X2[img, i, q, j*16 + (dy*2+dx)*3 + ci] = x[img, ci, 2*(i-2)+dy, 2*(q-2+j)+dx]   (0 outside the image, channels
//   12..15 of every 16 are 0), i in [0, H/2+3), q in [0, W/2), j in [0, 4)
// i.e. a 2x2 space-to-depth of the image with the four horizontally adjacent s2d pixels of each output column
// laid side by side (64 "channels" = one 128-byte row), so the implicit GEMM runs with K = 4 x 64 instead of the
// 49 x 8 a channel-padded 7x7 window would need.  One thread per (img, i, q, j): 6 float2 reads, 32 bytes written.
// The same kernel serves SEMDIFF_INPUT_S2D_ROW2 (3x3 stride-2 pad-1 stems, CLIP): window of 2 s2d pixels starting one
// to the left, one padding row on top (i -> y = 2*(i-1)+dy), slots j = 2, 3 zero (so the row is still 64 wide).
// kSplit (split precisions): the pixel row is 128 stored values, [64 hi | 64 lo]; slot j writes 16 hi values at j*16 and
// the 16 lo values (x - hi, rounded) at 64 + j*16.
template <typename T, typename TIn, bool kSplit = false>
__global__ void __launch_bounds__(256) pack_s2d_kernel(const TIn* __restrict__ gt, const TIn* __restrict__ sr,
                                                       int n_pairs, int img0, int H, int W, T* __restrict__ out,
                                                       int j_real, int off) {
  pdl_trigger();
  pdl_wait();  // inputs are the previous kernel's output
  const int H2 = H / 2 + (j_real == 4 ? 3 : 1), W2 = W / 2;
  const int per_img = H2 * W2 * 4;
  const int img = img0 + blockIdx.y;
  const int plane = H * W;
  const TIn* src = img < n_pairs ? gt + (int64_t)img * 3 * plane : sr + (int64_t)(img - n_pairs) * 3 * plane;
  T* out_img = out + (int64_t)blockIdx.y * per_img * (kSplit ? 32 : 16);
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < per_img; t += gridDim.x * blockDim.x) {
    const int j = t & 3;
    const int r = t >> 2;
    const int i = r / W2, q = r - i * W2;
    const int y0 = 2 * (i - off), x0 = 2 * (q - off + j);
    float f[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) f[k] = 0.f;
    if (j < j_real && x0 >= 0 && x0 < W) {
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        const int y = y0 + dy;
        if (y < 0 || y >= H) continue;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
          load2<TIn>(src + ci * plane + y * W + x0, f[(dy * 2 + 0) * 3 + ci], f[(dy * 2 + 1) * 3 + ci]);
        }
      }
    }
    if constexpr (kSplit) {
      T* dst = out_img + (int64_t)r * 128 + j * 16;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float v[8], h[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = f[half * 8 + k];
        const uint4 qh = pack8<T>(v);
        unpack8<T>(qh, h);
#pragma unroll
        for (int k = 0; k < 8; ++k) h[k] = v[k] - h[k];
        reinterpret_cast<uint4*>(dst)[half] = qh;
        reinterpret_cast<uint4*>(dst + 64)[half] = pack8<T>(h);
      }
      continue;
    }
    T* dst = out_img + (int64_t)t * 16;
    if constexpr (sizeof(T) == 2) {
      float lo[8], hi[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) { lo[k] = f[k]; hi[k] = f[8 + k]; }
      reinterpret_cast<uint4*>(dst)[0] = pack8<T>(lo);
      reinterpret_cast<uint4*>(dst)[1] = pack8<T>(hi);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) reinterpret_cast<float4*>(dst)[k] = make_float4(f[4 * k], f[4 * k + 1], f[4 * k + 2], f[4 * k + 3]);
    }
  }
}

// SEMDIFF_INPUT_S2D16: plain 2x2 space-to-depth, out[img, i, q, (dy*2+dx)*3+ci] = x[img, ci, 2i+dy, 2q+dx], channels 12..15
// zero; one thread per s2d pixel: 6 two-element reads, 32 bytes written.  4x smaller than the row-window layouts: the
// strip conv kernel forms the windows as shifted shared-memory views instead of materialising them.
template <typename T, typename TIn>
__global__ void __launch_bounds__(256) pack_s2d16_kernel(const TIn* __restrict__ gt, const TIn* __restrict__ sr, int n_pairs,
                                                         int img0, int H, int W, T* __restrict__ out) {
  pdl_trigger();
  pdl_wait();  // inputs are the previous kernel's output
  const int H2 = H / 2, W2 = W / 2;
  const int per_img = H2 * W2;
  const int img = img0 + blockIdx.y;
  const int plane = H * W;
  const TIn* src = img < n_pairs ? gt + (int64_t)img * 3 * plane : sr + (int64_t)(img - n_pairs) * 3 * plane;
  T* out_img = out + (int64_t)blockIdx.y * per_img * 16;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < per_img; t += gridDim.x * blockDim.x) {
    const int i = t / W2, q = t - i * W2;
    float f[16];
#pragma unroll
    for (int k = 12; k < 16; ++k) f[k] = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int ci = 0; ci < 3; ++ci)
        load2<TIn>(src + ci * plane + (2 * i + dy) * W + 2 * q, f[(dy * 2 + 0) * 3 + ci], f[(dy * 2 + 1) * 3 + ci]);
    T* dst = out_img + (int64_t)t * 16;
    if constexpr (sizeof(T) == 2) {
      float lo[8], hi[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) { lo[k] = f[k]; hi[k] = f[8 + k]; }
      reinterpret_cast<uint4*>(dst)[0] = pack8<T>(lo);
      reinterpret_cast<uint4*>(dst)[1] = pack8<T>(hi);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) reinterpret_cast<float4*>(dst)[k] = make_float4(f[4 * k], f[4 * k + 1], f[4 * k + 2], f[4 * k + 3]);
    }
  }
}

template <typename T> struct Vec8 {
  static __device__ __forceinline__ void load(const T* p, float (&f)[8]) {
    if constexpr (sizeof(T) == 2) {
      uint4 q = *reinterpret_cast<const uint4*>(p);
      unpack8<T>(q, f);
    } else {
      float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
      f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
  }
  static __device__ __forceinline__ void store(T* p, const float (&f)[8]) {
    if constexpr (sizeof(T) == 2) {
      *reinterpret_cast<uint4*>(p) = pack8<T>(f);
    } else {
      reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
      reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
    }
  }
};

// 3x3 stride-2 pad-1 max pool (torch MaxPool2d semantics: padding never wins).  grid.y = image; 32-bit index math;
// 16-bit types take the maximum on packed pairs without widening.
template <typename T> __device__ __forceinline__ uint4 max8(const uint4& a, const uint4& b) {
  uint4 r;
  if constexpr (sizeof(T) == 2) {
    using T2 = typename std::conditional<std::is_same<T, __half>::value, __half2, __nv_bfloat162>::type;
    const T2* pa = reinterpret_cast<const T2*>(&a);
    const T2* pb = reinterpret_cast<const T2*>(&b);
    T2* pr = reinterpret_cast<T2*>(&r);
#pragma unroll
    for (int k = 0; k < 4; ++k) pr[k] = __hmax2(pa[k], pb[k]);
  }
  return r;
}
template <typename T>
__global__ void __launch_bounds__(256) maxpool_kernel(const T* __restrict__ in, T* __restrict__ out, int H, int W, int C,
                                                      int OH, int OW) {
  pdl_trigger();
  pdl_wait();  // inputs are the previous kernel's output
  const int cv = C / 8;
  const int per_img = OH * OW * cv;
  const T* img_in = in + (int64_t)blockIdx.y * H * W * C;
  T* img_out = out + (int64_t)blockIdx.y * per_img * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per_img; i += gridDim.x * blockDim.x) {
    const int c8 = i % cv;
    const int p = i / cv;
    const int oh = p / OW, ow = p - oh * OW;
    if constexpr (sizeof(T) == 2) {
      uint4 m;
      bool first = true;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int ih = oh * 2 - 1 + r;
        if (ih < 0 || ih >= H) continue;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int iw = ow * 2 - 1 + s;
          if (iw < 0 || iw >= W) continue;
          const uint4 v = *reinterpret_cast<const uint4*>(img_in + ((int64_t)ih * W + iw) * C + c8 * 8);
          m = first ? v : max8<T>(m, v);
          first = false;
        }
      }
      *reinterpret_cast<uint4*>(img_out + (int64_t)i * 8) = m;
    } else {
      float m[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) m[k] = -INFINITY;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int ih = oh * 2 - 1 + r;
        if (ih < 0 || ih >= H) continue;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int iw = ow * 2 - 1 + s;
          if (iw < 0 || iw >= W) continue;
          float f[8];
          Vec8<T>::load(img_in + ((int64_t)ih * W + iw) * C + c8 * 8, f);
#pragma unroll
          for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], f[k]);
        }
      }
      Vec8<T>::store(img_out + (int64_t)i * 8, m);
    }
  }
}

// window x window average pool, stride = window (CLIP anti-aliased strides); trailing rows/cols are dropped
template <typename T>
__global__ void __launch_bounds__(256) avgpool_kernel(const T* __restrict__ in, T* __restrict__ out, int n_img, int H,
                                                      int W, int C, int win) {
  pdl_trigger();
  pdl_wait();  // inputs are the previous kernel's output
  const int cv = C / 8, OH = H / win, OW = W / win;
  const int64_t total = (int64_t)n_img * OH * OW * cv;
  const float inv = 1.f / (float)(win * win);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cv);
    int64_t p = i / cv;
    const int ow = (int)(p % OW); p /= OW;
    const int oh = (int)(p % OH);
    const int n = (int)(p / OH);
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < win; ++r)
      for (int s = 0; s < win; ++s) {
        float f[8];
        Vec8<T>::load(in + (((int64_t)n * H + oh * win + r) * W + ow * win + s) * C + c8 * 8, f);
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] += f[k];
      }
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] *= inv;
    Vec8<T>::store(out + i * 8, a);
  }
}

// ---------------------------------------------------------------------------------------------
// split precisions: a pixel's C logical channels are stored as C/64 blocks of [64 hi | 64 lo]
// ---------------------------------------------------------------------------------------------
// stored offset of the hi half of the 8-channel group c8 (the lo half is 64 elements further)
__device__ __forceinline__ int split_off(int c8) { return (c8 >> 3) * 128 + (c8 & 7) * 8; }

// max of (hi, lo) pairs: hi = round(x) is monotone in x, so the order of x is the lexicographic order of (hi, lo)
template <typename T>
__global__ void __launch_bounds__(256) maxpool_split_kernel(const T* __restrict__ in, T* __restrict__ out, int H, int W, int C,
                                                            int OH, int OW) {
  pdl_trigger();
  pdl_wait();
  const int cv = C / 8;
  const int per_img = OH * OW * cv;
  const T* img_in = in + (int64_t)blockIdx.y * H * W * C * 2;
  T* img_out = out + (int64_t)blockIdx.y * OH * OW * C * 2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per_img; i += gridDim.x * blockDim.x) {
    const int c8 = i % cv;
    const int p = i / cv;
    const int oh = p / OW, ow = p - oh * OW;
    const int so = split_off(c8);
    // all 18 loads are unconditional and in flight together: a tap outside the image is replaced by the window centre,
    // which is always valid (max is idempotent)
    uint4 qh[9], ql[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      int ih = oh * 2 - 1 + r;
      ih = (ih < 0 || ih >= H) ? oh * 2 : ih;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        int iw = ow * 2 - 1 + s;
        iw = (iw < 0 || iw >= W) ? ow * 2 : iw;
        const T* px = img_in + ((int64_t)ih * W + iw) * C * 2 + so;
        qh[r * 3 + s] = __ldg(reinterpret_cast<const uint4*>(px));        // neighbouring outputs share taps: keep them in L1
        ql[r * 3 + s] = __ldg(reinterpret_cast<const uint4*>(px + 64));
      }
    }
    float mh[8], ml[8];
    unpack8<T>(qh[0], mh);
    unpack8<T>(ql[0], ml);
#pragma unroll
    for (int t = 1; t < 9; ++t) {
      float h[8], l[8];
      unpack8<T>(qh[t], h);
      unpack8<T>(ql[t], l);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const bool better = h[k] > mh[k] || (h[k] == mh[k] && l[k] > ml[k]);
        mh[k] = better ? h[k] : mh[k];
        ml[k] = better ? l[k] : ml[k];
      }
    }
    T* dst = img_out + (int64_t)p * C * 2 + so;
    *reinterpret_cast<uint4*>(dst) = pack8<T>(mh);
    *reinterpret_cast<uint4*>(dst + 64) = pack8<T>(ml);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) avgpool_split_kernel(const T* __restrict__ in, T* __restrict__ out, int n_img, int H,
                                                            int W, int C, int win) {
  pdl_trigger();
  pdl_wait();
  const int cv = C / 8, OH = H / win, OW = W / win;
  const int64_t total = (int64_t)n_img * OH * OW * cv;
  const float inv = 1.f / (float)(win * win);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cv);
    int64_t p = i / cv;
    const int ow = (int)(p % OW); p /= OW;
    const int oh = (int)(p % OH);
    const int n = (int)(p / OH);
    const int so = split_off(c8);
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < win; ++r)
      for (int s = 0; s < win; ++s) {
        const T* px = in + (((int64_t)n * H + oh * win + r) * W + ow * win + s) * C * 2 + so;
        float h[8], l[8];
        unpack8<T>(*reinterpret_cast<const uint4*>(px), h);
        unpack8<T>(*reinterpret_cast<const uint4*>(px + 64), l);
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] += h[k] + l[k];
      }
    float h[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] *= inv;
    const uint4 qh = pack8<T>(a);
    unpack8<T>(qh, h);
#pragma unroll
    for (int k = 0; k < 8; ++k) h[k] = a[k] - h[k];
    T* dst = out + (((int64_t)n * OH + oh) * OW + ow) * C * 2 + so;
    *reinterpret_cast<uint4*>(dst) = qh;
    *reinterpret_cast<uint4*>(dst + 64) = pack8<T>(h);
  }
}

static int grid_for(int64_t total, int block) {
  int64_t g = (total + block - 1) / block;
  const int64_t cap = 148 * 32;  // a few waves of 256-thread CTAs over 148 SMs; grid-stride covers the rest
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

template <typename T, typename TIn, bool kSplit = false>
static int pack_t(const void* gt_, const void* sr_, int n_pairs, int img0, int n_imgs, int H, int W, void* out,
                  int layout, cudaStream_t st) {
  const TIn* gt = (const TIn*)gt_;
  const TIn* sr = (const TIn*)sr_;
  if constexpr (kSplit) {
    if (layout != SEMDIFF_INPUT_S2D_ROW4 && layout != SEMDIFF_INPUT_S2D_ROW2) {
      set_error("pack: the split precisions take the row-window stem layouts only (even image sizes)");
      return SEMDIFF_ERR_UNSUPPORTED;
    }
    const bool row4 = layout == SEMDIFF_INPUT_S2D_ROW4;
    const int per_img = (H / 2 + (row4 ? 3 : 1)) * (W / 2) * 4;
    dim3 grid((unsigned)std::min((per_img + 255) / 256, 64), (unsigned)n_imgs);
    launch_pdl(pack_s2d_kernel<T, TIn, true>, dim3(grid), dim3(256), 0, st, gt, sr, n_pairs, img0, H, W, (T*)out, row4 ? 4 : 2, row4 ? 2 : 1);
    SEMDIFF_CUDA_OK(cudaGetLastError());
    return 0;
  } else
  if (layout == SEMDIFF_INPUT_S2D16) {
    const int per_img = (H / 2) * (W / 2);
    dim3 grid((unsigned)std::min((per_img + 255) / 256, 64), (unsigned)n_imgs);
    launch_pdl(pack_s2d16_kernel<T, TIn>, dim3(grid), dim3(256), 0, st, gt, sr, n_pairs, img0, H, W, (T*)out);
  } else if (layout == SEMDIFF_INPUT_S2D_ROW4 || layout == SEMDIFF_INPUT_S2D_ROW2) {
    const bool row4 = layout == SEMDIFF_INPUT_S2D_ROW4;
    const int per_img = (H / 2 + (row4 ? 3 : 1)) * (W / 2) * 4;
    dim3 grid((unsigned)std::min((per_img + 255) / 256, 64), (unsigned)n_imgs);
    launch_pdl(pack_s2d_kernel<T, TIn>, dim3(grid), dim3(256), 0, st, gt, sr, n_pairs, img0, H, W, (T*)out, row4 ? 4 : 2, row4 ? 2 : 1);
  } else {
    const int64_t total = (int64_t)n_imgs * H * W;
    launch_pdl(pack_kernel<T, TIn>, dim3(grid_for(total, 256)), dim3(256), 0, st, gt, sr, n_pairs, img0, n_imgs, H * W, (T*)out);
  }
  SEMDIFF_CUDA_OK(cudaGetLastError());
  return 0;
}
template <typename T, bool kSplit = false>
static int pack_in(const void* gt, const void* sr, int n_pairs, int img0, int n_imgs, int H, int W, void* out, int layout,
                   int in_precision, cudaStream_t st) {
  switch (in_precision) {
    case SEMDIFF_FP32: return pack_t<T, float, kSplit>(gt, sr, n_pairs, img0, n_imgs, H, W, out, layout, st);
    case SEMDIFF_BF16: return pack_t<T, __nv_bfloat16, kSplit>(gt, sr, n_pairs, img0, n_imgs, H, W, out, layout, st);
    case SEMDIFF_FP16: return pack_t<T, __half, kSplit>(gt, sr, n_pairs, img0, n_imgs, H, W, out, layout, st);
  }
  set_error("pack: bad input precision %d", in_precision);
  return SEMDIFF_ERR_ARG;
}
int launch_pack(const void* gt, const void* sr, int in_precision, int n_pairs, int img0, int n_imgs, int H, int W, void* out,
                int precision, int layout, cudaStream_t st) {
  if (n_pairs <= 0 || H <= 0 || W <= 0 || img0 < 0 || n_imgs <= 0 || img0 + n_imgs > 2 * n_pairs) {
    set_error("pack: bad shape");
    return SEMDIFF_ERR_ARG;
  }
  if (layout != SEMDIFF_INPUT_NHWC8 && ((H | W) & 1)) { set_error("pack: the s2d stem layouts need even H and W"); return SEMDIFF_ERR_ARG; }
  if (layout < SEMDIFF_INPUT_NHWC8 || layout > SEMDIFF_INPUT_S2D16) { set_error("pack: bad layout %d", layout); return SEMDIFF_ERR_ARG; }
  switch (precision) {
    case SEMDIFF_BF16: return pack_in<__nv_bfloat16>(gt, sr, n_pairs, img0, n_imgs, H, W, out, layout, in_precision, st);
    case SEMDIFF_FP16: return pack_in<__half>(gt, sr, n_pairs, img0, n_imgs, H, W, out, layout, in_precision, st);
    case SEMDIFF_FP32: return pack_in<float>(gt, sr, n_pairs, img0, n_imgs, H, W, out, layout, in_precision, st);
    case SEMDIFF_FP16X3: return pack_in<__half, true>(gt, sr, n_pairs, img0, n_imgs, H, W, out, layout, in_precision, st);
    case SEMDIFF_BF16X3: return pack_in<__nv_bfloat16, true>(gt, sr, n_pairs, img0, n_imgs, H, W, out, layout, in_precision, st);
  }
  set_error("pack: bad precision %d", precision);
  return SEMDIFF_ERR_ARG;
}

template <typename T, bool kSplit = false>
static int maxpool_t(const void* in, void* out, int n, int H, int W, int C, cudaStream_t st) {
  const int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
  const int per_img = OH * OW * (C / 8);
  dim3 grid((unsigned)std::min((per_img + 255) / 256, 128), (unsigned)n);
  if constexpr (kSplit) launch_pdl(maxpool_split_kernel<T>, dim3(grid), dim3(256), 0, st, (const T*)in, (T*)out, H, W, C, OH, OW);
  else
  launch_pdl(maxpool_kernel<T>, dim3(grid), dim3(256), 0, st, (const T*)in, (T*)out, H, W, C, OH, OW);
  SEMDIFF_CUDA_OK(cudaGetLastError());
  return 0;
}
int launch_maxpool3x3s2(const void* in, void* out, int n, int H, int W, int C, int precision, cudaStream_t st) {
  if (C % 8 != 0 || n <= 0) { set_error("maxpool: C %% 8 != 0 or empty"); return SEMDIFF_ERR_ARG; }
  if (is_split(precision) && C % 64 != 0) { set_error("maxpool: split precisions need C %% 64 == 0"); return SEMDIFF_ERR_ARG; }
  switch (precision) {
    case SEMDIFF_FP16X3: return maxpool_t<__half, true>(in, out, n, H, W, C, st);
    case SEMDIFF_BF16X3: return maxpool_t<__nv_bfloat16, true>(in, out, n, H, W, C, st);
    case SEMDIFF_BF16: return maxpool_t<__nv_bfloat16>(in, out, n, H, W, C, st);
    case SEMDIFF_FP16: return maxpool_t<__half>(in, out, n, H, W, C, st);
    case SEMDIFF_FP32: return maxpool_t<float>(in, out, n, H, W, C, st);
  }
  set_error("maxpool: bad precision %d", precision);
  return SEMDIFF_ERR_ARG;
}

template <typename T, bool kSplit = false>
static int avgpool_t(const void* in, void* out, int n, int H, int W, int C, int win, cudaStream_t st) {
  const int64_t total = (int64_t)n * (H / win) * (W / win) * (C / 8);
  if constexpr (kSplit) launch_pdl(avgpool_split_kernel<T>, dim3(grid_for(total, 256)), dim3(256), 0, st, (const T*)in, (T*)out, n, H, W, C, win);
  else
  launch_pdl(avgpool_kernel<T>, dim3(grid_for(total, 256)), dim3(256), 0, st, (const T*)in, (T*)out, n, H, W, C, win);
  SEMDIFF_CUDA_OK(cudaGetLastError());
  return 0;
}
int launch_avgpool(const void* in, void* out, int n, int H, int W, int C, int win, int precision, cudaStream_t st) {
  if (C % 8 != 0 || win < 1 || H < win || W < win || n <= 0) {  // floor semantics like torch AvgPool2d
    set_error("avgpool: need C %% 8 == 0 and H, W >= window");
    return SEMDIFF_ERR_ARG;
  }
  if (is_split(precision) && C % 64 != 0) { set_error("avgpool: split precisions need C %% 64 == 0"); return SEMDIFF_ERR_ARG; }
  switch (precision) {
    case SEMDIFF_FP16X3: return avgpool_t<__half, true>(in, out, n, H, W, C, win, st);
    case SEMDIFF_BF16X3: return avgpool_t<__nv_bfloat16, true>(in, out, n, H, W, C, win, st);
    case SEMDIFF_BF16: return avgpool_t<__nv_bfloat16>(in, out, n, H, W, C, win, st);
    case SEMDIFF_FP16: return avgpool_t<__half>(in, out, n, H, W, C, win, st);
    case SEMDIFF_FP32: return avgpool_t<float>(in, out, n, H, W, C, win, st);
  }
  set_error("avgpool: bad precision %d", precision);
  return SEMDIFF_ERR_ARG;
}

// ---------------------------------------------------------------------------------------------
// Local-map decoder helpers (the reference's U-Net over squared feature differences,
// /root/reference/models/local_eval_models.py:109-125): (a - b)^2 tensors, channel concat, bilinear x2
// upsampling with align_corners = True (nn.UpsamplingBilinear2d, :84), final upsample + sigmoid.
// ---------------------------------------------------------------------------------------------
// 8 logical channels <-> stored representation, any precision: T = fp32 (plain), 16-bit (plain) or 16-bit split
template <typename T, bool kSplit> struct Val8 {
  // logical 8-channel chunk i of a tensor whose pixels hold C logical channels (C % 64 == 0 when split)
  static __device__ __forceinline__ int64_t offset(int64_t i) { return kSplit ? (i >> 3) * 128 + (i & 7) * 8 : i * 8; }
  static __device__ __forceinline__ void load(const T* base, int64_t i, float (&f)[8]) {
    const T* p = base + offset(i);
    Vec8<T>::load(p, f);
    if constexpr (kSplit) {
      float l[8];
      Vec8<T>::load(p + 64, l);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] += l[k];
    }
  }
  static __device__ __forceinline__ void store(T* base, int64_t i, const float (&f)[8]) {
    T* p = base + offset(i);
    if constexpr (kSplit) {
      const uint4 qh = pack8<T>(f);
      float h[8];
      unpack8<T>(qh, h);
#pragma unroll
      for (int k = 0; k < 8; ++k) h[k] = f[k] - h[k];
      *reinterpret_cast<uint4*>(p) = qh;
      *reinterpret_cast<uint4*>(p + 64) = pack8<T>(h);
    } else {
      Vec8<T>::store(p, f);
    }
  }
};

// in = [2n, HW, C] (GT images first) -> out[n, HW, C] = (gt - sr)^2; chunks = n * HW * C / 8
template <typename T, bool kSplit>
__global__ void __launch_bounds__(256) sqdiff_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t chunks) {
  pdl_trigger();
  pdl_wait();
  const T* b = in + chunks * 8 * (kSplit ? 2 : 1);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < chunks; i += (int64_t)gridDim.x * blockDim.x) {
    float d[8];
    if constexpr (kSplit) {   // (a_hi - b_hi) + (a_lo - b_lo): the hi difference is exact for nearby values
      const int64_t o = Val8<T, true>::offset(i);
      float ah[8], al[8], bh[8], bl[8];
      Vec8<T>::load(in + o, ah); Vec8<T>::load(in + o + 64, al); Vec8<T>::load(b + o, bh); Vec8<T>::load(b + o + 64, bl);
#pragma unroll
      for (int k = 0; k < 8; ++k) d[k] = (ah[k] - bh[k]) + (al[k] - bl[k]);
    } else {
      float fa[8], fb[8];
      Val8<T, false>::load(in, i, fa);
      Val8<T, false>::load(b, i, fb);
#pragma unroll
      for (int k = 0; k < 8; ++k) d[k] = fa[k] - fb[k];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k] *= d[k];
    Val8<T, kSplit>::store(out, i, d);
  }
}

// channel concat of two NHWC tensors with q1 / q2 16-byte chunks per pixel (any element type / split storage)
__global__ void __launch_bounds__(256) concat_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ out,
                                                     int64_t pixels, int q1, int q2) {
  pdl_trigger();
  pdl_wait();
  const int q = q1 + q2;
  const int64_t total = pixels * q;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / q;
    const int c = (int)(i - p * q);
    out[i] = c < q1 ? __ldg(a + p * q1 + c) : __ldg(b + p * q2 + (c - q1));
  }
}

// torch's bilinear source index for align_corners = True: src = dst * (in - 1) / (out - 1), in fp32 like ATen
struct Lerp { int i0, i1; float w0, w1; };
__device__ __forceinline__ Lerp lerp_at(int o, int in_size, float scale) {
  const float s = scale * (float)o;
  int i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  const int i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  const float w1 = s - (float)i0;
  return Lerp{i0, i1, 1.f - w1, w1};
}

// nn.UpsamplingBilinear2d(scale_factor=2): [n, H, W, C] -> [n, 2H, 2W, C]
template <typename T, bool kSplit>
__global__ void __launch_bounds__(256) upsample2x_kernel(const T* __restrict__ in, T* __restrict__ out, int n_img, int H, int W, int C) {
  pdl_trigger();
  pdl_wait();
  const int cv = C / 8, OH = 2 * H, OW = 2 * W;
  const float sy = OH > 1 ? (float)(H - 1) / (float)(OH - 1) : 0.f, sx = OW > 1 ? (float)(W - 1) / (float)(OW - 1) : 0.f;
  const int64_t total = (int64_t)n_img * OH * OW * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cv);
    int64_t p = i / cv;
    const int ox = (int)(p % OW); p /= OW;
    const int oy = (int)(p % OH);
    const int64_t n = p / OH;
    const Lerp ly = lerp_at(oy, H, sy), lx = lerp_at(ox, W, sx);
    float v00[8], v01[8], v10[8], v11[8], r[8];
    Val8<T, kSplit>::load(in, ((n * H + ly.i0) * W + lx.i0) * cv + c8, v00);
    Val8<T, kSplit>::load(in, ((n * H + ly.i0) * W + lx.i1) * cv + c8, v01);
    Val8<T, kSplit>::load(in, ((n * H + ly.i1) * W + lx.i0) * cv + c8, v10);
    Val8<T, kSplit>::load(in, ((n * H + ly.i1) * W + lx.i1) * cv + c8, v11);
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = ly.w0 * (lx.w0 * v00[k] + lx.w1 * v01[k]) + ly.w1 * (lx.w0 * v10[k] + lx.w1 * v11[k]);
    Val8<T, kSplit>::store(out, i, r);
  }
}

// final step of the decoder: channel 0 of [n, H, W, C] -> bilinear x2 -> sigmoid -> fp32 [n, 1, 2H, 2W]
template <typename T, bool kSplit>
__global__ void __launch_bounds__(256) map_out_kernel(const T* __restrict__ in, float* __restrict__ out, int n_img, int H, int W, int C) {
  pdl_trigger();
  pdl_wait();
  const int OH = 2 * H, OW = 2 * W, CS = kSplit ? 2 * C : C;
  const float sy = OH > 1 ? (float)(H - 1) / (float)(OH - 1) : 0.f, sx = OW > 1 ? (float)(W - 1) / (float)(OW - 1) : 0.f;
  const int64_t total = (int64_t)n_img * OH * OW;
  auto at = [&](int64_t n, int y, int x) {
    const T* p = in + ((n * H + y) * W + x) * CS;
    float v = Elem<T>::to_f(p[0]);
    if constexpr (kSplit) v += Elem<T>::to_f(p[64]);
    return v;
  };
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = i;
    const int ox = (int)(p % OW); p /= OW;
    const int oy = (int)(p % OH);
    const int64_t n = p / OH;
    const Lerp ly = lerp_at(oy, H, sy), lx = lerp_at(ox, W, sx);
    const float v = ly.w0 * (lx.w0 * at(n, ly.i0, lx.i0) + lx.w1 * at(n, ly.i0, lx.i1)) +
                    ly.w1 * (lx.w0 * at(n, ly.i1, lx.i0) + lx.w1 * at(n, ly.i1, lx.i1));
    out[i] = 1.f / (1.f + expf(-v));
  }
}

template <typename T, bool kSplit>
static int decoder_op_t(int what, const void* in, const void* in2, void* out, int n_img, int H, int W, int C, int C2, cudaStream_t st) {
  const int64_t chunks = (int64_t)n_img * H * W * C / 8;
  switch (what) {
    case 0: launch_pdl(sqdiff_kernel<T, kSplit>, dim3(grid_for(chunks, 256)), dim3(256), 0, st, (const T*)in, (T*)out, chunks); break;
    case 1: {
      const int eb = (int)sizeof(T) * (kSplit ? 2 : 1);
      const int64_t pixels = (int64_t)n_img * H * W;
      launch_pdl(concat_kernel, dim3(grid_for(pixels * (C + C2) * eb / 16, 256)), dim3(256), 0, st, (const uint4*)in, (const uint4*)in2,
                 (uint4*)out, pixels, C * eb / 16, C2 * eb / 16);
      break;
    }
    case 2: launch_pdl(upsample2x_kernel<T, kSplit>, dim3(grid_for(chunks * 4, 256)), dim3(256), 0, st, (const T*)in, (T*)out, n_img, H, W, C); break;
    case 3: launch_pdl(map_out_kernel<T, kSplit>, dim3(grid_for((int64_t)n_img * H * W * 4, 256)), dim3(256), 0, st, (const T*)in, (float*)out, n_img, H, W, C); break;
  }
  SEMDIFF_CUDA_OK(cudaGetLastError());
  return 0;
}

// what: 0 = squared difference of the two halves of a stacked batch (n_img = PAIRS), 1 = channel concat (C | C2),
// 2 = bilinear x2 upsampling, 3 = channel 0 -> bilinear x2 -> sigmoid -> fp32 map
int launch_decoder_op(int what, const void* in, const void* in2, void* out, int n_img, int H, int W, int C, int C2, int precision,
                      cudaStream_t st) {
  if (n_img <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 8 != 0 || (what == 1 && (C2 <= 0 || C2 % 8 != 0))) {
    set_error("decoder op %d: bad shape n=%d %dx%dx%d (+%d)", what, n_img, H, W, C, C2);
    return SEMDIFF_ERR_ARG;
  }
  if (is_split(precision) && (C % 64 != 0 || (what == 1 && C2 % 64 != 0))) { set_error("decoder op: split precisions need C %% 64 == 0"); return SEMDIFF_ERR_ARG; }
  switch (precision) {
    case SEMDIFF_BF16: return decoder_op_t<__nv_bfloat16, false>(what, in, in2, out, n_img, H, W, C, C2, st);
    case SEMDIFF_FP16: return decoder_op_t<__half, false>(what, in, in2, out, n_img, H, W, C, C2, st);
    case SEMDIFF_FP32: return decoder_op_t<float, false>(what, in, in2, out, n_img, H, W, C, C2, st);
    case SEMDIFF_FP16X3: return decoder_op_t<__half, true>(what, in, in2, out, n_img, H, W, C, C2, st);
    case SEMDIFF_BF16X3: return decoder_op_t<__nv_bfloat16, true>(what, in, in2, out, n_img, H, W, C, C2, st);
  }
  set_error("decoder op: bad precision %d", precision);
  return SEMDIFF_ERR_ARG;
}

}  // namespace semdiff
