// Fused per-layer feature distance + aggregation head.
//   reference: /root/reference/models/global_eval_models.py:379-395 (twin :755-771)
//     diff = (a-b)**2 ; w_layers[j](diff) ; mean over W, H ; mean over layers ; ReLU
// One pass over each tapped GT/SR activation: 128-bit streaming loads, fp32 math, channel weights held in
// registers, warp-shuffle + fixed-order block reduction, NO atomics: a pair's partial sums depend only on
// (hw, c), never on the batch size or its position in the batch, so 1-GPU and N-GPU sweeps agree bit for bit.
#include "common.cuh"
#include "kernels.h"

namespace semdiff {

constexpr int DIST_THREADS = 256;
constexpr int DIST_CHUNK_TARGET = 4096;  // 16-byte-equivalent chunks (8 elements) per CTA

int distance_parts(int hw, int c) {
  const int64_t chunks = (int64_t)hw * c / 8;
  int64_t parts = (chunks + DIST_CHUNK_TARGET - 1) / DIST_CHUNK_TARGET;
  if (parts < 1) parts = 1;
  if (parts > SEMDIFF_MAX_PARTS) parts = SEMDIFF_MAX_PARTS;
  return (int)parts;
}

template <typename T> __device__ __forceinline__ void load8_stream(const T* p, float (&f)[8]) {
  if constexpr (sizeof(T) == 2) {
    uint4 q = ld_nc_u4(p);
    unpack8<T>(q, f);
  } else {
    uint4 q0 = ld_nc_u4(p), q1 = ld_nc_u4(p + 4);
    f[0] = __uint_as_float(q0.x); f[1] = __uint_as_float(q0.y); f[2] = __uint_as_float(q0.z); f[3] = __uint_as_float(q0.w);
    f[4] = __uint_as_float(q1.x); f[5] = __uint_as_float(q1.y); f[6] = __uint_as_float(q1.z); f[7] = __uint_as_float(q1.w);
  }
}

// 8 logical channels starting at channel c (multiple of 8) of the pixel whose stored channel vector starts at px.
// Split precisions store a pixel as blocks of [64 hi | 64 lo]; the value is hi + lo.
template <typename T, bool kSplit> __device__ __forceinline__ void load8_px(const T* px, int c, float (&f)[8]) {
  if constexpr (kSplit) {
    const T* q = px + (c >> 6) * 128 + (c & 63);
    float l[8];
    unpack8<T>(*reinterpret_cast<const uint4*>(q), f);
    unpack8<T>(*reinterpret_cast<const uint4*>(q + 64), l);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] += l[k];
  } else if constexpr (sizeof(T) == 2) {
    unpack8<T>(*reinterpret_cast<const uint4*>(px + c), f);
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = Elem<T>::to_f(px[c + k]);
  }
}

__device__ __forceinline__ float block_reduce_fixed(float v, float* smem) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < DIST_THREADS / 32; ++w) t += smem[w];  // fixed order
  }
  return t;
}

// kRegW: 8 * DIST_THREADS % C == 0, so a thread sees the same 8 channels on every iteration and keeps
// their weights in registers.  grid = (n_parts, n_pairs).
template <typename T, bool kRegW>
__global__ void __launch_bounds__(DIST_THREADS) distance_kernel(const T* __restrict__ act, int n_pairs, int64_t elems,
                                                                int C, const float* __restrict__ w,
                                                                int64_t chunks_per_part, float* __restrict__ partial) {
  pdl_trigger();
  pdl_wait();  // inputs are the previous kernel's output
  __shared__ float red[DIST_THREADS / 32];
  const int pair = blockIdx.y, part = blockIdx.x;
  const T* a = act + (int64_t)pair * elems;
  const T* b = act + (int64_t)(pair + n_pairs) * elems;
  const int64_t total_chunks = elems / 8;
  const int64_t c_begin = (int64_t)part * chunks_per_part;
  int64_t c_end = c_begin + chunks_per_part;
  if (c_end > total_chunks) c_end = total_chunks;

  float wr[8];
  if constexpr (kRegW) {
    const int c0 = (int)(((c_begin + threadIdx.x) * 8) % C);
#pragma unroll
    for (int k = 0; k < 8; ++k) wr[k] = __ldg(w + c0 + k);
  }
  float acc = 0.f;
  int64_t i = c_begin + threadIdx.x;
  // 4 independent 128-bit loads per operand in flight per thread
  for (; i + 3 * DIST_THREADS < c_end; i += 4 * DIST_THREADS) {
    float fa[4][8], fb[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      load8_stream<T>(a + (i + u * DIST_THREADS) * 8, fa[u]);
      load8_stream<T>(b + (i + u * DIST_THREADS) * 8, fb[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if constexpr (!kRegW) {
        const int c0 = (int)(((i + u * DIST_THREADS) * 8) % C);
#pragma unroll
        for (int k = 0; k < 8; ++k) wr[k] = __ldg(w + c0 + k);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) { const float d = fa[u][k] - fb[u][k]; acc = fmaf(wr[k], d * d, acc); }
    }
  }
  for (; i < c_end; i += DIST_THREADS) {
    float fa[8], fb[8];
    load8_stream<T>(a + i * 8, fa);
    load8_stream<T>(b + i * 8, fb);
    if constexpr (!kRegW) {
      const int c0 = (int)((i * 8) % C);
#pragma unroll
      for (int k = 0; k < 8; ++k) wr[k] = __ldg(w + c0 + k);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) { const float d = fa[k] - fb[k]; acc = fmaf(wr[k], d * d, acc); }
  }
  const float t = block_reduce_fixed(acc, red);
  if (threadIdx.x == 0) partial[(int64_t)pair * SEMDIFF_MAX_PARTS + part] = t;
}

// Split precisions: logical 8-channel chunk i of an image lives at stored offset (i / 8) * 128 + (i % 8) * 8 (hi) and 64
// elements further (lo).  d = (ah - bh) + (al - bl): the hi difference is exact in fp32 for nearby values, so the
// difference keeps the full hi + lo resolution even when a and b agree in their leading bits (SR ~ GT pairs).
template <typename T, bool kRegW>
__global__ void __launch_bounds__(DIST_THREADS) distance_split_kernel(const T* __restrict__ act, int n_pairs, int64_t elems,
                                                                      int C, const float* __restrict__ w,
                                                                      int64_t chunks_per_part, float* __restrict__ partial) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[DIST_THREADS / 32];
  const int pair = blockIdx.y, part = blockIdx.x;
  const T* a = act + (int64_t)pair * elems * 2;
  const T* b = act + (int64_t)(pair + n_pairs) * elems * 2;
  const int64_t total_chunks = elems / 8;
  const int64_t c_begin = (int64_t)part * chunks_per_part;
  int64_t c_end = c_begin + chunks_per_part;
  if (c_end > total_chunks) c_end = total_chunks;
  float wr[8];
  if constexpr (kRegW) {
    const int c0 = (int)(((c_begin + threadIdx.x) * 8) % C);
#pragma unroll
    for (int k = 0; k < 8; ++k) wr[k] = __ldg(w + c0 + k);
  }
  float acc = 0.f;
  auto accumulate = [&](int64_t i, const uint4& ah, const uint4& al, const uint4& bh, const uint4& bl) {
    float fah[8], fal[8], fbh[8], fbl[8];
    unpack8<T>(ah, fah); unpack8<T>(al, fal); unpack8<T>(bh, fbh); unpack8<T>(bl, fbl);
    if constexpr (!kRegW) {
      const int c0 = (int)((i * 8) % C);
#pragma unroll
      for (int k = 0; k < 8; ++k) wr[k] = __ldg(w + c0 + k);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) { const float d = (fah[k] - fbh[k]) + (fal[k] - fbl[k]); acc = fmaf(wr[k], d * d, acc); }
  };
  int64_t i = c_begin + threadIdx.x;
  for (; i + DIST_THREADS < c_end; i += 2 * DIST_THREADS) {
    uint4 q[2][4];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t j = i + u * DIST_THREADS;
      const int64_t off = (j >> 3) * 128 + (j & 7) * 8;
      q[u][0] = ld_nc_u4(a + off); q[u][1] = ld_nc_u4(a + off + 64);
      q[u][2] = ld_nc_u4(b + off); q[u][3] = ld_nc_u4(b + off + 64);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) accumulate(i + u * DIST_THREADS, q[u][0], q[u][1], q[u][2], q[u][3]);
  }
  for (; i < c_end; i += DIST_THREADS) {
    const int64_t off = (i >> 3) * 128 + (i & 7) * 8;
    accumulate(i, ld_nc_u4(a + off), ld_nc_u4(a + off + 64), ld_nc_u4(b + off), ld_nc_u4(b + off + 64));
  }
  const float t = block_reduce_fixed(acc, red);
  if (threadIdx.x == 0) partial[(int64_t)pair * SEMDIFF_MAX_PARTS + part] = t;
}

// Optional LPIPS-style variant (default OFF; the reference does not normalise): each
// pixel's channel vector is scaled to unit L2 norm (eps 1e-10) before the difference.  One warp per
// pixel, two passes over the pixel's channels (second pass hits L1).  grid = (n_parts, n_pairs).
template <typename T, bool kSplit>
__global__ void __launch_bounds__(DIST_THREADS) distance_norm_kernel(const T* __restrict__ act, int n_pairs, int hw,
                                                                     int C, const float* __restrict__ w,
                                                                     int pix_per_part, float* __restrict__ partial) {
  pdl_trigger();
  pdl_wait();  // inputs are the previous kernel's output
  __shared__ float red[DIST_THREADS / 32];
  const int pair = blockIdx.y, part = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int CS = kSplit ? 2 * C : C;   // stored elements per pixel
  const T* a = act + (int64_t)pair * hw * CS;
  const T* b = act + (int64_t)(pair + n_pairs) * hw * CS;
  int p_end = (part + 1) * pix_per_part;
  if (p_end > hw) p_end = hw;
  float acc = 0.f;
  for (int p = part * pix_per_part + warp; p < p_end; p += DIST_THREADS / 32) {
    const T* pa = a + (int64_t)p * CS;
    const T* pb = b + (int64_t)p * CS;
    float sa = 0.f, sb = 0.f;
    for (int c = lane * 8; c < C; c += 256) {
      float fa[8], fb[8];
      load8_px<T, kSplit>(pa, c, fa);
      load8_px<T, kSplit>(pb, c, fb);
#pragma unroll
      for (int k = 0; k < 8; ++k) { sa = fmaf(fa[k], fa[k], sa); sb = fmaf(fb[k], fb[k], sb); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, o); sb += __shfl_xor_sync(0xffffffffu, sb, o); }
    const float ia = 1.f / (sqrtf(sa) + 1e-10f), ib = 1.f / (sqrtf(sb) + 1e-10f);
    for (int c = lane * 8; c < C; c += 256) {
      float fa[8], fb[8];
      load8_px<T, kSplit>(pa, c, fa);
      load8_px<T, kSplit>(pb, c, fb);
#pragma unroll
      for (int k = 0; k < 8; ++k) { const float d = fa[k] * ia - fb[k] * ib; acc = fmaf(__ldg(w + c + k), d * d, acc); }
    }
  }
  const float t = block_reduce_fixed(acc, red);
  if (threadIdx.x == 0) partial[(int64_t)pair * SEMDIFF_MAX_PARTS + part] = t;
}

// Per-channel spatial mean of (a-b)^2: chan_mean[pair][c].  Feeds d(score)/d(w_layers) for callers that
// train the head (sweep script :55-69).  grid = (C/8/32 .. , n_pairs): a thread owns 8 channels and
// walks the pixels in order (fixed order -> deterministic).
template <typename T, bool kSplit>
__global__ void __launch_bounds__(128) chan_mean_kernel(const T* __restrict__ act, int n_pairs, int hw, int C,
                                                        float* __restrict__ chan_mean, int chan_stride) {
  pdl_trigger();
  pdl_wait();  // inputs are the previous kernel's output
  const int pair = blockIdx.y;
  const int c8 = blockIdx.x * blockDim.x + threadIdx.x;
  if (c8 * 8 >= C) return;
  const int CS = kSplit ? 2 * C : C;   // stored elements per pixel
  const T* a = act + (int64_t)pair * hw * CS;
  const T* b = act + (int64_t)(pair + n_pairs) * hw * CS;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int p = 0; p < hw; ++p) {
    float fa[8], fb[8];
    load8_px<T, kSplit>(a + (int64_t)p * CS, c8 * 8, fa);
    load8_px<T, kSplit>(b + (int64_t)p * CS, c8 * 8, fb);
#pragma unroll
    for (int k = 0; k < 8; ++k) { const float d = fa[k] - fb[k]; acc[k] = fmaf(d, d, acc[k]); }
  }
  const float inv = 1.f / (float)hw;
#pragma unroll
  for (int k = 0; k < 8; ++k) chan_mean[(int64_t)pair * chan_stride + c8 * 8 + k] = acc[k] * inv;
}

template <typename T, bool kSplit = false>
static int distance_t(const void* act, int n_pairs, int hw, int c, const float* w, int normalize, float* partial,
                      float* chan_mean, int chan_stride, cudaStream_t st) {
  const int parts = distance_parts(hw, c);
  dim3 grid(parts, n_pairs);
  if (normalize) {
    const int ppp = (hw + parts - 1) / parts;
    launch_pdl(distance_norm_kernel<T, kSplit>, dim3(grid), dim3(DIST_THREADS), 0, st, (const T*)act, n_pairs, hw, c, w, ppp, partial);
  } else if constexpr (kSplit) {
    const int64_t elems = (int64_t)hw * c;
    const int64_t cpp = (elems / 8 + parts - 1) / parts;
    if ((8 * DIST_THREADS) % c == 0)
      launch_pdl(distance_split_kernel<T, true>, dim3(grid), dim3(DIST_THREADS), 0, st, (const T*)act, n_pairs, elems, c, w, cpp, partial);
    else
      launch_pdl(distance_split_kernel<T, false>, dim3(grid), dim3(DIST_THREADS), 0, st, (const T*)act, n_pairs, elems, c, w, cpp, partial);
  } else {
    const int64_t elems = (int64_t)hw * c;
    const int64_t cpp = (elems / 8 + parts - 1) / parts;
    if ((8 * DIST_THREADS) % c == 0)
      launch_pdl(distance_kernel<T, true>, dim3(grid), dim3(DIST_THREADS), 0, st, (const T*)act, n_pairs, elems, c, w, cpp, partial);
    else
      launch_pdl(distance_kernel<T, false>, dim3(grid), dim3(DIST_THREADS), 0, st, (const T*)act, n_pairs, elems, c, w, cpp, partial);
  }
  SEMDIFF_CUDA_OK(cudaGetLastError());
  if (chan_mean != nullptr) {
    dim3 g2((c / 8 + 127) / 128, n_pairs);
    launch_pdl(chan_mean_kernel<T, kSplit>, dim3(g2), dim3(128), 0, st, (const T*)act, n_pairs, hw, c, chan_mean, chan_stride);
    SEMDIFF_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

int launch_distance(const void* act, int n_pairs, int hw, int c, const float* w, int normalize, float* partial,
                    float* chan_mean, int chan_stride, int precision, cudaStream_t st) {
  if (n_pairs <= 0 || hw <= 0 || c <= 0 || c % 8 != 0) { set_error("distance: need c %% 8 == 0, non-empty"); return SEMDIFF_ERR_ARG; }
  if (n_pairs > 65535) { set_error("distance: n_pairs > 65535 per launch"); return SEMDIFF_ERR_ARG; }
  if (is_split(precision) && c % 64 != 0) { set_error("distance: split precisions need c %% 64 == 0"); return SEMDIFF_ERR_ARG; }
  switch (precision) {
    case SEMDIFF_FP16X3: return distance_t<__half, true>(act, n_pairs, hw, c, w, normalize, partial, chan_mean, chan_stride, st);
    case SEMDIFF_BF16X3: return distance_t<__nv_bfloat16, true>(act, n_pairs, hw, c, w, normalize, partial, chan_mean, chan_stride, st);
    case SEMDIFF_BF16: return distance_t<__nv_bfloat16>(act, n_pairs, hw, c, w, normalize, partial, chan_mean, chan_stride, st);
    case SEMDIFF_FP16: return distance_t<__half>(act, n_pairs, hw, c, w, normalize, partial, chan_mean, chan_stride, st);
    case SEMDIFF_FP32: return distance_t<float>(act, n_pairs, hw, c, w, normalize, partial, chan_mean, chan_stride, st);
  }
  set_error("distance: bad precision %d", precision);
  return SEMDIFF_ERR_ARG;
}

// ---------------------------------------------------------------------------------------------
// head: mean over tapped layers + ReLU (:385-395).  One thread per pair, fixed summation order.
// ---------------------------------------------------------------------------------------------
constexpr int MAX_TAPS = 16;
struct HeadArgs { int n_parts[MAX_TAPS]; float inv_hw[MAX_TAPS]; };

__global__ void __launch_bounds__(128) head_kernel(const float* __restrict__ partials, int n_taps, int n_pairs,
                                                   HeadArgs args, const float* __restrict__ head_b,
                                                   float* __restrict__ out, float* __restrict__ pre) {
  pdl_trigger();
  pdl_wait();  // inputs are the previous kernel's output
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  float total = 0.f;
  for (int j = 0; j < n_taps; ++j) {
    const float* src = partials + ((int64_t)j * n_pairs + p) * SEMDIFF_MAX_PARTS;
    float s = 0.f;
    for (int q = 0; q < args.n_parts[j]; ++q) s += src[q];
    total += s * args.inv_hw[j] + __ldg(head_b + j);
  }
  const float v = n_taps == 1 ? total : total / (float)n_taps;
  if (pre != nullptr) pre[p] = v;
  out[p] = v > 0.f ? v : (v != v ? v : 0.f);   // torch.relu propagates NaN (fmaxf would turn it into a perfect score)
}

int launch_head(const float* partials, int n_taps, int n_pairs, const int* n_parts, const int* hw, const float* head_b,
                float* out_scores, float* out_pre_relu, cudaStream_t st) {
  if (n_taps < 1 || n_taps > MAX_TAPS || n_pairs <= 0) { set_error("head: n_taps must be in [1,%d]", MAX_TAPS); return SEMDIFF_ERR_ARG; }
  HeadArgs a;
  for (int j = 0; j < n_taps; ++j) { a.n_parts[j] = n_parts[j]; a.inv_hw[j] = 1.f / (float)hw[j]; }
  launch_pdl(head_kernel, dim3((n_pairs + 127) / 128), dim3(128), 0, st, partials, n_taps, n_pairs, a, head_b, out_scores, out_pre_relu);
  SEMDIFF_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace semdiff
