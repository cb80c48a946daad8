// CUDA-core implicit-GEMM convolution: the fp32 parity path (SEMDIFF_FP32) and the on-device
// cross-check for the tcgen05 kernels (same NHWC layout, same [Cout][KH][KW][Cin] weights, same fused
// bias + residual + ReLU epilogue).  64x64 output tile per CTA, 4x4 per thread, K staged 16 at a time.
#include "common.cuh"
#include "kernels.h"

namespace semdiff {

constexpr int TM = 64, TN = 64, TK = 16;

template <typename T>
__global__ void __launch_bounds__(256) conv_simt_kernel(const T* __restrict__ in, const T* __restrict__ in2,
                                                        const T* __restrict__ wgt, const float* __restrict__ bias,
                                                        const T* __restrict__ res, T* __restrict__ out, ConvShape s,
                                                        int OH, int OW, int M, int K, int K1) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  // loader mapping: 64 rows x 4 k-quads
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  const int gm = m0 + lrow;
  int n = 0, oh = 0, ow = 0;
  const bool row_ok = gm < M;
  if (row_ok) { n = gm / (OH * OW); int r = gm - n * OH * OW; oh = r / OW; ow = r - oh * OW; }
  const int gn = n0 + lrow;  // weight row handled by this thread in the B loader
  const bool col_ok = gn < s.cout;

  const int tx = tid & 15, ty = tid >> 4;  // compute mapping: rows ty*4.., cols tx*4..
  float acc[4][4] = {};

  for (int k0 = 0; k0 < K; k0 += TK) {
    // ---- A tile: 4 consecutive k of one row (same tap: cin % 4 == 0) ----
    {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      const int k = k0 + lk;
      if (row_ok && k >= K1 && k < K) {
        // fused second source: 1x1 conv, stride2, no padding
        const T* p = in2 + (((int64_t)n * s.H2 + oh * s.stride2) * s.W2 + ow * s.stride2) * s.cin2 + (k - K1);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = Elem<T>::to_f(p[j]);
      } else if (row_ok && k < K1) {
        const int tap = k / s.cin, c = k - tap * s.cin;
        const int r = tap / s.kw, q = tap - r * s.kw;
        const int ih = oh * s.stride - s.pad + r, iw = ow * s.stride - s.pad + q;
        if (ih >= 0 && ih < s.H && iw >= 0 && iw < s.W) {
          const T* p = in + (((int64_t)n * s.H + ih) * s.W + iw) * s.cin + c;
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = Elem<T>::to_f(p[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) As[lk + j][lrow] = v[j];
    }
    // ---- B tile ----
    {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      const int k = k0 + lk;
      if (col_ok && k < K) {
        const T* p = wgt + (int64_t)gn * K + k;
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = Elem<T>::to_f(p[j]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) Bs[lk + j][lrow] = v[j];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  // ---- epilogue: + bias (+ residual) -> ReLU -> store ----
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + tx * 4 + j;
      if (c >= s.cout) continue;
      float v = acc[i][j] + __ldg(bias + c);
      if (res != nullptr) v += Elem<T>::to_f(res[(int64_t)m * s.cout + c]);
      if (s.relu) v = fmaxf(v, 0.f);
      out[(int64_t)m * s.cout + c] = Elem<T>::from_f(v);
    }
  }
}

template <typename T>
static int conv_simt_t(const ConvPtrs& q, const ConvShape& s, cudaStream_t st) {
  const int OH = s.OH(), OW = s.OW();
  const int64_t M = s.M();
  if (M > 0x7fffffffLL) { set_error("conv_simt: M too large"); return SEMDIFF_ERR_ARG; }
  dim3 grid((unsigned)((M + TM - 1) / TM), (unsigned)((s.cout + TN - 1) / TN));
  conv_simt_kernel<T><<<grid, 256, 0, st>>>((const T*)q.in, (const T*)q.in2, (const T*)q.w, q.bias, (const T*)q.res,
                                            (T*)q.out, s, OH, OW, (int)M, s.K(), s.K1());
  SEMDIFF_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_conv_simt(const ConvPtrs& q, const ConvShape& s, int precision, cudaStream_t st) {
  if (s.cin % 4 != 0 || s.cin2 % 4 != 0 || s.n_img <= 0 || s.OH() <= 0 || s.OW() <= 0) {
    set_error("conv_simt: need cin %% 4 == 0 and a non-empty output (cin=%d)", s.cin);
    return SEMDIFF_ERR_ARG;
  }
  switch (precision) {
    case SEMDIFF_BF16: return conv_simt_t<__nv_bfloat16>(q, s, st);
    case SEMDIFF_FP16: return conv_simt_t<__half>(q, s, st);
    case SEMDIFF_FP32: return conv_simt_t<float>(q, s, st);
  }
  set_error("conv_simt: bad precision %d", precision);
  return SEMDIFF_ERR_ARG;
}

}  // namespace semdiff
