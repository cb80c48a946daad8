// Shared between conv_tc.cu (16-bit tensor-core conv + host-side descriptor building) and conv_tc_split.cu (the
// split-precision kernel): tile constants, the kernel parameter block, swizzle helper.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"

namespace semdiff {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 x 16-bit = one 128-byte swizzle row
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
enum { A_TMA = 0, A_GATHER = 1, A_IM2COL = 2 };

struct alignas(64) ConvTcParams {
  CUtensorMap tmA;  // activations: [M, Cin] tiled (A_TMA) or NHWC im2col (A_IM2COL)
  CUtensorMap tmA2; // fused second source (1x1 conv): [M, Cin2] tiled when stride2 == 1, else NHWC im2col
  CUtensorMap tmB;  // weights as [Cout, K]
  CUtensorMap tmC;  // output as [M, Cout]
  CUtensorMap tmR;  // residual as [M, Cout]
  const void* in;
  const float* bias;
  int H, W, Cin, OH, OW, Cout, KH, KW, stride, pad, relu, has_res;
  int M, num_kb, m_tiles, n_tiles, cpt, taps;
  int num_kb1, a2_im2col, stride2;  // k-blocks [num_kb1, num_kb) come from the second source
  int has_src2;
  // split precisions (conv_tc_split.cu): stored tensors carry 2 x 64 columns (hi, lo) per k-block / 64 output channels
  int chunk_kb;                     // k-blocks accumulated in TMEM before the sum is promoted to registers
  int stages, ring;                 // operand-ring fills / C-ring groups (split_ring_config)
  float acc_scale;                  // 1 / wscale, applied to the accumulator before the bias
};
static_assert(sizeof(ConvTcParams) <= 896, "ConvTcLaunch::params too small");

// 16-byte chunk position inside a swizzled row of ROW_BYTES (128 B -> SWIZZLE_128B, 64 B -> SWIZZLE_64B)
template <int ROW_BYTES> __device__ __forceinline__ uint32_t swz_chunk(uint32_t chunk, uint32_t row) {
  if constexpr (ROW_BYTES == 128) return chunk ^ (row & 7);
  else return chunk ^ ((row >> 1) & 3);
}

// per-device caches (a process may drive several GPUs; function attributes and SM counts are per device)
constexpr int MAX_DEVICES = 64;
int current_device();
int num_sms();

// split-precision launch (conv_tc_split.cu); p prepared by conv_tc_prepare
int launch_conv_split(const ConvTcParams& p, int block_n, int a_mode, int precision, cudaStream_t st);
void split_ring_config(int block_n, bool has_res, int num_kb, int* stages, int* ring);

}  // namespace semdiff
