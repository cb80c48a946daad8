// Internal launcher declarations shared by the translation units of libsemdiff_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "semdiff_b200.h"

namespace semdiff {

// A conv may read a SECOND activation tensor through a fused 1x1 conv (stride2, no padding) whose products are
// accumulated into the same output: out = act(conv(in, W[:, :K1]) + conv1x1(in2, W[:, K1:]) + bias (+ res)).
// This is how a ResNet projection shortcut (downsample conv) is folded into the block's last conv.
struct ConvShape {
  int n_img, H, W, cin, cout, kh, kw, stride, pad, relu;
  int cin2 = 0, stride2 = 1, H2 = 0, W2 = 0;
  int pad_hi = -1;  // padding after the last row / column when it differs from `pad` (-1: symmetric)
  float wscale = 1.f;  // split precisions: the weights were multiplied by this power of two; the epilogue divides it out
  int pad_after() const { return pad_hi < 0 ? pad : pad_hi; }
  int OH() const { return (H + pad + pad_after() - kh) / stride + 1; }
  int OW() const { return (W + pad + pad_after() - kw) / stride + 1; }
  int64_t M() const { return (int64_t)n_img * OH() * OW(); }
  int K1() const { return kh * kw * cin; }
  int K() const { return kh * kw * cin + cin2; }
};
struct ConvPtrs {
  const void* in; const void* in2; const void* w; const float* bias; const void* res; void* out;
};

// launch with programmatic stream serialization (see common.cuh pdl_wait / pdl_trigger); only with SEMDIFF_PDL=1 (measured: no gain), else plain launches
bool pdl_enabled();
template <typename... P, typename... A>
inline cudaError_t launch_pdl(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<A&&>(args)...);
}

// split precisions (SEMDIFF_FP16X3 / BF16X3): a logical element is a (hi, lo) pair of 16-bit numbers
inline bool is_split(int precision) { return precision == SEMDIFF_FP16X3 || precision == SEMDIFF_BF16X3; }
inline int base_precision(int precision) {
  return precision == SEMDIFF_FP16X3 ? SEMDIFF_FP16 : (precision == SEMDIFF_BF16X3 ? SEMDIFF_BF16 : precision);
}
inline size_t elem_bytes(int precision) { return precision == SEMDIFF_FP32 || is_split(precision) ? 4 : 2; }

// each returns 0 / negative error code; all asynchronous on `stream`
// images [img0, img0 + n_imgs) of the stacked (GT..., SR...) batch -> out[0 .. n_imgs)
// gt / sr element type = in_precision (SEMDIFF_FP32 / BF16 / FP16), NCHW planes
int launch_pack(const void* gt, const void* sr, int in_precision, int n_pairs, int img0, int n_imgs, int H, int W,
                void* out, int precision, int layout, cudaStream_t stream);
int launch_maxpool3x3s2(const void* in, void* out, int n_img, int H, int W, int c, int precision, cudaStream_t stream);
int launch_avgpool(const void* in, void* out, int n_img, int H, int W, int c, int window, int precision,
                   cudaStream_t stream);
// local-map decoder helpers (elementwise.cu): what = 0 squared difference of the GT / SR halves (n_img = pairs), 1 channel
// concat (c | c2), 2 bilinear x2 upsampling (align_corners), 3 channel 0 -> bilinear x2 -> sigmoid -> fp32 map
int launch_decoder_op(int what, const void* in, const void* in2, void* out, int n_img, int H, int W, int c, int c2, int precision,
                      cudaStream_t stream);
int launch_conv_simt(const ConvPtrs& ptr, const ConvShape& s, int precision, cudaStream_t stream);
// tcgen05 path; use_tma lets the activation tile come through TMA (tiled for 1x1 stride 1, im2col mode otherwise);
// false forces the cp.async software-im2col gather
int launch_conv_tc(const ConvPtrs& ptr, const ConvShape& s, int precision, bool use_tma, cudaStream_t stream);
bool conv_tc_supported(const ConvShape& s, int precision, bool use_tma);
// prepared launch (TMA descriptors encoded once, reused while pointers and shapes stay the same)
struct alignas(64) ConvTcLaunch {
  unsigned char params[1280];
  int block_n, a_mode, precision;
};
int conv_tc_prepare(ConvTcLaunch* L, const ConvPtrs& ptr, const ConvShape& s, int precision, bool use_tma);
int conv_tc_launch(const ConvTcLaunch* L, cudaStream_t stream);
// strip variant for 3x3 stride-1 pad-1 64 -> 64 convs (conv3x3_strip.cu); conv_tc_launch dispatches to it
bool conv_strip_supported(const ConvShape& s, int precision);
int conv_strip_prepare(ConvTcLaunch* L, const ConvPtrs& ptr, const ConvShape& s, int precision);
int conv_strip_launch(const ConvTcLaunch* L, cudaStream_t stream);
// the space-to-depth stem conv with the following 3x3 stride-2 pad-1 max pool in its epilogue; ptr.out = POOLED output
bool conv_strip_pool_supported(const ConvShape& s, int precision);
int conv_strip_pool_prepare(ConvTcLaunch* L, const ConvPtrs& ptr, const ConvShape& s, int precision);
// a 3x3 pad-1 64 -> 64 conv with the following 2x2 average pool in its epilogue (CLIP stem); ptr.out = POOLED output
bool conv_strip_avgpool_supported(const ConvShape& s, int precision);
int conv_strip_avgpool_prepare(ConvTcLaunch* L, const ConvPtrs& ptr, const ConvShape& s, int precision);

// conv1x1 (-> 256 channels, residual or fused shortcut) chained with the next conv1x1 (256 -> 64 | 128): the 256-channel
// tile is written out AND consumed from shared memory by the second GEMM (conv_chain.cu); conv_tc_launch dispatches to it
bool conv_chain_supported(const ConvShape& s1, const ConvShape& s2, bool has_res, int precision);
int conv_chain_prepare(ConvTcLaunch* L, const ConvPtrs& ptr1, const ConvShape& s1, const ConvPtrs& ptr2, const ConvShape& s2,
                       int precision);
int conv_chain_launch(const ConvTcLaunch* L, cudaStream_t stream);

int distance_parts(int hw, int c);
int launch_distance(const void* act, int n_pairs, int hw, int c, const float* w, int normalize, float* partial,
                    float* chan_mean, int chan_stride, int precision, cudaStream_t stream);
int launch_head(const float* partials, int n_taps, int n_pairs, const int* n_parts, const int* hw, const float* head_b,
                float* out_scores, float* out_pre_relu, cudaStream_t stream);

}  // namespace semdiff
