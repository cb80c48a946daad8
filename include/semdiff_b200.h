/*
 * semdiff_b200 -- C-ABI of the B200-native global semantic-fidelity scorer ("CLIP-LPIPS regressor").
 *
 * The reference has no FFI: its scorer is a torch.nn.Module
 * (/root/reference/models/global_eval_models.py:308-429 CLIP_lpips_stages_cnn,
 *  :682-812 CLIP_lpips_stages_cnn_clsbckb).  This header is the boundary a maintainer binds to
 * replace that module's forward(): plain pointers and sizes, no torch types.  Every pointer named
 * "device" is CUDA device memory owned by the caller (PyTorch's caching allocator in the shipped
 * host code); the library allocates nothing on the device.  Every entry point returns 0 on success
 * and a negative code on failure, with a message available from semdiff_last_error().
 * Launches are asynchronous on the given stream.  A plan is not thread-safe.
 *
 * Layouts: images in  = fp32 NCHW [n,3,H,W]           (what forward(a, b) receives, :341 / :717)
 *          activations = NHWC, element type per `precision`, GT images first then SR images
 *          conv weights = [Cout][KH][KW][Cin] (K-major rows), BatchNorm already folded
 *          scores out  = fp32 [n]                      (what forward returns, :397 / :773)
 */
#ifndef SEMDIFF_B200_H_
#define SEMDIFF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* semdiff_stream_t; /* == cudaStream_t */

enum { SEMDIFF_OK = 0, SEMDIFF_ERR_ARG = -1, SEMDIFF_ERR_CUDA = -2, SEMDIFF_ERR_UNSUPPORTED = -3 };

/* storage/compute type of the trunk.  Accumulation is always fp32.
 * The two "x3" types are SPLIT types: every activation and weight is the unevaluated sum hi + lo of two 16-bit numbers
 * (hi = round(x), lo = round(x - hi): 22 significant bits for fp16, 16 for bf16), and every conv runs as three tensor-core
 * products per K block, Al*Wh + Ah*Wl + Ah*Wh (small terms first; the Al*Wl term is below fp32 resolution), whose chunk sums
 * are added in registers with round-to-nearest because the tensor core truncates its fp32 accumulator (csrc/conv_tc_split.cu).
 * Split activations are stored NHWC with 2*C 16-bit channels: for each block of 64 logical channels, 64 hi values then
 * 64 lo values.  Split conv weights are [Cout][2*K] with the same interleave along K, pre-multiplied by the power of two
 * semdiff_op.wscale (keeps the lo halves in the normal fp16 range); the epilogue multiplies the accumulator by 1/wscale. */
enum { SEMDIFF_BF16 = 0, SEMDIFF_FP16 = 1, SEMDIFF_FP32 = 2, SEMDIFF_FP16X3 = 3, SEMDIFF_BF16X3 = 4 };

/* conv implementation selector for semdiff_conv2d (the plan picks AUTO) */
enum {
  SEMDIFF_CONV_AUTO = 0,
  SEMDIFF_CONV_SIMT = 1,      /* CUDA-core implicit GEMM, any precision (the only fp32 path) */
  SEMDIFF_CONV_TC_GATHER = 2, /* tcgen05 + cp.async software im2col (any k/stride/pad, Cin % 8 == 0) */
  SEMDIFF_CONV_TC_TMA = 3     /* tcgen05 + TMA activations: tiled loads for 1x1 stride 1, im2col mode otherwise (Cin % 64 == 0) */
};

/* layout of buffer 0 (what the pack kernel makes of the fp32 NCHW images) */
enum {
  SEMDIFF_INPUT_NHWC8 = 0,    /* [2n, H, W, 8], channels 3..7 zero */
  SEMDIFF_INPUT_S2D_ROW4 = 1, /* [2n, H/2+3, W/2, 64]: 2x2 space-to-depth row windows for 7x7/2 pad-3 stems (elementwise.cu) */
  SEMDIFF_INPUT_S2D_ROW2 = 2, /* [2n, H/2+1, W/2, 64]: the same for 3x3/2 pad-1 stems (window of 2, upper 32 channels zero) */
  SEMDIFF_INPUT_S2D16 = 3     /* [2n, H/2, W/2, 16]: plain 2x2 space-to-depth, channel (dy*2+dx)*3+ci, 12..15 zero; a 7x7/2
                                 pad-3 stem is a 4x4 stride-1 conv over it with padding 2 before / 1 after */
};

/* One step of the trunk program.  Buffers are logical ids in [0, n_bufs); buffer 0 is the packed
 * input image batch.  The program is what /root/reference/models/global_eval_models.py:364,371
 * (self.clip(a), self.clip(b)) executes inside timm, with BatchNorm folded. */
enum {
  SEMDIFF_OP_CONV = 0, SEMDIFF_OP_MAXPOOL3S2 = 1, SEMDIFF_OP_AVGPOOL = 2, SEMDIFF_OP_TAP = 3,
  /* local-map decoder (/root/reference/models/local_eval_models.py:109-125): buffers written by SQDIFF hold ONE image per
   * pair (everything downstream of it too); all other buffers hold two (GT image, SR image) */
  SEMDIFF_OP_SQDIFF = 4,     /* dst[pair] = (src[GT image] - src[SR image])^2                      (:115) */
  SEMDIFF_OP_CONCAT = 5,     /* dst = channel concat (src | src2); channel counts come from the buffers (:121) */
  SEMDIFF_OP_UPSAMPLE2X = 6, /* nn.UpsamplingBilinear2d(scale_factor=2): bilinear, align_corners   (:84, :119, :123) */
  SEMDIFF_OP_MAP_OUT = 7     /* channel 0 of src -> bilinear x2 -> sigmoid -> the fp32 output map   (:123-125) */
};

typedef struct semdiff_op {
  int32_t kind;      /* SEMDIFF_OP_* */
  int32_t src;       /* input buffer id */
  int32_t dst;       /* output buffer id (unused for TAP) */
  int32_t res;       /* residual buffer id added before the activation, or -1 */
  int32_t cin;       /* input channels as stored (multiple of 8; the image is padded 3 -> 8) */
  int32_t cout;      /* output channels (conv) */
  int32_t kh, kw;    /* kernel size (conv); pooling window for AVGPOOL (kh == kw == stride) */
  int32_t stride;
  int32_t pad;
  int32_t relu;      /* apply ReLU in the epilogue */
  int32_t tap;       /* for TAP: index j of w_layers[j] (:336) this activation feeds */
  int32_t src2;      /* CONV: second input buffer read through a fused 1x1 conv (projection shortcut), or -1 */
  int32_t cin2;      /* channels of src2; weight rows are [kh*kw*cin | cin2] */
  int32_t stride2;   /* spatial stride of the fused 1x1 conv over src2 */
  int32_t pad_hi;    /* CONV: padding after the last row/column if it differs from `pad`, else -1 (symmetric) */
  const void* weight; /* device, [cout][kh][kw][cin] in the plan's precision (split types: [cout][2*K], see above) */
  const float* bias;  /* device, [cout] fp32 (folded BN shift) */
  float wscale;       /* split types: the power of two the weights were multiplied by (0 is read as 1) */
  int32_t reserved;
} semdiff_op;

typedef struct semdiff_plan semdiff_plan;

/* Build an execution plan for a trunk program.  Copies the op list (not the weights).
 * head_ops: the first `head_ops` ops (stem convs + first pooling op, no residuals/taps) may be run in chunks of images
 * sized so that their activations stay L2-resident between kernels (0 = never chunk; results are identical). */
int semdiff_plan_create(const semdiff_op* ops, int32_t n_ops, int32_t n_bufs, int32_t precision,
                        int32_t input_layout, int32_t head_ops, semdiff_plan** out_plan);
int semdiff_plan_destroy(semdiff_plan* plan);
/* Force the conv implementation of every conv op (SEMDIFF_CONV_*; AUTO = best supported). Testing aid.
 * With AUTO the plan also fuses adjacent ops into single launches where a fused kernel exists (stem conv + pooling
 * op, conv3 of a bottleneck + conv1 of the next: semdiff_conv2d_maxpool / _avgpool / semdiff_conv1x1_chain below); the
 * results are bit-identical to one launch per op, which any explicit choice here restores. */
int semdiff_plan_set_conv_impl(semdiff_plan* plan, int32_t impl);

/* Bytes of device workspace semdiff_score needs for `microbatch_pairs` pairs of HxW images. */
int64_t semdiff_workspace_bytes(const semdiff_plan* plan, int32_t microbatch_pairs, int32_t H, int32_t W);

/* forward(a, b) -> score, :341-397 / :717-773.
 *   gt, sr        device NCHW [n_pairs,3,H,W]; element type in_precision = SEMDIFF_FP32 (what the reference passes),
 *                 or SEMDIFF_BF16 / SEMDIFF_FP16 when the caller already holds 16-bit images (halves the PCIe bytes)
 *   head_w        device fp32, w_layers[j].weight concatenated in tap order (sum_j C_j floats)
 *   head_b        device fp32 [n_taps], w_layers[j].bias
 *   out_scores    device fp32 [n_pairs]   = relu(mean_j(b_j + mean_hw sum_c w_j[c] (A-B)^2))
 *   out_pre_relu  device fp32 [n_pairs] or NULL (same value before the final ReLU)
 *   out_chan_mean device fp32 [n_pairs, sum_j C_j] or NULL: per-channel spatial means of (A-B)^2
 *                 (d score / d w_j[c] * n_taps; lets the caller train w_layers, :55-69 of the sweep script)
 * Pairs are processed in passes of `microbatch_pairs` pairs (bounds the workspace; a pair's score does not depend on it). */
int semdiff_score(semdiff_plan* plan, const void* gt, const void* sr, int32_t in_precision, int32_t n_pairs, int32_t H,
                  int32_t W, int32_t microbatch_pairs, const float* head_w, const float* head_b, int32_t normalize,
                  void* workspace, int64_t workspace_bytes, float* out_scores, float* out_pre_relu,
                  float* out_chan_mean, semdiff_stream_t stream);

/* Local maps (CLIP_lpips_Unet, /root/reference/models/local_eval_models.py:86-126): runs a program that ends in
 * SEMDIFF_OP_MAP_OUT.  out_map: device fp32 [n_pairs, 1, Hm, Wm], Hm x Wm = twice the size of MAP_OUT's source (= H x W for
 * the reference's decoders).  Other arguments as semdiff_score. */
int semdiff_score_map(semdiff_plan* plan, const void* gt, const void* sr, int32_t in_precision, int32_t n_pairs, int32_t H,
                      int32_t W, int32_t microbatch_pairs, void* workspace, int64_t workspace_bytes, float* out_map,
                      semdiff_stream_t stream);

/* Profiling: when enabled, semdiff_score brackets every op with CUDA events on `stream`. */
int semdiff_plan_set_profiling(semdiff_plan* plan, int32_t enable);
/* Per-op accumulated milliseconds and launch counts since the last reset (arrays of n_ops + 3:
 * the three extra slots are pack, distance (all taps), head). Synchronises the recorded events. */
int semdiff_plan_get_profile(semdiff_plan* plan, float* out_ms, int32_t* out_launches, int32_t n, int32_t reset);
/* Number of kernels launched by the last semdiff_score call on this plan. */
int64_t semdiff_plan_last_launches(const semdiff_plan* plan);

/* ---- single kernels (unit tests call these; the plan calls the same launchers) ------------- */

/* NCHW [n,3,H,W] x2 (element type in_precision) -> buffer 0 in `layout` (SEMDIFF_INPUT_*), GT images first */
int semdiff_pack_input(const void* gt, const void* sr, int32_t in_precision, int32_t n_pairs, int32_t H, int32_t W,
                       void* out, int32_t precision, int32_t layout, semdiff_stream_t stream);

/* out = act(conv(in, weight[:, :kh*kw*cin]) (+ conv1x1_stride2(in2, weight[:, kh*kw*cin:])) + bias (+ residual));
 * NHWC; in2 may be NULL; pad_hi = padding after the last row/column (-1: same as pad); impl = SEMDIFF_CONV_* */
int semdiff_conv2d(const void* in, const void* weight, const float* bias, const void* residual, void* out,
                   int32_t n_img, int32_t H, int32_t W, int32_t cin, int32_t cout, int32_t kh, int32_t kw,
                   int32_t stride, int32_t pad, int32_t relu, const void* in2, int32_t H2, int32_t W2, int32_t cin2,
                   int32_t stride2, int32_t pad_hi, int32_t precision, int32_t impl, semdiff_stream_t stream);

/* Stem conv over the SEMDIFF_INPUT_S2D16 layout (cin = 16, cout = 64, stride 1: 4x4 pad 2|1 or 2x2 pad 1|0) with the
 * following max_pool2d(kernel 3, stride 2, padding 1) computed in the conv epilogue (timm resnet50 conv1+bn1+act1+maxpool):
 * out = NHWC [n_img, (OH-1)/2+1, (OW-1)/2+1, 64]; the un-pooled conv output is never written.  16-bit precisions, output
 * width OW <= 128 - (kw - 1).  Same values as semdiff_conv2d followed by semdiff_maxpool3x3s2.  The plan applies this fusion by itself. */
int semdiff_conv2d_maxpool(const void* in, const void* weight, const float* bias, void* out, int32_t n_img, int32_t H,
                           int32_t W, int32_t cin, int32_t cout, int32_t kh, int32_t kw, int32_t pad, int32_t pad_hi,
                           int32_t relu, int32_t precision, semdiff_stream_t stream);

/* 3x3 stride-1 pad-1 conv, cin (32 | 64) -> 64 channels, with the following avg_pool2d(2) computed in the conv epilogue (timm
 * resnet50_clip stem.conv3 + stem.pool): out = NHWC [n_img, H/2, W/2, 64]; the un-pooled conv output is never written.
 * 16-bit precisions, H even, 62 <= W <= 126.  Same values as semdiff_conv2d followed by semdiff_avgpool(window 2).
 * The plan applies this fusion by itself. */
int semdiff_conv2d_avgpool(const void* in, const void* weight, const float* bias, void* out, int32_t n_img, int32_t H,
                           int32_t W, int32_t cin, int32_t relu, int32_t precision, semdiff_stream_t stream);

/* Two chained pointwise convs at a bottleneck boundary, one launch (16-bit precisions only):
 *   out1[m, cout1] = act1(in[m, cin] * w1[:, :cin]^T (+ in2[m, cin2] * w1[:, cin:]^T) + bias1 (+ residual[m, cout1]))
 *   out2[m, cout2] = act2(out1 * w2^T + bias2)
 * cout1 = 256: cout2 = 64 | 128, cin + cin2 <= 128 (multiples of 64), in2 and residual exclusive (either may be NULL);
 * cout1 = 512: cin = 128, cout2 = 128, residual required, no in2 (identity blocks of the 512-channel stage).
 * out1 is written in full and consumed by the second conv from shared memory: same values as two semdiff_conv2d
 * calls (timm Bottleneck conv3+bn3+add+act3 followed by the next block's conv1+bn1+act1), one HBM read less.
 * The plan applies this fusion by itself. */
int semdiff_conv1x1_chain(const void* in, const void* in2, const void* w1, const float* bias1, const void* residual,
                          void* out1, const void* w2, const float* bias2, void* out2, int64_t m, int32_t cin,
                          int32_t cin2, int32_t cout1, int32_t cout2, int32_t relu1, int32_t relu2, int32_t precision,
                          semdiff_stream_t stream);

int semdiff_maxpool3x3s2(const void* in, void* out, int32_t n_img, int32_t H, int32_t W, int32_t c,
                         int32_t precision, semdiff_stream_t stream);
int semdiff_avgpool(const void* in, void* out, int32_t n_img, int32_t H, int32_t W, int32_t c, int32_t window,
                    int32_t precision, semdiff_stream_t stream);

/* One decoder helper (unit tests; the plan calls the same launcher): what = 0 SQDIFF (n_img = pairs; in holds 2 * n_img
 * images), 1 CONCAT (c | c2), 2 UPSAMPLE2X, 3 MAP_OUT (out = fp32 [n_img, 1, 2H, 2W]).  NHWC, element type per precision. */
int semdiff_decoder_op(int32_t what, const void* in, const void* in2, void* out, int32_t n_img, int32_t H, int32_t W,
                       int32_t c, int32_t c2, int32_t precision, semdiff_stream_t stream);

/* Fused per-layer distance (:379-384 / :755-760): act = NHWC [2*n_pairs, HW, C] with GT image i at
 * index i and SR image i at index n_pairs + i.  Writes partial[pair * SEMDIFF_MAX_PARTS + part] =
 * sum over the part's elements of w[c] * (a-b)^2 (fixed order, no atomics); n_parts via
 * semdiff_distance_parts().  chan_mean (nullable): [n_pairs, chan_stride] slice for this layer. */
#define SEMDIFF_MAX_PARTS 64
int32_t semdiff_distance_parts(int32_t hw, int32_t c);
int semdiff_layer_distance(const void* act, int32_t n_pairs, int32_t hw, int32_t c, const float* w,
                           int32_t normalize, float* partial, float* chan_mean, int32_t chan_stride,
                           int32_t precision, semdiff_stream_t stream);

/* Head (:385-395 / :761-771): score[p] = relu(mean_j(bias[j] + sum_parts partial_j[p] / hw_j)).
 * partials: device [n_taps][n_pairs][SEMDIFF_MAX_PARTS]; n_parts/hw: host arrays [n_taps]. */
int semdiff_head(const float* partials, int32_t n_taps, int32_t n_pairs, const int32_t* n_parts,
                 const int32_t* hw, const float* head_b, float* out_scores, float* out_pre_relu,
                 semdiff_stream_t stream);

/* ---- on-device preprocessing: the reference's `model.processor` (timm eval transform on PIL images, :333-334) ----
 * Pillow's 8-bit bicubic resize (separable, antialiased, Q22 fixed point, uint8 after each pass) -> center crop ->
 * /255 -> (x - mean) / std -> NCHW, bit-exact against Pillow + torchvision.
 * semdiff_resize_ksize / semdiff_resize_coeffs are HOST functions: bounds [out_size][2] = (first tap, taps),
 * coeffs [out_size][ksize] Q22.  The caller uploads the tables and passes device pointers to semdiff_preprocess_u8:
 *   src   device uint8 [n, Hs, Ws, 3] (decoded RGB, HWC);  the image is (virtually) resized to Hr x Wr and the window
 *         [top, top+crop_h) x [left, left+crop_w) is produced;  tmp: device uint8 [n, Hs, crop_w, 3] scratch
 *   out   device [n, 3, crop_h, crop_w] in out_precision (SEMDIFF_FP32 = what the reference's processor returns)
 *   mean, stdv: HOST float[3] */
int32_t semdiff_resize_ksize(int32_t in_size, int32_t out_size);
int semdiff_resize_coeffs(int32_t in_size, int32_t out_size, int32_t* bounds, int32_t* coeffs);
int semdiff_preprocess_u8(const uint8_t* src, int32_t n, int32_t Hs, int32_t Ws, int32_t Hr, int32_t Wr, int32_t top,
                          int32_t left, int32_t crop_h, int32_t crop_w, const int32_t* bounds_x, const int32_t* coeffs_x,
                          int32_t ksize_x, const int32_t* bounds_y, const int32_t* coeffs_y, int32_t ksize_y,
                          const float* mean, const float* stdv, uint8_t* tmp, void* out, int32_t out_precision,
                          semdiff_stream_t stream);

const char* semdiff_last_error(void);
/* "semdiff_b200 <version> sm_100a" */
const char* semdiff_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SEMDIFF_B200_H_ */
