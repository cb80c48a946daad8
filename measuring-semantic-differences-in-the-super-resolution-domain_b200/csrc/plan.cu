// Host side of libsemdiff_b200.so: the trunk "program" executor (plan) and the C-ABI (include/semdiff_b200.h).
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <tuple>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace semdiff {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// Programmatic dependent launch (every kernel starts with pdl_trigger / guards its first dependent access with pdl_wait):
// the next kernel's prologue - barrier init, TMEM allocation, descriptor prefetch, resident weights - overlaps the tail of the
// previous one.  Measured on B200: at 256 pairs per pass it changes nothing (-1 %: the persistent kernels all end within a
// tile of each other, profiles/r1_pdl.txt), but small passes are chains of ~10 us kernels whose prologues are a third of
// their run time: 5 pairs 0.617 -> 0.516 ms (bf16), 1.10 -> 1.00 ms (fp16x3); 32 pairs 1.12 -> 1.02 ms (profiles/r2_pdl_small_batch.txt).
// So the executor turns it on per pass when the pass is small (<= 8 M input pixels ~ 80 pairs of 224x224; split precisions 2 M).
// SEMDIFF_PDL=1 / =0 forces it on / off.
static thread_local bool g_pdl_auto = false;
bool pdl_enabled() {
  static const int forced = getenv("SEMDIFF_PDL") == nullptr ? -1 : (atoi(getenv("SEMDIFF_PDL")) != 0 || getenv("SEMDIFF_PDL")[0] == '\0' ? 1 : 0);
  return forced >= 0 ? forced != 0 : g_pdl_auto;
}

struct BufShape { int h = 0, w = 0, c = 0, pp = 2; };   // pp: images per pair held by the buffer (1 downstream of SQDIFF)
struct Prepared;

// everything that depends on (workspace address, images in the micro-batch, H, W)
struct ShapePlan {
  std::vector<BufShape> op_src, op_src2, op_dst;  // per op
  std::vector<int64_t> buf_offset;             // per buffer, bytes from the workspace base
  std::vector<ConvTcLaunch> tc;                // per op (valid where impl is a tcgen05 one)
  std::vector<ConvTcLaunch> tc_tail;           // head ops only: launch for the ragged last chunk
  std::vector<int> impl;                       // per op: SEMDIFF_CONV_*
  std::vector<char> fused_away;                // per op: 1 = computed by the chained launch of an earlier op (conv_chain.cu)
  int chunk_imgs = 0;                          // images per head chunk (0 = head ops run on the whole micro-batch)
  int map_h = 0, map_w = 0;                    // size of the output map (programs ending in SEMDIFF_OP_MAP_OUT)
  int64_t partial_offset = 0;
  int64_t total_bytes = 0;
  bool prepared = false;
  void* prepared_ws = nullptr;
};

}  // namespace semdiff

using namespace semdiff;

struct semdiff_plan {
  std::vector<semdiff_op> ops;
  int n_bufs = 0, precision = 0, n_taps = 0, conv_impl = SEMDIFF_CONV_AUTO, input_layout = SEMDIFF_INPUT_NHWC8;
  bool has_map = false;   // the program ends in SEMDIFF_OP_MAP_OUT (local-map U-Net) instead of feeding TAP ops to the head
  // The first `head_ops` ops (stem + first pool) stream the widest activations of the trunk.  They run in chunks of
  // images small enough that pack output and stem output stay L2-resident between the three kernels (the buffers are
  // reused by every chunk, so the lines are overwritten in cache instead of travelling to HBM and back).
  int head_ops = 0;
  int64_t head_l2_bytes = 0;  // measured on B200 (profiles/r1_head_chunking.txt): chunking loses 3-6 %, so it is off unless SEMDIFF_HEAD_L2_MB is set
  std::vector<int> tap_c, tap_off;  // channels and head_w offset per tap
  int chan_total = 0;
  std::map<std::tuple<int, int, int>, ShapePlan> shapes;  // key: (pairs in micro-batch, H, W)
  bool profiling = false;
  struct Ev { int slot; cudaEvent_t a, b; };
  std::vector<Ev> events;
  std::vector<double> prof_ms;
  std::vector<int> prof_launches;
  int64_t last_launches = 0;
};

namespace semdiff {

static int infer_shapes(const semdiff_plan* P, int pairs, int H, int W, ShapePlan* S) {
  const int n_ops = (int)P->ops.size();
  std::vector<BufShape> cur(P->n_bufs);
  std::vector<int64_t> buf_elems(P->n_bufs, 0);
  const int64_t n_img = 2 * (int64_t)pairs;
  if (P->input_layout == SEMDIFF_INPUT_S2D16) {
    if ((H | W) & 1) { set_error("the s2d stem layouts need even H and W (got %dx%d)", H, W); return SEMDIFF_ERR_ARG; }
    cur[0] = BufShape{H / 2, W / 2, 16};
  } else if (P->input_layout != SEMDIFF_INPUT_NHWC8) {
    if ((H | W) & 1) { set_error("the s2d stem layouts need even H and W (got %dx%d)", H, W); return SEMDIFF_ERR_ARG; }
    cur[0] = BufShape{H / 2 + (P->input_layout == SEMDIFF_INPUT_S2D_ROW4 ? 3 : 1), W / 2, 64};
  } else {
    cur[0] = BufShape{H, W, 8};
  }
  S->chunk_imgs = 0;
  if (P->head_ops > 0) {
    // per-image bytes of the buffers that live only inside a head chunk: the packed input + every head output but the last
    int64_t per_img = (int64_t)cur[0].h * cur[0].w * cur[0].c;
    std::vector<BufShape> t(P->n_bufs);
    t[0] = cur[0];
    for (int i = 0; i + 1 < P->head_ops; ++i) {
      const semdiff_op& op = P->ops[i];
      const BufShape in = t[op.src];
      BufShape o = in;
      if (op.kind == SEMDIFF_OP_CONV) { const int ph = op.pad + (op.pad_hi < 0 ? op.pad : op.pad_hi); o = BufShape{(in.h + ph - op.kh) / op.stride + 1, (in.w + ph - op.kw) / op.stride + 1, op.cout}; }
      else if (op.kind == SEMDIFF_OP_MAXPOOL3S2) o = BufShape{(in.h - 1) / 2 + 1, (in.w - 1) / 2 + 1, in.c};
      else if (op.kind == SEMDIFF_OP_AVGPOOL) o = BufShape{in.h / op.stride, in.w / op.stride, in.c};
      t[op.dst] = o;
      per_img += (int64_t)o.h * o.w * o.c;
    }
    per_img *= (int64_t)elem_bytes(P->precision);
    int64_t c = P->head_l2_bytes / (per_img > 0 ? per_img : 1);
    if (c < 1) c = 1;
    if (c < n_img) S->chunk_imgs = (int)c;   // otherwise the whole micro-batch already fits: no chunking
  }
  const int64_t head_img = S->chunk_imgs > 0 ? S->chunk_imgs : n_img;
  buf_elems[0] = head_img * cur[0].h * cur[0].w * cur[0].c;
  S->op_src.assign(n_ops, BufShape());
  S->op_src2.assign(n_ops, BufShape());
  S->op_dst.assign(n_ops, BufShape());
  for (int i = 0; i < n_ops; ++i) {
    const semdiff_op& op = P->ops[i];
    if (op.src < 0 || op.src >= P->n_bufs) { set_error("op %d: bad src buffer %d", i, op.src); return SEMDIFF_ERR_ARG; }
    const BufShape in = cur[op.src];
    if (in.c == 0) { set_error("op %d reads buffer %d before it is written", i, op.src); return SEMDIFF_ERR_ARG; }
    S->op_src[i] = in;
    BufShape out;
    switch (op.kind) {
      case SEMDIFF_OP_CONV:
        if (in.c != op.cin) { set_error("op %d: cin %d != buffer channels %d", i, op.cin, in.c); return SEMDIFF_ERR_ARG; }
        {
          const int ph = op.pad + (op.pad_hi < 0 ? op.pad : op.pad_hi);
          out = BufShape{(in.h + ph - op.kh) / op.stride + 1, (in.w + ph - op.kw) / op.stride + 1, op.cout, in.pp};
        }
        if (op.src2 >= 0) {
          if (op.src2 >= P->n_bufs || cur[op.src2].c == 0) { set_error("op %d: bad second source buffer", i); return SEMDIFF_ERR_ARG; }
          const BufShape b2 = cur[op.src2];
          S->op_src2[i] = b2;
          const int st2 = op.stride2 < 1 ? 1 : op.stride2;
          if (b2.c != op.cin2 || (b2.h - 1) / st2 + 1 != out.h || (b2.w - 1) / st2 + 1 != out.w) {
            set_error("op %d: second source %dx%dx%d (stride %d) does not map onto output %dx%d", i, b2.h, b2.w, b2.c, st2, out.h, out.w);
            return SEMDIFF_ERR_ARG;
          }
        }
        if (op.res >= 0) {
          if (op.res >= P->n_bufs) { set_error("op %d: bad residual buffer", i); return SEMDIFF_ERR_ARG; }
          const BufShape r = cur[op.res];
          if (r.h != out.h || r.w != out.w || r.c != out.c) {
            set_error("op %d: residual shape %dx%dx%d != output %dx%dx%d", i, r.h, r.w, r.c, out.h, out.w, out.c);
            return SEMDIFF_ERR_ARG;
          }
        }
        break;
      case SEMDIFF_OP_MAXPOOL3S2: out = BufShape{(in.h - 1) / 2 + 1, (in.w - 1) / 2 + 1, in.c, in.pp}; break;
      case SEMDIFF_OP_AVGPOOL: out = BufShape{in.h / op.stride, in.w / op.stride, in.c, in.pp}; break;
      case SEMDIFF_OP_TAP: continue;
      case SEMDIFF_OP_SQDIFF:
        if (in.pp != 2) { set_error("op %d: SQDIFF needs a buffer holding GT and SR images", i); return SEMDIFF_ERR_ARG; }
        out = BufShape{in.h, in.w, in.c, 1};
        break;
      case SEMDIFF_OP_CONCAT: {
        if (op.src2 < 0 || op.src2 >= P->n_bufs || cur[op.src2].c == 0) { set_error("op %d: bad second source buffer", i); return SEMDIFF_ERR_ARG; }
        const BufShape b2 = cur[op.src2];
        if (b2.h != in.h || b2.w != in.w || b2.pp != in.pp || op.dst == op.src2) { set_error("op %d: CONCAT of %dx%d and %dx%d tensors", i, in.h, in.w, b2.h, b2.w); return SEMDIFF_ERR_ARG; }
        S->op_src2[i] = b2;
        out = BufShape{in.h, in.w, in.c + b2.c, in.pp};
        break;
      }
      case SEMDIFF_OP_UPSAMPLE2X: out = BufShape{2 * in.h, 2 * in.w, in.c, in.pp}; break;
      case SEMDIFF_OP_MAP_OUT:
        if (in.pp != 1) { set_error("op %d: MAP_OUT needs a one-image-per-pair buffer", i); return SEMDIFF_ERR_ARG; }
        S->map_h = 2 * in.h; S->map_w = 2 * in.w;
        continue;
      default: set_error("op %d: unknown kind %d", i, op.kind); return SEMDIFF_ERR_ARG;
    }
    if (out.h <= 0 || out.w <= 0) { set_error("op %d: empty output (input %dx%d too small)", i, in.h, in.w); return SEMDIFF_ERR_ARG; }
    if (op.dst <= 0 || op.dst >= P->n_bufs || op.dst == op.src || op.dst == op.res ||
        (op.kind == SEMDIFF_OP_CONV && op.dst == op.src2)) {
      set_error("op %d: bad dst buffer %d", i, op.dst);
      return SEMDIFF_ERR_ARG;
    }
    S->op_dst[i] = out;
    cur[op.dst] = out;
    const int64_t e = (i + 1 < P->head_ops ? head_img : (int64_t)pairs * out.pp) * out.h * out.w * out.c;
    if (e > buf_elems[op.dst]) buf_elems[op.dst] = e;
  }
  S->buf_offset.assign(P->n_bufs, 0);
  int64_t off = 0;
  const int64_t eb = (int64_t)elem_bytes(P->precision);
  for (int b = 0; b < P->n_bufs; ++b) {
    S->buf_offset[b] = off;
    off += (buf_elems[b] * eb + 1023) / 1024 * 1024;
  }
  S->partial_offset = off;
  off += (int64_t)(P->n_taps > 0 ? P->n_taps : 1) * pairs * SEMDIFF_MAX_PARTS * 4;
  S->total_bytes = (off + 1023) / 1024 * 1024;
  return 0;
}

static int choose_impl(const semdiff_plan* P, const ConvShape& cs) {
  if (is_split(P->precision)) return conv_tc_supported(cs, P->precision, true) ? SEMDIFF_CONV_TC_TMA : -1;   // tensor cores or nothing
  if (P->precision == SEMDIFF_FP32 || P->conv_impl == SEMDIFF_CONV_SIMT) return SEMDIFF_CONV_SIMT;
  if (P->conv_impl != SEMDIFF_CONV_TC_GATHER && cs.cin < 64 && conv_strip_supported(cs, P->precision)) return SEMDIFF_CONV_TC_TMA;
  if ((P->conv_impl != SEMDIFF_CONV_TC_GATHER || cs.cin2 != 0) && conv_tc_supported(cs, P->precision, true))
    return SEMDIFF_CONV_TC_TMA;
  if (conv_tc_supported(cs, P->precision, false)) return SEMDIFF_CONV_TC_GATHER;
  return SEMDIFF_CONV_SIMT;
}

static ConvShape conv_shape(const semdiff_op& op, const BufShape& in, const BufShape& in2, int n_img) {
  ConvShape cs;
  cs.n_img = n_img; cs.H = in.h; cs.W = in.w; cs.cin = op.cin; cs.cout = op.cout; cs.kh = op.kh; cs.kw = op.kw;
  cs.stride = op.stride; cs.pad = op.pad; cs.relu = op.relu; cs.pad_hi = op.pad_hi;
  cs.wscale = op.wscale > 0.f ? op.wscale : 1.f;
  if (op.src2 >= 0) { cs.cin2 = op.cin2; cs.stride2 = op.stride2 < 1 ? 1 : op.stride2; cs.H2 = in2.h; cs.W2 = in2.w; }
  return cs;
}
static ConvPtrs conv_ptrs(const semdiff_op& op, const ShapePlan& S, char* ws) {
  ConvPtrs q;
  q.in = ws + S.buf_offset[op.src];
  q.in2 = op.src2 >= 0 ? ws + S.buf_offset[op.src2] : nullptr;
  q.w = op.weight; q.bias = op.bias;
  q.res = op.res >= 0 ? ws + S.buf_offset[op.res] : nullptr;
  q.out = ws + S.buf_offset[op.dst];
  return q;
}

static int prepare(semdiff_plan* P, ShapePlan* S, int pairs, char* ws) {
  const int n_ops = (int)P->ops.size();
  S->tc.resize(n_ops);
  S->tc_tail.resize(n_ops);
  S->impl.assign(n_ops, 0);
  S->fused_away.assign(n_ops, 0);
  // SEMDIFF_NO_CHAIN=1 keeps every conv in its own launch (A/B testing)
  static const bool chain_ok = getenv("SEMDIFF_NO_CHAIN") == nullptr;
  static const bool pool_ok = getenv("SEMDIFF_NO_POOL_FUSION") == nullptr;
  const int chunk = S->chunk_imgs, tail = chunk > 0 ? (2 * pairs) % chunk : 0;
  for (int i = 0; i < n_ops; ++i) {
    const semdiff_op& op = P->ops[i];
    if (op.kind != SEMDIFF_OP_CONV || S->fused_away[i]) continue;
    const bool in_head = chunk > 0 && i < P->head_ops;
    if (in_head && i + 1 == P->head_ops) { set_error("the last head op must be a pooling op"); return SEMDIFF_ERR_UNSUPPORTED; }
    const ConvShape cs = conv_shape(op, S->op_src[i], S->op_src2[i], in_head ? chunk : pairs * S->op_src[i].pp);
    const int impl = choose_impl(P, cs);
    if (impl < 0) {
      set_error("op %d: conv %dx%d cin=%d cout=%d stride=%d has no split-precision kernel (needs whole 64-channel blocks)", i, op.kh, op.kw,
                op.cin, op.cout, op.stride);
      return SEMDIFF_ERR_UNSUPPORTED;
    }
    S->impl[i] = impl;
    const bool next_max = i + 1 < n_ops && P->ops[i + 1].kind == SEMDIFF_OP_MAXPOOL3S2 && conv_strip_pool_supported(cs, P->precision);
    const bool next_avg = i + 1 < n_ops && P->ops[i + 1].kind == SEMDIFF_OP_AVGPOOL && P->ops[i + 1].stride == 2 &&
                          conv_strip_avgpool_supported(cs, P->precision);
    if (impl == SEMDIFF_CONV_TC_TMA && pool_ok && P->conv_impl == SEMDIFF_CONV_AUTO && chunk == 0 && (next_max || next_avg) &&
        P->ops[i + 1].src == op.dst && op.res < 0 && op.src2 < 0) {
      // stem conv + max pool (ImageNet trunk) or + 2x2 average pool (CLIP trunk) in one launch (conv3x3_strip.cu, kPool)
      // - only if nothing else reads the un-pooled conv output, which is then never written
      bool dead = true;
      for (int j = i + 2; j < n_ops && dead; ++j) {
        const semdiff_op& o = P->ops[j];
        if (o.src == op.dst || ((o.kind == SEMDIFF_OP_CONV || o.kind == SEMDIFF_OP_CONCAT) && (o.res == op.dst || o.src2 == op.dst))) dead = false;
        else if (o.kind != SEMDIFF_OP_TAP && o.dst == op.dst) break;   // overwritten before any read
      }
      if (dead) {
        ConvPtrs q = conv_ptrs(op, *S, ws);
        q.out = ws + S->buf_offset[P->ops[i + 1].dst];
        int rc = next_max ? conv_strip_pool_prepare(&S->tc[i], q, cs, P->precision) : conv_strip_avgpool_prepare(&S->tc[i], q, cs, P->precision);
        if (rc != 0) return rc;
        S->fused_away[i + 1] = 1;
        continue;
      }
    }
    if (impl == SEMDIFF_CONV_TC_TMA && chain_ok && P->conv_impl == SEMDIFF_CONV_AUTO && !in_head) {
      // block boundary in the 256-channel stage: this conv's output tile feeds the next block's first 1x1 conv from
      // shared memory (TAP ops in between only read the output, which is still written in full)
      int j = i + 1;
      while (j < n_ops && P->ops[j].kind == SEMDIFF_OP_TAP) ++j;
      if (j < n_ops && P->ops[j].kind == SEMDIFF_OP_CONV && P->ops[j].src == op.dst && P->ops[j].res < 0 && P->ops[j].src2 < 0 &&
          P->ops[j].dst != op.src && P->ops[j].dst != op.res && P->ops[j].dst != op.src2 && P->ops[j].dst != op.dst) {
        const ConvShape c2 = conv_shape(P->ops[j], S->op_src[j], S->op_src2[j], pairs * S->op_src[j].pp);
        if (conv_chain_supported(cs, c2, op.res >= 0, P->precision)) {
          int rc = conv_chain_prepare(&S->tc[i], conv_ptrs(op, *S, ws), cs, conv_ptrs(P->ops[j], *S, ws), c2, P->precision);
          if (rc != 0) return rc;
          S->impl[j] = SEMDIFF_CONV_TC_TMA;
          S->fused_away[j] = 1;
          continue;
        }
      }
    }
    if (impl == SEMDIFF_CONV_TC_TMA || impl == SEMDIFF_CONV_TC_GATHER) {
      int rc = conv_tc_prepare(&S->tc[i], conv_ptrs(op, *S, ws), cs, P->precision, impl == SEMDIFF_CONV_TC_TMA);
      if (rc != 0) return rc;
      if (in_head && tail > 0) {
        const ConvShape ct = conv_shape(op, S->op_src[i], S->op_src2[i], tail);
        rc = conv_tc_prepare(&S->tc_tail[i], conv_ptrs(op, *S, ws), ct, P->precision, impl == SEMDIFF_CONV_TC_TMA);
        if (rc != 0) return rc;
      }
    }
  }
  S->prepared = true;
  S->prepared_ws = ws;
  return 0;
}

struct ProfScope {
  semdiff_plan* P; int slot; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr;
  ProfScope(semdiff_plan* p, int s, cudaStream_t stream) : P(p), slot(s), st(stream) {
    if (P->profiling) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, st); }
  }
  ~ProfScope() {
    if (P->profiling) { cudaEventRecord(b, st); P->events.push_back({slot, a, b}); }
  }
};

}  // namespace semdiff

extern "C" {
#pragma GCC visibility push(default)

const char* semdiff_last_error(void) { return g_err; }
const char* semdiff_version(void) { return "semdiff_b200 0.1 sm_100a"; }

int semdiff_plan_create(const semdiff_op* ops, int32_t n_ops, int32_t n_bufs, int32_t precision, int32_t input_layout,
                        int32_t head_ops, semdiff_plan** out) {
  if (ops == nullptr || out == nullptr || n_ops <= 0 || n_bufs < 2) { set_error("plan_create: bad arguments"); return SEMDIFF_ERR_ARG; }
  if (precision < SEMDIFF_BF16 || precision > SEMDIFF_BF16X3) { set_error("plan_create: bad precision %d", precision); return SEMDIFF_ERR_ARG; }
  if (is_split(precision) && input_layout != SEMDIFF_INPUT_S2D_ROW4 && input_layout != SEMDIFF_INPUT_S2D_ROW2) {
    set_error("plan_create: the split precisions need a row-window stem layout (SEMDIFF_INPUT_S2D_ROW4 / _ROW2)");
    return SEMDIFF_ERR_UNSUPPORTED;
  }
  if (input_layout < SEMDIFF_INPUT_NHWC8 || input_layout > SEMDIFF_INPUT_S2D16) { set_error("plan_create: bad input layout %d", input_layout); return SEMDIFF_ERR_ARG; }
  semdiff_plan* P = new semdiff_plan();
  P->input_layout = input_layout;
  P->head_ops = head_ops < 0 || head_ops > n_ops ? 0 : head_ops;
  if (const char* e = getenv("SEMDIFF_HEAD_L2_MB")) P->head_l2_bytes = (int64_t)atoll(e) << 20;  // 0 disables chunking
  if (P->head_l2_bytes <= 0) P->head_ops = 0;
  for (int i = 0; i < P->head_ops; ++i)
    if (ops[i].kind == SEMDIFF_OP_TAP || ops[i].res >= 0 || ops[i].src2 >= 0) P->head_ops = 0;  // keep the head simple
  P->ops.assign(ops, ops + n_ops);
  P->n_bufs = n_bufs;
  P->precision = precision;
  for (const semdiff_op& op : P->ops) {
    if (op.kind == SEMDIFF_OP_TAP && op.tap + 1 > P->n_taps) P->n_taps = op.tap + 1;
    if (op.kind == SEMDIFF_OP_MAP_OUT) P->has_map = true;
  }
  if (P->has_map && (P->n_taps != 0 || P->ops.back().kind != SEMDIFF_OP_MAP_OUT)) {
    set_error("plan_create: a map program ends in its single MAP_OUT op and has no TAP ops");
    delete P;
    return SEMDIFF_ERR_ARG;
  }
  if (!P->has_map && (P->n_taps < 1 || P->n_taps > 16)) {
    set_error("plan_create: need 1..16 TAP ops, got %d", P->n_taps);
    delete P;
    return SEMDIFF_ERR_ARG;
  }
  // tap channel counts come from a dry shape inference at a nominal size
  ShapePlan S;
  int rc = infer_shapes(P, 1, 224, 224, &S);
  if (rc != 0) { delete P; return rc; }
  P->tap_c.assign(P->n_taps, 0);
  for (size_t i = 0; i < P->ops.size(); ++i)
    if (P->ops[i].kind == SEMDIFF_OP_TAP) P->tap_c[P->ops[i].tap] = S.op_src[i].c;
  P->tap_off.assign(P->n_taps, 0);
  int off = 0;
  for (int j = 0; j < P->n_taps; ++j) {
    if (P->tap_c[j] == 0) { delete P; set_error("plan_create: tap %d missing", j); return SEMDIFF_ERR_ARG; }
    P->tap_off[j] = off;
    off += P->tap_c[j];
  }
  P->chan_total = off;
  P->prof_ms.assign(n_ops + 3, 0.0);
  P->prof_launches.assign(n_ops + 3, 0);
  *out = P;
  return 0;
}

int semdiff_plan_destroy(semdiff_plan* P) {
  if (P == nullptr) return 0;
  for (auto& e : P->events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  delete P;
  return 0;
}

int semdiff_plan_set_conv_impl(semdiff_plan* P, int32_t impl) {
  if (P == nullptr || impl < SEMDIFF_CONV_AUTO || impl > SEMDIFF_CONV_TC_TMA) { set_error("set_conv_impl: bad argument"); return SEMDIFF_ERR_ARG; }
  P->conv_impl = impl;
  P->shapes.clear();
  return 0;
}

int64_t semdiff_workspace_bytes(const semdiff_plan* P, int32_t pairs, int32_t H, int32_t W) {
  if (P == nullptr || pairs <= 0 || H <= 0 || W <= 0) { set_error("workspace_bytes: bad arguments"); return SEMDIFF_ERR_ARG; }
  ShapePlan S;
  int rc = infer_shapes(P, pairs, H, W, &S);
  if (rc != 0) return rc;
  return S.total_bytes;
}

int semdiff_plan_set_profiling(semdiff_plan* P, int32_t enable) {
  if (P == nullptr) return SEMDIFF_ERR_ARG;
  P->profiling = enable != 0;
  return 0;
}

int semdiff_plan_get_profile(semdiff_plan* P, float* out_ms, int32_t* out_launches, int32_t n, int32_t reset) {
  if (P == nullptr || n < (int)P->prof_ms.size()) { set_error("get_profile: need arrays of n_ops + 3"); return SEMDIFF_ERR_ARG; }
  for (auto& e : P->events) {
    SEMDIFF_CUDA_OK(cudaEventSynchronize(e.b));
    float ms = 0.f;
    SEMDIFF_CUDA_OK(cudaEventElapsedTime(&ms, e.a, e.b));
    P->prof_ms[e.slot] += ms;
    P->prof_launches[e.slot] += 1;
    cudaEventDestroy(e.a);
    cudaEventDestroy(e.b);
  }
  P->events.clear();
  for (size_t i = 0; i < P->prof_ms.size(); ++i) {
    if (out_ms) out_ms[i] = (float)P->prof_ms[i];
    if (out_launches) out_launches[i] = P->prof_launches[i];
  }
  if (reset) {
    std::fill(P->prof_ms.begin(), P->prof_ms.end(), 0.0);
    std::fill(P->prof_launches.begin(), P->prof_launches.end(), 0);
  }
  return 0;
}

int64_t semdiff_plan_last_launches(const semdiff_plan* P) { return P ? P->last_launches : -1; }

// the executor behind semdiff_score (programs with TAP ops -> head) and semdiff_score_map (programs ending in MAP_OUT)
static int run_program(semdiff_plan* P, const void* gt, const void* sr, int32_t in_precision, int32_t n_pairs, int32_t H,
                       int32_t W, int32_t mb, const float* head_w, const float* head_b, int32_t normalize, void* workspace,
                       int64_t workspace_bytes, float* out_scores, float* out_pre_relu, float* out_chan_mean, float* out_map,
                       cudaStream_t st) {
  if (n_pairs < 0 || H <= 0 || W <= 0 || mb <= 0) { set_error("score: bad sizes"); return SEMDIFF_ERR_ARG; }
  if (in_precision < SEMDIFF_BF16 || in_precision > SEMDIFF_FP32) { set_error("score: bad input precision %d", in_precision); return SEMDIFF_ERR_ARG; }
  P->last_launches = 0;
  if (n_pairs == 0) return 0;  // empty batch -> empty result, like the reference
  if (mb > n_pairs) mb = n_pairs;
  const int n_ops = (int)P->ops.size();
  const int64_t img_elems = (int64_t)3 * H * W;
  char* ws = reinterpret_cast<char*>(workspace);

  for (int p0 = 0; p0 < n_pairs; p0 += mb) {
    const int cur = n_pairs - p0 < mb ? n_pairs - p0 : mb;
    // small pass: overlap every kernel's prologue with its predecessor's tail (the split kernels' longer main loops amortise
    // their prologue sooner: 4 pairs of 1024x1024 in fp16x3 are 3.6 % SLOWER with it)
    g_pdl_auto = 2ll * cur * H * W <= (is_split(P->precision) ? 2000000ll : 8000000ll);
    ShapePlan& S = P->shapes[std::make_tuple(cur, H, W)];
    if (S.total_bytes == 0) {
      int rc = infer_shapes(P, cur, H, W, &S);
      if (rc != 0) return rc;
    }
    if (S.total_bytes > workspace_bytes) {
      set_error("score: workspace too small (%lld needed, %lld given)", (long long)S.total_bytes, (long long)workspace_bytes);
      return SEMDIFF_ERR_ARG;
    }
    if (!S.prepared || S.prepared_ws != ws) {
      int rc = prepare(P, &S, cur, ws);
      if (rc != 0) return rc;
    }
    float* partials = reinterpret_cast<float*>(ws + S.partial_offset);
    int tap_parts[16], tap_hw[16];
    const int n_img = 2 * cur;
    const int64_t in_bytes = (int64_t)elem_bytes(in_precision);
    const char* gt_mb = reinterpret_cast<const char*>(gt) + p0 * img_elems * in_bytes;
    const char* sr_mb = reinterpret_cast<const char*>(sr) + p0 * img_elems * in_bytes;
    // one op on `imgs` images; dst_img0 offsets the destination (head chunks write into the full-batch buffer)
    auto run_op = [&](int i, int imgs, int dst_img0, bool tail_launch) -> int {
      const semdiff_op& op = P->ops[i];
      const BufShape in = S.op_src[i];
      char* src = ws + S.buf_offset[op.src];
      int rc = 0;
      if (S.fused_away[i]) return 0;
      if (op.kind == SEMDIFF_OP_TAP) {
        ProfScope ps(P, n_ops + 1, st);
        const int j = op.tap, hw = in.h * in.w;
        tap_parts[j] = distance_parts(hw, in.c);
        tap_hw[j] = hw;
        float* cm = out_chan_mean ? out_chan_mean + (int64_t)p0 * P->chan_total + P->tap_off[j] : nullptr;
        rc = launch_distance(src, cur, hw, in.c, head_w + P->tap_off[j], normalize,
                             partials + (int64_t)j * cur * SEMDIFF_MAX_PARTS, cm, P->chan_total, P->precision, st);
        P->last_launches += cm ? 2 : 1;
        return rc;
      }
      ProfScope ps(P, i, st);
      if (op.kind == SEMDIFF_OP_MAP_OUT) {
        P->last_launches++;
        return launch_decoder_op(3, src, nullptr, out_map + (int64_t)p0 * S.map_h * S.map_w, cur, in.h, in.w, in.c, 0, P->precision, st);
      }
      const BufShape o = S.op_dst[i];
      char* dst = ws + S.buf_offset[op.dst] + (int64_t)dst_img0 * o.h * o.w * o.c * (int64_t)elem_bytes(P->precision);
      if (op.kind >= SEMDIFF_OP_SQDIFF) {
        const char* src2 = op.kind == SEMDIFF_OP_CONCAT ? ws + S.buf_offset[op.src2] : nullptr;
        P->last_launches++;
        return launch_decoder_op(op.kind - SEMDIFF_OP_SQDIFF, src, src2, dst, cur * o.pp, in.h, in.w, in.c,
                                 op.kind == SEMDIFF_OP_CONCAT ? S.op_src2[i].c : 0, P->precision, st);
      }
      if (op.kind == SEMDIFF_OP_CONV) {
        if (S.impl[i] == SEMDIFF_CONV_SIMT) {
          ConvPtrs q = conv_ptrs(op, S, ws);
          q.out = dst;
          rc = launch_conv_simt(q, conv_shape(op, in, S.op_src2[i], in.pp == 2 ? imgs : cur), P->precision, st);
        } else {
          rc = conv_tc_launch(tail_launch ? &S.tc_tail[i] : &S.tc[i], st);
        }
      } else if (op.kind == SEMDIFF_OP_MAXPOOL3S2) {
        rc = launch_maxpool3x3s2(src, dst, in.pp == 2 ? imgs : cur, in.h, in.w, in.c, P->precision, st);
      } else {
        rc = launch_avgpool(src, dst, in.pp == 2 ? imgs : cur, in.h, in.w, in.c, op.stride, P->precision, st);
      }
      P->last_launches++;
      return rc;
    };
    int first_op = 0;
    if (S.chunk_imgs > 0) {
      // head: pack -> stem -> pool per L2-sized chunk of images
      for (int c0 = 0; c0 < n_img; c0 += S.chunk_imgs) {
        const int cn = n_img - c0 < S.chunk_imgs ? n_img - c0 : S.chunk_imgs;
        {
          ProfScope ps(P, n_ops + 0, st);
          int rc = launch_pack(gt_mb, sr_mb, in_precision, cur, c0, cn, H, W, ws + S.buf_offset[0], P->precision, P->input_layout, st);
          if (rc != 0) return rc;
          P->last_launches++;
        }
        for (int i = 0; i < P->head_ops; ++i) {
          int rc = run_op(i, cn, i + 1 == P->head_ops ? c0 : 0, cn != S.chunk_imgs);
          if (rc != 0) return rc;
        }
      }
      first_op = P->head_ops;
    } else {
      ProfScope ps(P, n_ops + 0, st);
      int rc = launch_pack(gt_mb, sr_mb, in_precision, cur, 0, n_img, H, W, ws + S.buf_offset[0], P->precision, P->input_layout, st);
      if (rc != 0) return rc;
      P->last_launches++;
    }
    for (int i = first_op; i < n_ops; ++i) {
      int rc = run_op(i, n_img, 0, false);
      if (rc != 0) return rc;
    }
    if (!P->has_map) {
      ProfScope ps(P, n_ops + 2, st);
      int rc = launch_head(partials, P->n_taps, cur, tap_parts, tap_hw, head_b, out_scores + p0,
                           out_pre_relu ? out_pre_relu + p0 : nullptr, st);
      if (rc != 0) return rc;
      P->last_launches++;
    }
  }
  g_pdl_auto = false;
  return 0;
}

int semdiff_score(semdiff_plan* P, const void* gt, const void* sr, int32_t in_precision, int32_t n_pairs, int32_t H,
                  int32_t W, int32_t mb, const float* head_w, const float* head_b, int32_t normalize, void* workspace,
                  int64_t workspace_bytes, float* out_scores, float* out_pre_relu, float* out_chan_mean,
                  semdiff_stream_t stream_) {
  if (P == nullptr || gt == nullptr || sr == nullptr || head_w == nullptr || head_b == nullptr || workspace == nullptr ||
      out_scores == nullptr) { set_error("score: null argument"); return SEMDIFF_ERR_ARG; }
  if (P->has_map) { set_error("score: this plan produces a map (use semdiff_score_map)"); return SEMDIFF_ERR_ARG; }
  return run_program(P, gt, sr, in_precision, n_pairs, H, W, mb, head_w, head_b, normalize, workspace, workspace_bytes, out_scores,
                     out_pre_relu, out_chan_mean, nullptr, reinterpret_cast<cudaStream_t>(stream_));
}

int semdiff_score_map(semdiff_plan* P, const void* gt, const void* sr, int32_t in_precision, int32_t n_pairs, int32_t H, int32_t W,
                      int32_t mb, void* workspace, int64_t workspace_bytes, float* out_map, semdiff_stream_t stream_) {
  if (P == nullptr || gt == nullptr || sr == nullptr || workspace == nullptr || out_map == nullptr) { set_error("score_map: null argument"); return SEMDIFF_ERR_ARG; }
  if (!P->has_map) { set_error("score_map: this plan has no MAP_OUT op (use semdiff_score)"); return SEMDIFF_ERR_ARG; }
  return run_program(P, gt, sr, in_precision, n_pairs, H, W, mb, nullptr, nullptr, 0, workspace, workspace_bytes, nullptr, nullptr,
                     nullptr, out_map, reinterpret_cast<cudaStream_t>(stream_));
}

int semdiff_decoder_op(int32_t what, const void* in, const void* in2, void* out, int32_t n_img, int32_t H, int32_t W, int32_t c,
                       int32_t c2, int32_t precision, semdiff_stream_t st) {
  if (what < 0 || what > 3 || in == nullptr || out == nullptr || (what == 1 && in2 == nullptr)) { set_error("decoder_op: bad arguments"); return SEMDIFF_ERR_ARG; }
  return launch_decoder_op(what, in, in2, out, n_img, H, W, c, c2, precision, reinterpret_cast<cudaStream_t>(st));
}

int semdiff_pack_input(const void* gt, const void* sr, int32_t in_precision, int32_t n_pairs, int32_t H, int32_t W,
                       void* out, int32_t precision, int32_t layout, semdiff_stream_t st) {
  return launch_pack(gt, sr, in_precision, n_pairs, 0, 2 * n_pairs, H, W, out, precision, layout,
                     reinterpret_cast<cudaStream_t>(st));
}

int semdiff_conv2d(const void* in, const void* weight, const float* bias, const void* residual, void* out, int32_t n_img,
                   int32_t H, int32_t W, int32_t cin, int32_t cout, int32_t kh, int32_t kw, int32_t stride, int32_t pad,
                   int32_t relu, const void* in2, int32_t H2, int32_t W2, int32_t cin2, int32_t stride2, int32_t pad_hi,
                   int32_t precision, int32_t impl, semdiff_stream_t st_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(st_);
  ConvShape cs;
  cs.n_img = n_img; cs.H = H; cs.W = W; cs.cin = cin; cs.cout = cout; cs.kh = kh; cs.kw = kw; cs.stride = stride;
  cs.pad = pad; cs.relu = relu; cs.pad_hi = pad_hi;
  if (n_img <= 0 || stride < 1 || cs.OH() <= 0 || cs.OW() <= 0) { set_error("conv2d: bad shape"); return SEMDIFF_ERR_ARG; }
  if (in2 != nullptr) {
    cs.cin2 = cin2; cs.stride2 = stride2; cs.H2 = H2; cs.W2 = W2;
    if (stride2 < 1 || cin2 <= 0 || (H2 - 1) / stride2 + 1 != cs.OH() || (W2 - 1) / stride2 + 1 != cs.OW()) {
      set_error("conv2d: second source does not map onto the output grid");
      return SEMDIFF_ERR_ARG;
    }
  }
  if (impl == SEMDIFF_CONV_AUTO) {
    semdiff_plan tmp;
    tmp.precision = precision;
    impl = choose_impl(&tmp, cs);
    if (impl < 0) { set_error("conv2d: no split-precision kernel for this shape"); return SEMDIFF_ERR_UNSUPPORTED; }
  }
  if (is_split(precision) && impl != SEMDIFF_CONV_TC_TMA) { set_error("conv2d: split precisions run on SEMDIFF_CONV_TC_TMA only"); return SEMDIFF_ERR_UNSUPPORTED; }
  ConvPtrs q{in, in2, weight, bias, residual, out};
  switch (impl) {
    case SEMDIFF_CONV_SIMT: return launch_conv_simt(q, cs, precision, st);
    case SEMDIFF_CONV_TC_GATHER: return launch_conv_tc(q, cs, precision, false, st);
    case SEMDIFF_CONV_TC_TMA: return launch_conv_tc(q, cs, precision, true, st);
  }
  set_error("conv2d: bad impl %d", impl);
  return SEMDIFF_ERR_ARG;
}

int semdiff_conv2d_maxpool(const void* in, const void* weight, const float* bias, void* out, int32_t n_img, int32_t H,
                           int32_t W, int32_t cin, int32_t cout, int32_t kh, int32_t kw, int32_t pad, int32_t pad_hi,
                           int32_t relu, int32_t precision, semdiff_stream_t st_) {
  if (in == nullptr || weight == nullptr || bias == nullptr || out == nullptr || n_img <= 0) { set_error("conv2d_maxpool: bad arguments"); return SEMDIFF_ERR_ARG; }
  ConvShape cs;
  cs.n_img = n_img; cs.H = H; cs.W = W; cs.cin = cin; cs.cout = cout; cs.kh = kh; cs.kw = kw; cs.stride = 1;
  cs.pad = pad; cs.relu = relu; cs.pad_hi = pad_hi;
  ConvTcLaunch L;
  int rc = conv_strip_pool_prepare(&L, ConvPtrs{in, nullptr, weight, bias, nullptr, out}, cs, precision);
  if (rc != 0) return rc;
  return conv_tc_launch(&L, reinterpret_cast<cudaStream_t>(st_));
}

int semdiff_conv2d_avgpool(const void* in, const void* weight, const float* bias, void* out, int32_t n_img, int32_t H,
                           int32_t W, int32_t cin, int32_t relu, int32_t precision, semdiff_stream_t st_) {
  if (in == nullptr || weight == nullptr || bias == nullptr || out == nullptr || n_img <= 0) { set_error("conv2d_avgpool: bad arguments"); return SEMDIFF_ERR_ARG; }
  ConvShape cs;
  cs.n_img = n_img; cs.H = H; cs.W = W; cs.cin = cin; cs.cout = 64; cs.kh = cs.kw = 3; cs.stride = 1; cs.pad = 1; cs.relu = relu;
  ConvTcLaunch L;
  int rc = conv_strip_avgpool_prepare(&L, ConvPtrs{in, nullptr, weight, bias, nullptr, out}, cs, precision);
  if (rc != 0) return rc;
  return conv_tc_launch(&L, reinterpret_cast<cudaStream_t>(st_));
}

int semdiff_conv1x1_chain(const void* in, const void* in2, const void* w1, const float* bias1, const void* residual, void* out1,
                          const void* w2, const float* bias2, void* out2, int64_t m, int32_t cin, int32_t cin2, int32_t cout1,
                          int32_t cout2, int32_t relu1, int32_t relu2, int32_t precision, semdiff_stream_t st_) {
  if (in == nullptr || w1 == nullptr || bias1 == nullptr || out1 == nullptr || w2 == nullptr || bias2 == nullptr || out2 == nullptr ||
      m <= 0 || m >= ((int64_t)1 << 31) || cin <= 0 || cin2 < 0 || (in2 == nullptr) != (cin2 == 0)) {
    set_error("conv1x1_chain: bad arguments");
    return SEMDIFF_ERR_ARG;
  }
  ConvShape s1;  // the pixel dimension is carried as one image of m x 1 pixels
  s1.n_img = 1; s1.H = (int)m; s1.W = 1; s1.cin = cin; s1.cout = cout1; s1.kh = s1.kw = 1; s1.stride = 1; s1.pad = 0; s1.relu = relu1;
  if (in2 != nullptr) { s1.cin2 = cin2; s1.stride2 = 1; s1.H2 = (int)m; s1.W2 = 1; }
  ConvShape s2 = s1;
  s2.cin = cout1; s2.cin2 = 0; s2.cout = cout2; s2.relu = relu2;
  ConvTcLaunch L;
  int rc = conv_chain_prepare(&L, ConvPtrs{in, in2, w1, bias1, residual, out1}, s1, ConvPtrs{out1, nullptr, w2, bias2, nullptr, out2}, s2,
                              precision);
  if (rc != 0) return rc;
  return conv_chain_launch(&L, reinterpret_cast<cudaStream_t>(st_));
}

int semdiff_maxpool3x3s2(const void* in, void* out, int32_t n, int32_t H, int32_t W, int32_t c, int32_t precision,
                         semdiff_stream_t st) {
  return launch_maxpool3x3s2(in, out, n, H, W, c, precision, reinterpret_cast<cudaStream_t>(st));
}
int semdiff_avgpool(const void* in, void* out, int32_t n, int32_t H, int32_t W, int32_t c, int32_t window,
                    int32_t precision, semdiff_stream_t st) {
  return launch_avgpool(in, out, n, H, W, c, window, precision, reinterpret_cast<cudaStream_t>(st));
}

int32_t semdiff_distance_parts(int32_t hw, int32_t c) { return distance_parts(hw, c); }

int semdiff_layer_distance(const void* act, int32_t n_pairs, int32_t hw, int32_t c, const float* w, int32_t normalize,
                           float* partial, float* chan_mean, int32_t chan_stride, int32_t precision, semdiff_stream_t st) {
  return launch_distance(act, n_pairs, hw, c, w, normalize, partial, chan_mean, chan_stride, precision,
                         reinterpret_cast<cudaStream_t>(st));
}

int semdiff_head(const float* partials, int32_t n_taps, int32_t n_pairs, const int32_t* n_parts, const int32_t* hw,
                 const float* head_b, float* out_scores, float* out_pre_relu, semdiff_stream_t st) {
  return launch_head(partials, n_taps, n_pairs, n_parts, hw, head_b, out_scores, out_pre_relu,
                     reinterpret_cast<cudaStream_t>(st));
}

#pragma GCC visibility pop
}  // extern "C"
