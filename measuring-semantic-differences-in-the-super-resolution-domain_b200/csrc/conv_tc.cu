// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (the trunk hot path).
//
//   out[m, co] = act( sum_k A[m, k] * Wt[co, k] + bias[co] (+ residual[m, co]) )
//   m = (image, oh, ow) in NHWC order, k = (r, s, ci) with ci fastest, Wt = [Cout][KH][KW][Cin]
//
// Replaces the cuDNN conv + batch_norm + relu (+ add) sequence timm runs under
// /root/reference/models/global_eval_models.py:364,371 (self.clip(a) / self.clip(b)); BatchNorm is folded
// into Wt / bias on the host, GT and SR images share one launch.
//
// CTA = one 128 x BLOCK_N output tile at a time, persistent over tiles, warp-specialised:
//   warp 0      TMA producer: weight tile (always) and, for 1x1 stride-1 convs, the activation tile
//   warp 1      tcgen05.mma issuer (one thread); owns the TMEM allocation (2 accumulator stages)
//   warps 2-5   epilogue: tcgen05.ld -> +bias (+residual) -> ReLU -> 16-bit -> global
//   warps 6-9   (gather variant) software im2col: cp.async 16 B chunks into the 128B-swizzled A tile,
//               zero-filling padding / M tail / K tail
// Pipelines: smem ring full/empty (producers <-> MMA), TMEM full/empty (MMA <-> epilogue).
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"

namespace semdiff {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 x 16-bit = one 128-byte swizzle row
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int GATHER_LAG = 2;  // cp.async groups kept in flight per gather thread

template <int BLOCK_N> struct TcCfg {
  static constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = (196608 / STAGE_BYTES) > 8 ? 8 : (196608 / STAGE_BYTES);
  static constexpr int TMEM_COLS = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
  static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // + slack for 1024 B alignment
};

struct alignas(64) ConvTcParams {
  CUtensorMap tmA;  // activations as [M, Cin] (TMA variant only)
  CUtensorMap tmB;  // weights as [Cout, K]
  const void* in;
  const void* res;
  void* out;
  const float* bias;
  int H, W, Cin, OH, OW, Cout, KH, KW, stride, pad, relu;
  int M, num_kb, m_tiles, n_tiles, cpt, taps;
};

template <typename T, int BLOCK_N, bool kTmaA>
__global__ void __launch_bounds__(kTmaA ? 192 : 320, 1) conv_tc_kernel(const __grid_constant__ ConvTcParams p) {
  using Cfg = TcCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    if (kTmaA) tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], kTmaA ? 1 : 1 + 4);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 4);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_tile = tile / p.n_tiles, n_tile = tile - m_tile * p.n_tiles;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::B_STAGE_BYTES + (kTmaA ? A_STAGE_BYTES : 0));
          if (kTmaA) tma_load_2d(&p.tmA, &full_bar[stage], smem_a + stage * A_STAGE_BYTES, kb * BLOCK_K, m_tile * BLOCK_M);
          tma_load_2d(&p.tmB, &full_bar[stage], smem_b + stage * Cfg::B_STAGE_BYTES, kb * BLOCK_K, n_tile * BLOCK_N);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(Elem<T>::kUmmaFormat, BLOCK_M, BLOCK_N);
      int stage = 0, phase = 0, local = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
        const int acc = local & 1, acc_phase = (local >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * A_STAGE_BYTES);
          const uint32_t b_addr = smem_u32(smem_b + stage * Cfg::B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k) {
            umma_f16_ss(tmem_d, umma_smem_desc_sw128(a_addr + k * 32), umma_smem_desc_sw128(b_addr + k * 32), idesc,
                        (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (kb == p.num_kb - 1) umma_commit(&tmem_full_bar[acc]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp < 6) {
    // ===================== epilogue =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may touch
    const int row = q * 32 + lane;
    const T* res = reinterpret_cast<const T*>(p.res);
    T* out = reinterpret_cast<T*>(p.out);
    int local = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      const int m_tile = tile / p.n_tiles, n_tile = tile - m_tile * p.n_tiles;
      const int acc = local & 1, acc_phase = (local >> 1) & 1;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
      const int gm = m_tile * BLOCK_M + row;
      const bool ok = gm < p.M;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + (uint32_t(q * 32) << 16) + acc * BLOCK_N + c * 32, v);
        tmem_ld_wait();
        if (ok) {
          const int col0 = n_tile * BLOCK_N + c * 32;
          const int64_t off = (int64_t)gm * p.Cout + col0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j * 8));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j * 8 + 4));
            float f[8] = {__uint_as_float(v[j * 8 + 0]) + b0.x, __uint_as_float(v[j * 8 + 1]) + b0.y,
                          __uint_as_float(v[j * 8 + 2]) + b0.z, __uint_as_float(v[j * 8 + 3]) + b0.w,
                          __uint_as_float(v[j * 8 + 4]) + b1.x, __uint_as_float(v[j * 8 + 5]) + b1.y,
                          __uint_as_float(v[j * 8 + 6]) + b1.z, __uint_as_float(v[j * 8 + 7]) + b1.w};
            if (res != nullptr) {
              float r[8];
              unpack8<T>(*reinterpret_cast<const uint4*>(res + off + j * 8), r);
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] += r[e];
            }
            if (p.relu) {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
            }
            *reinterpret_cast<uint4*>(out + off + j * 8) = pack8<T>(f);
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
    }
  } else if (!kTmaA) {
    // ===================== software im2col gather (128 threads, one A row each) =====================
    const int row = (warp - 6) * 32 + lane;
    const T* in = reinterpret_cast<const T*>(p.in);
    const uint32_t row_off = row * 128, sw = row & 7;
    int stage = 0, phase = 0, arr_stage = 0;
    int issued = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.n_tiles;
      const int gm = m_tile * BLOCK_M + row;
      const bool row_ok = gm < p.M;
      int n = 0, oh = 0, ow = 0;
      if (row_ok) { n = gm / (p.OH * p.OW); const int r = gm - n * p.OH * p.OW; oh = r / p.OW; ow = r - oh * p.OW; }
      const int ih0 = oh * p.stride - p.pad, iw0 = ow * p.stride - p.pad;
      const T* img = in + (int64_t)n * p.H * p.W * p.Cin;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        const uint32_t dst = smem_u32(smem_a + stage * A_STAGE_BYTES) + row_off;
        if (p.cpt >= 8) {
          // one tap per k-block: 128 contiguous bytes of one input pixel
          const int blocks_per_tap = p.cpt >> 3;
          const int tap = kb / blocks_per_tap, cc = (kb - tap * blocks_per_tap) * 64;
          const int r = tap / p.KW, s = tap - r * p.KW;
          const int ih = ih0 + r, iw = iw0 + s;
          const bool ok = row_ok && tap < p.taps && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W;
          const T* src = ok ? img + ((int64_t)ih * p.W + iw) * p.Cin + cc : in;
#pragma unroll
          for (int j = 0; j < 8; ++j) cp_async_16(dst + ((j ^ sw) << 4), src + (ok ? j * 8 : 0), ok ? 16u : 0u);
        } else {
          // several taps per k-block (Cin = 8 .. 56): each 16 B chunk is its own (tap, channel group)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int g = kb * 8 + j;
            const int tap = g / p.cpt, cc = (g - tap * p.cpt) * 8;
            const int r = tap / p.KW, s = tap - r * p.KW;
            const int ih = ih0 + r, iw = iw0 + s;
            const bool ok = row_ok && tap < p.taps && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W;
            const T* src = ok ? img + ((int64_t)ih * p.W + iw) * p.Cin + cc : in;
            cp_async_16(dst + ((j ^ sw) << 4), src, ok ? 16u : 0u);
          }
        }
        cp_async_commit();
        ++issued;
        if (issued > GATHER_LAG) {
          cp_async_wait<GATHER_LAG>();
          fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
          __syncwarp();
          if (lane == 0) mbar_arrive(&full_bar[arr_stage]);
          if (++arr_stage == STAGES) arr_stage = 0;
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    cp_async_wait<0>();
    fence_proxy_async_smem();
    __syncwarp();
    const int pending = issued < GATHER_LAG ? issued : GATHER_LAG;
    for (int i = 0; i < pending; ++i) {
      if (lane == 0) mbar_arrive(&full_bar[arr_stage]);
      if (++arr_stage == STAGES) arr_stage = 0;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 2-D K-major 16-bit tensor [rows, cols] (cols contiguous), box = 64 cols x box_rows, 128B swizzle, OOB -> 0
static int make_tmap_2d(CUtensorMap* m, const void* base, int precision, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_tiled();
  if (enc == nullptr) { set_error("cuTensorMapEncodeTiled entry point not found"); return SEMDIFF_ERR_CUDA; }
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * 2};
  const cuuint32_t box[2] = {BLOCK_K, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = precision == SEMDIFF_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = enc(m, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu box_rows=%u base=%p", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, box_rows, base);
    return SEMDIFF_ERR_CUDA;
  }
  return 0;
}

static int pick_block_n(int cout) {
  if (cout % 256 == 0) return 256;
  if (cout % 128 == 0) return 128;
  if (cout % 64 == 0) return 64;
  if (cout % 32 == 0) return 32;
  return 0;
}

bool conv_tc_supported(const ConvShape& s, int precision, bool use_tma) {
  if (precision != SEMDIFF_BF16 && precision != SEMDIFF_FP16) return false;
  if (s.cin % 8 != 0 || pick_block_n(s.cout) == 0) return false;
  if (s.cin > 64 && s.cin % 64 != 0) return false;
  if (s.cin < 64 && 64 % s.cin != 0) return false;
  if (use_tma && !(s.kh == 1 && s.kw == 1 && s.stride == 1 && s.pad == 0 && s.cin % 64 == 0)) return false;
  return s.M() > 0 && s.M() < (int64_t)1 << 31;
}

static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <typename T, int BLOCK_N, bool kTmaA>
static int launch_t(const ConvTcParams& p, cudaStream_t st) {
  using Cfg = TcCfg<BLOCK_N>;
  static bool configured = false;
  auto kern = conv_tc_kernel<T, BLOCK_N, kTmaA>;
  if (!configured) {
    SEMDIFF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  const int tiles = p.m_tiles * p.n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, kTmaA ? 192 : 320, Cfg::SMEM_BYTES, st>>>(p);
  SEMDIFF_CUDA_OK(cudaGetLastError());
  return 0;
}

template <typename T, bool kTmaA>
static int launch_n(const ConvTcParams& p, int block_n, cudaStream_t st) {
  switch (block_n) {
    case 256: return launch_t<T, 256, kTmaA>(p, st);
    case 128: return launch_t<T, 128, kTmaA>(p, st);
    case 64: return launch_t<T, 64, kTmaA>(p, st);
    case 32: return launch_t<T, 32, kTmaA>(p, st);
  }
  set_error("conv_tc: unsupported BLOCK_N %d", block_n);
  return SEMDIFF_ERR_UNSUPPORTED;
}

int conv_tc_prepare(ConvTcLaunch* L, const void* in, const void* w, const float* bias, const void* res, void* out,
                    const ConvShape& s, int precision, bool use_tma) {
  if (!conv_tc_supported(s, precision, use_tma)) {
    set_error("conv_tc: unsupported shape cin=%d cout=%d k=%dx%d stride=%d pad=%d tma=%d precision=%d", s.cin, s.cout,
              s.kh, s.kw, s.stride, s.pad, (int)use_tma, precision);
    return SEMDIFF_ERR_UNSUPPORTED;
  }
  static_assert(sizeof(ConvTcParams) <= sizeof(L->params), "ConvTcLaunch::params too small");
  ConvTcParams& p = *reinterpret_cast<ConvTcParams*>(L->params);
  memset(&p, 0, sizeof(p));
  const int block_n = pick_block_n(s.cout);
  p.in = in; p.res = res; p.out = out; p.bias = bias;
  p.H = s.H; p.W = s.W; p.Cin = s.cin; p.OH = s.OH(); p.OW = s.OW(); p.Cout = s.cout;
  p.KH = s.kh; p.KW = s.kw; p.stride = s.stride; p.pad = s.pad; p.relu = s.relu;
  p.M = (int)s.M();
  p.num_kb = (s.K() + BLOCK_K - 1) / BLOCK_K;
  p.m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
  p.n_tiles = s.cout / block_n;
  p.cpt = s.cin / 8;
  p.taps = s.kh * s.kw;
  int rc = make_tmap_2d(&p.tmB, w, precision, (uint64_t)s.cout, (uint64_t)s.K(), (uint32_t)block_n);
  if (rc != 0) return rc;
  if (use_tma) {
    rc = make_tmap_2d(&p.tmA, in, precision, (uint64_t)p.M, (uint64_t)s.cin, BLOCK_M);
    if (rc != 0) return rc;
  }
  L->block_n = block_n;
  L->use_tma = use_tma ? 1 : 0;
  L->precision = precision;
  return 0;
}

int conv_tc_launch(const ConvTcLaunch* L, cudaStream_t st) {
  const ConvTcParams& p = *reinterpret_cast<const ConvTcParams*>(L->params);
  if (L->precision == SEMDIFF_BF16)
    return L->use_tma ? launch_n<__nv_bfloat16, true>(p, L->block_n, st) : launch_n<__nv_bfloat16, false>(p, L->block_n, st);
  return L->use_tma ? launch_n<__half, true>(p, L->block_n, st) : launch_n<__half, false>(p, L->block_n, st);
}

int launch_conv_tc(const void* in, const void* w, const float* bias, const void* res, void* out, const ConvShape& s,
                   int precision, bool use_tma, cudaStream_t st) {
  ConvTcLaunch L;
  int rc = conv_tc_prepare(&L, in, w, bias, res, out, s, precision, use_tma);
  if (rc != 0) return rc;
  return conv_tc_launch(&L, st);
}

}  // namespace semdiff
