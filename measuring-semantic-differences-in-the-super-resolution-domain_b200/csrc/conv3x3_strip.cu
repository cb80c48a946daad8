// 3x3 stride-1 pad-1 convolution, 64 -> 64 channels, for the wide early layers (ResNet layer1, CLIP stem): the
// "strip" variant of the tcgen05 implicit GEMM.
//
// The generic kernel (conv_tc.cu) fetches the activation tile of every filter tap separately, i.e. every input byte
// crosses L2 -> SM nine times, and for 64 output channels that traffic (not DRAM, not the tensor pipe) bounds the
// kernel.  Here a CTA loads, ONCE per tile, the strip of input pixels its outputs need - (RT + 2) image rows of P
// pixels (P = power of two >= W + 2, RT = 128 / P output rows), one 128-byte row of shared memory per pixel - and the
// nine taps are nine VIEWS of that strip: the A operand of tap (r, s) is the 128 consecutive shared-memory rows
// starting at row r * P + s.  "Virtual" output pixel v = oy * P + ox (ox >= W are throw-away columns) keeps the view a
// plain row-contiguous K-major tile.  A start that is not a multiple of 8 rows needs nothing special: the 128-byte
// swizzle is a function of the absolute shared-memory address (measured: descriptor base-offset 0 is correct, a
// non-zero base offset corrupts the result), so TMA's swizzled strip is read consistently at any row offset.
// Weights (9 x 64 x 64) stay resident in shared memory.  Measured on B200 (512 images): 56x56 0.207 -> 0.122 ms,
// 112x112 0.794 -> 0.487 ms against the generic im2col kernel.
#include <cuda.h>

#include <type_traits>

#include "common.cuh"
#include "kernels.h"

namespace semdiff {

// Two instantiations share the code (64 -> 64 channels, stride 1):
//   RG = 1: 3x3 pad 1 (ResNet layer1 / CLIP stage 0 and stem convs): strip = RT + 2 rows, 9 taps
//   RG = 2: KHx1 pad 0 "row-window" stems (4x1 over SEMDIFF_INPUT_S2D_ROW4, 2x1 over ..._ROW2): no overlap between
//           the taps in W, but every input row feeds KH output rows, so a tile computes RG = 2 output row groups
//           (two TMEM accumulators) from one strip of 2 * RT + KH - 1 rows: 5 rows instead of 8 for the 7x7 stem.
//   RG = 4, ROWB = 32: the same (four row groups per strip) with 16-channel pixels (SEMDIFF_INPUT_S2D16, 32-byte rows, SWIZZLE_32B, one K = 16 MMA
//           per tap): the 7x7/2 stem as a 4x4 conv (pad 2 before / 1 after) whose 16 taps are views shifted by
//           (r rows, j pixels) - the 4x window expansion of the row-window layouts is never materialised.
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// kPool = 2 (RG = 1, 64-channel pixels, P = 128): the 2 x 2 average pool that follows the CLIP stem's third conv, in
// registers: row pairs are consecutive tiles of one CTA (ranges start on even rows), the top row's (a00 + a01) waits in
// registers for the bottom row, the horizontal neighbour comes by warp shuffle; summation order and the single rounding
// are those of avgpool_kernel, so the result is bit-identical and the 112 x 112 x 64 conv output is never written.
// kPool = 1 (RG = 4, ROWB = 32 only): the 3x3 stride-2 pad-1 max pool that follows the ResNet stem runs inside the epilogue.
// The RG * RT conv rows of a tile are staged (16-bit, post-ReLU) in shared memory as before, but instead of being stored
// they are pooled together with the LAST conv row of the previous tile (kept in a carry buffer): a CTA walks a
// contiguous range of tiles, top to bottom through each image, so the carry is always the row above - only the first
// tile of a range that starts inside an image is computed twice (once as a warm-up for its last row).  The
// 112 x 112 x 64 stem output (822 MB per 256 pairs) never exists in HBM.
template <int RG, int ROWB = 128, int kPool = 0, int NC = 64> struct StripCfg {
  static_assert(NC == 64 || (NC == 32 && kPool == 0), "32 output channels: plain variants only");
  // P = 128 worst case; + slack for the rows that taps shifted in W read beyond the strip (not needed for KW = 1)
  static constexpr int MAX_STRIP = (RG == 1 ? 3 : RG + 3) * 128 * ROWB + ((RG == 1 || ROWB == 32) ? 1024 : 0);
  static constexpr int B_TAP_BYTES = NC * ROWB;
  static constexpr int B_BYTES = (RG == 1 ? 9 : (ROWB == 32 ? 16 : 4)) * B_TAP_BYTES;
  static constexpr int C_BUFS = kPool ? 1 : ((RG == 1 || ROWB == 32) ? 2 : 1);  // the 64-channel row-window variant has no room for two
  static constexpr int C_ROW = NC * 2;               // staged bytes per output pixel
  static constexpr int C_BYTES = RG * 128 * C_ROW;
  static constexpr int POOL_BYTES = kPool == 2 ? 64 * 128 : RG * 128 * 32;   // max: (RG * RT / 2) pooled rows of P / 2 pixels (RG * RT * P = RG * 128); avg: one row of 64
  static constexpr int CARRY_BYTES = 128 * 128;      // one conv row, P <= 128 pixels
  static constexpr int TMEM_COLS = 2 * RG * NC;   // 2 stages x RG accumulators x NC columns
  static constexpr int SMEM = 2 * MAX_STRIP + B_BYTES + C_BUFS * C_BYTES + (kPool ? 2 * POOL_BYTES : 0) + (kPool == 1 ? 2 * CARRY_BYTES : 0) + 16 * 8 + 16 + 1024;
  static_assert(SMEM <= 232448, "shared memory budget");
};

struct alignas(64) StripParams {
  CUtensorMap tmX;  // input NHWC as (C, W, H, N), box (64, P, RT + 2, 1)
  CUtensorMap tmB;  // weights [64, 576], box (64, 64)
  CUtensorMap tmC;  // output as (64, OW, n_img * OH), box (64, CW, 1): a partial last column block is clipped in OW
                    // (kPool: the POOLED output as (64, POW, n_img * POH), box (64, POW, 1))
  const float* bias;
  int H, W, OH, OW, P, RT, KH, KW, pad, n_img, tiles_per_img, relu;
  int POH, POW;     // kPool: pooled output size
  int CW, col_blocks;  // images wider than one strip are cut into column blocks of CW output pixels (then P = 128, RT = 1)
};
static_assert(sizeof(StripParams) <= 896, "ConvTcLaunch::params too small");

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// K-major operand descriptor for ROWB-byte rows: 128B swizzle (8-row groups 1024 B apart) or 32B swizzle (256 B apart)
template <int ROWB> __device__ __forceinline__ uint64_t strip_desc(uint32_t smem_addr) {
  if constexpr (ROWB == 128) return umma_smem_desc_sw128(smem_addr);
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>((8 * ROWB) >> 4) << 32;   // SBO: 8 rows x ROWB
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(ROWB == 64 ? 4 : 6) << 61;   // SWIZZLE_64B | SWIZZLE_32B
  return d;
}

template <typename T> __device__ __forceinline__ uint4 strip_max8(const uint4& a, const uint4& b) {
  using T2 = typename std::conditional<std::is_same<T, __half>::value, __half2, __nv_bfloat162>::type;
  uint4 r;
  const T2* pa = reinterpret_cast<const T2*>(&a);
  const T2* pb = reinterpret_cast<const T2*>(&b);
  T2* pr = reinterpret_cast<T2*>(&r);
#pragma unroll
  for (int k = 0; k < 4; ++k) pr[k] = __hmax2(pa[k], pb[k]);
  return r;
}
template <typename T> __device__ __forceinline__ uint32_t strip_max2(uint32_t a, uint32_t b) {
  using T2 = typename std::conditional<std::is_same<T, __half>::value, __half2, __nv_bfloat162>::type;
  const T2 r = __hmax2(*reinterpret_cast<const T2*>(&a), *reinterpret_cast<const T2*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& o) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
}

template <typename T, int RG, int ROWB, int kPool = 0, int NC = 64>
__global__ void __launch_bounds__(320, 1) conv3x3_strip_kernel(const __grid_constant__ StripParams p) {
  using Cfg = StripCfg<RG, ROWB, kPool, NC>;
  constexpr int C_ROW = Cfg::C_ROW;
  constexpr int KSTEPS = ROWB / 32;   // K = 16 MMAs per tap
  constexpr int STRIP_MAX_BYTES = Cfg::MAX_STRIP, STRIP_B_BYTES = Cfg::B_BYTES, STRIP_C_BYTES = Cfg::C_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* strip = smem;                                   // [2][STRIP_MAX_BYTES]
  uint8_t* smem_b = smem + 2 * STRIP_MAX_BYTES;            // [taps][64 x 128 B]
  uint8_t* smem_c = smem_b + STRIP_B_BYTES;                // [C_BUFS][RG][128 x 128 B]
  uint8_t* smem_pool = smem_c + Cfg::C_BUFS * STRIP_C_BYTES;            // kPool: [2][POOL_BYTES] pooled rows, then
  uint8_t* smem_carry = smem_pool + (kPool ? 2 * Cfg::POOL_BYTES : 0);   //        [2][CARRY_BYTES] last conv row of a tile
  uint64_t* strip_full = reinterpret_cast<uint64_t*>(smem_carry + (kPool == 1 ? 2 * Cfg::CARRY_BYTES : 0));
  uint64_t* strip_empty = strip_full + 2;
  uint64_t* tmem_full = strip_empty + 2;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* b_bar = tmem_empty + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(b_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool leader = elect_one();
  const int total_tiles = p.n_img * p.tiles_per_img;
  const int taps = p.KH * p.KW;
  const uint32_t strip_bytes = (uint32_t)(RG * p.RT + p.KH - 1) * p.P * ROWB;
  // tile schedule: round-robin, or (kPool) a contiguous range per CTA preceded by one warm-up tile when the range
  // starts below the top of an image
  int t_first = blockIdx.x, t_begin = blockIdx.x, t_end = total_tiles, t_step = gridDim.x;
  if (kPool == 1) {
    t_begin = (int)((int64_t)total_tiles * blockIdx.x / gridDim.x);
    t_end = (int)((int64_t)total_tiles * (blockIdx.x + 1) / gridDim.x);
    t_step = 1;
    t_first = t_begin - ((t_begin < t_end && t_begin % p.tiles_per_img != 0) ? 1 : 0);
  } else if (kPool == 2) {   // ranges of whole row pairs (one conv row per tile, OH even)
    t_begin = 2 * (int)((int64_t)(total_tiles / 2) * blockIdx.x / gridDim.x);
    t_end = 2 * (int)((int64_t)(total_tiles / 2) * (blockIdx.x + 1) / gridDim.x);
    t_step = 1;
    t_first = t_begin;
  }

  if (warp == 0 && leader) {
    tma_prefetch_desc(&p.tmX); tma_prefetch_desc(&p.tmB); tma_prefetch_desc(&p.tmC);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&strip_full[i], 1); mbar_init(&strip_empty[i], 1);
      mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 8);
    }
    mbar_init(b_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_trigger();

  if (warp == 0) {
    if (leader) {
      mbar_arrive_expect_tx(b_bar, taps * Cfg::B_TAP_BYTES);
      for (int t = 0; t < taps; ++t) tma_load_2d(&p.tmB, b_bar, smem_b + t * Cfg::B_TAP_BYTES, t * (ROWB / 2), 0);
      pdl_wait();  // the strips are the previous kernel's output
      int local = 0;
      for (int tile = t_first; tile < t_end; tile += t_step, ++local) {
        const int b = local & 1, ph = (local >> 1) & 1;
        const int n = tile / p.tiles_per_img, t_in = tile - n * p.tiles_per_img;
        const int rg = t_in / p.col_blocks, cb = t_in - rg * p.col_blocks;
        const int oy0 = rg * p.RT * RG, ox0 = cb * p.CW;
        mbar_wait(&strip_empty[b], ph ^ 1);
        mbar_arrive_expect_tx(&strip_full[b], strip_bytes);
        tma_load_4d(&p.tmX, &strip_full[b], strip + b * STRIP_MAX_BYTES, 0, ox0 - p.pad, oy0 - p.pad, n);  // halo: OOB -> zeros
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_f16(Elem<T>::kUmmaFormat, 128, NC);
    const uint64_t b_desc0 = strip_desc<ROWB>(smem_u32(smem_b));
    int local = 0;
    if (t_first < t_end) mbar_wait(b_bar, 0);
    for (int tile = t_first; tile < t_end; tile += t_step, ++local) {
      const int b = local & 1, ph = (local >> 1) & 1;
      mbar_wait(&tmem_empty[b], ph ^ 1);
      mbar_wait(&strip_full[b], ph);
      tcgen05_fence_after();
      if (leader) {
        // The issuing thread's scalar code is on the critical path (one lane, dependent uniform-datapath instructions):
        // descriptors are built once and advanced by adds - the start-address field (addr >> 4) never carries out of
        // its 14 bits inside the 227 KB of shared memory.  Before: a division and a full descriptor rebuild per tap and
        // row group (57 instructions per 4 MMAs in the stem) bounded the kernel.
        const uint64_t a_base = strip_desc<ROWB>(smem_u32(strip + b * STRIP_MAX_BYTES));
        const uint32_t s_step = ROWB >> 4, r_step = (uint32_t)(p.P * ROWB) >> 4, g_step = (uint32_t)(p.RT * p.P * ROWB) >> 4;
        const uint32_t tmem_d0 = tmem_base + b * RG * NC;
        uint64_t b_desc = b_desc0, a_row = a_base;
        uint32_t acc_flag = 0;
#pragma unroll 1
        for (int r = 0; r < p.KH; ++r, a_row += r_step) {
          uint64_t a_tap = a_row;
#pragma unroll 1
          for (int sx = 0; sx < p.KW; ++sx, a_tap += s_step, b_desc += (uint64_t)(Cfg::B_TAP_BYTES >> 4)) {
            uint64_t a_desc = a_tap;   // view of the strip shifted by (g row groups + r rows, sx pixels)
#pragma unroll
            for (int g = 0; g < RG; ++g, a_desc += g_step) {
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k)
                umma_f16_ss(tmem_d0 + g * NC, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, k == 0 ? acc_flag : 1u);
            }
            acc_flag = 1u;
          }
        }
        umma_commit(&strip_empty[b]);
        umma_commit(&tmem_full[b]);
      }
      __syncwarp();
    }
  } else {
    // epilogue, 8 warps = 4 TMEM lane quarters x 2: the second set of four takes the second row group (RG = 2) or the
    // second half of the channels (RG = 1); with RG = 4 each set takes two row groups.  Virtual pixel v = oy * P + ox -> one staged 128-byte row; one TMA store
    // per output image row.
    const int q = warp & 3, v = q * 32 + lane;
    const int set = (warp - 2) >> 2;
    const bool store_thread = (warp == 2 && leader);
    const int pool_prl0 = kPool == 1 ? (int)((threadIdx.x - 64) >> 3) / p.POW : 0;
    const int pool_pc0 = kPool == 1 ? (int)((threadIdx.x - 64) >> 3) - pool_prl0 * p.POW : 0;
    float avg_top[32];        // kPool == 2: (a00 + a01) of the top row of the current 2 x 2 windows
#pragma unroll
    for (int k = 0; k < 32; ++k) avg_top[k] = 0.f;
    uint32_t carry_reg[16];   // kPool == 1, RT == 1: this thread's pixel of the previous tile's last conv row (32 channels, packed)
#pragma unroll
    for (int k = 0; k < 16; ++k) carry_reg[k] = 0;
    int local = 0;
    for (int tile = t_first; tile < t_end; tile += t_step, ++local) {
      const int b = local & 1, ph = (local >> 1) & 1;
      const int n = tile / p.tiles_per_img, t_in = tile - n * p.tiles_per_img;
      const int rg = t_in / p.col_blocks, cb = t_in - rg * p.col_blocks;
      const int oy0 = rg * p.RT * RG, ox0 = cb * p.CW;
      uint8_t* cbuf = smem_c + (Cfg::C_BUFS == 2 ? b : 0) * STRIP_C_BYTES;
      if (store_thread) bulk_wait_read<kPool ? 1 : Cfg::C_BUFS - 1>();   // kPool: the pooled buffer of two tiles ago
      named_bar_sync(1, 256);
      mbar_wait_backoff(&tmem_full[b], ph);
      tcgen05_fence_after();
      if constexpr (kPool == 2) {
        // ---- 2 x 2 average pool in registers (one conv row per tile; rows 2k / 2k + 1 are consecutive tiles of this CTA)
        uint32_t acc[32], cur[16];
        tmem_ld_32x32b_x32(tmem_base + (uint32_t(q * 32) << 16) + b * 64 + set * 32, acc);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[b]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c0 = set * 32 + j * 8;
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + c0));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + 4));
          float f[8] = {__uint_as_float(acc[j * 8 + 0]) + b0.x, __uint_as_float(acc[j * 8 + 1]) + b0.y,
                        __uint_as_float(acc[j * 8 + 2]) + b0.z, __uint_as_float(acc[j * 8 + 3]) + b0.w,
                        __uint_as_float(acc[j * 8 + 4]) + b1.x, __uint_as_float(acc[j * 8 + 5]) + b1.y,
                        __uint_as_float(acc[j * 8 + 6]) + b1.z, __uint_as_float(acc[j * 8 + 7]) + b1.w};
          if (p.relu) {
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
          }
          const uint4 o = pack8<T>(f);   // the 16-bit conv output the un-fused pool would read
          cur[j * 4 + 0] = o.x; cur[j * 4 + 1] = o.y; cur[j * 4 + 2] = o.z; cur[j * 4 + 3] = o.w;
        }
        const bool bottom = (oy0 & 1) != 0;
        uint32_t outp[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const uint32_t nb = __shfl_down_sync(0xffffffffu, cur[k], 1);   // pixel v + 1
          const T* me = reinterpret_cast<const T*>(&cur[k]);
          const T* ri = reinterpret_cast<const T*>(&nb);
          float r2[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            // avgpool_kernel's order: ((0 + a00) + a01) + a10) + a11, then * 0.25, one rounding
            const float t = ((bottom ? avg_top[2 * k + e] : 0.f) + Elem<T>::to_f(me[e])) + Elem<T>::to_f(ri[e]);
            avg_top[2 * k + e] = t;
            r2[e] = t * 0.25f;
          }
          T* op = reinterpret_cast<T*>(&outp[k]);
          op[0] = Elem<T>::from_f(r2[0]); op[1] = Elem<T>::from_f(r2[1]);
        }
        if (bottom && (v & 1) == 0 && v + 1 < p.OW) {
          const uint32_t pool_addr = smem_u32(smem_pool + ((local >> 1) & 1) * Cfg::POOL_BYTES);
          const uint32_t prow = (uint32_t)(v >> 1);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts128(pool_addr + prow * 128 + ((uint32_t)((set * 4 + j) ^ (prow & 7)) << 4),
                   make_uint4(outp[4 * j], outp[4 * j + 1], outp[4 * j + 2], outp[4 * j + 3]));
        }
        if (bottom) {
          fence_proxy_async_smem();
          named_bar_sync(1, 256);
          if (store_thread) {
            tma_store_3d(&p.tmC, smem_pool + ((local >> 1) & 1) * Cfg::POOL_BYTES, 0, 0, n * p.POH + (oy0 >> 1));
            bulk_commit();
          }
        }
        continue;
      }
      if constexpr (kPool == 1) {
        if (p.RT == 1) {
          // ---- one conv row per row group (P = 128, the 224 x 224 geometry): pool in registers.  A thread owns virtual
          // pixel v of all four conv rows of the tile (32 channels: `set` picks the half), so the vertical 3-max is a
          // register max against the row carried from the previous tile, and the horizontal one two warp shuffles (the
          // left neighbour of lane 0 comes from the previous warp through 128 bytes of shared memory).  Only the pooled
          // tile is staged: the MMA operand reads keep the shared-memory bandwidth this kernel is bound by.
          const bool warm = tile < t_begin;
          uint32_t va[16], vb[16];   // vertical maxima: rows (-1, 0, 1) -> pooled row oy0 / 2, rows (1, 2, 3) -> the next
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (warm && g < 3) continue;   // a warm-up tile only provides the carried row
            uint32_t acc[32], cur[16];
            tmem_ld_32x32b_x32(tmem_base + (uint32_t(q * 32) << 16) + (b * RG + g) * 64 + set * 32, acc);
            tmem_ld_wait();
            if (g == 3) {
              tcgen05_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tmem_empty[b]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int c0 = set * 32 + j * 8;
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + c0));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + 4));
              float f[8] = {__uint_as_float(acc[j * 8 + 0]) + b0.x, __uint_as_float(acc[j * 8 + 1]) + b0.y,
                            __uint_as_float(acc[j * 8 + 2]) + b0.z, __uint_as_float(acc[j * 8 + 3]) + b0.w,
                            __uint_as_float(acc[j * 8 + 4]) + b1.x, __uint_as_float(acc[j * 8 + 5]) + b1.y,
                            __uint_as_float(acc[j * 8 + 6]) + b1.z, __uint_as_float(acc[j * 8 + 7]) + b1.w};
              if (p.relu) {
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
              }
              const uint4 o = pack8<T>(f);
              cur[j * 4 + 0] = o.x; cur[j * 4 + 1] = o.y; cur[j * 4 + 2] = o.z; cur[j * 4 + 3] = o.w;
            }
            const bool valid = oy0 + g < p.OH;   // rows below the image never win (row 0 and, when used, row 2 are valid)
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              if (g == 0) va[k] = oy0 > 0 ? strip_max2<T>(carry_reg[k], cur[k]) : cur[k];
              if (g == 1) { va[k] = valid ? strip_max2<T>(va[k], cur[k]) : va[k]; vb[k] = cur[k]; }
              if (g == 2) vb[k] = strip_max2<T>(vb[k], cur[k]);
              if (g == 3) { vb[k] = valid ? strip_max2<T>(vb[k], cur[k]) : vb[k]; carry_reg[k] = cur[k]; }
            }
          }
          if (!warm) {
            const uint32_t xch = smem_u32(smem_carry) + (uint32_t)(((local & 1) * 8 + set * 4 + q) * 128);
            if (lane == 31) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                sts128(xch + k * 16, make_uint4(va[4 * k], va[4 * k + 1], va[4 * k + 2], va[4 * k + 3]));
                sts128(xch + 64 + k * 16, make_uint4(vb[4 * k], vb[4 * k + 1], vb[4 * k + 2], vb[4 * k + 3]));
              }
            }
            named_bar_sync(2 + set, 128);
            uint32_t la[16], lb[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              la[k] = __shfl_up_sync(0xffffffffu, va[k], 1);
              lb[k] = __shfl_up_sync(0xffffffffu, vb[k], 1);
            }
            if (lane == 0 && q > 0) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint4 xa = lds128(xch - 128 + k * 16), xb = lds128(xch - 128 + 64 + k * 16);
                la[4 * k] = xa.x; la[4 * k + 1] = xa.y; la[4 * k + 2] = xa.z; la[4 * k + 3] = xa.w;
                lb[4 * k] = xb.x; lb[4 * k + 1] = xb.y; lb[4 * k + 2] = xb.z; lb[4 * k + 3] = xb.w;
              }
            }
            const bool lv = v >= 1, rv = v + 1 < p.OW;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              const uint32_t ra = __shfl_down_sync(0xffffffffu, va[k], 1), rb = __shfl_down_sync(0xffffffffu, vb[k], 1);
              uint32_t ma = va[k], mb = vb[k];
              if (lv) { ma = strip_max2<T>(ma, la[k]); mb = strip_max2<T>(mb, lb[k]); }
              if (rv) { ma = strip_max2<T>(ma, ra); mb = strip_max2<T>(mb, rb); }
              va[k] = ma; vb[k] = mb;
            }
            if ((v & 1) == 0 && v < p.OW) {   // v is the centre of pooled column v / 2
              const uint32_t pool_addr = smem_u32(smem_pool + (local & 1) * Cfg::POOL_BYTES);
              const uint32_t prow_a = (uint32_t)(v >> 1), prow_b = 64u + (uint32_t)(v >> 1);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                sts128(pool_addr + prow_a * 128 + ((uint32_t)((set * 4 + j) ^ (prow_a & 7)) << 4),
                       make_uint4(va[4 * j], va[4 * j + 1], va[4 * j + 2], va[4 * j + 3]));
                sts128(pool_addr + prow_b * 128 + ((uint32_t)((set * 4 + j) ^ (prow_b & 7)) << 4),
                       make_uint4(vb[4 * j], vb[4 * j + 1], vb[4 * j + 2], vb[4 * j + 3]));
              }
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(1, 256);
          if (store_thread && !warm) {
            const int pr0 = oy0 >> 1;
            for (int r = 0; r < 2; ++r)
              if (pr0 + r < p.POH) tma_store_3d(&p.tmC, smem_pool + (local & 1) * Cfg::POOL_BYTES + r * 64 * 128, 0, 0, n * p.POH + pr0 + r);
            bulk_commit();
          }
          continue;
        }
      }
      const int g_lo = RG >= 2 ? set * (RG / 2) : 0, g_hi = RG >= 2 ? g_lo + RG / 2 : 1;
      const int u_lo = RG >= 2 ? 0 : set, u_hi = RG >= 2 ? 2 : set + 1;
#pragma unroll 1
      for (int g = g_lo; g < g_hi; ++g) {
        const uint32_t row_addr = smem_u32(cbuf + g * 128 * C_ROW) + v * C_ROW;
#pragma unroll 1
        for (int u = u_lo; u < u_hi; ++u) {
          uint32_t acc[32];
          const bool live = u * 32 < NC;   // 32 output channels: the second column unit does not exist
          if (live) {
            tmem_ld_32x32b_x32(tmem_base + (uint32_t(q * 32) << 16) + (b * RG + g) * NC + u * 32, acc);
            tmem_ld_wait();
          }
          if (g == g_hi - 1 && u == u_hi - 1) {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[b]);
          }
          if (!live) continue;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c0 = u * 32 + j * 8;
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + c0));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + 4));
            float f[8] = {__uint_as_float(acc[j * 8 + 0]) + b0.x, __uint_as_float(acc[j * 8 + 1]) + b0.y,
                          __uint_as_float(acc[j * 8 + 2]) + b0.z, __uint_as_float(acc[j * 8 + 3]) + b0.w,
                          __uint_as_float(acc[j * 8 + 4]) + b1.x, __uint_as_float(acc[j * 8 + 5]) + b1.y,
                          __uint_as_float(acc[j * 8 + 6]) + b1.z, __uint_as_float(acc[j * 8 + 7]) + b1.w};
            if (p.relu) {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
            }
            const uint4 o = pack8<T>(f);
            const uint32_t addr = row_addr + ((uint32_t)((u * 4 + j) ^ (C_ROW == 128 ? (v & 7) : ((v >> 1) & 3))) << 4);
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
          }
        }
      }
      if constexpr (kPool == 1) {
        // ---- 3x3 / 2 max pool over the staged conv rows (+ the carried row above), torch semantics: padding never wins
        named_bar_sync(1, 256);   // all conv rows of the tile are staged
        const int R = RG * p.RT, tid = threadIdx.x - 64;
        const uint32_t c_addr = smem_u32(cbuf);
        const uint32_t carry_rd = smem_u32(smem_carry + (local & 1) * Cfg::CARRY_BYTES);
        const uint32_t carry_wr = smem_u32(smem_carry + ((local + 1) & 1) * Cfg::CARRY_BYTES);
        const uint32_t pool_addr = smem_u32(smem_pool + (local & 1) * Cfg::POOL_BYTES);
        const bool warm = tile < t_begin;
        if (!warm) {
          // Invalid taps (padding, rows below the image) are replaced by the window centre, which is always valid: max is
          // idempotent, so the nine loads are unconditional and can all be in flight before the first comparison.
          const int half_p = p.P >> 1, items = (R >> 1) * p.POW * 8;
          const uint32_t j = tid & 7;
          int prl = pool_prl0, pc = pool_pc0;     // pooled row within the tile / pooled column of this thread's item
          for (int i = tid; i < items; i += 256) {
            const int r0 = 2 * prl, c0 = 2 * pc;
            const int rm = oy0 + r0 > 0 ? r0 - 1 : r0, rp = oy0 + r0 + 1 < p.OH ? r0 + 1 : r0;
            const int cm = c0 > 0 ? c0 - 1 : c0, cp = c0 + 1 < p.OW ? c0 + 1 : c0;
            const uint32_t row_a[3] = {rm < 0 ? carry_rd : c_addr + (uint32_t)(rm * p.P) * 128, c_addr + (uint32_t)(r0 * p.P) * 128,
                                       c_addr + (uint32_t)(rp * p.P) * 128};
            const uint32_t col_o[3] = {(uint32_t)cm * 128 + ((j ^ ((uint32_t)cm & 7)) << 4), (uint32_t)c0 * 128 + ((j ^ ((uint32_t)c0 & 7)) << 4),
                                       (uint32_t)cp * 128 + ((j ^ ((uint32_t)cp & 7)) << 4)};
            uint4 v[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) v[k] = lds128(row_a[k / 3] + col_o[k % 3]);
            uint4 m = strip_max8<T>(strip_max8<T>(strip_max8<T>(v[0], v[1]), strip_max8<T>(v[2], v[3])),
                                    strip_max8<T>(strip_max8<T>(v[4], v[5]), strip_max8<T>(v[6], v[7])));
            m = strip_max8<T>(m, v[8]);
            const uint32_t prow = (uint32_t)(prl * half_p + pc);
            sts128(pool_addr + prow * 128 + ((j ^ (prow & 7)) << 4), m);
            pc += 32;
            while (pc >= p.POW) { pc -= p.POW; ++prl; }
          }
        }
        // the tile's last conv row becomes the next tile's row above (other buffer: this tile's pooling may still read its own)
        for (int i = tid; i < p.P * 8; i += 256) {
          const uint32_t px = i >> 3, off = ((uint32_t)((i & 7) ^ (px & 7)) << 4);
          sts128(carry_wr + px * 128 + off, lds128(c_addr + ((uint32_t)(R - 1) * p.P + px) * 128 + off));
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 256);
        if (store_thread && !warm) {
          const int pr0 = oy0 >> 1;
          for (int r = 0; r < (R >> 1); ++r)
            if (pr0 + r < p.POH) tma_store_3d(&p.tmC, smem_pool + (local & 1) * Cfg::POOL_BYTES + r * (p.P >> 1) * 128, 0, 0, n * p.POH + pr0 + r);
          bulk_commit();
        }
      } else {
      fence_proxy_async_smem();
      named_bar_sync(1, 256);
      if (store_thread) {
        for (int oy = 0; oy < RG * p.RT; ++oy)
          if (oy0 + oy < p.OH) tma_store_3d(&p.tmC, cbuf + oy * p.P * C_ROW, 0, ox0, n * p.OH + oy0 + oy);
        bulk_commit();
      }
      }
    }
    if (store_thread) bulk_wait<0>();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn4)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool strip_is_3x3(const ConvShape& s) { return (s.cin == 64 || s.cin == 32) && s.kh == 3 && s.kw == 3 && s.pad == 1 && s.pad_after() == 1; }
static bool strip_is_rowwin(const ConvShape& s) { return s.cin == 64 && s.kw == 1 && s.kh >= 2 && s.kh <= 4 && s.pad == 0 && s.pad_after() == 0; }
// space-to-depth stems: 7x7/2 pad 3 -> 4x4 (pad 2 | 1), 3x3/2 pad 1 -> 2x2 (pad 1 | 0)
static bool strip_is_s2d16(const ConvShape& s) {
  return s.cin == 16 && ((s.kh == 4 && s.kw == 4 && s.pad == 2 && s.pad_after() == 1) || (s.kh == 2 && s.kw == 2 && s.pad == 1 && s.pad_after() == 0));
}
bool conv_strip_supported(const ConvShape& s, int precision) {
  // 32 output channels (CLIP stem conv1 / conv2) only where the input rows are narrower than 128 bytes
  const bool cout_ok = s.cout == 64 || (s.cout == 32 && (s.cin == 32 || s.cin == 16));
  return (precision == SEMDIFF_BF16 || precision == SEMDIFF_FP16) && (strip_is_3x3(s) || strip_is_rowwin(s) || strip_is_s2d16(s)) &&
         s.stride == 1 && cout_ok && s.cin2 == 0 && s.W >= 6 && s.OH() >= 1 && s.OW() >= 1 &&
         (int64_t)s.n_img * s.H * s.W < ((int64_t)1 << 31);
}

static int strip_prepare(ConvTcLaunch* L, const ConvPtrs& q, const ConvShape& s, int precision, int pooled);
int conv_strip_prepare(ConvTcLaunch* L, const ConvPtrs& q, const ConvShape& s, int precision) {
  return strip_prepare(L, q, s, precision, 0);
}
// 3x3 pad-1 64 -> 64 conv followed by avg_pool2d(2): one conv row per tile (P = 128), row pairs inside one CTA
bool conv_strip_avgpool_supported(const ConvShape& s, int precision) {
  return conv_strip_supported(s, precision) && strip_is_3x3(s) && s.cout == 64 && s.OH() % 2 == 0 && s.OW() >= 62 && s.OW() <= 126;
}
int conv_strip_avgpool_prepare(ConvTcLaunch* L, const ConvPtrs& q, const ConvShape& s, int precision) {
  if (!conv_strip_avgpool_supported(s, precision)) { set_error("conv_strip_avgpool: unsupported shape"); return SEMDIFF_ERR_UNSUPPORTED; }
  return strip_prepare(L, q, s, precision, 2);
}
// stem conv over SEMDIFF_INPUT_S2D16 followed by max_pool2d(3, 2, 1): one strip (no column blocks) per image row
bool conv_strip_pool_supported(const ConvShape& s, int precision) {
  return conv_strip_supported(s, precision) && strip_is_s2d16(s) && s.cout == 64 && s.OW() <= 128 - (s.kw - 1);
}
int conv_strip_pool_prepare(ConvTcLaunch* L, const ConvPtrs& q, const ConvShape& s, int precision) {
  if (!conv_strip_pool_supported(s, precision)) { set_error("conv_strip_pool: unsupported shape"); return SEMDIFF_ERR_UNSUPPORTED; }
  return strip_prepare(L, q, s, precision, 1);
}

static int strip_prepare(ConvTcLaunch* L, const ConvPtrs& q, const ConvShape& s, int precision, int pooled) {
  if (!conv_strip_supported(s, precision) || q.res != nullptr) { set_error("conv_strip: unsupported shape"); return SEMDIFF_ERR_UNSUPPORTED; }
  static EncodeTiledFn4 enc = nullptr;
  if (enc == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      enc = reinterpret_cast<EncodeTiledFn4>(ptr);
  }
  if (enc == nullptr) { set_error("cuTensorMapEncodeTiled entry point not found"); return SEMDIFF_ERR_CUDA; }
  StripParams& p = *reinterpret_cast<StripParams*>(L->params);
  memset(&p, 0, sizeof(p));
  const int RG = strip_is_3x3(s) ? 1 : (strip_is_s2d16(s) ? 4 : 2);
  const int ROWB = s.cin * 2;
  // Strip geometry: P pixels per strip row (power of two), RT = 128 / P output rows per row group, column blocks of
  // CW <= P - (kw - 1) output pixels.  Pick the P that wastes the fewest of the 128 "virtual" pixels per tile
  // (ties -> the larger P: fewer, longer TMA rows).  224x224: one block per row; 1024x1024 layer1: P = 32, 9 blocks.
  int P = 0;
  double best = -1.0;
  for (int cand = 128; cand >= 16; cand >>= 1) {
    const int cw_max = cand - (s.kw - 1);
    if (cw_max < 1) continue;
    if (pooled && cw_max < s.OW()) continue;   // the pooled epilogue needs whole image rows in one strip
    if (pooled == 2 && cand != 128) continue;  // register avg pool: one conv row per tile
    const int rt = 128 / cand;
    const int64_t blocks = (s.OW() + cw_max - 1) / cw_max;
    const int64_t tiles = blocks * ((s.OH() + rt * RG - 1) / (rt * RG));
    const double eff = (double)s.OW() * s.OH() / ((double)tiles * 128 * RG);
    if (eff > best + 1e-9) { best = eff; P = cand; }
  }
  const int cw_max = P - (s.kw - 1);
  p.col_blocks = (s.OW() + cw_max - 1) / cw_max;
  p.CW = (s.OW() + p.col_blocks - 1) / p.col_blocks;  // balanced blocks
  p.P = P; p.RT = 128 / P; p.H = s.H; p.W = s.W; p.OH = s.OH(); p.OW = s.OW(); p.KH = s.kh; p.KW = s.kw; p.pad = s.pad;
  p.n_img = s.n_img; p.relu = s.relu; p.bias = q.bias;
  p.tiles_per_img = ((p.OH + p.RT * RG - 1) / (p.RT * RG)) * p.col_blocks;
  const CUtensorMapDataType dt = precision == SEMDIFF_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  {
    const CUtensorMapSwizzle swz = ROWB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (ROWB == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    const cuuint64_t dims[4] = {(cuuint64_t)s.cin, (cuuint64_t)s.W, (cuuint64_t)s.H, (cuuint64_t)s.n_img};
    const cuuint64_t strides[3] = {(cuuint64_t)ROWB, (cuuint64_t)s.W * ROWB, (cuuint64_t)s.H * s.W * ROWB};
    const cuuint32_t box[4] = {(cuuint32_t)s.cin, (cuuint32_t)P, (cuuint32_t)(RG * p.RT + s.kh - 1), 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&p.tmX, dt, 4, const_cast<void*>(q.in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv_strip: strip tensor map failed (%d) W=%d H=%d P=%d", (int)r, s.W, s.H, P); return SEMDIFF_ERR_CUDA; }
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)s.K(), (cuuint64_t)s.cout};
    const cuuint64_t strides[1] = {(cuuint64_t)s.K() * 2};
    const cuuint32_t box[2] = {(cuuint32_t)s.cin, (cuuint32_t)s.cout};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&p.tmB, dt, 2, const_cast<void*>(q.w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     ROWB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (ROWB == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv_strip: weight tensor map failed (%d)", (int)r); return SEMDIFF_ERR_CUDA; }
  }
  if (pooled && p.col_blocks != 1) { set_error("conv_strip_pool: image wider than one strip"); return SEMDIFF_ERR_UNSUPPORTED; }
  p.POH = pooled == 2 ? p.OH / 2 : (p.OH - 1) / 2 + 1;
  p.POW = pooled == 2 ? p.OW / 2 : (p.OW - 1) / 2 + 1;
  {
    const int ow = pooled ? p.POW : p.OW, oh = pooled ? p.POH : p.OH;
    const cuuint64_t crow = (cuuint64_t)s.cout * 2;
    const cuuint64_t dims[3] = {(cuuint64_t)s.cout, (cuuint64_t)ow, (cuuint64_t)s.n_img * oh};
    const cuuint64_t strides[2] = {crow, (cuuint64_t)ow * crow};
    const cuuint32_t box[3] = {(cuuint32_t)s.cout, (cuuint32_t)(pooled ? p.POW : p.CW), 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&p.tmC, dt, 3, q.out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     s.cout == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv_strip: output tensor map failed (%d)", (int)r); return SEMDIFF_ERR_CUDA; }
  }
  // 101 / 102: 64-channel strips, RG = 1 / 2; 103: 16-channel; 104: 16-channel + max pool; 105: 3x3 + 2x2 average pool;
  // 106-108: 32-channel input rows (3x3: 32 -> 32, 32 -> 64, 32 -> 64 + average pool); 109: 16-channel, 32 outputs
  int mode = pooled == 2 ? 105 : pooled ? 104 : (ROWB == 32 ? 103 : 100 + RG);
  if (ROWB == 64) mode = s.cout == 32 ? 106 : (pooled == 2 ? 108 : 107);
  if (ROWB == 32 && s.cout == 32) mode = 109;
  L->block_n = s.cout; L->a_mode = mode; L->precision = precision;
  return 0;
}

template <typename T, int RG, int ROWB, int kPool = 0, int NC = 64>
static int strip_launch_t(const StripParams& p, int dev, int sms, cudaStream_t st) {
  static bool configured[64] = {};
  auto kern = conv3x3_strip_kernel<T, RG, ROWB, kPool, NC>;
  if (!configured[dev]) {
    SEMDIFF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, StripCfg<RG, ROWB, kPool, NC>::SMEM));
    configured[dev] = true;
  }
  // one CTA per SM, never more CTAs than work units (kPool == 2 hands out row PAIRS): a CTA with an empty range would
  // exit with its resident-weight TMA load still in flight
  const int units = kPool == 2 ? (p.n_img * p.tiles_per_img) / 2 : p.n_img * p.tiles_per_img;
  SEMDIFF_CUDA_OK(launch_pdl(kern, dim3(units < sms ? units : sms), dim3(320), StripCfg<RG, ROWB, kPool, NC>::SMEM, st, p));
  return 0;
}

int conv_strip_launch(const ConvTcLaunch* L, cudaStream_t st) {
  const StripParams& p = *reinterpret_cast<const StripParams*>(L->params);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (dev < 0 || dev >= 64) dev = 0;
  const bool bf = L->precision == SEMDIFF_BF16;
  switch (L->a_mode) {
    case 109: return bf ? strip_launch_t<__nv_bfloat16, 4, 32, 0, 32>(p, dev, sms, st) : strip_launch_t<__half, 4, 32, 0, 32>(p, dev, sms, st);
    case 108: return bf ? strip_launch_t<__nv_bfloat16, 1, 64, 2>(p, dev, sms, st) : strip_launch_t<__half, 1, 64, 2>(p, dev, sms, st);
    case 107: return bf ? strip_launch_t<__nv_bfloat16, 1, 64>(p, dev, sms, st) : strip_launch_t<__half, 1, 64>(p, dev, sms, st);
    case 106: return bf ? strip_launch_t<__nv_bfloat16, 1, 64, 0, 32>(p, dev, sms, st) : strip_launch_t<__half, 1, 64, 0, 32>(p, dev, sms, st);
    case 105: return bf ? strip_launch_t<__nv_bfloat16, 1, 128, 2>(p, dev, sms, st) : strip_launch_t<__half, 1, 128, 2>(p, dev, sms, st);
    case 104: return bf ? strip_launch_t<__nv_bfloat16, 4, 32, 1>(p, dev, sms, st) : strip_launch_t<__half, 4, 32, 1>(p, dev, sms, st);
    case 103: return bf ? strip_launch_t<__nv_bfloat16, 4, 32>(p, dev, sms, st) : strip_launch_t<__half, 4, 32>(p, dev, sms, st);
    case 102: return bf ? strip_launch_t<__nv_bfloat16, 2, 128>(p, dev, sms, st) : strip_launch_t<__half, 2, 128>(p, dev, sms, st);
  }
  return bf ? strip_launch_t<__nv_bfloat16, 1, 128>(p, dev, sms, st) : strip_launch_t<__half, 1, 128>(p, dev, sms, st);
}

}  // namespace semdiff
