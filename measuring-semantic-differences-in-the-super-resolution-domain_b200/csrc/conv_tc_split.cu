// Split-precision ("x3") tcgen05 implicit-GEMM convolution for sm_100a: SEMDIFF_FP16X3 / SEMDIFF_BF16X3.
//
// Every activation and weight is the unevaluated sum hi + lo of two 16-bit numbers (include/semdiff_b200.h), stored per
// block of 64 channels as [64 hi | 64 lo].  Per 64-wide K block the accumulator receives THREE tensor-core products
//     A_lo * W_hi  +  A_hi * W_lo  +  A_hi * W_hi          (A_lo * W_lo is below fp32 resolution and dropped)
// which restores fp32-level operands on the 16-bit tensor pipe - the trunk precision the reference's fp32 path
// (/root/reference/models/global_eval_models.py:364,371 under torch fp32) needs on SR ~ GT pairs, whose feature difference
// (:379) is far below 16-bit resolution.  Same output contract as conv_tc.cu: out = act(conv(in) + bias (+ residual)).
//
// Two measured properties of tcgen05.mma shape the kernel (tools/x3_accuracy.py, profiles/r2_x3_accuracy.txt):
//  * every MMA output is TRUNCATED to fp32 (error biased towards zero, growing linearly with the number of MMAs that
//    touch a large accumulator): one long accumulation over K = 4608 is off by -1.2e-5 relative, and through 50 layers
//    that systematic shrink costs 1e-4 of the score.  So (1) within a K block the two small products are issued FIRST -
//    while the accumulator only holds 2^-11-sized terms their truncation is free - and (2) the accumulation is PROMOTED:
//    every `chunk_kb` K blocks (default 1 = 12 MMAs) the chunk sum is drained from TMEM and added in registers with
//    round-to-nearest by the epilogue warps, while the next chunks accumulate from zero in the other TMEM stages.
//    Residual bias per conv: -9e-8; scores: <= 7e-6 of the fp32 oracle (tests/test_scorer_gpu.py).
//  * an N = 128 SS-mode MMA reads 8 KB of operands per 64 tensor cycles = the whole shared-memory bandwidth, so the wide
//    layers use 256-column tiles (12 KB per 128 cycles) and a ring fill is ONE K block = four tiles (A hi, A lo, W hi,
//    W lo) feeding twelve MMAs: 2/3 of the fill traffic of three separate (A, W) pairs.
//
// CTA = one 128 x BLOCK_N tile at a time, persistent, warp-specialised like conv_tc.cu:
//   warp 0     TMA producer (four tile loads per ring fill; tiled 2-D or im2col mode for A)
//   warp 1     tcgen05.mma issuer; chunk c accumulates into TMEM stage c % ACC_STAGES (4 stages for tiles up to 128 wide)
//   warps 2-9  drain + epilogue: warp (q, half) owns TMEM lanes 32q..32q+31 and, of every 64-column group, the 32 columns
//              of its half; it sums the chunk accumulators in registers, then + bias (+ residual) -> ReLU -> split into
//              hi / lo -> the group's two swizzled 64-column boxes
//   warp 10    C-ring I/O: TMA store of staged groups, slot grants + residual (hi and lo box) prefetch
#include <cuda.h>
#include <stdlib.h>

#include "conv_tc.h"

namespace semdiff {

template <int BLOCK_N> struct SplitCfg {
  static constexpr int B_TILE_BYTES = BLOCK_N * BLOCK_K * 2;                      // one weight tile (hi or lo)
  static constexpr int STAGE_BYTES = 2 * A_STAGE_BYTES + 2 * B_TILE_BYTES;        // A hi | A lo | W hi | W lo
  static constexpr int GROUPS = BLOCK_N / 64;            // output groups: 64 accumulator columns = [64 hi | 64 lo] stored
  static constexpr int BOX_BYTES = BLOCK_M * 64 * 2;     // one 64-column 16-bit box of 128 rows
  static constexpr int GROUP_BYTES = 2 * BOX_BYTES;
  // The operand ring (p.stages fills) and the C ring (p.ring output groups) share 224 KB; the host picks the split per conv
  // (split_ring_config): a deep operand ring hides the TMA latency of the im2col / short-K tiles, a deep C ring keeps the
  // residual prefetch of the HBM-bound residual convs ahead.
  static constexpr int MAX_STAGES = BLOCK_N == 64 ? 4 : (BLOCK_N == 128 ? 3 : 2);
  static constexpr int MAX_RING = 3;
  static constexpr int DATA_BYTES = 224 * 1024;
  static constexpr int ACC_STAGES = BLOCK_N == 256 ? 2 : 4;   // chunk accumulators in flight: the MMA warp runs this far ahead of the drain
  static constexpr int TMEM_COLS = ACC_STAGES * BLOCK_N;      // 256 / 512 / 512 columns
  static constexpr int NUM_BARS = 2 * MAX_STAGES + 2 * ACC_STAGES + 2 * MAX_RING;
  static constexpr int SMEM_BYTES = DATA_BYTES + NUM_BARS * 8 + 16 + 1024;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

constexpr int SPLIT_THREADS = 11 * 32;

template <typename T, int BLOCK_N, int kAMode>
__global__ void __launch_bounds__(SPLIT_THREADS, 1) conv_tc_split_kernel(const __grid_constant__ ConvTcParams p) {
  using Cfg = SplitCfg<BLOCK_N>;
  constexpr int GROUPS = Cfg::GROUPS, EPI_WARPS = 8, ACC = Cfg::ACC_STAGES;
  const int STAGES = p.stages, RING = p.ring;   // stages * STAGE_BYTES + ring * GROUP_BYTES <= DATA_BYTES (host-checked)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_c = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::DATA_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::MAX_STAGES;
  uint64_t* tmem_full_bar = empty_bar + Cfg::MAX_STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + ACC;
  uint64_t* res_full_bar = tmem_empty_bar + ACC;
  uint64_t* staged_bar = res_full_bar + Cfg::MAX_RING;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(staged_bar + Cfg::MAX_RING);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;
  const bool leader = elect_one();

  if (warp == 0 && leader) {
    tma_prefetch_desc(&p.tmA);
    if (p.has_src2) tma_prefetch_desc(&p.tmA2);
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmC);
    if (p.has_res) tma_prefetch_desc(&p.tmR);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < ACC; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], EPI_WARPS);
    }
    for (int i = 0; i < RING; ++i) {
      mbar_init(&res_full_bar[i], 1);
      mbar_init(&staged_bar[i], EPI_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_trigger();

  if (warp == 0) {
    // ===================== TMA producer: one ring fill = one K block = A hi, A lo, W hi, W lo =====================
    if (leader) {
      int stage = 0, phase = 0;
      const int kb_per_tap = p.Cin >> 6;
      pdl_wait();
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_tile = tile / p.n_tiles, n_tile = tile - m_tile * p.n_tiles;
        int n = 0, h0 = 0, w0 = 0, h2 = 0, w2 = 0;
        if (kAMode == A_IM2COL || p.a2_im2col) {
          const int gm = m_tile * BLOCK_M;
          n = gm / (p.OH * p.OW);
          const int r = gm - n * p.OH * p.OW;
          const int oh = r / p.OW, ow = r - oh * p.OW;
          h0 = oh * p.stride - p.pad;
          w0 = ow * p.stride - p.pad;
          h2 = oh * p.stride2;
          w2 = ow * p.stride2;
        }
        int tap = 0, cb = 0;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          uint8_t* a_hi = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* a_lo = a_hi + A_STAGE_BYTES;
          uint8_t* w_hi = a_lo + A_STAGE_BYTES;
          uint8_t* w_lo = w_hi + Cfg::B_TILE_BYTES;
          if (kb >= p.num_kb1) {
            const int c2 = 2 * (kb - p.num_kb1) * BLOCK_K;   // fused 1x1 conv over the second activation tensor
            if (p.a2_im2col) {
              tma_load_im2col_4d(&p.tmA2, &full_bar[stage], a_hi, c2, w2, h2, n, 0, 0);
              tma_load_im2col_4d(&p.tmA2, &full_bar[stage], a_lo, c2 + BLOCK_K, w2, h2, n, 0, 0);
            } else {
              tma_load_2d(&p.tmA2, &full_bar[stage], a_hi, c2, m_tile * BLOCK_M);
              tma_load_2d(&p.tmA2, &full_bar[stage], a_lo, c2 + BLOCK_K, m_tile * BLOCK_M);
            }
          } else if (kAMode == A_TMA) {
            tma_load_2d(&p.tmA, &full_bar[stage], a_hi, 2 * kb * BLOCK_K, m_tile * BLOCK_M);
            tma_load_2d(&p.tmA, &full_bar[stage], a_lo, (2 * kb + 1) * BLOCK_K, m_tile * BLOCK_M);
          } else {
            const int r = tap / p.KW, s = tap - r * p.KW;
            tma_load_im2col_4d(&p.tmA, &full_bar[stage], a_hi, 2 * cb * BLOCK_K, w0, h0, n, (uint16_t)s, (uint16_t)r);
            tma_load_im2col_4d(&p.tmA, &full_bar[stage], a_lo, (2 * cb + 1) * BLOCK_K, w0, h0, n, (uint16_t)s, (uint16_t)r);
            if (++cb == kb_per_tap) { cb = 0; ++tap; }
          }
          tma_load_2d(&p.tmB, &full_bar[stage], w_hi, 2 * kb * BLOCK_K, n_tile * BLOCK_N);
          tma_load_2d(&p.tmB, &full_bar[stage], w_lo, (2 * kb + 1) * BLOCK_K, n_tile * BLOCK_N);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: 12 MMAs per fill, small products first; chunk c -> TMEM stage c % ACC =====================
    constexpr uint32_t idesc = umma_idesc_f16(Elem<T>::kUmmaFormat, BLOCK_M, BLOCK_N);
    const uint64_t desc0 = umma_smem_desc_sw128(smem_u32(smem));
    int stage = 0, phase = 0, chunk = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < p.num_kb;) {
        const int acc = chunk % ACC, acc_phase = (chunk / ACC) & 1;
        ++chunk;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        const int kend = kb + p.chunk_kb < p.num_kb ? kb + p.chunk_kb : p.num_kb;
        for (int j = 0; kb < kend; ++kb, ++j) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          if (leader) {
            const uint64_t a_hi = desc0 + (uint64_t)((stage * Cfg::STAGE_BYTES) >> 4);
            const uint64_t a_lo = a_hi + (uint64_t)(A_STAGE_BYTES >> 4);
            const uint64_t w_hi = a_lo + (uint64_t)(A_STAGE_BYTES >> 4);
            const uint64_t w_lo = w_hi + (uint64_t)(Cfg::B_TILE_BYTES >> 4);
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k) umma_f16_ss(tmem_d, a_lo + (uint64_t)(k * 2), w_hi + (uint64_t)(k * 2), idesc, (j | k) != 0 ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k) umma_f16_ss(tmem_d, a_hi + (uint64_t)(k * 2), w_lo + (uint64_t)(k * 2), idesc, 1u);
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k) umma_f16_ss(tmem_d, a_hi + (uint64_t)(k * 2), w_hi + (uint64_t)(k * 2), idesc, 1u);
            umma_commit(&empty_bar[stage]);
            if (kb == kend - 1) umma_commit(&tmem_full_bar[acc]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp < 2 + EPI_WARPS) {
    // ===================== drain (promoted accumulation) + epilogue from registers =====================
    const int q = warp & 3;                 // TMEM lane quarter this warp may touch
    const int half = (warp - 2) >> 2;       // which 32 columns of every 64-column group
    const int row = q * 32 + lane;
    const int n_chunks = (p.num_kb + p.chunk_kb - 1) / p.chunk_kb;
    const float sc = p.acc_scale;           // a power of two: exact
    int chunk = 0, slot = 0, sphase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.n_tiles, n_tile = tile - m_tile * p.n_tiles;
      float racc[GROUPS][32];
      for (int c = 0; c < n_chunks; ++c, ++chunk) {
        const int acc = chunk % ACC;
        if (c == 0) mbar_wait_backoff(&tmem_full_bar[acc], (chunk / ACC) & 1);
        else mbar_wait_short(&tmem_full_bar[acc], (chunk / ACC) & 1);
        tcgen05_fence_after();
#pragma unroll
        for (int g = 0; g < GROUPS; ++g) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + (uint32_t(q * 32) << 16) + acc * BLOCK_N + g * 64 + half * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) racc[g][i] = c == 0 ? __uint_as_float(v[i]) : racc[g][i] + __uint_as_float(v[i]);
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      }
#pragma unroll
      for (int g = 0; g < GROUPS; ++g) {
        uint8_t* cbuf = smem_c + slot * Cfg::GROUP_BYTES;
        mbar_wait_short(&res_full_bar[slot], sphase);  // the slot is ours (and the residual boxes, if any, have landed)
        const int col0 = n_tile * BLOCK_N + g * 64 + half * 32;
        const uint32_t row_addr = smem_u32(cbuf) + row * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j * 8));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j * 8 + 4));
          float f[8] = {fmaf(racc[g][j * 8 + 0], sc, b0.x), fmaf(racc[g][j * 8 + 1], sc, b0.y), fmaf(racc[g][j * 8 + 2], sc, b0.z),
                        fmaf(racc[g][j * 8 + 3], sc, b0.w), fmaf(racc[g][j * 8 + 4], sc, b1.x), fmaf(racc[g][j * 8 + 5], sc, b1.y),
                        fmaf(racc[g][j * 8 + 6], sc, b1.z), fmaf(racc[g][j * 8 + 7], sc, b1.w)};
          // chunks half*4 .. half*4+3 of the group's hi box and, one box further, of its lo box
          const uint32_t a_hi = row_addr + (swz_chunk<128>(half * 4 + j, row) << 4), a_lo = a_hi + Cfg::BOX_BYTES;
          if (p.has_res) {
            uint4 rh, rl;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(rh.x), "=r"(rh.y), "=r"(rh.z), "=r"(rh.w) : "r"(a_hi));
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(rl.x), "=r"(rl.y), "=r"(rl.z), "=r"(rl.w) : "r"(a_lo));
            float h[8], l[8];
            unpack8<T>(rh, h);
            unpack8<T>(rl, l);
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] += h[e] + l[e];
          }
          if (p.relu) {
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
          }
          const uint4 oh = pack8<T>(f);
          float h[8];
          unpack8<T>(oh, h);
#pragma unroll
          for (int e = 0; e < 8; ++e) h[e] = f[e] - h[e];   // exact: hi is f rounded to fewer bits
          const uint4 ol = pack8<T>(h);
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a_hi), "r"(oh.x), "r"(oh.y), "r"(oh.z), "r"(oh.w) : "memory");
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a_lo), "r"(ol.x), "r"(ol.y), "r"(ol.z), "r"(ol.w) : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&staged_bar[slot]);
        if (++slot == RING) { slot = 0; sphase ^= 1; }
      }
    }
  } else {
    // ===================== C-ring I/O: TMA stores of staged groups + slot grants / residual prefetch =====================
    if (leader && blockIdx.x < total_tiles) {
      pdl_wait();
      const int total = GROUPS * ((total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x);
      int k_load = 0, l_slot = 0, l_g = 0, l_tile = blockIdx.x;
      int k_store = 0, s_slot = 0, s_phase = 0, s_g = 0, s_tile = blockIdx.x;
      while (k_store < total) {
        while (k_load < total && k_load < k_store + RING) {
          if (p.has_res) {
            const int m_tile = l_tile / p.n_tiles, n_tile = l_tile - m_tile * p.n_tiles;
            mbar_arrive_expect_tx(&res_full_bar[l_slot], Cfg::GROUP_BYTES);
            uint8_t* cbuf = smem_c + l_slot * Cfg::GROUP_BYTES;
            const int c0 = 2 * (n_tile * BLOCK_N + l_g * 64);      // stored column of the group's hi box
            tma_load_2d(&p.tmR, &res_full_bar[l_slot], cbuf, c0, m_tile * BLOCK_M);
            tma_load_2d(&p.tmR, &res_full_bar[l_slot], cbuf + Cfg::BOX_BYTES, c0 + 64, m_tile * BLOCK_M);
          } else {
            mbar_arrive(&res_full_bar[l_slot]);
          }
          ++k_load;
          if (++l_slot == RING) l_slot = 0;
          if (++l_g == GROUPS) { l_g = 0; l_tile += gridDim.x; }
        }
        mbar_wait(&staged_bar[s_slot], s_phase);
        {
          const int m_tile = s_tile / p.n_tiles, n_tile = s_tile - m_tile * p.n_tiles;
          uint8_t* cbuf = smem_c + s_slot * Cfg::GROUP_BYTES;
          const int c0 = 2 * (n_tile * BLOCK_N + s_g * 64);
          tma_store_2d(&p.tmC, cbuf, c0, m_tile * BLOCK_M);
          tma_store_2d(&p.tmC, cbuf + Cfg::BOX_BYTES, c0 + 64, m_tile * BLOCK_M);
        }
        bulk_commit();
        bulk_wait_read<0>();
        ++k_store;
        if (++s_slot == RING) { s_slot = 0; s_phase ^= 1; }
        if (++s_g == GROUPS) { s_g = 0; s_tile += gridDim.x; }
      }
      bulk_wait<0>();  // smem must stay valid until the last store has completed
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <typename T, int BLOCK_N, int kAMode>
static int launch_split_t(const ConvTcParams& p, cudaStream_t st) {
  using Cfg = SplitCfg<BLOCK_N>;
  if (p.stages < 2 || p.stages > Cfg::MAX_STAGES || p.ring < 1 || p.ring > Cfg::MAX_RING ||
      p.stages * Cfg::STAGE_BYTES + p.ring * Cfg::GROUP_BYTES > Cfg::DATA_BYTES) {
    set_error("conv_tc (split): bad ring configuration %d / %d for a %d-wide tile", p.stages, p.ring, BLOCK_N);
    return SEMDIFF_ERR_ARG;
  }
  static bool configured[MAX_DEVICES] = {};
  auto kern = conv_tc_split_kernel<T, BLOCK_N, kAMode>;
  const int dev = current_device();
  if (!configured[dev]) {
    SEMDIFF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured[dev] = true;
  }
  const int tiles = p.m_tiles * p.n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  SEMDIFF_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(SPLIT_THREADS), Cfg::SMEM_BYTES, st, p));
  return 0;
}

template <typename T>
static int launch_split_n(const ConvTcParams& p, int block_n, int a_mode, cudaStream_t st) {
  if (a_mode == A_TMA) {
    switch (block_n) {
      case 256: return launch_split_t<T, 256, A_TMA>(p, st);
      case 128: return launch_split_t<T, 128, A_TMA>(p, st);
      case 64: return launch_split_t<T, 64, A_TMA>(p, st);
    }
  } else if (a_mode == A_IM2COL) {
    switch (block_n) {
      case 256: return launch_split_t<T, 256, A_IM2COL>(p, st);
      case 128: return launch_split_t<T, 128, A_IM2COL>(p, st);
      case 64: return launch_split_t<T, 64, A_IM2COL>(p, st);
    }
  }
  set_error("conv_tc (split): unsupported tile %d / mode %d", block_n, a_mode);
  return SEMDIFF_ERR_UNSUPPORTED;
}

// operand-ring fills and C-ring groups of a split conv (224 KB in total): 64-wide tiles 4 + 1 or 3 + 2, 128-wide 3 + 1 or
// 2 + 3, 256-wide 2 + 1.  A deep operand ring hides the TMA latency of long K loops (measured, profiles/r2_x3_ring_ab.txt:
// 3x3 convs -3..4 %, 512 -> 128 -10 %); short K loops (< 8 K blocks) are store-bound and residual convs prefetch their
// residual tiles RING - 1 groups ahead, so both keep the deep C ring (64 -> 256 with one slot: +50 %).
// SEMDIFF_X3_DEEP_RING=0 restores the shallow operand ring everywhere.
void split_ring_config(int block_n, bool has_res, int num_kb, int* stages, int* ring) {
  static const bool deep = getenv("SEMDIFF_X3_DEEP_RING") == nullptr || atoi(getenv("SEMDIFF_X3_DEEP_RING")) != 0;
  const bool d = deep && !has_res && num_kb >= 8;
  if (block_n == 64) { *stages = d ? 4 : 3; *ring = d ? 1 : 2; }
  else if (block_n == 128) { *stages = d ? 3 : 2; *ring = d ? 1 : 3; }
  else { *stages = 2; *ring = 1; }
}

int launch_conv_split(const ConvTcParams& p, int block_n, int a_mode, int precision, cudaStream_t st) {
  if (precision == SEMDIFF_FP16X3) return launch_split_n<__half>(p, block_n, a_mode, st);
  if (precision == SEMDIFF_BF16X3) return launch_split_n<__nv_bfloat16>(p, block_n, a_mode, st);
  set_error("conv_tc (split): bad precision %d", precision);
  return SEMDIFF_ERR_ARG;
}

}  // namespace semdiff
