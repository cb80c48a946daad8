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
//   warp 0      TMA producer: weight tile (always) + activation tile:
//                 A_TMA     1x1 stride-1 convs: plain 2-D tiled loads of [M, Cin]
//                 A_IM2COL  kxk / strided convs with Cin % 64 == 0: TMA im2col mode over the NHWC tensor
//   warp 1      tcgen05.mma issuer (one thread); owns the TMEM allocation (2 accumulator stages)
//   warps 2-9   epilogue (8 warps = 4 TMEM lane quarters x 2 column halves; 4 warps in the gather variant and for 32-wide tiles):
//               tcgen05.ld -> +bias (+residual) -> ReLU -> 16-bit -> swizzled smem slot; the warps never synchronise
//               with each other, each publishes its part of the slot on an mbarrier
//   warp 10     C-ring I/O: TMA store of every staged slot, and as soon as the store has READ the slot, the grant of
//               the slot to the column group RING ahead - with the TMA prefetch of its residual tile, which the
//               epilogue then overwrites in place with the result
//   warps 6-9   (A_GATHER only; the residual loader is then warp 10) software im2col for Cin < 64 (stems): cp.async 16 B chunks into the swizzled
//               A tile, zero-filling padding / M tail / K tail
// Pipelines: smem ring full/empty (producers <-> MMA), TMEM full/empty (MMA <-> epilogue),
//            C ring res_full/staged (I/O warp <-> epilogue).
#include <cuda.h>
#include <stdlib.h>

#include "conv_tc.h"

namespace semdiff {

// kBRes ("B resident"): the whole weight matrix of the conv (one n-tile, <= MAX_RES_KB k-blocks) is loaded into shared
// memory ONCE per CTA and the ring only cycles activation tiles - for the 64-output-channel convs (stems, layer1 3x3s)
// whose tiles are bound by L2->SM operand traffic, this removes a third of it and doubles the ring depth.
constexpr int MAX_RES_KB = 9;  // 9 x 64 = 576 = 3x3x64

template <int BLOCK_N, bool kBRes = false> struct TcCfg {
  static constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + (kBRes ? 0 : B_STAGE_BYTES);
  static constexpr int STAGES = kBRes ? 6 : (BLOCK_N >= 128 ? 4 : 6);
  static constexpr int B_REGION_BYTES = kBRes ? MAX_RES_KB * B_STAGE_BYTES : STAGES * B_STAGE_BYTES;
  static constexpr int GATHER_LAG = STAGES - 2;        // cp.async groups in flight per gather thread
  static constexpr int TMEM_COLS = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
  // epilogue staging: the tile is written out in column groups, each group = BOXES TMA boxes of BOX_COLS columns
  static constexpr int BOX_COLS = BLOCK_N < 64 ? BLOCK_N : 64;
  static constexpr int GROUP_COLS = BLOCK_N < 128 ? BLOCK_N : 128;
  static constexpr int GROUPS = BLOCK_N / GROUP_COLS;
  static constexpr int BOXES = GROUP_COLS / BOX_COLS;   // per group
  static constexpr int BOX_BYTES = BLOCK_M * BOX_COLS * 2;
  static constexpr int GROUP_BYTES = BOXES * BOX_BYTES;
  static constexpr int RING = BLOCK_N == 256 ? 1 : 3;   // BLOCK_N == 256 never carries a residual
  static constexpr int NUM_BARS = 2 * STAGES + 4 + 2 * RING + 1;
  static constexpr int SMEM_BYTES = STAGES * A_STAGE_BYTES + B_REGION_BYTES + RING * GROUP_BYTES + NUM_BARS * 8 + 16 + 1024;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

// epilogue warps: 8 (two per TMEM lane quarter, splitting the columns) except where warps 6-9 are the gather producers
template <int BLOCK_N, int kAMode> __host__ __device__ constexpr int epi_warps() { return (kAMode == A_GATHER || BLOCK_N < 64) ? 4 : 8; }
template <int BLOCK_N, int kAMode> __host__ __device__ constexpr int cta_threads() {
  return kAMode == A_GATHER ? 352 : (2 + epi_warps<BLOCK_N, kAMode>() + 1) * 32;
}

template <typename T, int BLOCK_N, int kAMode, bool kBRes = false>
__global__ void __launch_bounds__(cta_threads<BLOCK_N, kAMode>(), 1) conv_tc_kernel(const __grid_constant__ ConvTcParams p) {
  using Cfg = TcCfg<BLOCK_N, kBRes>;
  constexpr int EPI_WARPS = epi_warps<BLOCK_N, kAMode>();
  constexpr int STAGES = Cfg::STAGES, RING = Cfg::RING;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint8_t* smem_c = smem_b + Cfg::B_REGION_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_c + RING * Cfg::GROUP_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* res_full_bar = tmem_empty_bar + 2;
  uint64_t* staged_bar = res_full_bar + RING;
  uint64_t* bres_bar = staged_bar + RING;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bres_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;
  // one lane per warp, elected once: ptxas then knows that the single-thread roles below (TMA / tcgen05 issue) are
  // single-lane and emits the uniform-datapath instructions directly instead of an elect-and-retry loop per instruction
  const bool leader = elect_one();

  if (warp == 0 && leader) {
    if (kAMode != A_GATHER) tma_prefetch_desc(&p.tmA);
    if (p.has_src2) tma_prefetch_desc(&p.tmA2);
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmC);
    if (p.has_res) tma_prefetch_desc(&p.tmR);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], kAMode == A_GATHER ? 1 + 4 : 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], EPI_WARPS);
    }
    for (int i = 0; i < RING; ++i) {
      mbar_init(&res_full_bar[i], 1);
      mbar_init(&staged_bar[i], EPI_WARPS);
    }
    mbar_init(bres_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_trigger();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (leader) {
      int stage = 0, phase = 0;
      const int kb_per_tap = p.Cin >> 6;
      if (kBRes) {  // the whole weight matrix, once
        mbar_arrive_expect_tx(bres_bar, p.num_kb * Cfg::B_STAGE_BYTES);
        for (int kb = 0; kb < p.num_kb; ++kb)
          tma_load_2d(&p.tmB, bres_bar, smem_b + kb * Cfg::B_STAGE_BYTES, kb * BLOCK_K, 0);
      }
      pdl_wait();  // activations of the previous kernel (everything downstream of these loads is ordered by mbarriers)
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
        int tap = 0, cb = 0;  // filter tap and 64-channel block of the current k-block (im2col)
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], (kBRes ? 0 : Cfg::B_STAGE_BYTES) + (kAMode != A_GATHER ? A_STAGE_BYTES : 0));
          if (kb >= p.num_kb1) {
            const int kb2 = kb - p.num_kb1;  // fused 1x1 conv over the second activation tensor
            if (p.a2_im2col)
              tma_load_im2col_4d(&p.tmA2, &full_bar[stage], smem_a + stage * A_STAGE_BYTES, kb2 * BLOCK_K, w2, h2, n, 0, 0);
            else
              tma_load_2d(&p.tmA2, &full_bar[stage], smem_a + stage * A_STAGE_BYTES, kb2 * BLOCK_K, m_tile * BLOCK_M);
          } else if (kAMode == A_TMA) {
            tma_load_2d(&p.tmA, &full_bar[stage], smem_a + stage * A_STAGE_BYTES, kb * BLOCK_K, m_tile * BLOCK_M);
          } else if (kAMode == A_IM2COL) {
            const int r = tap / p.KW, s = tap - r * p.KW;
            tma_load_im2col_4d(&p.tmA, &full_bar[stage], smem_a + stage * A_STAGE_BYTES, cb * BLOCK_K, w0, h0, n,
                               (uint16_t)s, (uint16_t)r);
            if (++cb == kb_per_tap) { cb = 0; ++tap; }
          }
          if (!kBRes)
            tma_load_2d(&p.tmB, &full_bar[stage], smem_b + stage * Cfg::B_STAGE_BYTES, kb * BLOCK_K, n_tile * BLOCK_N);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp waits, the elected lane issues) =====================
    constexpr uint32_t idesc = umma_idesc_f16(Elem<T>::kUmmaFormat, BLOCK_M, BLOCK_N);
    // descriptors differ between stages / K steps only in the 14-bit start-address field: precompute and add
    const uint64_t a_desc0 = umma_smem_desc_sw128(smem_u32(smem_a));
    const uint64_t b_desc0 = umma_smem_desc_sw128(smem_u32(smem_b));
    int stage = 0, phase = 0, local = 0;
    if (kBRes && blockIdx.x < total_tiles) mbar_wait(bres_bar, 0);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      const int acc = local & 1, acc_phase = (local >> 1) & 1;
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        if (leader) {
          const uint64_t a_desc = a_desc0 + (uint64_t)((stage * A_STAGE_BYTES) >> 4);
          const uint64_t b_desc = b_desc0 + (uint64_t)(((kBRes ? kb : stage) * Cfg::B_STAGE_BYTES) >> 4);
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k)
            umma_f16_ss(tmem_d, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (kb == p.num_kb - 1) umma_commit(&tmem_full_bar[acc]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp < 2 + EPI_WARPS) {
    // ===================== epilogue (EPI_WARPS warps: 4 TMEM lane quarters x HALVES column halves) =====================
    constexpr int ROW_BYTES = Cfg::BOX_COLS * 2;
    constexpr int HALVES = EPI_WARPS / 4;
    constexpr int UNITS = Cfg::GROUP_COLS / 32;          // 32-column units per group (one tcgen05.ld.x32 each)
    const int q = warp & 3;                              // TMEM lane quarter this warp may touch
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    int local = 0, slot = 0, sphase = 0;  // slot / sphase: position in the C ring
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      const int m_tile = tile / p.n_tiles, n_tile = tile - m_tile * p.n_tiles;
      const int acc = local & 1, acc_phase = (local >> 1) & 1;
      bool tmem_ready = false;
#pragma unroll 1
      for (int g = 0; g < Cfg::GROUPS; ++g) {
        uint8_t* cbuf = smem_c + slot * Cfg::GROUP_BYTES;
        mbar_wait_short(&res_full_bar[slot], sphase);  // the slot is ours (and the residual tile, if any, has landed)
        if (!tmem_ready) {
          mbar_wait_backoff(&tmem_full_bar[acc], acc_phase);
          tcgen05_fence_after();
          tmem_ready = true;
        }
#pragma unroll 1
        for (int u = half; u < UNITS; u += HALVES) {
          const int col_in_group = u * 32;
          const int col_in_tile = g * Cfg::GROUP_COLS + col_in_group;
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + (uint32_t(q * 32) << 16) + acc * BLOCK_N + col_in_tile, v);
          tmem_ld_wait();
          if (g == Cfg::GROUPS - 1 && u + HALVES >= UNITS) {
            // this warp's last TMEM read of the accumulator stage has retired: hand it back to the MMA warp early
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
          }
          const int col0 = n_tile * BLOCK_N + col_in_tile;
          const int box = col_in_group / Cfg::BOX_COLS, j0 = (col_in_group % Cfg::BOX_COLS) / 8;
          const uint32_t row_addr = smem_u32(cbuf + box * Cfg::BOX_BYTES) + row * ROW_BYTES;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j * 8));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j * 8 + 4));
            float f[8] = {__uint_as_float(v[j * 8 + 0]) + b0.x, __uint_as_float(v[j * 8 + 1]) + b0.y,
                          __uint_as_float(v[j * 8 + 2]) + b0.z, __uint_as_float(v[j * 8 + 3]) + b0.w,
                          __uint_as_float(v[j * 8 + 4]) + b1.x, __uint_as_float(v[j * 8 + 5]) + b1.y,
                          __uint_as_float(v[j * 8 + 6]) + b1.z, __uint_as_float(v[j * 8 + 7]) + b1.w};
            const uint32_t addr = row_addr + (swz_chunk<ROW_BYTES>(j0 + j, row) << 4);
            if (p.has_res) {
              uint4 rq;
              asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(rq.x), "=r"(rq.y), "=r"(rq.z), "=r"(rq.w) : "r"(addr));
              float r[8];
              unpack8<T>(rq, r);
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] += r[e];
            }
            if (p.relu) {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
            }
            const uint4 o = pack8<T>(f);
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
          }
        }
        // this warp's part of the slot is staged: make it visible to the async proxy (TMA store), then publish
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&staged_bar[slot]);
        if (++slot == RING) { slot = 0; sphase ^= 1; }
      }
    }
  } else if (warp == (kAMode == A_GATHER ? 10 : 2 + EPI_WARPS)) {
    // ===================== C-ring I/O: TMA stores of staged slots + slot grants / residual prefetch =====================
    // One thread owns both directions: a slot is granted to the column group RING ahead as soon as its store has READ
    // it, and that group's residual tile starts loading at once - RING - 1 groups of HBM latency hidden.
    if (leader && blockIdx.x < total_tiles) {
      pdl_wait();
      const int total = Cfg::GROUPS * ((total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x);
      int k_load = 0, l_slot = 0, l_g = 0, l_tile = blockIdx.x;
      int k_store = 0, s_slot = 0, s_phase = 0, s_g = 0, s_tile = blockIdx.x;
      while (k_store < total) {
        while (k_load < total && k_load < k_store + RING) {
          if (p.has_res) {
            const int m_tile = l_tile / p.n_tiles, n_tile = l_tile - m_tile * p.n_tiles;
            mbar_arrive_expect_tx(&res_full_bar[l_slot], Cfg::GROUP_BYTES);
            uint8_t* cbuf = smem_c + l_slot * Cfg::GROUP_BYTES;
#pragma unroll
            for (int b = 0; b < Cfg::BOXES; ++b)
              tma_load_2d(&p.tmR, &res_full_bar[l_slot], cbuf + b * Cfg::BOX_BYTES,
                          n_tile * BLOCK_N + l_g * Cfg::GROUP_COLS + b * Cfg::BOX_COLS, m_tile * BLOCK_M);
          } else {
            mbar_arrive(&res_full_bar[l_slot]);
          }
          ++k_load;
          if (++l_slot == RING) l_slot = 0;
          if (++l_g == Cfg::GROUPS) { l_g = 0; l_tile += gridDim.x; }
        }
        mbar_wait(&staged_bar[s_slot], s_phase);
        {
          const int m_tile = s_tile / p.n_tiles, n_tile = s_tile - m_tile * p.n_tiles;
          uint8_t* cbuf = smem_c + s_slot * Cfg::GROUP_BYTES;
#pragma unroll
          for (int b = 0; b < Cfg::BOXES; ++b)
            tma_store_2d(&p.tmC, cbuf + b * Cfg::BOX_BYTES, n_tile * BLOCK_N + s_g * Cfg::GROUP_COLS + b * Cfg::BOX_COLS,
                         m_tile * BLOCK_M);
        }
        bulk_commit();
        bulk_wait_read<0>();
        ++k_store;
        if (++s_slot == RING) { s_slot = 0; s_phase ^= 1; }
        if (++s_g == Cfg::GROUPS) { s_g = 0; s_tile += gridDim.x; }
      }
      bulk_wait<0>();  // smem must stay valid until the last store has completed
    }
  } else {
    // ===================== software im2col gather (128 threads, one A row each) =====================
    constexpr int LAG = Cfg::GATHER_LAG;
    const int row = (warp - 6) * 32 + lane;
    const T* in = reinterpret_cast<const T*>(p.in);
    const uint32_t row_off = row * 128, sw = row & 7;
    int stage = 0, phase = 0, arr_stage = 0;
    int issued = 0;
    pdl_wait();
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
          // several taps per k-block (Cin = 8 .. 32): each 16 B chunk is its own (tap, channel group)
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
        if (issued > LAG) {
          cp_async_wait<LAG>();
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
    const int pending = issued < LAG ? issued : LAG;
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
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);
static void* driver_entry(const char* name) {
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint(name, &ptr, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
    return nullptr;
  return ptr;
}
static CUtensorMapDataType tmap_dtype(int precision) {
  return base_precision(precision) == SEMDIFF_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
}

// 2-D 16-bit tensor [rows, cols] (cols contiguous), box = box_cols x box_rows; swizzle span = box_cols * 2 bytes; OOB -> 0
static int make_tmap_2d(CUtensorMap* m, const void* base, int precision, uint64_t rows, uint64_t cols, uint32_t box_cols,
                        uint32_t box_rows) {
  static EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(driver_entry("cuTensorMapEncodeTiled"));
  if (enc == nullptr) { set_error("cuTensorMapEncodeTiled entry point not found"); return SEMDIFF_ERR_CUDA; }
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * 2};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = box_cols * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = enc(m, tmap_dtype(precision), 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu box=%ux%u base=%p", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, box_cols, box_rows, base);
    return SEMDIFF_ERR_CUDA;
  }
  return 0;
}

// NHWC activation as (C, W, H, N) in TMA im2col mode: one load = 128 consecutive output pixels x 64 channels of one
// filter tap.  Corner arithmetic as in CUTLASS's fprop (conv/collective/detail.hpp compute_{lower,upper}_corner_whd):
// lower = -pad, upper = pad - (k - 1); traversal stride = conv stride.
static int make_tmap_im2col(CUtensorMap* m, const void* base, int precision, const ConvShape& s_in, bool second) {
  ConvShape s = s_in;
  if (second) { s.H = s_in.H2; s.W = s_in.W2; s.cin = s_in.cin2; s.kh = s.kw = 1; s.stride = s_in.stride2; s.pad = 0; }
  static EncodeIm2colFn enc = reinterpret_cast<EncodeIm2colFn>(driver_entry("cuTensorMapEncodeIm2col"));
  if (enc == nullptr) { set_error("cuTensorMapEncodeIm2col entry point not found"); return SEMDIFF_ERR_CUDA; }
  const cuuint64_t cs = (cuuint64_t)s.cin * (is_split(precision) ? 2 : 1);   // stored 16-bit channels per pixel
  const cuuint64_t dims[4] = {cs, (cuuint64_t)s.W, (cuuint64_t)s.H, (cuuint64_t)s.n_img};
  const cuuint64_t strides[3] = {cs * 2, (cuuint64_t)s.W * cs * 2, (cuuint64_t)s.H * s.W * cs * 2};
  const int lower[2] = {-s.pad, -s.pad};
  const int upper[2] = {s.pad - (s.kw - 1), s.pad - (s.kh - 1)};
  const cuuint32_t estr[4] = {1, (cuuint32_t)s.stride, (cuuint32_t)s.stride, 1};
  CUresult r = enc(m, tmap_dtype(precision), 4, const_cast<void*>(base), dims, strides, lower, upper, BLOCK_K, BLOCK_M,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeIm2col failed (%d) n=%d H=%d W=%d C=%d k=%d stride=%d pad=%d", (int)r, s.n_img, s.H, s.W,
              s.cin, s.kh, s.stride, s.pad);
    return SEMDIFF_ERR_CUDA;
  }
  // Same workaround CUTLASS applies (cute/atom/copy_traits_sm90_im2col.hpp) for drivers <= 13.1: small tensors
  int drv = 0;
  cudaDriverGetVersion(&drv);
  if (drv <= 13010 && (uint64_t)s.n_img * s.H * s.W * cs * 2 < 131072)
    reinterpret_cast<uint64_t*>(m)[1] &= ~(1ull << 21);
  return 0;
}

static int pick_block_n(int cout, bool has_res, int num_kb, int m_tiles, int sms) {
  if (cout % 128 != 0) return cout % 64 == 0 ? 64 : (cout % 32 == 0 ? 32 : 0);
  if (has_res || cout % 256 != 0) return 128;
  // 256-wide tiles raise the FLOP per operand byte fetched from L2 (and halve the activation re-loads of a conv with
  // several n-tiles) but stage their output through a single smem slot; short-K convs are store-bound and want the
  // ring of the 128-wide tile.  Measured (profiles/r1_knobs.txt): K = 384 (layer2's conv3 + strided shortcut) 0.217 ->
  // 0.184 ms with 256-wide tiles.
  static const int min_kb = getenv("SEMDIFF_N256_MIN_KB") ? atoi(getenv("SEMDIFF_N256_MIN_KB")) : 6;
  if (num_kb < min_kb) return 128;
  const int64_t t256 = (int64_t)m_tiles * (cout / 256);
  if (t256 < sms) return 128;
  // (falling back to 128-wide tiles when the last wave of 256-wide ones is poorly filled was measured slower: 5.58 -> 5.83 ms)
  return 256;
}

static int a_mode_for(const ConvShape& s, int requested) {
  const bool pointwise = s.kh == 1 && s.kw == 1 && s.stride == 1 && s.pad == 0 && s.cin % 64 == 0;
  if (requested == SEMDIFF_CONV_TC_GATHER) return A_GATHER;
  if (pointwise) return A_TMA;
  if (s.cin % 64 == 0 && s.pad <= 8 && s.kh <= 9 && s.kw <= 9) return A_IM2COL;
  return A_GATHER;
}

bool conv_tc_supported(const ConvShape& s, int precision, bool use_tma) {
  if (is_split(precision)) {   // split precisions: TMA-fed only, whole 64-channel blocks on both sides
    if (!use_tma || s.cin % 64 != 0 || s.cout % 64 != 0 || s.cin2 % 64 != 0) return false;
    precision = base_precision(precision);
  }
  if (precision != SEMDIFF_BF16 && precision != SEMDIFF_FP16) return false;
  if (s.pad_hi >= 0 && s.pad_hi != s.pad) return false;  // asymmetric padding: strip kernel or SIMT only
  if (s.cin % 8 != 0 || s.cout % 32 != 0) return false;
  if (s.cin > 64 && s.cin % 64 != 0) return false;
  if (s.cin < 64 && 64 % s.cin != 0) return false;
  if (use_tma && s.cin % 64 != 0) return false;
  if (s.cin2 != 0 && (!use_tma || s.cin2 % 64 != 0 || s.K1() % 64 != 0)) return false;
  return s.M() > 0 && s.M() < (int64_t)1 << 31;
}

int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev >= 0 && dev < MAX_DEVICES ? dev : 0;
}
int num_sms() {
  static int sms[MAX_DEVICES] = {};
  const int dev = current_device();
  if (sms[dev] == 0) {
    cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
    if (sms[dev] <= 0) sms[dev] = 148;
  }
  return sms[dev];
}

template <typename T, int BLOCK_N, int kAMode, bool kBRes = false>
static int launch_t(const ConvTcParams& p, cudaStream_t st) {
  using Cfg = TcCfg<BLOCK_N, kBRes>;
  static bool configured[MAX_DEVICES] = {};
  auto kern = conv_tc_kernel<T, BLOCK_N, kAMode, kBRes>;
  const int dev = current_device();
  if (!configured[dev]) {
    SEMDIFF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured[dev] = true;
  }
  const int tiles = p.m_tiles * p.n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  SEMDIFF_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(cta_threads<BLOCK_N, kAMode>()), Cfg::SMEM_BYTES, st, p));
  return 0;
}

template <typename T, int kAMode>
static int launch_n(const ConvTcParams& p, int block_n, cudaStream_t st) {
  switch (block_n) {
    case 256: return launch_t<T, 256, kAMode>(p, st);
    case 128: return launch_t<T, 128, kAMode>(p, st);
    case 64:
      if constexpr (kAMode != A_GATHER) {
        if (p.n_tiles == 1 && p.num_kb <= MAX_RES_KB) return launch_t<T, 64, kAMode, true>(p, st);
      }
      return launch_t<T, 64, kAMode>(p, st);
    case 32: return launch_t<T, 32, kAMode>(p, st);
  }
  set_error("conv_tc: unsupported BLOCK_N %d", block_n);
  return SEMDIFF_ERR_UNSUPPORTED;
}

template <typename T>
static int launch_mode(const ConvTcParams& p, int block_n, int a_mode, cudaStream_t st) {
  switch (a_mode) {
    case A_TMA: return launch_n<T, A_TMA>(p, block_n, st);
    case A_GATHER: return launch_n<T, A_GATHER>(p, block_n, st);
    case A_IM2COL: return launch_n<T, A_IM2COL>(p, block_n, st);
  }
  return SEMDIFF_ERR_UNSUPPORTED;
}

int conv_tc_prepare(ConvTcLaunch* L, const ConvPtrs& q, const ConvShape& s, int precision, bool use_tma) {
  // 64 -> 64 spatial convs of the wide early layers and the space-to-depth stems: strip kernel (conv3x3_strip.cu);
  // SEMDIFF_NO_STRIP=1 keeps them on the generic im2col path (A/B testing)
  static const bool strip_ok = getenv("SEMDIFF_NO_STRIP") == nullptr;
  if ((strip_ok || s.cin < 64) && use_tma && q.res == nullptr && conv_strip_supported(s, precision)) return conv_strip_prepare(L, q, s, precision);
  const void* in = q.in; const void* w = q.w; const float* bias = q.bias; const void* res = q.res; void* out = q.out;
  if (!conv_tc_supported(s, precision, use_tma)) {
    set_error("conv_tc: unsupported shape cin=%d cout=%d k=%dx%d stride=%d pad=%d tma=%d precision=%d", s.cin, s.cout,
              s.kh, s.kw, s.stride, s.pad, (int)use_tma, precision);
    return SEMDIFF_ERR_UNSUPPORTED;
  }
  static_assert(sizeof(ConvTcParams) <= sizeof(L->params), "ConvTcLaunch::params too small");
  ConvTcParams& p = *reinterpret_cast<ConvTcParams*>(L->params);
  memset(&p, 0, sizeof(p));
  const int a_mode = a_mode_for(s, use_tma ? SEMDIFF_CONV_TC_TMA : SEMDIFF_CONV_TC_GATHER);
  const bool split = is_split(precision);
  const int kmul = split ? 2 : 1;    // stored 16-bit columns per logical channel
  p.M = (int)s.M();
  p.m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
  p.num_kb = (s.K() + BLOCK_K - 1) / BLOCK_K;
  int block_n = pick_block_n(s.cout, res != nullptr, p.num_kb, p.m_tiles, num_sms());
  if (split && block_n == 256) {
    // the 256-wide split tile stages its output through one slot while the same warps must keep draining the next tile's
    // chunk sums, which only amortises over a long K loop: measured 15.19k / 15.31k / 15.41k pairs/s for a threshold of
    // 6 / 8 / 10 K blocks (profiles/r2_x3_per_op.txt)
    static const int min_kb = getenv("SEMDIFF_X3_N256_MIN_KB") ? atoi(getenv("SEMDIFF_X3_N256_MIN_KB")) : 10;
    if (p.num_kb < min_kb) block_n = 128;
  }
  if (block_n == 0) { set_error("conv_tc: cout %d not a multiple of 32", s.cout); return SEMDIFF_ERR_UNSUPPORTED; }
  p.in = in; p.bias = bias; p.has_res = res != nullptr;
  p.H = s.H; p.W = s.W; p.Cin = s.cin; p.OH = s.OH(); p.OW = s.OW(); p.Cout = s.cout;
  p.KH = s.kh; p.KW = s.kw; p.stride = s.stride; p.pad = s.pad; p.relu = s.relu;
  p.n_tiles = s.cout / block_n;
  p.acc_scale = split && s.wscale > 0.f ? 1.f / s.wscale : 1.f;
  // chunk length of the promoted accumulation in k-blocks; SEMDIFF_X3_CHUNK_KB overrides (A/B testing)
  static const int chunk_kb = getenv("SEMDIFF_X3_CHUNK_KB") ? atoi(getenv("SEMDIFF_X3_CHUNK_KB")) : 1;
  p.chunk_kb = chunk_kb < 1 ? 1 : chunk_kb;
  if (split) split_ring_config(block_n, res != nullptr, p.num_kb, &p.stages, &p.ring);
  p.cpt = s.cin / 8;
  p.taps = s.kh * s.kw;
  const uint32_t box_cols = block_n < 64 ? block_n : 64;
  int rc = make_tmap_2d(&p.tmB, w, precision, (uint64_t)s.cout, (uint64_t)s.K() * kmul, BLOCK_K, (uint32_t)block_n);
  if (rc == 0 && a_mode == A_TMA) rc = make_tmap_2d(&p.tmA, in, precision, (uint64_t)p.M, (uint64_t)s.cin * kmul, BLOCK_K, BLOCK_M);
  if (rc == 0 && a_mode == A_IM2COL) rc = make_tmap_im2col(&p.tmA, in, precision, s, false);
  p.num_kb1 = p.num_kb;
  p.stride2 = 1;
  if (rc == 0 && s.cin2 != 0) {
    p.num_kb1 = s.K1() / BLOCK_K;
    p.has_src2 = 1;
    p.stride2 = s.stride2;
    p.a2_im2col = s.stride2 != 1;
    rc = p.a2_im2col ? make_tmap_im2col(&p.tmA2, q.in2, precision, s, true)
                     : make_tmap_2d(&p.tmA2, q.in2, precision, (uint64_t)p.M, (uint64_t)s.cin2 * kmul, BLOCK_K, BLOCK_M);
  }
  if (rc == 0) rc = make_tmap_2d(&p.tmC, out, precision, (uint64_t)p.M, (uint64_t)s.cout * kmul, box_cols, BLOCK_M);
  if (rc == 0 && res != nullptr) rc = make_tmap_2d(&p.tmR, res, precision, (uint64_t)p.M, (uint64_t)s.cout * kmul, box_cols, BLOCK_M);
  if (rc != 0) return rc;
  L->block_n = block_n;
  L->a_mode = a_mode;
  L->precision = precision;
  return 0;
}

int conv_tc_launch(const ConvTcLaunch* L, cudaStream_t st) {
  if (L->a_mode > 200) return conv_chain_launch(L, st);
  if (L->a_mode > 100) return conv_strip_launch(L, st);
  const ConvTcParams& p = *reinterpret_cast<const ConvTcParams*>(L->params);
  switch (L->precision) {
    case SEMDIFF_BF16: return launch_mode<__nv_bfloat16>(p, L->block_n, L->a_mode, st);
    case SEMDIFF_FP16: return launch_mode<__half>(p, L->block_n, L->a_mode, st);
    case SEMDIFF_FP16X3:
    case SEMDIFF_BF16X3: return launch_conv_split(p, L->block_n, L->a_mode, L->precision, st);
  }
  set_error("conv_tc: bad precision %d", L->precision);
  return SEMDIFF_ERR_ARG;
}

int launch_conv_tc(const ConvPtrs& q, const ConvShape& s, int precision, bool use_tma, cudaStream_t st) {
  ConvTcLaunch L;
  int rc = conv_tc_prepare(&L, q, s, precision, use_tma);
  if (rc != 0) return rc;
  return conv_tc_launch(&L, st);
}

}  // namespace semdiff
