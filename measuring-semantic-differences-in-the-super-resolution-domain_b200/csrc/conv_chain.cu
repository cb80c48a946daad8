// Chained pointwise convs of a bottleneck boundary, one kernel, for sm_100a:
//
//   y  = relu1( x1 * W1a^T (+ x2 * W1b^T) + b1 (+ res) )        [M, N1]   (conv3 of block i, shortcut fused or residual)
//   t  = relu2( y * W2^T + b2 )                                  [M, N2]   (conv1 of block i + 1)
//   N1 = 256 (G = 2 chunks of 128 columns; W1 and W2 resident in shared memory; N2 = 64 | 128), or
//   N1 = 512 (G = 4; identity blocks of the 512-channel stage: K1 = 128, N2 = 128; the two 128 KB weight matrices are
//             streamed per chunk from L2 through one-chunk slots, W2 by its own producer warp)
//
// In the 256-channel stage (56 x 56 at 224 x 224 input) both convs are bound by HBM: y is written by the first and read
// straight back by the second.  Here the 128 x 256 tile of y never leaves the SM between the two: the epilogue of the
// first GEMM stages the 16-bit tile in shared memory for its TMA store, and that staging buffer - already in the
// K-major SWIZZLE_128B layout - is the A operand of the second GEMM.  Numerically nothing changes: the second GEMM
// consumes exactly the rounded 16-bit values the unfused path would have re-read from HBM.
// Reference: the timm bottleneck the scorer runs under /root/reference/models/global_eval_models.py:364,371.
//
// Roles (384 threads; warp 11 only loads W2 chunks in the streamed variant): warp 0 TMA producer (x tiles; W1 / W2 once per CTA, resident), warp 1 tcgen05 issuer
// (GEMM1 per 128-column chunk into 2 TMEM stages, GEMM2 one chunk behind into 2 more), warps 2-9 epilogue
// (TMEM -> +bias (+residual, in place) -> ReLU -> 16-bit staged tile; no barrier between the warps), warp 10 C-ring I/O
// (TMA store of each staged slot, residual prefetch into it as soon as the store has read it).  C ring slots cycle
// chunk0, chunk1, t-tile  per pixel tile; a slot is reusable when the TMA store has read it AND (for the y chunks)
// GEMM2 has consumed it.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "kernels.h"

namespace semdiff {

namespace chain {
constexpr int BLOCK_M = 128, CHUNK = 128;
constexpr int A_STAGE = BLOCK_M * 64 * 2;       // one 64-channel k-block of a pixel tile
constexpr int W1_BLOCK = CHUNK * 64 * 2;        // [128 output channels][64 k]
constexpr int SLOT = BLOCK_M * CHUNK * 2;       // one staged chunk: 2 boxes of [128 px][64 ch]
constexpr int BOX = BLOCK_M * 64 * 2;
constexpr int MAX_STAGES = 6, MAX_RING = 4, EPI_WARPS = 8, THREADS = (2 + EPI_WARPS + 2) * 32;
constexpr int NUM_BARS = 2 * MAX_STAGES + 8 + 4 * MAX_RING + 1 + 4;
constexpr int SMEM_LIMIT = 232448;
}  // namespace chain

struct alignas(64) ChainParams {
  CUtensorMap tmA, tmA2, tmW1, tmW2, tmC, tmR, tmC2;
  const float* bias1;
  const float* bias2;
  int M, m_tiles, nkb1, nkb_a, has_res, relu1, relu2, stages, ring, smem_bytes, n2, g;
};
static_assert(sizeof(ChainParams) <= sizeof(ConvTcLaunch::params), "ConvTcLaunch::params too small");

template <typename T, int N2, int G>
__global__ void __launch_bounds__(chain::THREADS, 1) conv_chain_kernel(const __grid_constant__ ChainParams p) {
  using namespace chain;
  constexpr int N1 = G * CHUNK;
  constexpr bool kStream = G > 2;        // weights streamed per chunk instead of resident
  constexpr int W2_BLOCK = N2 * 64 * 2;  // [N2][64 k]
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_w1 = smem;
  uint8_t* smem_w2 = smem_w1 + (kStream ? 1 : G) * p.nkb1 * W1_BLOCK;
  uint8_t* smem_a = smem_w2 + (kStream ? CHUNK / 64 : N1 / 64) * W2_BLOCK;
  uint8_t* smem_c = smem_a + p.stages * A_STAGE;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_c + p.ring * SLOT);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* t1_full = empty_bar + MAX_STAGES;
  uint64_t* t1_empty = t1_full + 2;
  uint64_t* t2_full = t1_empty + 2;
  uint64_t* t2_empty = t2_full + 2;
  uint64_t* res_full = t2_empty + 2;
  uint64_t* c_free = res_full + MAX_RING;
  uint64_t* c_ready = c_free + MAX_RING;
  uint64_t* staged = c_ready + MAX_RING;
  uint64_t* w_bar = staged + MAX_RING;
  uint64_t* w1_full = w_bar + 1;   // streamed weights: one chunk slot each, full / empty
  uint64_t* w1_empty = w1_full + 1;
  uint64_t* w2_full = w1_empty + 1;
  uint64_t* w2_empty = w2_full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(w2_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool leader = elect_one();
  const int S = p.stages, R = p.ring;

  if (warp == 0 && leader) {
    tma_prefetch_desc(&p.tmA);
    if (p.nkb_a < p.nkb1) tma_prefetch_desc(&p.tmA2);
    tma_prefetch_desc(&p.tmW1);
    tma_prefetch_desc(&p.tmW2);
    tma_prefetch_desc(&p.tmC);
    tma_prefetch_desc(&p.tmC2);
    if (p.has_res) tma_prefetch_desc(&p.tmR);
    for (int i = 0; i < MAX_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&t1_full[i], 1); mbar_init(&t1_empty[i], EPI_WARPS);
      mbar_init(&t2_full[i], 1); mbar_init(&t2_empty[i], EPI_WARPS);
    }
    for (int i = 0; i < MAX_RING; ++i) {
      mbar_init(&res_full[i], 1); mbar_init(&c_free[i], 1);
      mbar_init(&c_ready[i], EPI_WARPS); mbar_init(&staged[i], EPI_WARPS);
    }
    mbar_init(w_bar, 1);
    mbar_init(w1_full, 1); mbar_init(w1_empty, 1); mbar_init(w2_full, 1); mbar_init(w2_empty, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_acc2 = tmem_base + 2 * CHUNK;
  pdl_trigger();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (leader) {
      int stage = 0, phase = 0;
      auto load_x = [&](int tile) {   // the x k-blocks of one pixel tile
        for (int kb = 0; kb < p.nkb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], A_STAGE);
          if (kb < p.nkb_a) tma_load_2d(&p.tmA, &full_bar[stage], smem_a + stage * A_STAGE, kb * 64, tile * BLOCK_M);
          else tma_load_2d(&p.tmA2, &full_bar[stage], smem_a + stage * A_STAGE, (kb - p.nkb_a) * 64, tile * BLOCK_M);
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      };
      if constexpr (!kStream) {
        mbar_arrive_expect_tx(w_bar, G * p.nkb1 * W1_BLOCK + (N1 / 64) * W2_BLOCK);
        for (int g = 0; g < G; ++g)
          for (int kb = 0; kb < p.nkb1; ++kb)
            tma_load_2d(&p.tmW1, w_bar, smem_w1 + (g * p.nkb1 + kb) * W1_BLOCK, kb * 64, g * CHUNK);
        for (int kb = 0; kb < N1 / 64; ++kb) tma_load_2d(&p.tmW2, w_bar, smem_w2 + kb * W2_BLOCK, kb * 64, 0);
        pdl_wait();  // x tiles are the previous kernel's output
        for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) load_x(tile);
      } else {
        // streamed: the W1 chunk of (tile, g) goes into the single slot as soon as GEMM1 of the previous chunk has
        // retired; the x tile of the NEXT pixel tile is requested right after the first chunk of this one
        pdl_wait();
        int n = 0;
        if (blockIdx.x < p.m_tiles) load_x(blockIdx.x);
        for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
          for (int g = 0; g < G; ++g, ++n) {
            mbar_wait(w1_empty, (n & 1) ^ 1);
            mbar_arrive_expect_tx(w1_full, p.nkb1 * W1_BLOCK);
            for (int kb = 0; kb < p.nkb1; ++kb) tma_load_2d(&p.tmW1, w1_full, smem_w1 + kb * W1_BLOCK, kb * 64, g * CHUNK);
            if (g == 0 && tile + (int)gridDim.x < p.m_tiles) load_x(tile + gridDim.x);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc1 = umma_idesc_f16(Elem<T>::kUmmaFormat, BLOCK_M, CHUNK);
    constexpr uint32_t idesc2 = umma_idesc_f16(Elem<T>::kUmmaFormat, BLOCK_M, N2);
    const uint64_t a_desc0 = umma_smem_desc_sw128(smem_u32(smem_a));
    const uint64_t w1_desc0 = umma_smem_desc_sw128(smem_u32(smem_w1));
    const uint64_t w2_desc0 = umma_smem_desc_sw128(smem_u32(smem_w2));
    const uint64_t c_desc0 = umma_smem_desc_sw128(smem_u32(smem_c));
    int stage = 0, phase = 0, slot = 0, c = 0, local = 0;
    uint32_t ready_phase = 0;
    int prev_slot = -1, prev_g = 0, prev_local = 0, c2 = 0;   // c2: GEMM2 chunk counter (streamed W2 slot parity)
    // GEMM2 over one staged chunk of y (runs one chunk behind GEMM1 so that the tensor pipe never waits for the epilogue)
    auto gemm2 = [&](int pslot, int pg, int plocal) {
      const int a2 = plocal & 1;
      if (pg == 0) mbar_wait(&t2_empty[a2], ((plocal >> 1) & 1) ^ 1);
      mbar_wait(&c_ready[pslot], (ready_phase >> pslot) & 1);
      ready_phase ^= 1u << pslot;
      if (kStream) mbar_wait(w2_full, c2 & 1);
      ++c2;
      tcgen05_fence_after();
      if (leader) {
        const uint32_t d = tmem_acc2 + a2 * N2;
#pragma unroll
        for (int j = 0; j < CHUNK / 16; ++j) {
          const uint64_t a_desc = c_desc0 + (uint64_t)((pslot * SLOT + (j >> 2) * BOX) >> 4) + (uint64_t)((j & 3) * 2);
          const uint64_t b_desc = w2_desc0 + (uint64_t)((((kStream ? 0 : pg * (CHUNK / 64)) + (j >> 2)) * W2_BLOCK) >> 4) + (uint64_t)((j & 3) * 2);
          umma_f16_ss(d, a_desc, b_desc, idesc2, (pg | j) != 0 ? 1u : 0u);
        }
        umma_commit(&c_free[pslot]);
        if (kStream) umma_commit(w2_empty);
        if (pg == G - 1) umma_commit(&t2_full[a2]);
      }
      __syncwarp();
    };
    if (!kStream && blockIdx.x < p.m_tiles) mbar_wait(w_bar, 0);
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++local) {
      const int stage0 = stage;
      for (int g = 0; g < G; ++g, ++c) {
        const int acc = c & 1;
        mbar_wait(&t1_empty[acc], ((c >> 1) & 1) ^ 1);
        if (kStream) mbar_wait(w1_full, c & 1);
        tcgen05_fence_after();
        int st = stage0;
        for (int kb = 0; kb < p.nkb1; ++kb) {
          if (g == 0) {
            mbar_wait(&full_bar[stage], phase);
            tcgen05_fence_after();
            if (++stage == S) { stage = 0; phase ^= 1; }
          }
          if (leader) {
            const uint64_t a_desc = a_desc0 + (uint64_t)((st * A_STAGE) >> 4);
            const uint64_t b_desc = w1_desc0 + (uint64_t)((((kStream ? 0 : g * p.nkb1) + kb) * W1_BLOCK) >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_ss(tmem_base + acc * CHUNK, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc1, (kb | k) != 0 ? 1u : 0u);
            if (g == G - 1) umma_commit(&empty_bar[st]);   // both chunks have read this x k-block
            if (kb == p.nkb1 - 1) {
              umma_commit(&t1_full[acc]);
              if (kStream) umma_commit(w1_empty);
            }
          }
          __syncwarp();
          if (++st == S) st = 0;
        }
        if (prev_slot >= 0) gemm2(prev_slot, prev_g, prev_local);
        prev_slot = slot; prev_g = g; prev_local = local;
        if (++slot == R) slot = 0;
        if (g == 0 && local > 0 && ++slot == R) slot = 0;  // the slot of t(i - 1), staged between the two y chunks
      }
    }
    if (prev_slot >= 0) gemm2(prev_slot, prev_g, prev_local);
  } else if (warp < 2 + EPI_WARPS) {
    // ===================== epilogue: 4 TMEM lane quarters x 2 column halves =====================
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    int slot = 0, sphase = 0, c = 0, local = 0;
    // bias + (residual) + relu + 16-bit pack of 32 accumulator columns into the swizzled staging rows
    auto emit = [&](const uint32_t (&v)[32], const float* bias, uint32_t row_addr, int j0, bool add_res, bool relu) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + j * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + j * 8 + 4));
        float f[8] = {__uint_as_float(v[j * 8 + 0]) + b0.x, __uint_as_float(v[j * 8 + 1]) + b0.y,
                      __uint_as_float(v[j * 8 + 2]) + b0.z, __uint_as_float(v[j * 8 + 3]) + b0.w,
                      __uint_as_float(v[j * 8 + 4]) + b1.x, __uint_as_float(v[j * 8 + 5]) + b1.y,
                      __uint_as_float(v[j * 8 + 6]) + b1.z, __uint_as_float(v[j * 8 + 7]) + b1.w};
        const uint32_t addr = row_addr + ((uint32_t)((j0 + j) ^ (row & 7)) << 4);
        if (add_res) {
          uint4 rq;
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(rq.x), "=r"(rq.y), "=r"(rq.z), "=r"(rq.w) : "r"(addr));
          float r[8];
          unpack8<T>(rq, r);
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] += r[e];
        }
        if (relu) {
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
        }
        const uint4 o = pack8<T>(f);
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
      }
    };
    // this warp's part of the slot is staged: visible to the async proxy (TMA store, GEMM2), then signal - the epilogue
    // warps never synchronise with each other
    auto publish = [&](int s, bool is_y) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&staged[s]);
        if (is_y) mbar_arrive(&c_ready[s]);
      }
    };
    // phase order per pixel tile i:  y chunk 0 of i,  t of i - 1,  y chunks 1.. of i  - GEMM2 of tile i - 1 retires while
    // chunk 0 of tile i is in the epilogue, so the t phase never waits for the tensor pipe
    auto y_phase = [&](int tile, int g) {
      uint8_t* cbuf = smem_c + slot * SLOT;
      const int acc = c & 1;
      mbar_wait_short(&res_full[slot], sphase);
      mbar_wait_short(&t1_full[acc], (c >> 1) & 1);
      tcgen05_fence_after();
#pragma unroll 1
      for (int u = half; u < CHUNK / 32; u += 2) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + lane_addr + acc * CHUNK + u * 32, v);
        tmem_ld_wait();
        if (u + 2 >= CHUNK / 32) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&t1_empty[acc]);
        }
        emit(v, p.bias1 + g * CHUNK + u * 32, smem_u32(cbuf + (u >> 1) * BOX) + row * 128, (u & 1) * 4, p.has_res != 0, p.relu1 != 0);
      }
      publish(slot, true);
      ++c;
      if (++slot == R) { slot = 0; sphase ^= 1; }
    };
    auto t_phase = [&](int tl) {   // tl: CTA-local index of the pixel tile
      uint8_t* cbuf = smem_c + slot * SLOT;
      const int a2 = tl & 1;
      mbar_wait_short(&res_full[slot], sphase);
      mbar_wait_short(&t2_full[a2], (tl >> 1) & 1);
      tcgen05_fence_after();
#pragma unroll 1
      for (int u = half; u < N2 / 32; u += 2) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_acc2 + lane_addr + a2 * N2 + u * 32, v);
        tmem_ld_wait();
        if (u + 2 >= N2 / 32) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&t2_empty[a2]);
        }
        emit(v, p.bias2 + u * 32, smem_u32(cbuf + (u >> 1) * BOX) + row * 128, (u & 1) * 4, false, p.relu2 != 0);
      }
      publish(slot, false);
      if (++slot == R) { slot = 0; sphase ^= 1; }
    };
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++local) {
#pragma unroll 1
      for (int g = 0; g < G; ++g) {
        y_phase(tile, g);
        if (g == 0 && local > 0) t_phase(local - 1);
      }
    }
    if (local > 0) t_phase(local - 1);
  } else if (warp == 2 + EPI_WARPS + 1) {
    // ===================== W2 chunk producer (streamed variant only) =====================
    if (kStream && leader) {
      int n = 0;
      for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x)
        for (int g = 0; g < G; ++g, ++n) {
          mbar_wait(w2_empty, (n & 1) ^ 1);
          mbar_arrive_expect_tx(w2_full, (CHUNK / 64) * W2_BLOCK);
          for (int b = 0; b < CHUNK / 64; ++b)
            tma_load_2d(&p.tmW2, w2_full, smem_w2 + b * W2_BLOCK, (g * (CHUNK / 64) + b) * 64, 0);
        }
    }
  } else {
    // ===================== C-ring I/O: TMA stores of staged slots, then residual prefetch into the freed slot ==========
    // One thread owns both directions, so a slot is handed back as soon as its store has READ it (plus, for y chunks,
    // GEMM2's commit) and the residual of the phase R ahead starts loading at once: R - 1 phases of HBM latency hidden.
    if (leader && blockIdx.x < p.m_tiles) {
      pdl_wait();
      const int n_local = (p.m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      const int total = (G + 1) * n_local;
      // cursor over the phase order: sub 0 = y chunk 0 of tile i, 1 = t of tile i - 1, k >= 2 = y chunk k - 1 of tile i
      struct Cursor {
        int i = 0, sub = 0;
        __device__ bool is_y() const { return sub != 1; }
        __device__ int g() const { return sub == 0 ? 0 : sub - 1; }
        __device__ int local() const { return sub == 1 ? i - 1 : i; }
        __device__ void next(int n) {
          if (sub == 0) sub = i > 0 ? 1 : 2;
          else if (sub < G) ++sub;
          else { ++i; sub = i < n ? 0 : 1; }
        }
      } ld, stc;
      int k_load = 0, l_slot = 0, k_store = 0, s_slot = 0, s_phase = 0;
      uint32_t free_phase = 0, last_y = 0;
      while (k_store < total) {
        while (k_load < total && k_load < k_store + R) {
          if ((last_y >> l_slot) & 1) {   // previous tenant was a y chunk: GEMM2 must have consumed it
            mbar_wait(&c_free[l_slot], (free_phase >> l_slot) & 1);
            free_phase ^= 1u << l_slot;
          }
          last_y = (last_y & ~(1u << l_slot)) | ((ld.is_y() ? 1u : 0u) << l_slot);
          if (ld.is_y() && p.has_res) {
            const int row0 = ((int)blockIdx.x + ld.local() * (int)gridDim.x) * BLOCK_M;
            mbar_arrive_expect_tx(&res_full[l_slot], SLOT);
            uint8_t* cbuf = smem_c + l_slot * SLOT;
#pragma unroll
            for (int b = 0; b < CHUNK / 64; ++b)
              tma_load_2d(&p.tmR, &res_full[l_slot], cbuf + b * BOX, ld.g() * CHUNK + b * 64, row0);
          } else {
            mbar_arrive(&res_full[l_slot]);
          }
          ++k_load;
          ld.next(n_local);
          if (++l_slot == R) l_slot = 0;
        }
        mbar_wait(&staged[s_slot], s_phase);
        uint8_t* cbuf = smem_c + s_slot * SLOT;
        const int row0 = ((int)blockIdx.x + stc.local() * (int)gridDim.x) * BLOCK_M;
        if (stc.is_y()) {
#pragma unroll
          for (int b = 0; b < CHUNK / 64; ++b) tma_store_2d(&p.tmC, cbuf + b * BOX, stc.g() * CHUNK + b * 64, row0);
        } else {
#pragma unroll
          for (int b = 0; b < N2 / 64; ++b) tma_store_2d(&p.tmC2, cbuf + b * BOX, b * 64, row0);
        }
        bulk_commit();
        bulk_wait_read<0>();
        ++k_store;
        stc.next(n_local);
        if (++s_slot == R) { s_slot = 0; s_phase ^= 1; }
      }
      bulk_wait<0>();  // smem must stay valid until the last store has completed
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFnC)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// [rows, cols] 16-bit, cols contiguous; box = 64 cols x box_rows, SWIZZLE_128B
static int chain_tmap(CUtensorMap* m, const void* base, int precision, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  static EncodeTiledFnC enc = nullptr;
  if (enc == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      enc = reinterpret_cast<EncodeTiledFnC>(ptr);
  }
  if (enc == nullptr) { set_error("cuTensorMapEncodeTiled entry point not found"); return SEMDIFF_ERR_CUDA; }
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * 2};
  const cuuint32_t box[2] = {64, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = precision == SEMDIFF_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = enc(m, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("conv_chain: tensor map failed (%d) rows=%llu cols=%llu", (int)r, (unsigned long long)rows, (unsigned long long)cols);
    return SEMDIFF_ERR_CUDA;
  }
  return 0;
}

bool conv_chain_supported(const ConvShape& s1, const ConvShape& s2, bool has_res, int precision) {
  if (precision != SEMDIFF_BF16 && precision != SEMDIFF_FP16) return false;
  const bool pw1 = s1.kh == 1 && s1.kw == 1 && s1.stride == 1 && s1.pad == 0 && s1.pad_after() == 0;
  const bool pw2 = s2.kh == 1 && s2.kw == 1 && s2.stride == 1 && s2.pad == 0 && s2.pad_after() == 0;
  if (!pw1 || !pw2 || s2.cin != s1.cout || s2.cin2 != 0) return false;
  if (s1.n_img != s2.n_img || s1.H != s2.H || s1.W != s2.W) return false;
  if (s1.M() <= 0 || s1.M() >= ((int64_t)1 << 31) - chain::BLOCK_M) return false;
  if (s1.cout == 256) {   // resident weights
    if (s2.cout != 64 && s2.cout != 128) return false;
    if (s1.cin % 64 != 0 || s1.cin2 % 64 != 0 || s1.K() > 128) return false;
    return s1.cin2 == 0 || (s1.stride2 == 1 && !has_res);
  }
  // 512-channel stage, identity blocks: streamed weights (SEMDIFF_NO_CHAIN512=1 keeps these pairs apart)
  static const bool wide_ok = getenv("SEMDIFF_NO_CHAIN512") == nullptr;
  return wide_ok && s1.cout == 512 && s2.cout == 128 && s1.cin == 128 && s1.cin2 == 0 && has_res;
}

int conv_chain_prepare(ConvTcLaunch* L, const ConvPtrs& q1, const ConvShape& s1, const ConvPtrs& q2, const ConvShape& s2,
                       int precision) {
  using namespace chain;
  if (!conv_chain_supported(s1, s2, q1.res != nullptr, precision) || q2.res != nullptr || q2.in != q1.out) {
    set_error("conv_chain: unsupported pair (cin=%d+%d cout=%d -> cout=%d)", s1.cin, s1.cin2, s1.cout, s2.cout);
    return SEMDIFF_ERR_UNSUPPORTED;
  }
  ChainParams& p = *reinterpret_cast<ChainParams*>(L->params);
  memset(&p, 0, sizeof(p));
  const int N1 = s1.cout;
  p.g = N1 / CHUNK;
  const bool stream = p.g > 2;
  p.M = (int)s1.M();
  p.m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
  p.nkb_a = s1.cin / 64;
  p.nkb1 = s1.K() / 64;
  p.has_res = q1.res != nullptr;
  p.relu1 = s1.relu; p.relu2 = s2.relu;
  p.bias1 = q1.bias; p.bias2 = q2.bias;
  p.n2 = s2.cout;
  // shared-memory split: W1 / W2 (resident, or one chunk slot each when streamed), then the C ring (residual prefetch
  // depth), the rest to the x ring
  const int fixed = stream ? p.nkb1 * W1_BLOCK + (CHUNK / 64) * p.n2 * 128 : p.g * p.nkb1 * W1_BLOCK + (N1 / 64) * p.n2 * 128;
  const int avail = SMEM_LIMIT - 1024 - NUM_BARS * 8 - 16 - fixed;
  // measured on B200 (profiles/r1_chain_conv.md): with a residual, 4 slots (3 residual tiles in flight) reach the HBM
  // roofline where 3 lose 13 %; without one the ring only buffers stores
  int ring = p.has_res ? MAX_RING : 2;
  if (const char* e = getenv("SEMDIFF_CHAIN_RING")) { const int v = atoi(e); if (v >= 2 && v <= MAX_RING) ring = v; }
  int stages = (avail - ring * SLOT) / A_STAGE;
  while (stages < 2 * p.nkb1 && ring > 2) { --ring; stages = (avail - ring * SLOT) / A_STAGE; }
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < p.nkb1 || stages < 2) { set_error("conv_chain: shared memory budget"); return SEMDIFF_ERR_UNSUPPORTED; }
  p.ring = ring; p.stages = stages;
  p.smem_bytes = fixed + stages * A_STAGE + ring * SLOT + NUM_BARS * 8 + 16 + 1024;
  int rc = chain_tmap(&p.tmA, q1.in, precision, (uint64_t)p.M, (uint64_t)s1.cin, BLOCK_M);
  if (rc == 0 && s1.cin2 != 0) rc = chain_tmap(&p.tmA2, q1.in2, precision, (uint64_t)p.M, (uint64_t)s1.cin2, BLOCK_M);
  if (rc == 0) rc = chain_tmap(&p.tmW1, q1.w, precision, (uint64_t)N1, (uint64_t)s1.K(), CHUNK);
  if (rc == 0) rc = chain_tmap(&p.tmW2, q2.w, precision, (uint64_t)p.n2, (uint64_t)N1, (uint32_t)p.n2);
  if (rc == 0) rc = chain_tmap(&p.tmC, q1.out, precision, (uint64_t)p.M, (uint64_t)N1, BLOCK_M);
  if (rc == 0 && p.has_res) rc = chain_tmap(&p.tmR, q1.res, precision, (uint64_t)p.M, (uint64_t)N1, BLOCK_M);
  if (rc == 0) rc = chain_tmap(&p.tmC2, q2.out, precision, (uint64_t)p.M, (uint64_t)p.n2, BLOCK_M);
  if (rc != 0) return rc;
  L->block_n = CHUNK; L->a_mode = 201; L->precision = precision;
  return 0;
}

template <typename T, int N2, int G>
static int chain_launch_t(const ChainParams& p, cudaStream_t st) {
  static int configured[64] = {};
  auto kern = conv_chain_kernel<T, N2, G>;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (configured[dev] < p.smem_bytes) {
    SEMDIFF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, chain::SMEM_LIMIT));
    configured[dev] = chain::SMEM_LIMIT;
  }
  SEMDIFF_CUDA_OK(launch_pdl(kern, dim3(p.m_tiles < sms ? p.m_tiles : sms), dim3(chain::THREADS), (size_t)p.smem_bytes, st, p));
  return 0;
}

int conv_chain_launch(const ConvTcLaunch* L, cudaStream_t st) {
  const ChainParams& p = *reinterpret_cast<const ChainParams*>(L->params);
  const bool bf = L->precision == SEMDIFF_BF16;
  if (p.g == 4) return bf ? chain_launch_t<__nv_bfloat16, 128, 4>(p, st) : chain_launch_t<__half, 128, 4>(p, st);
  if (p.n2 == 64) return bf ? chain_launch_t<__nv_bfloat16, 64, 2>(p, st) : chain_launch_t<__half, 64, 2>(p, st);
  return bf ? chain_launch_t<__nv_bfloat16, 128, 2>(p, st) : chain_launch_t<__half, 128, 2>(p, st);
}

}  // namespace semdiff
