// Fused ResBlock1 pair for the low-channel decoder stages (C = 16 / 32), bf16 tensor-core mode:
//
//     x_new = conv2(leaky_relu(conv1(leaky_relu(x)) + b1)) + b2 + x            (modules/visinger/decoder.py:91-104)
//
// in ONE kernel: the intermediate `xt` never leaves the SM.  Un-fused, a pair moves 6 tensor passes through HBM
// (conv1: read + write, conv2: read + residual read + raw write + activated write); fused it moves 4.  These
// stages are HBM / per-tile-overhead bound (AI 24-176 FLOP/B), so this is where fusion pays.
//
// Per tile of V = 256 - (k-1) output rows (tile stride V; the k-1 halo rows of conv2 are recomputed by the neighbour):
//   producer (warp 0)      TMA: leaky_relu(x) rows [t0 - h2 - pad1, +256 + (k-1)*d1) -> A1 ring; W1, W2 resident
//   MMA front (warp 1)     conv1 over 2 blocks of 128 rows -> accumulator bank 0 in TMEM
//   mid warps (7..10)      TMEM -> + b1 -> leaky_relu -> zero rows outside [0, L) (conv2's zero padding) -> bf16 ->
//                          swizzled K-major tile in shared memory = conv2's A operand (never written to HBM)
//   MMA back (warp 6)      conv2 from that tile through row-shifted descriptors -> accumulator bank 1
//   epilogue (warps 2..5   the ordinary conv_tc epilogue: + b2 + residual (+ running resblock sum, x 1/3) -> raw and
//             and 11..14)  leaky_relu'd bf16 -> one bulk TMA store each, V rows per tile.  Two sets of 4 warps, one
//                          128-row block of the tile each: ncu showed one set (one warp per SM sub-partition) busy
//                          ~80 % of the time with everything upstream waiting on it.
// All five stages run concurrently on different tiles (two-deep rings between them).
#pragma once
#include "conv_tc.cuh"

namespace vsg {

namespace tc {
constexpr int kPairThreads = 480;   // 15 warps
__device__ __forceinline__ void mid_bar_sync() { asm volatile("bar.sync 3, 128;" ::: "memory"); }
}  // namespace tc

// conv1's epilogue: accumulator -> activated bf16 A-operand tile of conv2.
template <int C>
__device__ __forceinline__ void pair_mid_epilogue(const ConvTC& p1, const ConvTC& p2, uint32_t smem_base, uint32_t bar_base,
                                                  uint32_t tmem_base, int warp, int lane) {
  using namespace tc;
  const int quarter = warp & 3;
  const bool leader = (warp == 7 && lane == 0);
  const uint32_t bar1 = bar_base + 8u * (uint32_t)p1.bar_slot0, bar2 = bar_base + 8u * (uint32_t)p2.bar_slot0;
  const uint32_t acc_full0 = bar1 + 8u * kBarAccFull, acc_empty0 = bar1 + 8u * kBarAccEmpty;
  const uint32_t t_full0 = bar2 + 8u * kBarAFull, t_empty0 = bar2 + 8u * kBarAEmpty;
  const uint32_t t_base = smem_base + p2.a_off;
  const int mb = p1.mb, total_tiles = p1.total_tiles, L = p1.Lq, stages_t = p2.stages_a;
  const float slope = p1.slope;
  int* const error_flag = p1.error_flag;
  const uint32_t swz_mask = C >= 64 ? 7u : C >= 32 ? 3u : 1u;
  float bias_r[C];
#pragma unroll
  for (int i = 0; i < C; i += 4) {
    const float4 bv = p1.bias ? __ldg(reinterpret_cast<const float4*>(p1.bias + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
    bias_r[i] = bv.x; bias_r[i + 1] = bv.y; bias_r[i + 2] = bv.z; bias_r[i + 3] = bv.w;
  }
  int as = 0, ts = 0;
  uint32_t pacc = 0, pt = 0;
  TileIter it;
  it.init((int)blockIdx.x, (int)gridDim.x, 1, p1.m_tiles_per_b);
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it.next()) {
    const int t_row0 = it.mt * p1.tile_stride + p1.t_row_off;   // time of row 0 of the intermediate tile
    mbar_wait(acc_full0 + 8u * as, pacc, error_flag);
    fence_after_sync();
    mbar_wait(t_empty0 + 8u * ts, pt ^ 1, error_flag);          // conv2 of the tile that last used this buffer is done
    const uint32_t taddr0 = tmem_base + (uint32_t)p1.tmem_col0 + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * mb * C);
    const uint32_t tbuf = t_base + (uint32_t)ts * p2.a_stage_bytes;
    for (int bi = 0; bi < mb; ++bi) {
      const int srow = bi * 128 + quarter * 32 + lane;
      float v[C];
      {
        uint32_t r[C];
#pragma unroll
        for (int i = 0; i < C / 16; ++i) tmem_ld16_nowait(taddr0 + (uint32_t)(bi * C + i * 16), r + 16 * i);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < C; ++i) v[i] = __uint_as_float(r[i]);
      }
      if (bi == mb - 1) {
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty0 + 8u * as);
      }
      const int t = t_row0 + srow;
      const bool inside = (t >= 0) && (t < L);                  // outside the utterance conv2 sees zeros, not conv1(pad)
#pragma unroll
      for (int i = 0; i < C; ++i) {
        const float a = v[i] + bias_r[i];
        v[i] = inside ? fmaxf(a, a * slope) : 0.f;
      }
      stage_out<C>(v, tbuf, (uint32_t)srow * (C * 2), swz_mask, 1, 0u);
    }
    fence_async_smem();          // generic-proxy writes -> visible to the tensor core's async-proxy reads
    mid_bar_sync();
    if (leader) mbar_arrive(t_full0 + 8u * ts);
    if (++ts == stages_t) { ts = 0; pt ^= 1; }
    if (++as == 2) { as = 0; pacc ^= 1; }
  }
}

template <int C, int SIG = EPI_SIG_GENERIC>
__global__ void __launch_bounds__(tc::kPairThreads, 1)
pair_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
               const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmAdd0,
               const __grid_constant__ CUtensorMap tmAdd1, const __grid_constant__ CUtensorMap tmRaw,
               const __grid_constant__ CUtensorMap tmAct, const ConvTC p1, const ConvTC p2) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + p2.bar_off;
  const uint32_t tmem_slot = bar_base + 8u * (2 * kNumBars);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t bar1 = bar_base + 8u * (uint32_t)p1.bar_slot0, bar2 = bar_base + 8u * (uint32_t)p2.bar_slot0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmW1); prefetch_tmap(&tmW2);
    if (p2.has_add0) prefetch_tmap(&tmAdd0);
    if (p2.has_add1) prefetch_tmap(&tmAdd1);
    if (p2.has_raw) prefetch_tmap(&tmRaw);
    if (p2.has_act) prefetch_tmap(&tmAct);
    for (int bank = 0; bank < 2; ++bank) {
      const uint32_t bb = bank ? bar2 : bar1;
      for (int s = 0; s < kMaxStages; ++s) {
        mbar_init(bb + 8u * (kBarAFull + s), 1); mbar_init(bb + 8u * (kBarAEmpty + s), 1);
        mbar_init(bb + 8u * (kBarWFull + s), 1); mbar_init(bb + 8u * (kBarWEmpty + s), 1);
      }
      // bank 0 accumulators are drained by the 4 mid warps, bank 1 by the 4 epilogue warps
      for (int s = 0; s < 2; ++s) { mbar_init(bb + 8u * (kBarAccFull + s), 1); mbar_init(bb + 8u * (kBarAccEmpty + s), (bank && p2.epi_sets > 1) ? 8 : 4); }
      for (int s = 0; s < kMaxAddBufs; ++s) mbar_init(bb + 8u * (kBarAdd + s), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, p2.tmem_cols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_ptr, 0);
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (warp == 0 && lane == 0) {   // both convs keep their weights resident: one TMA burst, independent of the previous kernel
    mbar_expect_tx(bar1 + 8u * kBarWFull, (uint32_t)p1.n_wtiles * p1.w_box_bytes);
    for (int j = 0; j < p1.ktaps; ++j)
      tma_load_2d(smem_base + p1.w_off + (uint32_t)j * p1.w_stage_bytes, &tmW1, bar1 + 8u * kBarWFull, 0, j * p1.CoutT);
    mbar_expect_tx(bar2 + 8u * kBarWFull, (uint32_t)p2.n_wtiles * p2.w_box_bytes);
    for (int j = 0; j < p2.ktaps; ++j)
      tma_load_2d(smem_base + p2.w_off + (uint32_t)j * p2.w_stage_bytes, &tmW2, bar2 + 8u * kBarWFull, 0, j * p2.CoutT);
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    // ===================== TMA producer: activated input tiles for conv1 =====================
    {
      int sa = 0;
      uint32_t pa = 0;
      const uint32_t a_base = smem_base + p1.a_off;
      TileIter it;
      it.init((int)blockIdx.x, (int)gridDim.x, 1, p1.m_tiles_per_b);
      for (int tile = blockIdx.x; tile < p1.total_tiles; tile += gridDim.x, it.next()) {
        const int row0 = it.mt * p1.tile_stride + p1.in_off0;
        mbar_wait(bar1 + 8u * (kBarAEmpty + sa), pa ^ 1, p1.error_flag);
        if (elect_one()) {
          mbar_expect_tx(bar1 + 8u * (kBarAFull + sa), (uint32_t)p1.a_n_boxes * p1.a_box_bytes);
          for (int bx = 0; bx < p1.a_n_boxes; ++bx)
            tma_load_3d(a_base + sa * p1.a_stage_bytes + bx * p1.a_box_bytes, &tmA, bar1 + 8u * (kBarAFull + sa), 0,
                        row0 + bx * p1.a_box_rows, it.b);
        }
        if (++sa == p1.stages_a) { sa = 0; pa ^= 1; }
      }
    }
  } else if (warp == 1) {
    conv_tc_mma_loop<true, true, C / 16>(p1, smem_base + p1.a_off, smem_base + p1.w_off, bar_base, tmem_base, 0, 1);
  } else if (warp == 6) {
    conv_tc_mma_loop<true, true, C / 16>(p2, smem_base + p2.a_off, smem_base + p2.w_off, bar_base, tmem_base, 0, 1);
  } else if (warp >= 7 && warp <= 10) {
    pair_mid_epilogue<C>(p1, p2, smem_base, bar_base, tmem_base, warp, lane);
  } else {
    const int set = warp >= 11 ? 1 : 0;
    if (set < p2.epi_sets)
      conv_tc_epilogue<C, EPI_TC_LINEAR, 1, SIG>(tmAdd0, tmAdd1, tmRaw, tmAct, p2, smem_base, bar_base, tmem_base, warp, lane, set, 2);
  }

  fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc(tmem_base, p2.tmem_cols);
  }
}

}  // namespace vsg
