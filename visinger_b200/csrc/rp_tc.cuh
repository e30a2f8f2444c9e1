// Row-packed whole-ResBlock1 kernel for the low-channel decoder stages (C = 16 / 32 [/ 64]), bf16 tensor-core mode:
//
//     for (c1, c2, d) in pairs:   x = c2(leaky_relu(c1_d(leaky_relu(x)))) + x          (modules/visinger/decoder.py:91-104)
//
// -- every convolution of one ResBlock1 in ONE kernel, like rb_tc.cuh, but the GEMM is re-shaped for what the tensor
// pipe and the barrier round trips cost at these widths.  A tcgen05.mma of M = 128 takes 40 cycles at N = 16 and 48 at
// N = 64 (tools/mma_bench2.cu), so a C = 16 convolution issued tap by tap runs the pipe at a quarter of its N = 64 rate,
// and a work unit of 128 time steps x 16 channels is so small that the MMA <-> epilogue hand-over (commit, mbarrier,
// tcgen05.ld, fence, arrive) dominates: rb_tc.cuh measured 8.6 % tensor-pipe activity with both sides waiting on each
// other half of the time.  Here S = 64 / C consecutive time steps are ONE 128-byte row of the channels-last tensor
// (which is just a different view of the same bytes: [L, C] == [L / S, 64]):
//   * GEMM rows are groups of S time steps: a 128-row block is 128 S time steps, its accumulator 64 columns
//     (column = sub-step * C + channel), a tile is up to 4 blocks = 2048 / 1024 time steps at C = 16 / 32;
//   * a dilation-1 convolution becomes Conv1d(64 -> 64) over rows with block-Toeplitz weights (pre-packed, see
//     pack_conv_rowpacked): one N = 64 MMA per 16-wide K slice of the k + S - 1 input sub-steps a row depends on,
//     instead of S k N = C MMAs -- 6 / 10 / 14 instead of 12 / 28 / 44 at C = 16, k = 3 / 7 / 11.  Consecutive slices
//     are consecutive 32-byte pieces of the tile and consecutive 2 KB weight tiles, so the issue loop is plain adds;
//   * a dilated convolution keeps its compact [tap][C][C] weights: per (tap, sub-step) an N = C MMA whose A operand
//     starts (sub-step + (tap - centre) * dilation) time steps into the row-packed tile (a byte offset of the UMMA
//     descriptor; the swizzle is a function of absolute shared-memory address bits, so any 32-byte shift is legal) and
//     whose accumulator is the sub-step's column range;
//   * the epilogue hand-over happens per (block, convolution): 4x fewer, 4x larger units than rb_tc.cuh at C = 16.
//   * the residual stream x IS the accumulator of every second convolution: x lives in tensor memory in fp32 and
//     c2's MMAs accumulate straight into it (x += W2 * leaky_relu(...)), so `x = xt + x` (decoder.py:102) costs no
//     instruction at all; the biases of the c2 convolutions are kept out of tensor memory and added as a running sum
//     (bias[c] of an odd c is b2_0 + ... + b2_q) wherever x is read;
//   * so both epilogues are the same ~60 instructions per 16 values: tcgen05.ld, + bias, pack to bf16, leaky_relu on
//     packed pairs, two 16-byte shared-memory stores.  The epilogue's instruction issue is what bounds these stages
//     (rb_tc.cuh and the first version of this kernel spent ~170 instructions per 16 values and ran at the same
//     ~400 us per resblock whatever the tap count).
// Everything else follows rb_tc.cuh: the activated stream ping-pongs between two swizzled tiles P and Q in shared
// memory, weights stream through a TMA ring (one stage = four
// Toeplitz K slices or a group of compact taps; blocks are the OUTER loop so block b's accumulator is complete
// -- and its epilogue running -- while the tensor pipe works on block b + 1; a convolution's stages stay resident until
// its last block has used them), a halo of H = sum_q ((k-1)/2 (d_q + 1)) time steps per side is recomputed by the
// neighbour tile.
//
// Warp roles (640 threads): 0 TMA producer (input tiles and the weight ring), 1 TMEM allocator + MMA issuer (converged
// warp, elect_one per instruction), 2..3 idle, 4..19 epilogue: warp % 4 = TMEM lane quarter, set (warp - 4) / 4 owns
// block `set` of a 4-block tile (all 64 columns of its rows; the four blocks are in flight together, staggered by the
// MMA order), or shares a block with other sets when the tile has fewer blocks.
//
// X3 = true is the same kernel in the split-bf16 arithmetic of the bf16x3 mode (fp32-tolerance results on tcgen05): every
// activation is two bf16 planes (hi = bf16(v), lo = bf16(v - hi)), every weight tile comes as W_hi and W_lo, and every
// product is three MMAs issued small-first (lo * W_hi, hi * W_lo, hi * W_hi).  The input and the running resblock sum
// are PLANAR two-plane tensors ([2][B, L, C]: each plane loads through the one-plane row-packed tensor map; planes
// interleaved per time step cannot -- a box narrower than the swizzle span is padded to it in shared memory -- and a
// first version that staged such rows unswizzled and re-laid them in the epilogue lost 10 k cycles per tile to the
// serialised load -> re-lay -> first MMA chain); the stage output keeps the decoder's rows [hi (C) | lo (C)].  The residual stream is still the fp32 accumulator in
// tensor memory -- more exact than the two-plane stream of the per-convolution kernels.  Four activation tiles (P / Q x
// hi / lo) leave room for 2-block tiles only and not for a whole convolution's weights, so the weight ring (16 KB
// stages: [W_hi | W_lo]) is streamed per BLOCK instead of per convolution (L2 -> shared memory twice per tile).
#pragma once
#include "conv_tc.cuh"
#include "rb_tc.cuh"

namespace vsg {

constexpr int kRpMaxConvs = 8;
constexpr int kRpMaxBlocks = 4;
constexpr int kRpMaxWStages = 12;
constexpr uint32_t kRpStageBytes = 8192;

struct RpTC {
  int B, L;                       // utterances, time steps per utterance (L % S == 0)
  int k;                          // taps of every convolution of the block
  int n_convs;                    // 2 * n_pairs: c1_0, c2_0, c1_1, c2_1, ...
  int dil[kRpMaxConvs];           // dilation per convolution (c2: 1)
  uint32_t packed_mask;           // bit c: convolution c (dilation 1) runs in the row-packed block-Toeplitz form
  int n_k, packed_stages;         // 16-wide K slices of the row-packed form ((k + S - 1) * C / 16), ring stages they fill (4 each)
  int tps, direct_stages;         // compact taps per ring stage, ring stages of one direct convolution
  int mb;                         // 128-row blocks per tile
  int spb;                        // epilogue warp sets per block (1, 2 or 4): 4 / spb blocks are drained concurrently
  int H, V;                       // halo time steps per side, valid time steps per tile (both multiples of S)
  int m_tiles_per_b, total_tiles;
  int n_wst;                      // weight ring stages
  uint32_t margin_bytes;          // zero margin before / after the tile inside P and Q (>= the largest tap reach)
  uint32_t buf_bytes;             // margin + 128 * 128 * mb + margin (X3: the lo plane of P / Q follows its hi plane at + buf_bytes)
  uint32_t p_off, q_off, w_off, bar_off, bias_off;   // shared-memory carve-up relative to the 1024-aligned base
  uint32_t tmem_cols;
  const float* bias[kRpMaxConvs];  // even c: bias of c1_q; odd c: b2_0 + ... + b2_q (the residual stream's running bias)
  const __nv_bfloat16* add1;      // running resblock sum [B, L, C] or null (read for the tile's valid steps only)
                                  // (X3: the input, add1 and out_raw are planar two-plane tensors, the lo plane
                                  // plane_stride elements after the hi plane; out_act has rows [hi (C) | lo (C)])
  long long plane_stride;
  __nv_bfloat16* out_raw;         // (x_out + add1) * scale as bf16, or null
  __nv_bfloat16* out_act;         // leaky_relu of the same, or null
  float* out_f32;                 // fp32 copy (parity hook), or null
  float scale, slope;
  int* error_flag;
  uint32_t* trace;                // tuning aid: CTA 0 logs clock() at its pipeline events ([5 roles][1024] words), or null
};

struct RpMaps { CUtensorMap w[kRpMaxConvs]; CUtensorMap add1; CUtensorMap a_lo, add1_lo; };   // (*_lo: X3, the lo planes)

namespace tc {

constexpr int kRpBarAccFull = 0;                                  // [kRpMaxBlocks] tcgen05.commit
constexpr int kRpBarReady = kRpMaxBlocks;                         // [2][kRpMaxBlocks] even / odd convolution steps (see rb_tc.cuh)
constexpr int kRpBarWFull = 3 * kRpMaxBlocks;                     // [kRpMaxWStages]
constexpr int kRpBarWEmpty = kRpBarWFull + kRpMaxWStages;         // [kRpMaxWStages]
constexpr int kRpBarAFull = kRpBarWEmpty + kRpMaxWStages;
constexpr int kRpBarPFree = kRpBarAFull + 1;
constexpr int kRpNumBars = kRpBarPFree + 1;
constexpr int kRpEpiWarps = 16;
constexpr int kRpThreads = (4 + kRpEpiWarps) * 32;                // 640

}  // namespace tc

// NSETS = 2: the two-CTAs-per-SM form -- 8 epilogue warps, tiles of <= 2 blocks (256 accumulator columns), weights
// streamed per block like X3: two independent hand-over chains share an SM's tensor pipe and fill each other's bubbles.
template <int C, bool X3 = false, int NSETS = 4>
__global__ void __launch_bounds__((4 + 4 * NSETS) * 32, NSETS == 2 ? 2 : 1)
rp_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ RpMaps wm, const RpTC p) {
  using namespace tc;
  constexpr int S = 64 / C;                      // time steps per 128-byte row
  constexpr int KK = C / 16;                     // 16-wide K slices per time step
  constexpr uint32_t kStep16 = (uint32_t)(C * 2) / 16u;     // 16-byte units per time step
  constexpr uint32_t kSlotBytes = C * C * 2 > 1024 ? C * C * 2 : 1024;   // one compact tap tile inside a ring stage
  constexpr uint32_t kTapBytes = X3 ? 2 * kSlotBytes : kSlotBytes;      // X3: [W_hi tile | W_lo tile] per tap
  constexpr uint32_t kStageBytes = X3 ? 2 * kRpStageBytes : kRpStageBytes;   // X3: [W_hi 8 KB | W_lo 8 KB]
  constexpr int NPL = X3 ? 2 : 1;                // bf16 planes per activation tile
  constexpr bool STREAM = X3 || NSETS == 2;      // the weight ring is streamed per block (not held for a whole convolution)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + p.bar_off;
  auto bar = [&](int slot) { return bar_base + 8u * (uint32_t)slot; };
  const uint32_t tmem_slot = bar_base + 8u * kRpNumBars;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_gen + p.bar_off + 8u * kRpNumBars);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int mb = p.mb, n_convs = p.n_convs, k = p.k, total_tiles = p.total_tiles, n_wst = p.n_wst;
  int* const error_flag = p.error_flag;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    for (int c = 0; c < n_convs; ++c) prefetch_tmap(&wm.w[c]);
    if (p.add1) prefetch_tmap(&wm.add1);
    if (X3) { prefetch_tmap(&wm.a_lo); if (p.add1) prefetch_tmap(&wm.add1_lo); }
    for (int b = 0; b < kRpMaxBlocks; ++b) {
      mbar_init(bar(kRpBarAccFull + b), 1);
      const uint32_t warps_per_block = 4u * (uint32_t)p.spb;                  // epilogue warps that drain one block
      mbar_init(bar(kRpBarReady + b), warps_per_block);
      mbar_init(bar(kRpBarReady + kRpMaxBlocks + b), warps_per_block);
    }
    for (int s = 0; s < kRpMaxWStages; ++s) { mbar_init(bar(kRpBarWFull + s), 1); mbar_init(bar(kRpBarWEmpty + s), 1); }
    mbar_init(bar(kRpBarAFull), 1);
    mbar_init(bar(kRpBarPFree), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  {  // zero margins of P and Q: rows the taps reach beyond the tile; never written afterwards
    const uint32_t margin16 = p.margin_bytes / 16u;
    const uint32_t tail16 = (p.margin_bytes + 16384u * (uint32_t)mb) / 16u;
    for (int pl = 0; pl < NPL; ++pl) {
      uint4* pg = reinterpret_cast<uint4*>(smem_gen + p.p_off + (uint32_t)pl * p.buf_bytes);
      uint4* qg = reinterpret_cast<uint4*>(smem_gen + p.q_off + (uint32_t)pl * p.buf_bytes);
      for (uint32_t i = threadIdx.x; i < margin16; i += blockDim.x) {
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        pg[i] = z; qg[i] = z;
        pg[tail16 + i] = z; qg[tail16 + i] = z;
      }
    }
    fence_async_smem();
    // biases -> shared memory [n_convs][C] (the epilogue re-reads them every convolution step)
    float* bs = reinterpret_cast<float*>(smem_gen + p.bias_off);
    for (int i = threadIdx.x; i < n_convs * C; i += blockDim.x) bs[i] = __ldg(p.bias[i / C] + (i % C));
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_ptr, 0);
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const uint32_t p_base = smem_base + p.p_off, q_base = smem_base + p.q_off, w_base = smem_base + p.w_off;
  const uint32_t tile_off = p.margin_bytes;                      // byte offset of tile row 0 inside P / Q
  const uint32_t lo_delta = p.buf_bytes;                         // X3: the lo plane of a tile, relative to its hi plane

  if (warp == 0) {
    // ===================== TMA producer: input tiles -> P, weight ring in the order the issuers consume it =====================
    asm volatile("griddepcontrol.wait;" ::: "memory");           // the previous kernel produced the input
    TileIter it;
    it.init((int)blockIdx.x, (int)gridDim.x, 1, p.m_tiles_per_b);
    uint32_t n_tile = 0;
    uint32_t st = 0, eph = 0;                                     // eph bit s: parity of stage s's next "empty" completion
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it.next(), ++n_tile) {
      // (the previous tile's weight loads were all issued above, so waiting for P here cannot starve the issuers)
      if (n_tile > 0) mbar_wait(bar(kRpBarPFree), (n_tile - 1) & 1u, error_flag);   // P no longer read by the previous tile
      const int row0 = (it.mt * p.V - p.H) / S;                  // exact: V and H are multiples of S
      if (elect_one()) {
        mbar_expect_tx(bar(kRpBarAFull), 16384u * (uint32_t)(mb * NPL));
        for (int bx = 0; bx < mb; ++bx) {
          tma_load_3d(p_base + tile_off + 16384u * (uint32_t)bx, &tmA, bar(kRpBarAFull), 0, row0 + 128 * bx, it.b);
          if (X3) tma_load_3d(p_base + lo_delta + tile_off + 16384u * (uint32_t)bx, &wm.a_lo, bar(kRpBarAFull), 0, row0 + 128 * bx, it.b);
        }
      }
      // L2 prefetch: the running sum this tile's last convolution will add, and the NEXT tile's input (its TMA load can
      // only be issued once P is free; from L2 it lands in a quarter of the HBM latency)
      if (p.add1) {
        if (elect_one())
          for (int bx = 0; bx < mb; ++bx) {
            tma_prefetch_3d(&wm.add1, 0, row0 + 128 * bx, it.b);
            if (X3) tma_prefetch_3d(&wm.add1_lo, 0, row0 + 128 * bx, it.b);
          }
      }
      if (tile + (int)gridDim.x < total_tiles) {
        TileIter nx = it;
        nx.next();
        const int nrow0 = (nx.mt * p.V - p.H) / S;
        if (elect_one())
          for (int bx = 0; bx < mb; ++bx) {
            tma_prefetch_3d(&tmA, 0, nrow0 + 128 * bx, nx.b);
            if (X3) tma_prefetch_3d(&wm.a_lo, 0, nrow0 + 128 * bx, nx.b);
          }
      }
      for (int c = 0; c < n_convs; ++c) {
        const bool packed = (p.packed_mask >> c) & 1u;
        const int ns = packed ? p.packed_stages : p.direct_stages;
        const int n_pass = STREAM ? mb : 1;                        // STREAM: the ring is streamed once per block
        for (int ii = 0; ii < ns * n_pass; ++ii) {
          const int i = STREAM ? ii % ns : ii;
          mbar_wait(bar(kRpBarWEmpty + (int)st), ((eph >> st) & 1u) ^ 1u, error_flag);
          eph ^= 1u << st;
          const uint32_t dst = w_base + st * kStageBytes;
          if (packed) {
            if (elect_one()) {
              mbar_expect_tx(bar(kRpBarWFull + (int)st), kStageBytes);
              tma_load_2d(dst, &wm.w[c], bar(kRpBarWFull + (int)st), 0, i * 64);
              if (X3) tma_load_2d(dst + kRpStageBytes, &wm.w[c], bar(kRpBarWFull + (int)st), 64, i * 64);
            }
          } else {
            const int j0 = i * p.tps, j1 = min(k, j0 + p.tps);
            if (elect_one()) {
              mbar_expect_tx(bar(kRpBarWFull + (int)st), (uint32_t)(j1 - j0) * (uint32_t)(C * C * 2 * NPL));
              for (int j = j0; j < j1; ++j) {
                tma_load_2d(dst + (uint32_t)(j - j0) * kTapBytes, &wm.w[c], bar(kRpBarWFull + (int)st), 0, j * C);
                if (X3) tma_load_2d(dst + (uint32_t)(j - j0) * kTapBytes + kSlotBytes, &wm.w[c], bar(kRpBarWFull + (int)st), C, j * C);
              }
            }
          }
          st = (st + 1 == (uint32_t)n_wst) ? 0u : st + 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp, elect_one per instruction) =====================
    // Blocks are the outer loop of a convolution: block b's accumulator is complete -- and its epilogue running --
    // while the tensor pipe works on block b + 1.  Per block the issuer executes one barrier poll, the MMAs in groups
    // of four with distinct descriptor registers (back-to-back UTCHMMA) and one commit.
    const uint32_t idesc_row = make_idesc_bf16(128, 64), idesc_dir = make_idesc_bf16(128, (uint32_t)C);
    constexpr uint32_t kDirCode = C == 64 ? 2u : C == 32 ? 4u : 6u;
    constexpr uint32_t hi_row = ((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);              // SWIZZLE_128B, 8 rows = 1 KB
    constexpr uint32_t hi_dir = (((8u * C * 2u) >> 4) & 0x3FFFu) | (1u << 14) | (kDirCode << 29);
    auto mk_row = [&](uint32_t lo) { return ((uint64_t)hi_row << 32) | (uint64_t)lo; };
    auto mk_dir = [&](uint32_t lo) { return ((uint64_t)hi_dir << 32) | (uint64_t)lo; };
    constexpr uint32_t lo_flag = 1u << 16, blk16 = 16384u >> 4, stage16 = kStageBytes >> 4, slot16 = kSlotBytes >> 4;
    constexpr uint32_t tap16 = kTapBytes >> 4, wlo16 = kRpStageBytes >> 4;   // X3: W_lo follows W_hi inside a tap slot / a packed stage
    const uint32_t alo16 = lo_delta >> 4;                                    // X3: lo plane of the A tile
    const uint32_t p16 = (p_base + tile_off) >> 4, q16 = (q_base + tile_off) >> 4, w16_0 = w_base >> 4;
    const int n_k = p.n_k, packed_stages = p.packed_stages, tps = p.tps, direct_stages = p.direct_stages, cen = (k - 1) / 2;
    const int full_groups = n_k >> 2, tail = n_k & 3;             // packed form: K slices in groups of 4 = one ring stage
    const uint32_t packed_mask = p.packed_mask;
    const int d0 = p.dil[0], d1 = p.dil[2], d2 = p.dil[4], d3 = p.dil[6];
    const uint32_t bar_acc = bar(kRpBarAccFull), bar_ready = bar(kRpBarReady), bar_wfull = bar(kRpBarWFull),
                   bar_wempty = bar(kRpBarWEmpty), bar_afull = bar(kRpBarAFull), bar_pfree = bar(kRpBarPFree);
    uint32_t ws0 = 0, wph = 0, n = 0, n_tile = 0;                 // wph bit s: parity of stage s's next "full" completion
    // one product of the GEMM: a single MMA, or (X3) the three plane products, smallest first
    auto mma_row = [&](uint32_t d, uint32_t a, uint32_t b, uint32_t acc) {
      if (X3) {
        if (elect_one()) umma_bf16(d, mk_row(a + alo16), mk_row(b), idesc_row, acc);
        if (elect_one()) umma_bf16(d, mk_row(a), mk_row(b + wlo16), idesc_row, 1u);
        if (elect_one()) umma_bf16(d, mk_row(a), mk_row(b), idesc_row, 1u);
      } else {
        if (elect_one()) umma_bf16(d, mk_row(a), mk_row(b), idesc_row, acc);
      }
    };
    auto mma_dir = [&](uint32_t d, uint32_t a, uint32_t b, uint32_t acc) {
      if (X3) {
        if (elect_one()) umma_bf16(d, mk_row(a + alo16), mk_dir(b), idesc_dir, acc);
        if (elect_one()) umma_bf16(d, mk_row(a), mk_dir(b + slot16), idesc_dir, 1u);
        if (elect_one()) umma_bf16(d, mk_row(a), mk_dir(b), idesc_dir, 1u);
      } else {
        if (elect_one()) umma_bf16(d, mk_row(a), mk_dir(b), idesc_dir, acc);
      }
    };
    uint32_t* const trace = (p.trace && blockIdx.x == 0 && lane == 0) ? p.trace : nullptr;
    uint32_t ntr = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++n_tile) {
      for (int c = 0; c < n_convs; ++c, ++n) {
        const int pr = c >> 1;
        const int d = (c & 1) ? 1 : (pr == 0 ? d0 : pr == 1 ? d1 : pr == 2 ? d2 : d3);
        const bool packed = (packed_mask >> c) & 1u;
        const uint32_t ns = (uint32_t)(packed ? packed_stages : direct_stages);
        const uint32_t prev_par = ((n - 1) >> 1) & 1u;
        const uint32_t bar_prev = bar_ready + (((n - 1) & 1u) ? 8u * kRpMaxBlocks : 0u);
        // c1 accumulates from zero into T (columns [0, 64 mb)); c2 accumulates ON TOP of the residual stream X
        // (columns [64 mb, 128 mb)): x += W2 * a
        const uint32_t acc0 = (c & 1) ? 1u : 0u;
        uint32_t blk_lo = (lo_flag | ((c & 1) ? q16 : p16)) - (packed ? (uint32_t)(cen * KK * 2) : 0u);
        uint32_t d_tmem = tmem_base + ((c & 1) ? (uint32_t)(64 * mb) : 0u);
        if (trace && ntr < 1022) trace[ntr++] = (uint32_t)clock();
        uint32_t st = ws0;
        for (int b = 0; b < mb; ++b) {
          // block b reads blocks b-1 .. b+1 of the previous convolution's output and overwrites accumulator b: all
          // released through `ready` of the previous step (b-1 and b were waited for by the previous iterations)
          if (c == 0) {
            if (b == 0) mbar_wait(bar_afull, n_tile & 1u, error_flag);
            if (n > 0) mbar_wait(bar_prev + 8u * b, prev_par, error_flag);
          } else {
            if (b == 0) mbar_wait(bar_prev, prev_par, error_flag);
            if (b + 1 < mb) mbar_wait(bar_prev + 8u * (b + 1), prev_par, error_flag);
          }
          fence_after_sync();
          if (!STREAM) st = ws0;                                   // (STREAM: the ring is streamed per block and just runs on)
          if (packed) {
            // K slice m = 0 .. n_k - 1: input sub-step -cen + m / KK, channels 16 (m % KK) ..; consecutive slices are
            // consecutive 32-byte pieces of the row-packed tile and of the stage's [64 x 64] SWIZZLE_128B weight tile
            uint32_t a_lo = blk_lo;
            uint32_t accumulate = acc0;
            for (int g = 0; g < full_groups; ++g) {
              if (STREAM || b == 0) {
                mbar_wait(bar_wfull + 8u * st, (wph >> st) & 1u, error_flag);
                wph ^= 1u << st;
                fence_after_sync();
              }
              const uint32_t b_lo = lo_flag | (w16_0 + st * stage16);
#pragma unroll
              for (int u = 0; u < 4; ++u) mma_row(d_tmem, a_lo + 2u * u, b_lo + 2u * u, u > 0 ? 1u : accumulate);
              accumulate = 1u;
              a_lo += 8u;
              if (STREAM || b == mb - 1) {
                if (elect_one()) umma_commit(bar_wempty + 8u * st);
              }
              st = (st + 1 == (uint32_t)n_wst) ? 0u : st + 1;
            }
            if (tail) {
              if (STREAM || b == 0) {
                mbar_wait(bar_wfull + 8u * st, (wph >> st) & 1u, error_flag);
                wph ^= 1u << st;
                fence_after_sync();
              }
              const uint32_t b_lo = lo_flag | (w16_0 + st * stage16);
              mma_row(d_tmem, a_lo, b_lo, accumulate);
              if (tail > 1) mma_row(d_tmem, a_lo + 2u, b_lo + 2u, 1u);
              if (tail > 2) mma_row(d_tmem, a_lo + 4u, b_lo + 4u, 1u);
              if (STREAM || b == mb - 1) {
                if (elect_one()) umma_commit(bar_wempty + 8u * st);
              }
              if (STREAM) st = (st + 1 == (uint32_t)n_wst) ? 0u : st + 1;
            }
          } else {
            int slot = 0;
            for (int j = 0; j < k; ++j) {
              if ((STREAM || b == 0) && slot == 0) {
                mbar_wait(bar_wfull + 8u * st, (wph >> st) & 1u, error_flag);
                wph ^= 1u << st;
                fence_after_sync();
              }
              const uint32_t b_lo = lo_flag | (w16_0 + st * stage16 + (uint32_t)slot * tap16);
              const uint32_t a_lo0 = blk_lo + (uint32_t)((j - cen) * d * (int)kStep16);
#pragma unroll
              for (int sp = 0; sp < S; ++sp) {
#pragma unroll
                for (int kk = 0; kk < KK; ++kk)
                  mma_dir(d_tmem + (uint32_t)(sp * C), a_lo0 + (uint32_t)sp * kStep16 + 2u * kk, b_lo + 2u * kk,
                          (j > 0 || kk > 0) ? 1u : acc0);
              }
              if (slot == tps - 1 || j == k - 1) {
                if (STREAM || b == mb - 1) {
                  if (elect_one()) umma_commit(bar_wempty + 8u * st);
                }
                st = (st + 1 == (uint32_t)n_wst) ? 0u : st + 1;
                slot = 0;
              } else {
                ++slot;
              }
            }
          }
          if (elect_one()) umma_commit(bar_acc + 8u * b);
          blk_lo += blk16;
          d_tmem += 64u;
        }
        if (STREAM) {
          ws0 = st;
        } else {
          ws0 += ns;
          if (ws0 >= (uint32_t)n_wst) ws0 -= (uint32_t)n_wst;
        }
        if (c == n_convs - 2) {
          if (elect_one()) umma_commit(bar_pfree);               // last reader of P (c even) is done
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: 4 sets of 4 warps (warp % 4 = TMEM lane quarter) =====================
    // `spb` sets share a block (each drains 64 / spb of its columns) and the 4 / spb groups of sets take the tile's blocks
    // round-robin: spb = 1 -> set s owns block s and drains all 64 columns of its rows (four tcgen05.ld in two waves, ONE
    // fence / arrive per block and convolution); spb = 2 on a 4-block tile -> sets {0, 1} drain blocks 0 and 2, sets
    // {2, 3} blocks 1 and 3, 32 columns per thread: a block's hand-over chain (accumulator full -> A tile of the next
    // convolution ready) is half as long, which is what bounds the kernel -- the issuer waits for block b + 1 of the
    // previous convolution before it may start block b.
    const int set = (warp - 4) >> 2, quarter = warp & 3;
    const int spb = p.spb;                                       // sets per block
    const int n_slots = NSETS / spb;                             // blocks drained concurrently
    const int slot = set / spb;                                  // my blocks: slot, slot + n_slots, ...
    const int nch = 4 / spb;                                     // my 16-column chunks: chunk0 .. chunk0 + nch - 1
    const int chunk0 = (set % spb) * nch;
    if (slot < mb) {
    constexpr int NB = C == 16 ? 1 : (C == 32 && !X3) ? 2 : 0;   // bias register sets (C = 64, X3 at C = 32: loaded per chunk)
    constexpr int MAXCH = X3 ? 2 : 4;                            // 16-column chunks per thread (X3 tiles have <= 2 blocks: spb >= 2)
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const uint32_t x_col0 = (uint32_t)(mb * 64);                 // residual stream x (fp32) lives after the c1 accumulators
    const float slope = p.slope, inv_slope = 1.0f / p.slope, scale = p.scale;
    const int L = p.L, H = p.H, V = p.V;
    const int row_in_blk = quarter * 32 + lane;
    const uint32_t swz_row = (uint32_t)(row_in_blk & 7) << 4;    // 16-byte chunk index of a row is XORed with (row & 7)
    const float* const bias_s = reinterpret_cast<const float*>(smem_gen + p.bias_off);
    asm volatile("griddepcontrol.wait;" ::: "memory");          // add1 / outputs belong to the stream's previous kernels
    TileIter it;
    it.init((int)blockIdx.x, (int)gridDim.x, 1, p.m_tiles_per_b);
    uint32_t n = 0, n_tile = 0;
    uint32_t* const trace = (p.trace && blockIdx.x == 0 && lane == 0 && quarter == 0) ? p.trace + 1024 * (1 + set) : nullptr;
    uint32_t ntr = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it.next(), ++n_tile) {
      const int t_tile0 = it.mt * V - H;                         // time step of tile row 0, sub-step 0
      // ---- x0 = inverse leaky_relu of the input tile (bf16 in P) -> fp32 in tensor memory
      mbar_wait(bar(kRpBarAFull), n_tile & 1u, error_flag);
      for (int blk = slot; blk < mb; blk += n_slots) {
        const uint32_t row_off = tile_off + (uint32_t)(blk * 128 + row_in_blk) * 128u;
        const uint32_t taddr0 = tmem_base + lane_addr + (uint32_t)(blk * 64 + chunk0 * 16);
#pragma unroll
        for (int j = 0; j < MAXCH; ++j) {
          if (j < nch) {
            const uint32_t cb = (uint32_t)((chunk0 + j) * 32);
            uint32_t r[16];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float f[8];
              unpack_bf16x8(lds128(p_base + row_off + ((cb + (uint32_t)(h * 16)) ^ swz_row)), f);
              if (X3) {
                float fl[8];
                unpack_bf16x8(lds128(p_base + lo_delta + row_off + ((cb + (uint32_t)(h * 16)) ^ swz_row)), fl);
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] += fl[i];
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) r[8 * h + i] = __float_as_uint(fminf(f[i], f[i] * inv_slope));
            }
            tmem_st16(taddr0 + x_col0 + (uint32_t)(j * 16), r);
          }
        }
      }
      for (int c = 0; c < n_convs; ++c, ++n) {
        const bool is_last = (c == n_convs - 1);
        float bias_r[NB > 0 ? NB : 1][16];
#pragma unroll
        for (int jb = 0; jb < NB; ++jb) {
          const float4* bp = reinterpret_cast<const float4*>(bias_s + c * C + ((chunk0 + jb) * 16) % C);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 bv = bp[i];
            bias_r[jb][4 * i] = bv.x; bias_r[jb][4 * i + 1] = bv.y; bias_r[jb][4 * i + 2] = bv.z; bias_r[jb][4 * i + 3] = bv.w;
          }
        }
        for (int blk = slot; blk < mb; blk += n_slots) {
        const int row = blk * 128 + row_in_blk;                    // my row of the tile
        const int t_row = row * S;                                 // its first time step inside the tile
        // rows (partly) outside the utterance: the next convolution must see zeros there
        const bool row_out = (t_tile0 + t_row < 0) || (t_tile0 + t_row + S > L);
        const uint32_t row_off = tile_off + (uint32_t)row * 128u;
        const uint32_t taddr0 = tmem_base + lane_addr + (uint32_t)(blk * 64 + chunk0 * 16);
        const uint32_t bar_acc = bar(kRpBarAccFull + blk), bar_ready = bar(kRpBarReady + blk);
        // conv c reads P (c even) / Q (c odd) and writes the other; its accumulator is T (c even) / X (c odd)
        const uint32_t dst = ((c & 1) ? p_base : q_base) + row_off;
        const uint32_t taddr = taddr0 + ((c & 1) ? x_col0 : 0u);
        const uint32_t bar_rdy = bar_ready + ((n & 1u) ? 8u * kRpMaxBlocks : 0u);
        // ---- last convolution: the running sum of my rows is fetched before the accumulator is waited for
        uint4 a0[MAXCH], a1[MAXCH], l0[X3 ? MAXCH : 1], l1[X3 ? MAXCH : 1];
        bool valid[MAXCH];
        long long g_row = 0;
        // X3: out_act has rows [hi (C) | lo (C)] per time step: chunk ch of my row starts at g_row * 2 + x3_off(ch)
        // (hi plane; the lo plane C elements further); add1 / out_raw are planar
        auto x3_off = [&](int ch) { return (long long)(((ch * 16) / C) * 2 * C + (ch * 16) % C); };
        if (is_last) {
          g_row = ((long long)it.b * L + t_tile0 + t_row) * C;   // my row's 64 values are contiguous in the output
#pragma unroll
          for (int j = 0; j < MAXCH; ++j) {
            const int t = t_row + ((chunk0 + j) * 16) / C;
            valid[j] = j < nch && t >= H && t < H + V && t_tile0 + t < L;
            a0[j] = make_uint4(0u, 0u, 0u, 0u); a1[j] = a0[j];
            if (X3) { l0[j] = a0[j]; l1[j] = a0[j]; }
            if (valid[j] && p.add1) {
              a0[j] = ldg128(p.add1 + g_row + (chunk0 + j) * 16);
              a1[j] = ldg128(p.add1 + g_row + (chunk0 + j) * 16 + 8);
              if (X3) {
                l0[j] = ldg128(p.add1 + p.plane_stride + g_row + (chunk0 + j) * 16);
                l1[j] = ldg128(p.add1 + p.plane_stride + g_row + (chunk0 + j) * 16 + 8);
              }
            }
          }
        }
        if (trace && ntr < 1021) trace[ntr++] = (uint32_t)clock();
        mbar_wait(bar_acc, n & 1u, error_flag);
        fence_after_sync();
        if (trace && ntr < 1021) trace[ntr++] = (uint32_t)clock();
#pragma unroll
        for (int j2 = 0; j2 < MAXCH; j2 += 2) {
          if (j2 < nch) {
            uint32_t ra[2][16];
            tmem_ld16_nowait(taddr + (uint32_t)(j2 * 16), ra[0]);
            if (j2 + 1 < nch) tmem_ld16_nowait(taddr + (uint32_t)(j2 * 16 + 16), ra[1]);
            tmem_wait_ld();
            if (is_last && j2 + 2 >= nch) {
              // accumulator drained into registers: the next tile's first convolution may overwrite it while the
              // output below goes to global memory
              fence_before_sync();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_rdy);
            }
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
              const int j = j2 + jj;
              if (j < nch) {
                float v[16];
                if (NB > 0) {
#pragma unroll
                  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(ra[jj][i]) + bias_r[NB > 0 ? j % NB : 0][i];
                } else {
                  const float4* bp = reinterpret_cast<const float4*>(bias_s + c * C + ((chunk0 + j) * 16) % C);
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float4 bv = bp[i];
                    v[4 * i] = __uint_as_float(ra[jj][4 * i]) + bv.x; v[4 * i + 1] = __uint_as_float(ra[jj][4 * i + 1]) + bv.y;
                    v[4 * i + 2] = __uint_as_float(ra[jj][4 * i + 2]) + bv.z; v[4 * i + 3] = __uint_as_float(ra[jj][4 * i + 3]) + bv.w;
                  }
                }
                const uint32_t cb = (uint32_t)((chunk0 + j) * 32);
                if (!is_last) {
                  uint32_t h[8], hl[X3 ? 8 : 1];
                  if (X3) {   // leaky_relu in fp32, then the two planes: hi = bf16(a), lo = bf16(a - hi)
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                      const float e0 = fmaxf(v[2 * i], v[2 * i] * slope), e1 = fmaxf(v[2 * i + 1], v[2 * i + 1] * slope);
                      h[i] = pack_bf16x2(e0, e1);
                      hl[i] = pack_bf16x2(e0 - __uint_as_float(h[i] << 16), e1 - __uint_as_float(h[i] & 0xffff0000u));
                    }
                  } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) h[i] = bf16x2_scale_max(pack_bf16x2(v[2 * i], v[2 * i + 1]), slope);
                  }
                  if (row_out) {
                    const int tg = t_tile0 + t_row + ((chunk0 + j) * 16) / C;
                    if (tg < 0 || tg >= L) {
#pragma unroll
                      for (int i = 0; i < 8; ++i) h[i] = 0u;
                      if (X3) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) hl[i] = 0u;
                      }
                    }
                  }
                  sts128(dst + (cb ^ swz_row), make_uint4(h[0], h[1], h[2], h[3]));
                  sts128(dst + ((cb + 16u) ^ swz_row), make_uint4(h[4], h[5], h[6], h[7]));
                  if (X3) {
                    sts128(dst + lo_delta + (cb ^ swz_row), make_uint4(hl[0], hl[1], hl[2], hl[3]));
                    sts128(dst + lo_delta + ((cb + 16u) ^ swz_row), make_uint4(hl[4], hl[5], hl[6], hl[7]));
                  }
                } else if (valid[j]) {
                  // ---- block output: (x_out [+ running sum]) * scale -> bf16 raw / leaky_relu'd, valid steps only
                  const long long g_off = g_row + (chunk0 + j) * 16;
                  float f[16];
                  unpack_bf16x8(a0[j], f);
                  unpack_bf16x8(a1[j], f + 8);
                  if (X3) {
                    float fl[16];
                    unpack_bf16x8(l0[j], fl);
                    unpack_bf16x8(l1[j], fl + 8);
#pragma unroll
                    for (int i = 0; i < 16; ++i) f[i] += fl[i];
                  }
#pragma unroll
                  for (int i = 0; i < 16; ++i) v[i] = (v[i] + f[i]) * scale;
                  if (p.out_f32) {
                    float4* o = reinterpret_cast<float4*>(p.out_f32 + g_off);
#pragma unroll
                    for (int i = 0; i < 4; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                  }
                  uint32_t h[8];
                  if (X3) {
                    const long long g2 = 2 * g_row + x3_off(chunk0 + j);
                    uint32_t hl[8];
                    if (p.out_raw) {
#pragma unroll
                      for (int i = 0; i < 8; ++i) {
                        h[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
                        hl[i] = pack_bf16x2(v[2 * i] - __uint_as_float(h[i] << 16), v[2 * i + 1] - __uint_as_float(h[i] & 0xffff0000u));
                      }
                      stg128(p.out_raw + g_off, make_uint4(h[0], h[1], h[2], h[3]));
                      stg128(p.out_raw + g_off + 8, make_uint4(h[4], h[5], h[6], h[7]));
                      stg128(p.out_raw + p.plane_stride + g_off, make_uint4(hl[0], hl[1], hl[2], hl[3]));
                      stg128(p.out_raw + p.plane_stride + g_off + 8, make_uint4(hl[4], hl[5], hl[6], hl[7]));
                    }
                    if (p.out_act) {
#pragma unroll
                      for (int i = 0; i < 8; ++i) {
                        const float e0 = fmaxf(v[2 * i], v[2 * i] * slope), e1 = fmaxf(v[2 * i + 1], v[2 * i + 1] * slope);
                        h[i] = pack_bf16x2(e0, e1);
                        hl[i] = pack_bf16x2(e0 - __uint_as_float(h[i] << 16), e1 - __uint_as_float(h[i] & 0xffff0000u));
                      }
                      stg128(p.out_act + g2, make_uint4(h[0], h[1], h[2], h[3]));
                      stg128(p.out_act + g2 + 8, make_uint4(h[4], h[5], h[6], h[7]));
                      stg128(p.out_act + g2 + C, make_uint4(hl[0], hl[1], hl[2], hl[3]));
                      stg128(p.out_act + g2 + C + 8, make_uint4(hl[4], hl[5], hl[6], hl[7]));
                    }
                  } else {
#pragma unroll
                  for (int i = 0; i < 8; ++i) h[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
                  if (p.out_raw) {
                    stg128(p.out_raw + g_off, make_uint4(h[0], h[1], h[2], h[3]));
                    stg128(p.out_raw + g_off + 8, make_uint4(h[4], h[5], h[6], h[7]));
                  }
                  if (p.out_act) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) h[i] = bf16x2_scale_max(h[i], slope);
                    stg128(p.out_act + g_off, make_uint4(h[0], h[1], h[2], h[3]));
                    stg128(p.out_act + g_off + 8, make_uint4(h[4], h[5], h[6], h[7]));
                  }
                  }
                }
              }
            }
          }
        }
        // release: accumulator columns drained, my part of the next A tile written (generic -> async proxy)
        if (trace && ntr < 1021) trace[ntr++] = (uint32_t)clock();
        if (!is_last) {
          if (c == 0) tmem_wait_st();                            // x0 is in tensor memory before c2's MMAs accumulate onto it
          fence_async_smem();
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_rdy);
        }
        if (trace && ntr < 1021) trace[ntr++] = (uint32_t)clock();
        }
      }
    }
    }
  }

  // ---- teardown: everyone done with TMEM before the allocating warp frees it ----
  fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

}  // namespace vsg
