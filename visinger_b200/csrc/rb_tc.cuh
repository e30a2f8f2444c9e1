// Whole-ResBlock1 fusion for the low-channel decoder stages (C = 16 / 32 / 64), bf16 tensor-core mode:
//
//     for (c1, c2, d) in pairs:   x = c2(leaky_relu(c1_d(leaky_relu(x)))) + x          (modules/visinger/decoder.py:91-104)
//
// -- all 2 * n_pairs convolutions of one ResBlock1 in ONE kernel.  Un-fused, a resblock streams ~16 tensor passes through
// HBM (13 with fused pairs); here it reads its input tile once and writes its result once, every intermediate lives on
// the SM:
//   * the activated stream leaky_relu(x) / leaky_relu(c1(.)) ping-pongs between two swizzled K-major tiles P and Q in
//     shared memory (bf16; each is the A operand of the next convolution; taps = row-shifted UMMA descriptors);
//   * the residual stream x stays in TENSOR MEMORY in fp32 next to the accumulators (tcgen05.ld / tcgen05.st by the
//     thread that owns the row) -- it is never rounded to bf16 inside a resblock, unlike the per-conv path;
//   * weights stream L2 -> shared memory through a TMA ring, one stage = a group of taps of one convolution.
// A tile is mb blocks of 128 rows of one utterance including a halo of H = sum_q ((k-1)/2 * (d_q + 1)) rows per side
// that is recomputed by the neighbour tile (12 / 36 / 60 rows for k = 3 / 7 / 11; valid rows V = 128 * mb - 2 H), so
// tiles are as tall as shared / tensor memory allow: 2048 rows at C = 16, 1024 at C = 32, 512 at C = 64.
//
// Pipelining is per 128-row BLOCK inside the tile (one CTA per SM owns all of TMEM): the MMA warp walks
// (conv, tap group, block); the accumulator of block b is complete in the last tap group and handed to the epilogue warps
// through acc_full[b]; they write block b of the next convolution's A tile and release it (and the accumulator columns)
// through buf_ready[b].  Convolution c+1 on block b needs blocks b-1, b, b+1 of convolution c (its halo rows), so while
// the epilogue of the last blocks of conv c runs, the tensor pipe already works on the first blocks of conv c+1.
//
// Warp roles: 0 input-tile TMA producer, 1 TMEM allocator + MMA issuer 0, 2 weight TMA producer, 3 MMA issuer 1,
// 4 .. 4 + 4 n_sets - 1 epilogue sets of 4 warps (warp % 4 = TMEM lane quarter), then MMA issuers 2, 3.  Issuer i takes
// blocks i, i + n_issuers, ... (converged warps, elect_one per instruction, descriptors advanced by plain adds so the loop
// is uniform-datapath code with back-to-back UTCHMMA).  An epilogue work unit is (block, 16-channel chunk); set s always
// owns the same chunk (its bias stays in registers) and walks the blocks, so the chunks of one block drain in parallel.
// The last convolution's epilogue adds the running resblock sum (plain 16-byte global loads of its own rows), scales and
// writes the result with 16-byte global stores: no staging tile, no store-side tensor maps.
#pragma once
#include "conv_tc.cuh"

namespace vsg {

constexpr int kRbMaxConvs = 8;
constexpr int kRbMaxBlocks = 16;
constexpr int kRbMaxWStages = 6;

struct RbTC {
  int B, L;                       // utterances, rows (time steps at this stage's rate) per utterance
  int k;                          // taps of every convolution of the block
  int n_convs;                    // 2 * n_pairs: c1_0, c2_0, c1_1, c2_1, ...
  int dil[kRbMaxConvs];           // dilation per convolution (c2: 1)
  int mb;                         // 128-row blocks per tile
  int H, V;                       // halo rows per side, valid rows per tile
  int M;                          // zero margin rows before / after the tile inside P and Q (>= largest tap reach)
  int m_tiles_per_b, total_tiles;
  int G, n_groups;                // taps per weight stage, stages (groups) per convolution
  int stages_w;
  uint32_t tap_bytes, w_stage_bytes;      // one tap tile [C x C] bf16 rounded up to 1 KB; one ring stage
  uint32_t buf_bytes;                     // one activation tile (M + 128 mb + M rows)
  uint32_t p_off, q_off, w_off, bar_off;  // shared-memory carve-up relative to the 1024-aligned base
  uint32_t tmem_cols;
  uint32_t swizzle_code, sbo_bytes;       // UMMA layout type (2 = 128B, 4 = 64B, 6 = 32B), 8 rows * row bytes
  int n_sets;                     // epilogue warp sets (4 warps each); a multiple of C / 16 or a divisor of it
  int n_issuers;                  // MMA issuer warps (1, 2 or 4)
  const float* bias[kRbMaxConvs];
  const __nv_bfloat16* add1;      // running resblock sum [B, L, C] or null (read for the tile's valid rows only)
  __nv_bfloat16* out_raw;         // (x_out + add1) * scale as bf16, or null
  __nv_bfloat16* out_act;         // leaky_relu of the same, or null
  float* out_f32;                 // fp32 copy (parity hook), or null
  float scale, slope;
  int* error_flag;
};

struct RbMaps { CUtensorMap w[kRbMaxConvs]; };

namespace tc {

// barrier slots of the resblock kernel
constexpr int kRbBarAccFull = 0;                                  // [kRbMaxBlocks] tcgen05.commit, count 1
// buf_ready has TWO barrier banks, used by even / odd convolution steps: with several issuer warps the owner of block b
// may finish step n+1 on it before the neighbour block's issuer has looked at step n -- a single parity bit would alias
constexpr int kRbBarBufReady = kRbMaxBlocks;                      // [2][kRbMaxBlocks] (quarter warp, chunk) units
constexpr int kRbBarWFull = 3 * kRbMaxBlocks;                     // [kRbMaxWStages]
constexpr int kRbBarWEmpty = kRbBarWFull + kRbMaxWStages;         // [kRbMaxWStages]
constexpr int kRbBarAFull = kRbBarWEmpty + kRbMaxWStages;         // input tile landed in P
constexpr int kRbBarPFree = kRbBarAFull + 1;                      // last convolution that reads P has completed
constexpr int kRbNumBars = kRbBarPFree + 1;
constexpr int kRbMaxSets = 4;
constexpr int kRbMaxIssuers = 4;
constexpr int kRbThreads = (4 + 4 * kRbMaxSets + 2) * 32;   // 704: producers, 4 issuers, 16 epilogue warps

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint4 ldg128(const void* p) {
  uint4 v;
  asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg128(void* p, const uint4& v) {
  asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

}  // namespace tc

template <int C>
__global__ void __launch_bounds__(tc::kRbThreads, 1)
rb_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ RbMaps wm, const RbTC p) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));     // generic pointer to the aligned base
  const uint32_t bar_base = smem_base + p.bar_off;
  auto bar = [&](int slot) { return bar_base + 8u * (uint32_t)slot; };
  const uint32_t tmem_slot = bar_base + 8u * kRbNumBars;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_gen + p.bar_off + 8u * kRbNumBars);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  constexpr uint32_t kRowBytes = C * 2;
  constexpr uint32_t kSwzMask = C >= 64 ? 7u : C >= 32 ? 3u : 1u;
  constexpr int KK = C / 16;
  const int mb = p.mb, n_convs = p.n_convs, k = p.k, total_tiles = p.total_tiles;
  int* const error_flag = p.error_flag;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    for (int c = 0; c < n_convs; ++c) prefetch_tmap(&wm.w[c]);
    // buf_ready[b]: every (quarter warp, chunk) unit of block b arrives once per convolution step
    for (int b = 0; b < kRbMaxBlocks; ++b) {
      mbar_init(bar(kRbBarAccFull + b), 1);
      mbar_init(bar(kRbBarBufReady + b), 4 * (C / 16));
      mbar_init(bar(kRbBarBufReady + kRbMaxBlocks + b), 4 * (C / 16));
    }
    for (int s = 0; s < kRbMaxWStages; ++s) { mbar_init(bar(kRbBarWFull + s), 1); mbar_init(bar(kRbBarWEmpty + s), p.n_issuers); }
    mbar_init(bar(kRbBarAFull), 1);
    mbar_init(bar(kRbBarPFree), p.n_issuers);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  {  // zero margins of P and Q: rows the taps reach beyond the tile; never written afterwards
    const uint32_t margin16 = (uint32_t)p.M * kRowBytes / 16u;        // 16-byte words per margin
    const uint32_t tail_off = ((uint32_t)p.M + 128u * (uint32_t)mb) * kRowBytes;
    uint4* pg = reinterpret_cast<uint4*>(smem_gen + p.p_off);
    uint4* qg = reinterpret_cast<uint4*>(smem_gen + p.q_off);
    for (uint32_t i = threadIdx.x; i < margin16; i += blockDim.x) {
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
      pg[i] = z; qg[i] = z;
      pg[tail_off / 16u + i] = z; qg[tail_off / 16u + i] = z;
    }
    fence_async_smem();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_ptr, 0);
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const uint32_t p_base = smem_base + p.p_off, q_base = smem_base + p.q_off, w_base = smem_base + p.w_off;
  const uint32_t tile_off = (uint32_t)p.M * kRowBytes;            // byte offset of tile row 0 inside P / Q

  if (warp == 0) {
    // ===================== input-tile producer: leaky_relu(x) rows [t0 - H, t0 - H + 128 mb) -> P =====================
    asm volatile("griddepcontrol.wait;" ::: "memory");           // the previous kernel produced the input
    TileIter it;
    it.init((int)blockIdx.x, (int)gridDim.x, 1, p.m_tiles_per_b);
    uint32_t n_tile = 0;
    const int R = 128 * mb, box_rows = R < 256 ? R : 256, n_boxes = R / box_rows;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it.next(), ++n_tile) {
      if (n_tile > 0) mbar_wait(bar(kRbBarPFree), (n_tile - 1) & 1u, error_flag);   // P no longer read by the previous tile
      const int row0 = it.mt * p.V - p.H;
      if (elect_one()) {
        mbar_expect_tx(bar(kRbBarAFull), (uint32_t)R * kRowBytes);
        for (int bx = 0; bx < n_boxes; ++bx)
          tma_load_3d(p_base + tile_off + (uint32_t)(bx * box_rows) * kRowBytes, &tmA, bar(kRbBarAFull), 0,
                      row0 + bx * box_rows, it.b);
      }
    }
  } else if (warp == 2) {
    // ===================== weight producer: ring of tap groups, same order as the MMA warp consumes them =====================
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x)
      for (int c = 0; c < n_convs; ++c)
        for (int gi = 0; gi < p.n_groups; ++gi) {
          const int j0 = gi * p.G, j1 = min(k, j0 + p.G);
          mbar_wait(bar(kRbBarWEmpty + s), ph ^ 1u, error_flag);
          if (elect_one()) {
            mbar_expect_tx(bar(kRbBarWFull + s), (uint32_t)(j1 - j0) * (uint32_t)(C * C * 2));
            for (int j = j0; j < j1; ++j)
              tma_load_2d(w_base + (uint32_t)s * p.w_stage_bytes + (uint32_t)(j - j0) * p.tap_bytes, &wm.w[c],
                          bar(kRbBarWFull + s), 0, j * C);
          }
          if (++s == p.stages_w) { s = 0; ph ^= 1u; }
        }
  } else if (warp == 1 || warp == 3 || warp >= 4 + 4 * p.n_sets) {
    // ===================== MMA issuers (converged warps) =====================
    // Every operand below derives from kernel parameters and loop counters through selects (no dynamically indexed
    // parameter arrays) and the descriptor low words are advanced by plain adds (a shared-memory address >> 4 never
    // carries out of its 14-bit field), so the loop compiles to uniform-datapath code with bare UTCHMMA instructions.
    const int issuer = warp == 1 ? 0 : warp == 3 ? 1 : 2 + (warp - (4 + 4 * p.n_sets));
    const int n_iss = p.n_issuers;
    if (issuer < n_iss) {
    const uint32_t idesc = make_idesc_bf16(128, (uint32_t)C);
    const uint32_t desc_hi = ((p.sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | ((p.swizzle_code & 7u) << 29);
    auto mk = [&](uint32_t lo) { return ((uint64_t)desc_hi << 32) | (uint64_t)lo; };
    constexpr uint32_t blk16 = (128u * kRowBytes) >> 4, row16 = kRowBytes >> 4, lo_flag = 1u << 16;
    const uint32_t tap16 = p.tap_bytes >> 4, wstage16 = p.w_stage_bytes >> 4;
    const uint32_t p16 = (p_base + tile_off) >> 4, q16 = (q_base + tile_off) >> 4, w16_0 = w_base >> 4;
    const int G = p.G, n_groups = p.n_groups, stages_w = p.stages_w;
    const int d0 = p.dil[0], d1 = p.dil[2], d2 = p.dil[4], d3 = p.dil[6];
    const uint32_t bar_acc = bar(kRbBarAccFull), bar_ready = bar(kRbBarBufReady), bar_wfull = bar(kRbBarWFull),
                   bar_wempty = bar(kRbBarWEmpty), bar_afull = bar(kRbBarAFull), bar_pfree = bar(kRbBarPFree);
    const uint32_t blk_step = blk16 * (uint32_t)n_iss, tm_step = (uint32_t)(C * n_iss);
    int s = 0;
    uint32_t wph = 0, n = 0, n_tile = 0;                          // n: convolution steps issued so far by this CTA
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++n_tile) {
      for (int c = 0; c < n_convs; ++c, ++n) {
        const uint32_t src16 = (c & 1) ? q16 : p16;
        const int pr = c >> 1;
        const int d = (c & 1) ? 1 : (pr == 0 ? d0 : pr == 1 ? d1 : pr == 2 ? d2 : d3);
        const uint32_t tap_step16 = (uint32_t)d * row16;
        // descriptor low word of tap 0, this issuer's first block
        const uint32_t first_lo = lo_flag | (src16 - (uint32_t)((k - 1) / 2 * d) * row16 + (uint32_t)issuer * blk16);
        // buf_ready of the previous step: bank (n-1) & 1, its ((n-1) >> 1)-th completion
        const uint32_t prev_par = ((n - 1) >> 1) & 1u;
        const uint32_t bar_prev = bar_ready + (((n - 1) & 1u) ? 8u * kRbMaxBlocks : 0u);
        for (int gi = 0; gi < n_groups; ++gi) {
          const int j0 = gi * G, nj = min(k, j0 + G) - j0;
          mbar_wait(bar_wfull + 8u * s, wph, error_flag);
          fence_after_sync();
          const uint32_t w_lo0 = lo_flag | (w16_0 + (uint32_t)s * wstage16);
          uint32_t a_blk_lo = first_lo + (uint32_t)j0 * tap_step16;
          uint32_t d_tmem = tmem_base + (uint32_t)(issuer * C);
          for (int bi = issuer; bi < mb; bi += n_iss) {
            if (gi == 0) {
              // block bi of this convolution reads blocks bi-1 .. bi+1 of the previous convolution's output and
              // overwrites accumulator bi: all released through buf_ready of the previous step.  This issuer's previous
              // block was bi - n_iss, so bi-1 has been waited for only when n_iss == 1.
              if (c == 0) {
                if (bi == issuer) mbar_wait(bar_afull, n_tile & 1u, error_flag);
                if (n > 0) mbar_wait(bar_prev + 8u * bi, prev_par, error_flag);
              } else {
                if (bi > 0 && (n_iss > 1 || bi == issuer)) mbar_wait(bar_prev + 8u * (bi - 1), prev_par, error_flag);
                if (n_iss > 1 || bi == issuer) mbar_wait(bar_prev + 8u * bi, prev_par, error_flag);
                if (bi + 1 < mb) mbar_wait(bar_prev + 8u * (bi + 1), prev_par, error_flag);
              }
              fence_after_sync();
            }
            uint32_t a_lo = a_blk_lo, b_lo = w_lo0;
            uint32_t accumulate = gi > 0 ? 1u : 0u;
            for (int j = 0; j < nj; ++j) {
#pragma unroll
              for (int kk = 0; kk < KK; ++kk)
                if (elect_one()) umma_bf16(d_tmem, mk(a_lo + 2u * kk), mk(b_lo + 2u * kk), idesc, kk > 0 ? 1u : accumulate);
              accumulate = 1u;
              a_lo += tap_step16;
              b_lo += tap16;
            }
            if (gi == n_groups - 1) {
              if (elect_one()) umma_commit(bar_acc + 8u * bi);
            }
            a_blk_lo += blk_step;
            d_tmem += tm_step;
          }
          if (elect_one()) umma_commit(bar_wempty + 8u * s);
          if (++s == stages_w) { s = 0; wph ^= 1u; }
        }
        if (c == n_convs - 2) {
          if (elect_one()) umma_commit(bar_pfree);               // last reader of P (c even) is done
        }
      }
    }
    }
  } else if (warp >= 4 && warp < 4 + 4 * p.n_sets) {
    // ===================== epilogue sets =====================
    // unit (block, chunk): set s owns chunk s % n_chunks of blocks blk0, blk0 + blk_stride, ...
    constexpr int n_chunks = C / 16;
    const int set = (warp - 4) >> 2, quarter = warp & 3, n_sets = p.n_sets;
    const int ch = set % n_chunks;
    const int blk0 = n_sets >= n_chunks ? set / n_chunks : 0;
    const int blk_stride = n_sets >= n_chunks ? n_sets / n_chunks : 1;
    const int n_my_chunks = n_sets >= n_chunks ? 1 : n_chunks / n_sets;      // fewer sets than chunks: a set takes several
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const uint32_t x_col0 = (uint32_t)(mb * C);                 // residual stream x (fp32) lives after the accumulators
    const float slope = p.slope, inv_slope = 1.0f / p.slope, scale = p.scale;
    const int L = p.L, H = p.H, V = p.V;
    asm volatile("griddepcontrol.wait;" ::: "memory");          // add1 / outputs belong to the stream's previous kernels
    TileIter it;
    it.init((int)blockIdx.x, (int)gridDim.x, 1, p.m_tiles_per_b);
    uint32_t n = 0, n_tile = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it.next(), ++n_tile) {
      const int t_row0 = it.mt * V - H;                          // time step of tile row 0
      // ---- x0 = inverse leaky_relu of the input tile (bf16 in P) -> fp32 in tensor memory
      mbar_wait(bar(kRbBarAFull), n_tile & 1u, error_flag);
      for (int bi = blk0; bi < mb; bi += blk_stride) {
        const int row = bi * 128 + quarter * 32 + lane;
        const uint32_t row_off = tile_off + (uint32_t)row * kRowBytes;
        for (int cc = 0; cc < n_my_chunks; ++cc) {
          const int chn = ch + cc * n_sets;
          uint32_t r[16];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float f[8];
            unpack_bf16x8(lds128(p_base + swz(row_off + (uint32_t)(chn * 32 + h * 16), kSwzMask)), f);
#pragma unroll
            for (int i = 0; i < 8; ++i) r[8 * h + i] = __float_as_uint(fminf(f[i], f[i] * inv_slope));
          }
          tmem_st16(tmem_base + lane_addr + x_col0 + (uint32_t)(bi * C + chn * 16), r);
        }
      }
      tmem_wait_st();
      for (int c = 0; c < n_convs; ++c, ++n) {
        const bool is_c2 = (c & 1) != 0, is_last = (c == n_convs - 1);
        const uint32_t dst = ((c & 1) ? p_base : q_base);        // conv c reads P (c even) / Q (c odd), writes the other
        const float* const bias = p.bias[c];
        float bias_r[16];
        if (n_my_chunks == 1) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + ch * 16 + i));
            bias_r[i] = bv.x; bias_r[i + 1] = bv.y; bias_r[i + 2] = bv.z; bias_r[i + 3] = bv.w;
          }
        }
        for (int bi = blk0; bi < mb; bi += blk_stride) {
          const int row = bi * 128 + quarter * 32 + lane;
          const int t = t_row0 + row;
          const bool inside = (t >= 0) && (t < L);               // outside the utterance the next conv must see zeros
          const uint32_t row_off = tile_off + (uint32_t)row * kRowBytes;
          mbar_wait(bar(kRbBarAccFull + bi), n & 1u, error_flag);
          fence_after_sync();
          for (int cc = 0; cc < n_my_chunks; ++cc) {
            const int chn = ch + cc * n_sets;
            uint32_t ra[16], rx[16];
            tmem_ld16_nowait(tmem_base + lane_addr + (uint32_t)(bi * C + chn * 16), ra);
            if (is_c2) tmem_ld16_nowait(tmem_base + lane_addr + x_col0 + (uint32_t)(bi * C + chn * 16), rx);
            if (n_my_chunks != 1) {
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + chn * 16 + i));
                bias_r[i] = bv.x; bias_r[i + 1] = bv.y; bias_r[i + 2] = bv.z; bias_r[i + 3] = bv.w;
              }
            }
            tmem_wait_ld();
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(ra[i]) + bias_r[i];
            if (is_c2) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += __uint_as_float(rx[i]);        // x = c2(...) + x   (decoder.py:102)
            }
            if (!is_last) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = inside ? v[i] : 0.f;
              if (is_c2) {
#pragma unroll
                for (int i = 0; i < 16; ++i) rx[i] = __float_as_uint(v[i]);
                tmem_st16(tmem_base + lane_addr + x_col0 + (uint32_t)(bi * C + chn * 16), rx);
              }
              uint32_t h[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float a0 = fmaxf(v[2 * i], v[2 * i] * slope), a1 = fmaxf(v[2 * i + 1], v[2 * i + 1] * slope);
                h[i] = pack_bf16x2(a0, a1);
              }
              sts128(dst + swz(row_off + (uint32_t)(chn * 32), kSwzMask), make_uint4(h[0], h[1], h[2], h[3]));
              sts128(dst + swz(row_off + (uint32_t)(chn * 32 + 16), kSwzMask), make_uint4(h[4], h[5], h[6], h[7]));
            } else if (row >= H && row < H + V && t < L) {
              // ---- block output: (x_out [+ running sum]) * scale -> bf16 raw / leaky_relu'd, valid rows only
              const long long g_off = ((long long)it.b * L + t) * C + chn * 16;
              if (p.add1) {
                float f[16];
                unpack_bf16x8(ldg128(p.add1 + g_off), f);
                unpack_bf16x8(ldg128(p.add1 + g_off + 8), f + 8);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] += f[i];
              }
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] *= scale;
              if (p.out_f32) {
                float4* o = reinterpret_cast<float4*>(p.out_f32 + g_off);
#pragma unroll
                for (int i = 0; i < 4; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
              }
              if (p.out_raw) {
                uint32_t h[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) h[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
                stg128(p.out_raw + g_off, make_uint4(h[0], h[1], h[2], h[3]));
                stg128(p.out_raw + g_off + 8, make_uint4(h[4], h[5], h[6], h[7]));
              }
              if (p.out_act) {
                uint32_t h[8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  h[i] = pack_bf16x2(fmaxf(v[2 * i], v[2 * i] * slope), fmaxf(v[2 * i + 1], v[2 * i + 1] * slope));
                stg128(p.out_act + g_off, make_uint4(h[0], h[1], h[2], h[3]));
                stg128(p.out_act + g_off + 8, make_uint4(h[4], h[5], h[6], h[7]));
              }
            }
            if (is_c2 && !is_last) tmem_wait_st();
            // release: this unit's accumulator columns drained, its part of the next A tile written (generic -> async proxy)
            fence_async_smem();
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(kRbBarBufReady + (int)(n & 1u) * kRbMaxBlocks + bi));
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
