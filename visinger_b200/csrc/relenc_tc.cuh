// bf16 (throughput-mode) kernels of the relative-position transformer encoder (reference modules/rel_transformer.py):
// tensor-core flash attention with windowed relative-position keys / values (:137-177) and the fused
// residual-sum -> channel LayerNorm -> condition -> mask step (:24-42, :293-303).  Layout: channels-last bf16 [B, T, C],
// the layout of the tcgen05 convolution kernels that run the projections and the FFN around them.
//
// The attention is a full T x T softmax per (utterance, head) with d_k = 96: M = queries, N = keys, K = d_k for the
// scores and K = keys for the values -- two chained GEMMs with an online softmax in between, whose operands change every
// (utterance, head): a register-resident flash-attention kernel on warp-level mma (m16n8k16, fp32 accumulate) fits that
// better than the persistent TMEM pipeline of the convolutions (at T = 1000 a score tile is consumed in place; there is
// no accumulator to hand to an epilogue).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace vsg {

namespace att {

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
// D (16x8, fp32) += A (16x16, bf16, row) * B (16x8, bf16, col)
__device__ __forceinline__ void mma_bf16(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

}  // namespace att

constexpr int kAttQ = 64;    // queries per CTA (4 warps x 16 rows)
constexpr int kAttK = 64;    // keys per tile

// qkv: [B, T, 3 * n_heads * DK] bf16 (q | k | v, head h at columns h * DK of each third); o: [B, T, n_heads * DK] bf16;
// Ek / Ev: fp32 [2w+1][DK]; mask: fp32 [B, T].  See relenc_attention_f32_kernel for the arithmetic.
template <int DK>
__global__ void __launch_bounds__(128) relenc_attention_bf16_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                    const float* __restrict__ mask, const float* __restrict__ Ek,
                                                                    const float* __restrict__ Ev, __nv_bfloat16* __restrict__ o,
                                                                    int n_heads, int T, int w) {
  using namespace att;
  constexpr int RS = DK + 8;                 // padded row (elements): conflict-free ldmatrix
  constexpr int KT = DK / 16;                // k steps of the score GEMM
  constexpr int NO = DK / 8;                 // n tiles of the output
  constexpr int NREL_MAX = 33;
  extern __shared__ __align__(16) uint8_t smraw[];
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smraw);     // [64][RS]
  __nv_bfloat16* Ks = Qs + kAttQ * RS;                              // [64][RS]
  __nv_bfloat16* Vs = Ks + kAttK * RS;                              // [64][RS]
  float* Rq = reinterpret_cast<float*>(Vs + kAttK * RS);            // [64][nrel]   q_i . Ek[m]
  const int nrel = 2 * w + 1;
  float* Pw = Rq + kAttQ * nrel;                                    // [64][nrel]   un-normalised p on the 2w+1 diagonals
  float* mq = Pw + kAttQ * nrel;                                    // [64]
  float* mk = mq + kAttQ;                                           // [64]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kAttQ;
  const int H = n_heads * DK, ld = 3 * H;
  const __nv_bfloat16* base = qkv + (long long)b * T * ld + h * DK;
  const float* mb = mask + (long long)b * T;
  const float scale = rsqrtf((float)DK);
  constexpr int VPR = DK / 8;                // 16-byte vectors per row

  for (int i = tid; i < kAttQ * VPR; i += 128) {
    const int r = i / VPR, v = i % VPR, t = q0 + r;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (t < T) val = *reinterpret_cast<const uint4*>(base + (long long)t * ld + v * 8);
    *reinterpret_cast<uint4*>(Qs + r * RS + v * 8) = val;
  }
  if (tid < kAttQ) mq[tid] = (q0 + tid < T) ? mb[q0 + tid] : 0.f;
  for (int i = tid; i < kAttQ * nrel; i += 128) Pw[i] = 0.f;
  __syncthreads();
  for (int i = tid; i < kAttQ * nrel; i += 128) {       // relative-position logits of my CTA's queries (fp32)
    const int r = i / nrel, m = i % nrel;
    float s = 0.f;
    for (int d = 0; d < DK; ++d) s = fmaf(__bfloat162float(Qs[r * RS + d]), __ldg(Ek + m * DK + d), s);
    Rq[i] = s;
  }
  // Q fragments of my 16 rows stay in registers
  uint32_t qf[KT][4];
  {
    const int row = warp * 16 + (lane & 15), col = (lane >> 4) * 8;
#pragma unroll
    for (int kt = 0; kt < KT; ++kt)
      ldmatrix_x4((uint32_t)__cvta_generic_to_shared(Qs + row * RS + kt * 16 + col), qf[kt][0], qf[kt][1], qf[kt][2], qf[kt][3]);
  }
  float oacc[NO][4];
#pragma unroll
  for (int n = 0; n < NO; ++n) { oacc[n][0] = oacc[n][1] = oacc[n][2] = oacc[n][3] = 0.f; }
  float row_max[2] = {-INFINITY, -INFINITY}, row_sum[2] = {0.f, 0.f};
  const int r_lo = warp * 16 + (lane >> 2), r_hi = r_lo + 8;       // my two rows inside the CTA tile
  const int c2 = (lane & 3) * 2;                                    // my two columns inside every 8-wide n tile

  for (int k0 = 0; k0 < T; k0 += kAttK) {
    __syncthreads();                                                 // previous K / V tile consumed (and Rq / Pw initialised)
    for (int i = tid; i < kAttK * VPR; i += 128) {
      const int r = i / VPR, v = i % VPR, t = k0 + r;
      uint4 kv = make_uint4(0u, 0u, 0u, 0u), vv = kv;
      if (t < T) {
        kv = *reinterpret_cast<const uint4*>(base + (long long)t * ld + H + v * 8);
        vv = *reinterpret_cast<const uint4*>(base + (long long)t * ld + 2 * H + v * 8);
      }
      *reinterpret_cast<uint4*>(Ks + r * RS + v * 8) = kv;
      *reinterpret_cast<uint4*>(Vs + r * RS + v * 8) = vv;
    }
    if (tid < kAttK) mk[tid] = (k0 + tid < T) ? mb[k0 + tid] : 0.f;
    __syncthreads();
    // ---- S = Q K^T for my 16 rows x 64 keys
    float s[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) { s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f; }
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {                               // two key n-tiles per ldmatrix.x4
        uint32_t b0, b1, b2, b3;
        const int krow = np * 16 + (lane & 7) + ((lane >> 4) << 3), kcol = kt * 16 + ((lane >> 3) & 1) * 8;
        ldmatrix_x4((uint32_t)__cvta_generic_to_shared(Ks + krow * RS + kcol), b0, b1, b2, b3);
        mma_bf16(s[2 * np], qf[kt], b0, b1);
        mma_bf16(s[2 * np + 1], qf[kt], b2, b3);
      }
    }
    // ---- scale, relative-position logits, masks, tile maxima
    float tmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int n = 0; n < 8; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = (e < 2) ? r_lo : r_hi, j = n * 8 + c2 + (e & 1);
        const int qi = q0 + r, kj = k0 + j, rel = kj - qi + w;
        float v = s[n][e];
        if (rel >= 0 && rel < nrel) v += Rq[r * nrel + rel];
        v *= scale;
        if (mq[r] * mk[j] == 0.f) v = -1e4f;
        if (kj >= T) v = -INFINITY;
        s[n][e] = v;
        tmax[e >> 1] = fmaxf(tmax[e >> 1], v);
      }
    }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      tmax[hh] = fmaxf(tmax[hh], __shfl_xor_sync(0xffffffffu, tmax[hh], 1));
      tmax[hh] = fmaxf(tmax[hh], __shfl_xor_sync(0xffffffffu, tmax[hh], 2));
    }
    float corr[2];
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const float nm = fmaxf(row_max[hh], tmax[hh]);
      corr[hh] = __expf(row_max[hh] - nm);
      row_max[hh] = nm;
    }
    // the diagonal weights collected so far follow the running maximum (one thread of the quad rescales a row)
    if ((lane & 3) == 0) {
      for (int m = 0; m < nrel; ++m) { Pw[r_lo * nrel + m] *= corr[0]; Pw[r_hi * nrel + m] *= corr[1]; }
    }
    __syncwarp();
    float psum[2] = {0.f, 0.f};
    uint32_t pf[4][4];                                               // P as the A operand of the value GEMM
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      float p[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        p[e] = __expf(s[n][e] - row_max[e >> 1]);
        psum[e >> 1] += p[e];
        const int r = (e < 2) ? r_lo : r_hi, j = n * 8 + c2 + (e & 1);
        const int rel = (k0 + j) - (q0 + r) + w;
        if (rel >= 0 && rel < nrel && k0 + j < T) Pw[r * nrel + rel] = p[e];
      }
      pf[n >> 1][(n & 1) * 2 + 0] = pack2(p[0], p[1]);
      pf[n >> 1][(n & 1) * 2 + 1] = pack2(p[2], p[3]);
    }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      psum[hh] += __shfl_xor_sync(0xffffffffu, psum[hh], 1);
      psum[hh] += __shfl_xor_sync(0xffffffffu, psum[hh], 2);
      row_sum[hh] = row_sum[hh] * corr[hh] + psum[hh];
    }
#pragma unroll
    for (int n = 0; n < NO; ++n) { oacc[n][0] *= corr[0]; oacc[n][1] *= corr[0]; oacc[n][2] *= corr[1]; oacc[n][3] *= corr[1]; }
    // ---- O += P V
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {                                 // 16 keys per step
#pragma unroll
      for (int np = 0; np < NO / 2; ++np) {                          // two d n-tiles per ldmatrix.x4.trans
        uint32_t b0, b1, b2, b3;
        const int vrow = kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, vcol = np * 16 + (lane >> 4) * 8;
        ldmatrix_x4_trans((uint32_t)__cvta_generic_to_shared(Vs + vrow * RS + vcol), b0, b1, b2, b3);
        mma_bf16(oacc[2 * np], pf[kt], b0, b1);
        mma_bf16(oacc[2 * np + 1], pf[kt], b2, b3);
      }
    }
  }
  __syncwarp();
  // ---- o = (O + sum_m pw[m] Ev[m]) / row_sum  -> bf16, channels-last
  const float inv[2] = {1.0f / row_sum[0], 1.0f / row_sum[1]};
#pragma unroll
  for (int n = 0; n < NO; ++n) {
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int r = hh ? r_hi : r_lo, t = q0 + r, d = n * 8 + c2;
      float v0 = oacc[n][2 * hh], v1 = oacc[n][2 * hh + 1];
      for (int m = 0; m < nrel; ++m) {
        const float pw = Pw[r * nrel + m];
        v0 = fmaf(pw, __ldg(Ev + m * DK + d), v0);
        v1 = fmaf(pw, __ldg(Ev + m * DK + d + 1), v1);
      }
      if (t < T)
        *reinterpret_cast<uint32_t*>(o + ((long long)b * T + t) * H + h * DK + d) = pack2(v0 * inv[hh], v1 * inv[hh]);
    }
  }
  (void)NREL_MAX;
}

inline size_t relenc_attention_bf16_smem(int dk, int w) {
  return (size_t)3 * 64 * (dk + 8) * 2 + sizeof(float) * ((size_t)2 * 64 * (2 * w + 1) + 128);
}

// x = LayerNorm(s) over the C channels of every row of s [B*T, C] (bf16; s already holds residual + branch), then
// x += g (next layer's condition: fp32 per-utterance [B, C] if g_t == 0, per-frame [B, C, T] otherwise), x *= mask;
// written back in place as bf16.  One warp per row, fp32 statistics (two passes over registers).
template <int MAXC>
__global__ void relenc_layernorm_bf16_kernel(__nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                                             const float* __restrict__ beta, const float* __restrict__ g, int g_t,
                                             const float* __restrict__ mask, float eps, int C, int T, long long rows) {
  constexpr int PER = MAXC / 64;              // bf16 pairs per lane
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int b = (int)(row / T), t = (int)(row % T);
  __nv_bfloat16* xr = x + row * C;
  float v[PER][2];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = 2 * lane + 64 * i;
    v[i][0] = v[i][1] = 0.f;
    if (c < C) {
      const __nv_bfloat162 p = *reinterpret_cast<const __nv_bfloat162*>(xr + c);
      v[i][0] = __low2float(p); v[i][1] = __high2float(p);
      s += v[i][0] + v[i][1];
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  const float mean = s / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = 2 * lane + 64 * i;
    if (c < C) { const float d0 = v[i][0] - mean, d1 = v[i][1] - mean; q = fmaf(d0, d0, fmaf(d1, d1, q)); }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) q += __shfl_xor_sync(0xffffffffu, q, off);
  const float rstd = rsqrtf(q / (float)C + eps);
  const float mkv = mask[row];
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = 2 * lane + 64 * i;
    if (c < C) {
      float y0 = (v[i][0] - mean) * rstd * gamma[c] + beta[c], y1 = (v[i][1] - mean) * rstd * gamma[c + 1] + beta[c + 1];
      if (g) {
        if (g_t) { y0 += g[((long long)b * C + c) * T + t]; y1 += g[((long long)b * C + c + 1) * T + t]; }
        else { y0 += g[(long long)b * C + c]; y1 += g[(long long)b * C + c + 1]; }
      }
      *reinterpret_cast<__nv_bfloat162*>(xr + c) = __floats2bfloat162_rn(y0 * mkv, y1 * mkv);
    }
  }
}

// x[b, t, c] = bf16((x_in[b, c, t] + g) * mask): the entry transpose of the bf16 encoder with the first layer's condition
__global__ void relenc_entry_bf16_kernel(const float* __restrict__ x, const float* __restrict__ g, int g_t,
                                         const float* __restrict__ mask, __nv_bfloat16* __restrict__ y, int C, int T) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, t = t0 + tx;
    float v = 0.f;
    if (c < C && t < T) {
      v = x[((long long)b * C + c) * T + t];
      if (g) v += g_t ? g[((long long)b * C + c) * T + t] : g[(long long)b * C + c];
      v *= mask[(long long)b * T + t];
    }
    tile[i][tx] = v;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, c = c0 + tx;
    if (t < T && c < C) y[((long long)b * T + t) * C + c] = __float2bfloat16(tile[tx][i]);
  }
}

}  // namespace vsg
