// bf16 (throughput-mode) glue kernels of the relative-position transformer encoder (reference modules/rel_transformer.py):
// the fused residual-sum -> channel LayerNorm -> condition -> mask step (:24-42, :293-303) and the entry transpose.  Layout: channels-last bf16 [B, T, C],
// the layout of the tcgen05 convolution kernels that run the projections and the FFN around them.
//
// The attention kernel itself (tcgen05, TMEM, TMA) lives in attn_tc.cuh.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace vsg {

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
