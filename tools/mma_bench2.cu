// Microbenchmark 2: tcgen05.mma (cta_group::1, kind::f16, M=128) throughput per SM as a function of N and of the number
// of issuing warps (one elected thread each, separate TMEM accumulators), with one or two CTAs per SM.
// Question answered: how many issuing warps does a small-N (16/32/64 channel) convolution need to hide the
// per-thread issue latency of ~80 cycles?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mma_bench2 tools/mma_bench2.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1u << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1u << 46;
  d |= (uint64_t)(layout & 7u) << 61;
  return d;
}

constexpr int kMaxIssuers = 8;

// mode bit 0: constant descriptors (no per-MMA address arithmetic)
__global__ void __launch_bounds__(32 * (kMaxIssuers + 1)) bench(int N, int n_mma, int kc, int n_issuers, int mode, int tmem_cols,
                                                               long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[kMaxIssuers];
  __shared__ uint32_t tslot;
  __shared__ long long t_end[kMaxIssuers], t_beg[kMaxIssuers];
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 40 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxIssuers; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMaxIssuers) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tslot;
  if (warp < n_issuers && lane == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t row_bytes = kc * 2, sbo = 8 * row_bytes, layout = kc == 64 ? 2u : kc == 32 ? 4u : 6u;
    const uint32_t a0 = base, b0 = base + 24 * 1024;
    const uint32_t acc = tmem + (uint32_t)(warp * (tmem_cols / n_issuers));
    const int mine = n_mma / n_issuers;
    const long long t0 = clock64();
    if (mode & 1) {
      const uint64_t ad = make_desc(a0, sbo, layout), bd = make_desc(b0, sbo, layout);
#pragma unroll 8
      for (int i = 0; i < mine; ++i)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(acc), "l"(ad), "l"(bd), "r"(idesc), "r"(i > 1 ? 1u : 0u) : "memory");
    } else {
      for (int i = 0; i < mine; ++i) {
        const uint32_t aa = a0 + (uint32_t)(i % 11) * row_bytes;
        const uint64_t ad = make_desc(aa, sbo, layout), bd = make_desc(b0, sbo, layout);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(acc), "l"(ad), "l"(bd), "r"(idesc), "r"(i > 1 ? 1u : 0u) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[warp])) : "memory");
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bars[warp])) : "memory");
      if (clock64() - t0 > 2000000000LL) break;
    }
    t_beg[warp] = t0;
    t_end[warp] = clock64();
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    long long b = t_beg[0], e = t_end[0];
    for (int i = 1; i < n_issuers; ++i) { b = min(b, t_beg[i]); e = max(e, t_end[i]); }
    out[0] = e - b;
  }
  if (warp == kMaxIssuers) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols));
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  const int n_mma = 8000;
  printf("%5s %4s %3s %8s %5s %12s %10s\n", "ctas", "N", "kc", "issuers", "mode", "cyc/mma/SM", "TF/s@1.9G");
  for (int ctas : {1, 2})
    for (int N : {16, 32, 64, 128})
      for (int kc : {64, 16})
        for (int ni : {1, 2, 4, 8})
          for (int mode : {0, 1}) {
            if (kc == 16 && (mode == 1 || N > 32)) continue;
            const int cols = 512 / ctas;
            if (N * ni > cols) continue;
            for (int rep = 0; rep < 2; ++rep) bench<<<148 * ctas, 32 * (kMaxIssuers + 1), 44 * 1024>>>(N, n_mma, kc, ni, mode, cols, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            long long h[2];
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            const double cyc = (double)h[0] / (n_mma * ctas);   // per SM: ctas CTAs issue n_mma each in the same time
            printf("%5d %4d %3d %8d %5d %12.1f %10.0f\n", ctas, N, kc, ni, mode, cyc,
                   2.0 * 128 * N * 16 / cyc * 1.9e9 * 148 / 1e12);
          }
  return 0;
}
