// Microbenchmark: issue rate of tcgen05.mma (cta_group::1, kind::f16, M=128) as a function of N, of the operand
// swizzle / row width, and of a tcgen05.fence::after_thread_sync between instructions.  One CTA per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mma_bench tools/mma_bench.cu
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

// mode bits: 0 = fence between MMAs, 1 = alternate between two accumulators, 2 = shift A start by (i%11) rows
__global__ void __launch_bounds__(128, 1) bench(int N, int M, int n_mma, int kc, int mode, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tslot;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 58 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tslot;
  if (warp == 1 && (mode & 8)) {
    // whole-warp issue: uniform control flow, elect.sync guards the instruction (CUTLASS style)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)N >> 3) << 17) | (((uint32_t)M >> 4) << 24);
    const uint32_t row_bytes = kc * 2, sbo = 8 * row_bytes, layout = kc == 64 ? 2u : kc == 32 ? 4u : 6u;
    const uint32_t a0 = base, b0 = base + 24 * 1024;
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const uint32_t aa = a0 + ((mode & 4) ? (uint32_t)(i % 11) * row_bytes : 0u);
      const uint64_t ad = make_desc(aa, sbo, layout), bd = make_desc(b0, sbo, layout);
      uint32_t elected;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(elected));
      if (elected)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(i > 1 ? 1u : 0u) : "memory");
      __syncwarp();
    }
    const long long t1 = clock64();
    if (lane == 0) {
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      uint32_t ok = 0;
      while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        if (clock64() - t0 > 2000000000LL) break;
      }
      const long long t2 = clock64();
      if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  } else if ((mode & 16) && (threadIdx.x == 0 || threadIdx.x == 64)) {
    const int who = threadIdx.x == 0 ? 0 : 1;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)N >> 3) << 17) | (((uint32_t)M >> 4) << 24);
    const uint32_t row_bytes = kc * 2, sbo = 8 * row_bytes, layout = kc == 64 ? 2u : kc == 32 ? 4u : 6u;
    const uint32_t a0 = base, b0 = base + 24 * 1024;
    const long long t0 = clock64();
    for (int i = 0; i < n_mma / 2; ++i) {
      const uint32_t aa = a0 + (uint32_t)(i % 11) * row_bytes;
      const uint64_t ad = make_desc(aa, sbo, layout), bd = make_desc(b0, sbo, layout);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem + who * 256), "l"(ad), "l"(bd), "r"(idesc), "r"(i > 1 ? 1u : 0u) : "memory");
    }
    const long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(who ? &bar2 : &bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(who ? &bar2 : &bar)) : "memory");
      if (clock64() - t0 > 2000000000LL) break;
    }
    const long long t2 = clock64();
    if (blockIdx.x == 0 && who == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  } else if (threadIdx.x == 0 && !(mode & 8) && !(mode & 16)) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)N >> 3) << 17) | (((uint32_t)M >> 4) << 24);
    const uint32_t row_bytes = kc * 2, sbo = 8 * row_bytes, layout = kc == 64 ? 2u : kc == 32 ? 4u : 6u;
    const uint32_t a0 = base, b0 = base + 24 * 1024;
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      if (mode & 1) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t aa = a0 + ((mode & 4) ? (uint32_t)(i % 11) * row_bytes : 0u);
      const uint32_t d = tmem + ((mode & 2) ? (uint32_t)((i & 1) * 256) : 0u);
      const uint64_t ad = make_desc(aa, sbo, layout), bd = make_desc(b0, sbo, layout);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(i > 1 ? 1u : 0u) : "memory");
    }
    const long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
      if (clock64() - t0 > 2000000000LL) break;
    }
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int n_mma = 2000;
  printf("%4s %4s %3s %5s %10s %10s\n", "M", "N", "kc", "mode", "issue/mma", "total/mma");
  for (int M : {128})
    for (int kc : {64})
      for (int N : {16, 64, 128, 256})
        for (int mode : {4, 16}) {
          for (int rep = 0; rep < 2; ++rep) bench<<<148, 128, 60 * 1024>>>(N, M, n_mma, kc, mode, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          long long h[2];
          cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          printf("%4d %4d %3d %5d %10.1f %10.1f\n", M, N, kc, mode, (double)h[0] / n_mma, (double)h[1] / n_mma);
        }
  return 0;
}
