// Microbenchmark: back-to-back tcgen05.mma issue/execute rate per SM (development tool).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
  return pred != 0;
}
// mode 0: A from TMEM (TS), 1: A from smem (SS)
__global__ void __launch_bounds__(32, 1) k(int N, int n_mma, int mode, long long* cycles) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tb;
  for (int i = threadIdx.x; i < (256 * 16 * 2 + 128 * 16 * 2) / 4; i += 32) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tb)) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tb;
  const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint64_t bd = make_desc(smem_u32(smem), N * 16, 128);
  const uint64_t ad = make_desc(smem_u32(smem) + 256 * 16 * 2, 128 * 16, 128);
  long long t0 = clock64();
  if (elect_one()) {
    for (int i = 0; i < n_mma; ++i) {
      if (mode == 0)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                     ::"r"(tmem), "r"(tmem + 256 + (i & 7) * 8), "l"(bd), "r"(idesc), "r"(1u) : "memory");
      else
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                     ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  __syncwarp();
  long long t1 = clock64();
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
  long long t2 = clock64();
  if (threadIdx.x == 0) { cycles[blockIdx.x * 2] = t1 - t0; cycles[blockIdx.x * 2 + 1] = t2 - t0; }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
int main() {
  long long* d; cudaMalloc(&d, 148 * 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  const int n = 2000;
  for (int mode = 0; mode < 2; ++mode)
    for (int N : {16, 64, 128, 256}) {
      k<<<148, 32, 16384>>>(N, n, mode, d); cudaDeviceSynchronize();
      k<<<148, 32, 16384>>>(N, n, mode, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long c[2]; cudaMemcpy(c, d, 16, cudaMemcpyDeviceToHost);
      printf("%s N=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (floor %.0f) %s\n", mode ? "SS" : "TS", N, (double)c[0] / n,
             (double)c[1] / n, N / 2.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
