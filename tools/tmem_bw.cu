// Microbenchmark: tcgen05.ld / tcgen05.st throughput per SM (development tool).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
#define R4(a, i) "=r"(a[i]), "=r"(a[i + 1]), "=r"(a[i + 2]), "=r"(a[i + 3])
#define R16(a, i) R4(a, i), R4(a, i + 4), R4(a, i + 8), R4(a, i + 12)
#define W4(a, i) "r"(a[i]), "r"(a[i + 1]), "r"(a[i + 2]), "r"(a[i + 3])
#define W16(a, i) W4(a, i), W4(a, i + 4), W4(a, i + 8), W4(a, i + 12)

// mode 0: ld x32, 1: ld x64, 2: st x32, 3: ld x32 with pack::16b (16-bit data), 4: ld 16x256b.x8 (32 regs)
template <int MODE>
__global__ void k(int iters, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t tb;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tb)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t t = tb + (((uint32_t)(warp & 3) * 32) << 16) + (warp >> 2) * 128;
  uint32_t r[64];
  for (int i = 0; i < 64; ++i) r[i] = i + threadIdx.x;
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : R16(r, 0), R16(r, 16) : "r"(t + (it & 1) * 32));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += r[0] ^ r[31];
    } else if (MODE == 1) {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
                   "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
                   : R16(r, 0), R16(r, 16), R16(r, 32), R16(r, 48) : "r"(t + (it & 1) * 64));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += r[0] ^ r[63];
    } else if (MODE == 2) {
      asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                   ::"r"(t + (it & 1) * 32), W16(r, 0), W16(r, 16) : "memory");
      if ((it & 7) == 7) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    } else if (MODE == 3) {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.pack::16b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : R16(r, 0), R16(r, 16) : "r"(t + (it & 1) * 64));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += r[0] ^ r[31];
    }
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

template <int MODE>
int run(const char* name, int warps, int regs_per_instr, int bytes_per_reg) {
  long long* d_c; uint32_t* d_s; const int iters = 20000;
  CK(cudaMalloc(&d_c, 148 * 8)); CK(cudaMalloc(&d_s, 148 * 1024 * 4));
  k<MODE><<<148, warps * 32>>>(iters, d_c, d_s);
  CK(cudaDeviceSynchronize());
  k<MODE><<<148, warps * 32>>>(iters, d_c, d_s);
  CK(cudaDeviceSynchronize());
  long long c[148]; CK(cudaMemcpy(c, d_c, sizeof(c), cudaMemcpyDeviceToHost));
  double cyc = (double)c[0] / iters;
  double bytes = (double)warps * 32 * regs_per_instr * bytes_per_reg;
  printf("%-28s warps=%2d : %7.1f cycles/iter, %7.1f B/clk/SM (%.1f B/clk per warp)\n", name, warps, cyc, bytes / cyc, bytes / cyc / warps);
  cudaFree(d_c); cudaFree(d_s);
  return 0;
}
int main() {
  for (int w : {4, 8, 16}) {
    run<0>("ld 32x32b.x32 + wait", w, 32, 4);
    run<1>("ld 32x32b.x64 + wait", w, 64, 4);
    run<2>("st 32x32b.x32 (wait/8)", w, 32, 4);
    run<3>("ld 32x32b.pack16.x32 + wait", w, 32, 4);
  }
  return 0;
}
