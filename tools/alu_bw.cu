// Microbenchmark: issue rate of the epilogue's instruction types, 1 warp per SM sub-partition (development tool).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
template <int MODE>
__global__ void k(int iters, long long* cycles, uint32_t* sink, float seed) {
  float a[32]; uint32_t h[16];
  for (int i = 0; i < 32; ++i) a[i] = seed * (i + threadIdx.x);
  for (int i = 0; i < 16; ++i) h[i] = i * 0x10001u + threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) { __half2 v = __floats2half2_rn(a[2 * i], a[2 * i + 1]); h[i] ^= *reinterpret_cast<uint32_t*>(&v); a[2 * i] += 1.0f; }
      else if (MODE == 1) { __half2 v = *reinterpret_cast<__half2*>(&h[i]); v = __hmul2(v, __floats2half2_rn(0.01f, 0.01f)); h[i] = *reinterpret_cast<uint32_t*>(&v); }
      else if (MODE == 2) { __half2 v = *reinterpret_cast<__half2*>(&h[i]); v = __hmax2(v, __floats2half2_rn(0.5f, 0.25f)); h[i] = *reinterpret_cast<uint32_t*>(&v) + 1; }
      else if (MODE == 3) { a[2 * i] = fmaxf(a[2 * i], 0.01f * a[2 * i]); }
      else if (MODE == 4) { a[2 * i] = fmaf(a[2 * i], 1.0001f, 0.5f); }
    }
  }
  long long t1 = clock64();
  uint32_t acc = 0;
  for (int i = 0; i < 16; ++i) acc ^= h[i] ^ __float_as_uint(a[2 * i]);
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> void run(const char* name, int warps, int per_iter) {
  long long* d_c; uint32_t* d_s; const int iters = 4000;
  cudaMalloc(&d_c, 148 * 8); cudaMalloc(&d_s, 148 * 1024 * 4);
  k<MODE><<<148, warps * 32>>>(iters, d_c, d_s, 1.5f); cudaDeviceSynchronize();
  k<MODE><<<148, warps * 32>>>(iters, d_c, d_s, 1.5f); cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
  printf("%-34s warps/SM=%2d: %6.2f cycles per warp-instruction-group (%d instr): %.2f cyc/instr/SMSP\n", name, warps,
         (double)c / iters, per_iter, (double)c / iters / per_iter * (warps / 4.0 > 1 ? 1.0 : 1.0));
  cudaFree(d_c); cudaFree(d_s);
}
int main() {
  for (int w : {4, 8}) {
    run<0>("F2FP pack (+FADD) x16", w, 16);
    run<1>("HMUL2 x16", w, 16);
    run<2>("HMNMX2 (+IADD) x16", w, 16);
    run<3>("FMUL+FMNMX x16", w, 16);
    run<4>("FFMA x16", w, 16);
  }
  return 0;
}
