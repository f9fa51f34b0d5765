// Microbenchmark (development tool, round 2): cost of the softplus epilogue of one 128-column accumulator row
// (fp32 accumulator -> softplus -> packed fp16) for 1 or 2 warps per SM sub-partition, plus raw pipe rates.
// Inputs are made opaque per iteration (empty asm) so that nothing is hoisted out of the timed loop
// (tools/softplus_bw.cu let the compiler hoist half of the MUFU work).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/epi_bw tools/epi_bw.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#define DI __device__ __forceinline__
DI float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
DI float lg2f(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
DI uint32_t ex2h2(uint32_t x) { uint32_t y; asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
DI uint32_t h2u(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
DI __half2 u2h(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
DI __half2 hc(float v) { return __floats2half2_rn(v, v); }

// log1p(u) ~ u*(Q1 + Q2 u + Q3 u^2 + Q4 u^3) on [0,1], max error 7.1e-5
#define Q1 9.974489612e-01f
#define Q2 -4.713012532e-01f
#define Q3 2.256856408e-01f
#define Q4 -5.875710231e-02f
// degree 5, max error 9.9e-6
#define R1 9.994943976e-01f
#define R2 -4.919007447e-01f
#define R3 2.894552203e-01f
#define R4 -1.360436010e-01f
#define R5 3.215182172e-02f

// V0: two MUFU, fp32 (the shipped form)
DI uint32_t v0(float a, float b) {
  const float ra = fmaf(0.6931471805599453f, lg2f(1.0f + ex2f(-1.4426950408889634f * fabsf(a))), fmaxf(a, 0.0f));
  const float rb = fmaf(0.6931471805599453f, lg2f(1.0f + ex2f(-1.4426950408889634f * fabsf(b))), fmaxf(b, 0.0f));
  return h2u(__floats2half2_rn(ra, rb));
}
// V1: one MUFU, fp32 Horner degree 4 (no constant), combine with max(x,0) in the last FMA
DI float sp1(float x) {
  const float u = ex2f(-1.4426950408889634f * fabsf(x));
  float p = fmaf(Q4, u, Q3); p = fmaf(p, u, Q2); p = fmaf(p, u, Q1);
  return fmaf(p, u, fmaxf(x, 0.0f));
}
DI uint32_t v1(float a, float b) { return h2u(__floats2half2_rn(sp1(a), sp1(b))); }
// V2: half2 everything: F2FP first, HMUL2, 2 x MUFU.EX2.F16, PRMT, HFMA2 Horner, HMNMX2
DI uint32_t v2(float a, float b) {
  const __half2 x = __floats2half2_rn(a, b);
  const __half2 t = __hmul2(__habs2(x), hc(-1.4426950408889634f));
  const __half2 u = u2h(ex2h2(h2u(t)));
  __half2 p = __hfma2(hc(Q4), u, hc(Q3)); p = __hfma2(p, u, hc(Q2)); p = __hfma2(p, u, hc(Q1));
  return h2u(__hfma2(p, u, __hmax2(x, hc(0.0f))));
}
// V3: fp32 exponent (FMUL + MUFU f32), then pack u and x and run the polynomial on half2
DI uint32_t v3(float a, float b) {
  const float ua = ex2f(-1.4426950408889634f * fabsf(a)), ub = ex2f(-1.4426950408889634f * fabsf(b));
  const __half2 u = __floats2half2_rn(ua, ub), x = __floats2half2_rn(a, b);
  __half2 p = __hfma2(hc(Q4), u, hc(Q3)); p = __hfma2(p, u, hc(Q2)); p = __hfma2(p, u, hc(Q1));
  return h2u(__hfma2(p, u, __hmax2(x, hc(0.0f))));
}
// V4: no MUFU: 2^t on the half2 FMA pipe (magic-number rounding, degree-3 polynomial of the fraction, exponent by integer add)
DI __half2 exp2_neg_h2(__half2 t) {   // t in [-15, 0]
  const __half2 magic = hc(1039.0f);                     // ulp 1 in [1024, 2048): s = round(t) + 1039
  const __half2 s = __hadd2(t, magic);
  const __half2 n = __hsub2(s, magic);
  const __half2 f = __hsub2(t, n);                       // [-0.5, 0.5]
  __half2 p = __hfma2(hc(5.5504109e-2f), f, hc(2.4022651e-1f)); p = __hfma2(p, f, hc(6.9314718e-1f)); p = __hfma2(p, f, hc(1.0f));
  // 2^n as a half: exponent field = n + 15 = (s bits & 0x1f) (s = 1024 + (n + 15): the low mantissa bits hold n + 15)
  const uint32_t e = (h2u(s) & 0x001f001fu) << 10;
  return __hmul2(p, u2h(e));
}
DI uint32_t v4(float a, float b) {
  const __half2 x = __floats2half2_rn(a, b);
  const __half2 t = __hmax2(__hmul2(__habs2(x), hc(-1.4426950408889634f)), hc(-14.0f));
  const __half2 u = exp2_neg_h2(t);
  __half2 p = __hfma2(hc(Q4), u, hc(Q3)); p = __hfma2(p, u, hc(Q2)); p = __hfma2(p, u, hc(Q1));
  return h2u(__hfma2(p, u, __hmax2(x, hc(0.0f))));
}
// V6: one MUFU, fp32 Horner degree 5
DI float sp6(float x) {
  const float u = ex2f(-1.4426950408889634f * fabsf(x));
  float p = fmaf(R5, u, R4); p = fmaf(p, u, R3); p = fmaf(p, u, R2); p = fmaf(p, u, R1);
  return fmaf(p, u, fmaxf(x, 0.0f));
}
DI uint32_t v6(float a, float b) { return h2u(__floats2half2_rn(sp6(a), sp6(b))); }
// V7: leaky relu on half2 (reference point: F2FP + HMUL2 + HMNMX2)
DI uint32_t v7(float a, float b) { const __half2 h = __floats2half2_rn(a, b); return h2u(__hmax2(h, __hmul2(h, hc(0.01f)))); }
// raw pipe rates (per pair)
DI uint32_t p_mufu(float a, float b) { return __float_as_uint(ex2f(a)) ^ __float_as_uint(ex2f(b)); }
DI uint32_t p_mufu16(float a, float b) { return ex2h2(__float_as_uint(a)) ^ ex2h2(__float_as_uint(b)); }   // 4 MUFU.F16 + 2 PRMT
DI uint32_t p_f2fp(float a, float b) { return h2u(__floats2half2_rn(a, b)); }
DI uint32_t p_hfma(float a, float b) {
  __half2 x = u2h(__float_as_uint(a)), y = u2h(__float_as_uint(b));
  x = __hfma2(x, y, y); x = __hfma2(x, y, y); x = __hfma2(x, y, y); x = __hfma2(x, y, y);
  return h2u(x);
}
DI uint32_t p_ffma(float a, float b) {
  float x = a; x = fmaf(x, b, b); x = fmaf(x, b, b); x = fmaf(x, b, b); x = fmaf(x, b, b);
  return __float_as_uint(x);
}

// V9: one MUFU, packed fp32 pairs: FMUL2, 2 x MUFU, 4 x FFMA2, 2 x FMNMX, F2FP
DI uint64_t pk2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
DI void up2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
DI uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
DI uint64_t mul2(uint64_t a, uint64_t b) { uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
DI uint32_t v9(float a, float b) {
  const uint64_t t = mul2(pk2(fabsf(a), fabsf(b)), pk2(-1.4426950408889634f, -1.4426950408889634f));
  float ta, tb; up2(t, ta, tb);
  const uint64_t u = pk2(ex2f(ta), ex2f(tb));
  uint64_t q = fma2(pk2(Q4, Q4), u, pk2(Q3, Q3));
  q = fma2(q, u, pk2(Q2, Q2)); q = fma2(q, u, pk2(Q1, Q1));
  q = fma2(q, u, pk2(fmaxf(a, 0.0f), fmaxf(b, 0.0f)));
  float ra, rb; up2(q, ra, rb);
  return h2u(__floats2half2_rn(ra, rb));
}
DI uint32_t p_ffma2(float a, float b) {
  uint64_t x = pk2(a, b), y = pk2(b, a);
  x = fma2(x, y, y); x = fma2(x, y, y); x = fma2(x, y, y); x = fma2(x, y, y);
  float ra, rb; up2(x, ra, rb);
  return __float_as_uint(ra) ^ __float_as_uint(rb);
}
template <int V> DI uint32_t conv(float a, float b) {
  if (V == 9) return v9(a, b); if (V == 15) return p_ffma2(a, b);
  if (V == 0) return v0(a, b); if (V == 1) return v1(a, b); if (V == 2) return v2(a, b); if (V == 3) return v3(a, b);
  if (V == 4) return v4(a, b); if (V == 6) return v6(a, b); if (V == 7) return v7(a, b);
  if (V == 10) return p_mufu(a, b); if (V == 11) return p_mufu16(a, b); if (V == 12) return p_f2fp(a, b);
  if (V == 13) return p_hfma(a, b); if (V == 14) return p_ffma(a, b);
  return 0;
}
// V5: three MUFU pairs + one FMA-pipe pair per four pairs
template <int V> DI void conv_row(const float* a, uint32_t* pk) {
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    if (V == 5) pk[i] = (i & 3) == 3 ? v4(a[2 * i], a[2 * i + 1]) : v2(a[2 * i], a[2 * i + 1]);
    else if (V == 8) pk[i] = (i & 3) == 3 ? v4(a[2 * i], a[2 * i + 1]) : v1(a[2 * i], a[2 * i + 1]);
    else pk[i] = conv<V>(a[2 * i], a[2 * i + 1]);
  }
}

template <int V>
__global__ void __launch_bounds__(256, 1) k(int iters, long long* cycles, uint32_t* sink, float seed) {
  // the inputs of every iteration come out of shared memory (32 x LDS.128 per row, like the tcgen05.ld of the real
  // epilogue): volatile, so the conversion cannot be hoisted out of the timed loop
  extern __shared__ float4 sm[];
  for (int i = 0; i < 32; ++i) {
    float4 v;
    v.x = seed * (float)(((4 * i) * 37 + threadIdx.x * 11) % 257 - 128) / 16.0f;
    v.y = seed * (float)(((4 * i + 1) * 37 + threadIdx.x * 11) % 257 - 128) / 16.0f;
    v.z = seed * (float)(((4 * i + 2) * 37 + threadIdx.x * 11) % 257 - 128) / 16.0f;
    v.w = seed * (float)(((4 * i + 3) * 37 + threadIdx.x * 11) % 257 - 128) / 16.0f;
    sm[i * blockDim.x + threadIdx.x] = v;
  }
  uint32_t acc = 0;
  __syncthreads();
  long long t0, t1;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0) :: "memory");
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    float a[128];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const uint32_t addr = (uint32_t)__cvta_generic_to_shared(&sm[i * blockDim.x + threadIdx.x]);
      asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a[4 * i]), "=f"(a[4 * i + 1]), "=f"(a[4 * i + 2]), "=f"(a[4 * i + 3]) : "r"(addr) : "memory");
    }
    uint32_t pk[64];
    conv_row<V>(a, pk);
#pragma unroll
    for (int i = 0; i < 64; i += 4) {   // consume 16 packed words per "tcgen05.st" (cheap: 1 op per 4 words)
      uint32_t w;
      asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(w) : "r"(pk[i]), "r"(pk[i + 1]), "r"(pk[i + 2]));
      asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(acc) : "r"(w), "r"(pk[i + 3]), "r"(acc));
    }
  }
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1) :: "memory");
  if ((threadIdx.x & 31) == 0) cycles[blockIdx.x * 8 + (threadIdx.x >> 5)] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int V> __global__ void kacc(float* err, float* err_abs) {
  float worst = 0.f, worst_abs = 0.f;
  for (int j = -24000; j <= 24000; ++j) {
    const float x = j * 0.001f;
    uint32_t pk[64]; float a[128];
    for (int i = 0; i < 128; ++i) a[i] = x;
    if (V == 5 || V == 8) { pk[0] = v4(x, x); } else pk[0] = conv<V>(x, x);
    const float got = __half2float(__ushort_as_half((unsigned short)(pk[0] & 0xffff)));
    const double ref = V == 7 ? (x > 0 ? x : 0.01 * x) : (x > 20 ? x : log1p(exp((double)x)));
    const float ulp = fmaxf(ldexpf(1.0f, (int)floor(log2(fmax(fabs(ref), 6.1e-5))) - 10), 5.96e-8f);
    worst = fmaxf(worst, fabsf(got - (float)ref) / ulp);
    worst_abs = fmaxf(worst_abs, fabsf(got - (float)ref));
  }
  *err = worst; *err_abs = worst_abs;
}
template <int V> void run(const char* name) {
  long long* d_c; uint32_t* d_s; float* d_e; const int iters = 2000;
  cudaMalloc(&d_c, 148 * 8 * 8); cudaMalloc(&d_s, 148 * 1024 * 4); cudaMalloc(&d_e, 8);
  double res[2]; double evt[2];
  cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 512);
  for (int w = 0; w < 2; ++w) {
    const int threads = w == 0 ? 128 : 256;   // 1 or 2 warps per SM sub-partition
    k<V><<<148, threads, threads * 512>>>(iters, d_c, d_s, 1.0f); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<V><<<148, threads, threads * 512>>>(iters, d_c, d_s, 1.0f);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1); evt[w] = ms * 1e3 / iters;
    long long c[8]; cudaMemcpy(c, d_c, 64, cudaMemcpyDeviceToHost);
    long long m = 0; for (int i = 0; i < threads / 32; ++i) m = c[i] > m ? c[i] : m;
    res[w] = (double)m / iters;
  }
  float e[2] = {0, 0};
  if (V < 10) { kacc<V><<<1, 1>>>(d_e, d_e + 1); cudaDeviceSynchronize(); cudaMemcpy(e, d_e, 8, cudaMemcpyDeviceToHost); }
  printf("%-58s %7.0f cycles/row alone (%.3f us)  %7.0f with 2 warps per SMSP (%.3f us)  worst err %.2f fp16 ulp, abs %.2e  [%s]\n", name, res[0], evt[0], res[1], evt[1],
         e[0], e[1], cudaGetErrorString(cudaGetLastError()));
  cudaFree(d_c); cudaFree(d_s); cudaFree(d_e);
}
int main() {
  run<0>("V0 ex2 + lg2 fp32 (2 MUFU, shipped)");
  run<1>("V1 ex2 + fp32 Horner deg4");
  run<6>("V6 ex2 + fp32 Horner deg5");
  run<2>("V2 half2: HMUL2, 2 MUFU.EX2.F16, PRMT, HFMA2 deg4");
  run<3>("V3 fp32 ex2, pack, HFMA2 deg4");
  run<4>("V4 no MUFU: half2 exp2 on the FMA pipe + HFMA2 deg4");
  run<5>("V5 3 x V2 + 1 x V4 per four pairs");
  run<8>("V8 3 x V1 + 1 x V4 per four pairs");
  run<9>("V9 FMUL2, 2 MUFU, 4 FFMA2 (packed fp32), 2 FMNMX, F2FP");
  run<7>("V7 leaky relu half2 (reference point)");
  run<10>("raw: 2 MUFU.EX2 f32 per pair");
  run<11>("raw: 4 MUFU.EX2.F16 + 2 PRMT per pair");
  run<12>("raw: 1 F2FP per pair");
  run<13>("raw: 4 dependent HFMA2 per pair");
  run<14>("raw: 4 dependent FFMA per pair");
  run<15>("raw: 4 dependent FFMA2 per pair (8 FMA)");
  return 0;
}
