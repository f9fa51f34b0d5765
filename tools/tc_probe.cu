// tcgen05 probe: validates the UMMA shared-memory descriptor convention, the TMEM A-operand
// packing and the instruction descriptor used by neural_raytracing_b200/csrc/nrt_tc.cu on a
// single 128 x N x K tile against a CPU reference.  Development tool (run under gpurun).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tools/tc_probe.cu
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // version = 1 (Blackwell)
  return d;                 // layout_type = 0 (no swizzle), base_offset = 0
}

__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  for (int it = 0; it < 20000000; ++it) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}

// mode 0: A from smem (SS); mode 1: A from TMEM (TS).  fmt 0: fp16, 1: bf16.
__global__ void __launch_bounds__(128, 1)
probe(const uint16_t* __restrict__ A, const uint16_t* __restrict__ B, float* __restrict__ D, int N, int K,
      uint32_t a_lbo, uint32_t a_sbo, uint32_t b_lbo, uint32_t b_sbo, int mode, int fmt, int* status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint16_t* sA = (uint16_t*)smem;                        // [K/8][128][8]
  uint16_t* sB = (uint16_t*)(smem + 128 * K * 2);        // [K/8][N][8]
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  // stage operands into the canonical no-swizzle K-major layout
  for (int i = tid; i < 128 * K; i += 128) { int m = i / K, k = i % K; sA[((k >> 3) * 128 + m) * 8 + (k & 7)] = A[i]; }
  for (int i = tid; i < N * K; i += 128) { int n = i / K, k = i % K; sB[((k >> 3) * N + n) * 8 + (k & 7)] = B[i]; }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic smem writes -> async proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t tmem_d = tmem;             // accumulator: columns [0, N)
  const uint32_t tmem_a = tmem + 256;       // A operand (TS mode): columns [256, 256 + K/2)

  if (mode == 1) {
    // thread t owns lane (row) t: pack two 16-bit values per 32-bit column and store 8 columns at a time
    const uint32_t lane_addr = tmem_a + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < K / 2; c0 += 8) {
      uint32_t r[8];
      for (int j = 0; j < 8; ++j) {
        const int k = (c0 + j) * 2;
        r[j] = (uint32_t)A[tid * K + k] | ((uint32_t)A[tid * K + k + 1] << 16);
      }
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                   ::"r"(lane_addr + c0), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }

  // instruction descriptor: D=f32, A/B format, K-major both, N>>3, M>>4
  uint32_t idesc = 0;
  idesc |= 1u << 4;                        // c_format = F32
  idesc |= (uint32_t)(fmt & 7) << 7;       // a_format
  idesc |= (uint32_t)(fmt & 7) << 10;      // b_format
  idesc |= (uint32_t)(N >> 3) << 17;
  idesc |= (uint32_t)(128 >> 4) << 24;

  if (tid == 0) {
    for (int kc = 0; kc < K / 16; ++kc) {
      const uint64_t bd = make_desc(smem_u32(sB) + kc * 2 * (N * 16), b_lbo, b_sbo);
      const uint32_t acc = kc > 0 ? 1u : 0u;
      if (mode == 0) {
        const uint64_t ad = make_desc(smem_u32(sA) + kc * 2 * (128 * 16), a_lbo, a_sbo);
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                     ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
      } else {
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                     ::"r"(tmem_d), "r"(tmem_a + kc * 8), "l"(bd), "r"(idesc), "r"(acc) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
  }
  const bool ok = mbar_wait(smem_u32(&mbar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!ok && tid == 0) *status = 1;
  if (ok) {
    const uint32_t lane_addr = tmem_d + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < N; c0 += 8) {
      uint32_t r[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(lane_addr + c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) D[tid * N + c0 + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

static uint16_t f2h(float f, int fmt) {
  if (fmt == 0) { __half h = __float2half(f); return *(uint16_t*)&h; }
  __nv_bfloat16 b = __float2bfloat16(f); return *(uint16_t*)&b;
}
static float h2f(uint16_t u, int fmt) {
  if (fmt == 0) { __half h = *(__half*)&u; return __half2float(h); }
  __nv_bfloat16 b = *(__nv_bfloat16*)&u; return __bfloat162float(b);
}

static double run(int N, int K, int mode, int fmt, bool swap_lbo_sbo) {
  std::vector<uint16_t> A(128 * K), B(N * K);
  srand(1234 + N + K);
  for (auto& v : A) v = f2h((rand() % 2001 - 1000) / 1000.0f, fmt);
  for (auto& v : B) v = f2h((rand() % 2001 - 1000) / 1000.0f, fmt);
  uint16_t *dA, *dB; float* dD; int* dS;
  CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, B.size() * 2)); CK(cudaMalloc(&dD, 128 * N * 4)); CK(cudaMalloc(&dS, 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, 128 * N * 4)); CK(cudaMemset(dS, 0, 4));
  // K-major no-swizzle: LBO = byte stride between the two 16-byte K chunks, SBO = stride between 8-row groups
  uint32_t a_lbo = 128 * 16, a_sbo = 128, b_lbo = N * 16, b_sbo = 128;
  if (swap_lbo_sbo) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
  size_t smem = (size_t)(128 + N) * K * 2 + 1024;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe<<<1, 128, smem>>>(dA, dB, dD, N, K, a_lbo, a_sbo, b_lbo, b_sbo, mode, fmt, dS);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  kernel failed: %s\n", cudaGetErrorString(e)); exit(3); }
  int st = 0; std::vector<float> D(128 * N);
  CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
    double acc = 0;
    for (int k = 0; k < K; ++k) acc += (double)h2f(A[m * K + k], fmt) * h2f(B[n * K + k], fmt);
    double err = fabs(acc - D[m * N + n]);
    if (!(err <= 1e30)) err = 1e30;
    if (err > maxerr) maxerr = err;
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS);
  printf("N=%3d K=%3d mode=%s fmt=%s swap=%d : %s max_abs_err=%.3e\n", N, K, mode ? "TS" : "SS", fmt ? "bf16" : "fp16",
         (int)swap_lbo_sbo, st ? "TIMEOUT" : "done", maxerr);
  return st ? 1e30 : maxerr;
}

int main() {
  int dev = 0; cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
  printf("device %s cc %d.%d SMs %d\n", p.name, p.major, p.minor, p.multiProcessorCount);
  int bad = 0;
  bad += run(128, 64, 0, 0, false) > 1e-2;
  bad += run(128, 64, 1, 0, false) > 1e-2;
  bad += run(128, 176, 0, 0, false) > 1e-2;
  bad += run(128, 176, 1, 1, false) > 1e-1;
  bad += run(64, 112, 1, 0, false) > 1e-2;
  bad += run(256, 64, 1, 0, false) > 1e-2;
  bad += run(16, 64, 1, 0, false) > 1e-2;
  bad += run(80, 128, 1, 0, false) > 1e-2;
  bad += run(32, 16, 1, 1, false) > 1e-1;
  printf(bad ? "PROBE FAILED (%d)\n" : "PROBE OK\n", bad);
  return bad ? 1 : 0;
}
