// a6: value and analytic gradient of the SphereSDF (replaces SDF.autograd_diff, shapes/sdfs.py:184-197, where no
// graph is needed).  The Jacobian d sdf / d p is propagated in forward mode through the same fused MLP tile
// evaluator: every point occupies four tile columns (value and three tangents), tangent columns see the Linear
// layers without bias and the activation's derivative at the value column.
#include <algorithm>

#include "mlp_tile_f32.cuh"

namespace nrt {

// smooth-min of the warped spheres with its gradient (utils.py:385-387, sdfs.py:37-46)
__device__ __forceinline__ void sphere_set_value_grad(const SdfDev& sd, const float p[3], float* value, float grad[3]) {
  float sum = 0.0f, gs[3] = {0.0f, 0.0f, 0.0f};
  for (int i = 0; i < sd.n; ++i) {
    const float* T = sd.tfs + i * 9;
    float A[9], q[3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int k = 0; k < 3; ++k) A[j * 3 + k] = __ldg(T + j * 3 + k) + (j == k ? 1.0f : 0.0f);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float a = A[j * 3] * p[0];
      a = nrt_fma(A[j * 3 + 1], p[1], a);
      a = nrt_fma(A[j * 3 + 2], p[2], a);
      q[j] = a - __ldg(sd.centers + i * 3 + j);
    }
    float n2 = q[0] * q[0];
    n2 = nrt_fma(q[1], q[1], n2);
    n2 = nrt_fma(q[2], q[2], n2);
    const float nq = sqrtf(n2);
    const float e = nrt_expf(-32.0f * (nq - __ldg(sd.radii + i)));
    sum = sum + e;
    if (nq > 0.0f) {
#pragma unroll
      for (int k = 0; k < 3; ++k) gs[k] += e * ((A[k] * q[0] + A[3 + k] * q[1] + A[6 + k] * q[2]) / nq);
    }
  }
  *value = -nrt_logf(fmaxf(sum, 1e-4f)) / 32.0f;
#pragma unroll
  for (int k = 0; k < 3; ++k) grad[k] = (sum >= 1e-4f) ? gs[k] / sum : 0.0f;   // clamp passes gradient where sum >= min
}

template <int H, int TM>
__global__ void __launch_bounds__(kThreads, 1)
k_sdf_value_grad(SdfDev sd, const float* __restrict__ p, int64_t M, float* __restrict__ value, float* __restrict__ grad) {
  extern __shared__ __align__(16) float smem[];
  TileSmem s;
  float* rest = carve_tile(s, smem, sd.mlp.dim_p, H, sd.mlp.out, TM);
  float* sv = rest; rest += TM;            // sphere value / gradient per tile column
  constexpr int PTS = TM / 4;
  const int64_t ntiles = (M + PTS - 1) / PTS;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t pb = tile * PTS;
    const int tid = threadIdx.x;
    if (tid < PTS) {
      float pt[3] = {0.0f, 0.0f, 0.0f};
      if (pb + tid < M) { pt[0] = p[(pb + tid) * 3]; pt[1] = p[(pb + tid) * 3 + 1]; pt[2] = p[(pb + tid) * 3 + 2]; }
      s.enc_raw[0 * TM + 4 * tid] = pt[0]; s.enc_raw[1 * TM + 4 * tid] = pt[1]; s.enc_raw[2 * TM + 4 * tid] = pt[2];
      float v, g[3];
      sphere_set_value_grad(sd, pt, &v, g);
      sv[4 * tid] = v; sv[4 * tid + 1] = g[0]; sv[4 * tid + 2] = g[1]; sv[4 * tid + 3] = g[2];
    }
    __syncthreads();
    mlp_tile_forward<H, TM, true>(sd.mlp, s, nullptr, 0, 0, TM);
    if (tid < PTS && pb + tid < M) {
      value[pb + tid] = sv[4 * tid] + s.outb[4 * tid];
      grad[(pb + tid) * 3 + 0] = sv[4 * tid + 1] + s.outb[4 * tid + 1];
      grad[(pb + tid) * 3 + 1] = sv[4 * tid + 2] + s.outb[4 * tid + 2];
      grad[(pb + tid) * 3 + 2] = sv[4 * tid + 3] + s.outb[4 * tid + 3];
    }
    __syncthreads();
  }
}


// MLP-only variant that also saves the post-activation states of the four-column network for
// nrt_mlp_value_jac_backward (training: SDF.autograd_diff with a graph, sdfs.py:184-197).
template <int H, int TM>
__global__ void __launch_bounds__(kThreads, 1)
k_mlp_value_jac(MlpDev m, const float* __restrict__ p, int64_t M, float* __restrict__ value, float* __restrict__ jac,
                float* __restrict__ acts) {
  extern __shared__ __align__(16) float smem[];
  TileSmem s;
  carve_tile(s, smem, m.dim_p, H, m.out, TM);
  constexpr int PTS = TM / 4;
  const int64_t ntiles = (M + PTS - 1) / PTS;
  const int OUT = m.out;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t pb = tile * PTS;
    const int tid = threadIdx.x;
    const int valid_pts = (int)min((int64_t)PTS, M - pb);
    if (tid < PTS) {
      float pt[3] = {0.0f, 0.0f, 0.0f};
      if (tid < valid_pts) { pt[0] = p[(pb + tid) * 3]; pt[1] = p[(pb + tid) * 3 + 1]; pt[2] = p[(pb + tid) * 3 + 2]; }
      s.enc_raw[0 * TM + 4 * tid] = pt[0]; s.enc_raw[1 * TM + 4 * tid] = pt[1]; s.enc_raw[2 * TM + 4 * tid] = pt[2];
    }
    __syncthreads();
    mlp_tile_forward<H, TM, true>(m, s, acts, M * 4, pb * 4, valid_pts * 4);
    for (int idx = tid; idx < OUT * TM; idx += kThreads) {
      const int n = idx / TM, mm = idx - n * TM;
      const int q = mm >> 2, c = mm & 3;
      if (q < valid_pts) {
        if (c == 0) value[(pb + q) * OUT + n] = s.outb[idx];
        else jac[((pb + q) * OUT + n) * 3 + c - 1] = s.outb[idx];
      }
    }
    __syncthreads();
  }
}

}  // namespace nrt
using namespace nrt;

extern "C" int nrt_sdf_value_grad(const nrt_sphere_sdf_t* s, const float* p, int64_t M, float* value, float* grad,
                                  void* stream) {
  SdfDev d;
  int rc = nrt_build_sdf_dev(s, &d);
  if (rc != NRT_OK) return rc;
  NRT_REQUIRE(M >= 0, "nrt_sdf_value_grad: negative M");
  if (M == 0) return NRT_OK;
  NRT_REQUIRE(p && value && grad, "nrt_sdf_value_grad: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
#define NRT_VG_CASE(HV, TMV)                                                                             \
  if (d.mlp.hidden == HV) {                                                                              \
    const size_t bytes = (tile_smem_floats(d.mlp.dim_p, HV, d.mlp.out, TMV) + TMV) * sizeof(float);      \
    NRT_REQUIRE(bytes <= 227 * 1024, "shared memory");                                                   \
    NRT_CUDA(cudaFuncSetAttribute(k_sdf_value_grad<HV, TMV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes)); \
    const int64_t ntiles = (M + TMV / 4 - 1) / (TMV / 4);                                                \
    const int grid = (int)std::min<int64_t>(ntiles, (int64_t)nrt_sm_count() * 4);                        \
    NrtProfScope _ps(TAG_SDF_GRAD_F32, st);                                                              \
    k_sdf_value_grad<HV, TMV><<<grid, kThreads, bytes, st>>>(d, p, M, value, grad);                      \
    NRT_CUDA(cudaGetLastError());                                                                        \
    return NRT_OK;                                                                                       \
  }
  NRT_VG_CASE(128, 64)
  NRT_VG_CASE(64, 64)
  NRT_VG_CASE(32, 64)
#undef NRT_VG_CASE
  nrt_set_error("nrt_sdf_value_grad: unsupported hidden size %d", d.mlp.hidden);
  return NRT_E_UNSUPPORTED;
}

extern "C" int nrt_mlp_value_jac_forward(const nrt_mlp_t* mm, const float* p, int64_t M, float* value, float* jac,
                                         float* acts, void* stream) {
  MlpDev d;
  int rc = nrt_build_mlp_dev(mm, &d);
  if (rc != NRT_OK) return rc;
  NRT_REQUIRE(M >= 0, "nrt_mlp_value_jac_forward: negative M");
  NRT_REQUIRE(d.in_size == 3 && d.latent == 0, "nrt_mlp_value_jac_forward: needs in_size 3 and no latent");
  if (M == 0) return NRT_OK;
  NRT_REQUIRE(p && value && jac, "nrt_mlp_value_jac_forward: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
#define NRT_VJ_CASE(HV, TMV)                                                                             \
  if (d.hidden == HV) {                                                                                  \
    const size_t bytes = tile_smem_floats(d.dim_p, HV, d.out, TMV) * sizeof(float);                      \
    NRT_REQUIRE(bytes <= 227 * 1024, "shared memory");                                                   \
    NRT_CUDA(cudaFuncSetAttribute(k_mlp_value_jac<HV, TMV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes)); \
    const int64_t ntiles = (M + TMV / 4 - 1) / (TMV / 4);                                                \
    const int grid = (int)std::min<int64_t>(ntiles, (int64_t)nrt_sm_count() * 4);                        \
    NrtProfScope _ps(TAG_SDF_GRAD_F32, st);                                                              \
    k_mlp_value_jac<HV, TMV><<<grid, kThreads, bytes, st>>>(d, p, M, value, jac, acts);                  \
    NRT_CUDA(cudaGetLastError());                                                                        \
    return NRT_OK;                                                                                       \
  }
  NRT_VJ_CASE(128, 64)
  NRT_VJ_CASE(64, 64)
  NRT_VJ_CASE(32, 64)
#undef NRT_VJ_CASE
  nrt_set_error("nrt_mlp_value_jac_forward: unsupported hidden size %d", d.hidden);
  return NRT_E_UNSUPPORTED;
}
