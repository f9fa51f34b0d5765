// The sphere set of SphereSDF under autograd (shapes/sdfs.py:37-45, utils.py:385-387):
//     q_i = (I + tfs_i) p - c_i,   d_i = |q_i| - r_i,   s = -log(max(sum_i exp(-k d_i), 1e-4)) / k,   k = 32,
// its gradient n = ds/dp = sum_i w_i T_i^T q_i / |q_i| (w = softmax(-k d); zero where the clamp is active), which the
// reference obtains with autograd.grad(create_graph=True) (SDF.autograd_diff, sdfs.py:184-197), and the reverse pass of
// BOTH outputs into centers / radii / tfs, i.e. the double backward that loss.backward() runs through that graph
// (eikonal_loss, utils.py:294; the shading normals; the 5-epsilon point offset, sdfs.py:196):
//     with a_i = g_n . T_i^T qh_i,  abar = sum_j w_j a_j,  G_i = g_s w_i - k w_i (a_i - abar),
//          e_i = (T_i g_n - (qh_i . T_i g_n) qh_i) / |q_i|,   Q_i = G_i qh_i + w_i e_i:
//     g_c_i = -sum Q_i,   g_r_i = -sum G_i,   g_tfs_i = sum (Q_i p^T + w_i qh_i g_n^T).
// The points carry no gradient (leaves made from a no_grad march).  In the eager mirror this was 17 torch launches per
// direction with [n, K, 3] intermediates (9 bmm + 8 sgemm kernels: 30 of 81 ms of a 262,144-ray DTU step).
// Thread = point, loop over the spheres (parameters in shared memory); the parameter gradients are reduced per warp with
// shuffles, per block in shared memory, and leave the block as one atomicAdd per parameter.
#include "nrt_common.cuh"

namespace nrt {

constexpr float kSmoothK = 32.0f;
constexpr int kSphThreads = 128;

struct SphereGeom { float q[3], len, qh[3], d; };

__device__ __forceinline__ SphereGeom sphere_geom(const float* __restrict__ sp, const float* p) {
  // sp: [c(3) | r | T (9, row major, identity added)]
  SphereGeom g;
#pragma unroll
  for (int a = 0; a < 3; ++a) g.q[a] = sp[4 + 3 * a] * p[0] + sp[4 + 3 * a + 1] * p[1] + sp[4 + 3 * a + 2] * p[2] - sp[a];
  g.len = sqrtf(g.q[0] * g.q[0] + g.q[1] * g.q[1] + g.q[2] * g.q[2]);
  const float inv = g.len > 0.0f ? 1.0f / g.len : 0.0f;     // torch: the subgradient of norm at 0 is 0
#pragma unroll
  for (int a = 0; a < 3; ++a) g.qh[a] = g.q[a] * inv;
  g.d = g.len - sp[3];
  return g;
}

__device__ __forceinline__ void load_spheres(float* sS, int n, const float* centers, const float* radii, const float* tfs) {
  for (int i = threadIdx.x; i < n * 13; i += blockDim.x) {
    const int s = i / 13, k = i - s * 13;
    float v;
    if (k < 3) v = centers[s * 3 + k];
    else if (k == 3) v = radii[s];
    else { const int e = k - 4; v = tfs[s * 9 + e] + ((e == 0 || e == 4 || e == 8) ? 1.0f : 0.0f); }
    sS[i] = v;
  }
}

__global__ void __launch_bounds__(kSphThreads)
k_sphere_set_fwd(int n, const float* __restrict__ centers, const float* __restrict__ radii, const float* __restrict__ tfs,
                 const float* __restrict__ p, int64_t K, float* __restrict__ value, float* __restrict__ grad) {
  extern __shared__ float sS[];
  load_spheres(sS, n, centers, radii, tfs);
  __syncthreads();
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < K; m += (int64_t)gridDim.x * blockDim.x) {
    const float pt[3] = {p[m * 3], p[m * 3 + 1], p[m * 3 + 2]};
    float S = 0.0f, nn[3] = {0.0f, 0.0f, 0.0f};
    for (int i = 0; i < n; ++i) {
      const float* sp = sS + i * 13;
      const SphereGeom g = sphere_geom(sp, pt);
      const float e = expf(-kSmoothK * g.d);
      S += e;
#pragma unroll
      for (int b = 0; b < 3; ++b) nn[b] += e * (sp[4 + b] * g.qh[0] + sp[7 + b] * g.qh[1] + sp[10 + b] * g.qh[2]);   // T^T qh
    }
    const bool clamped = !(S >= 1e-4f);
    value[m] = -logf(clamped ? 1e-4f : S) / kSmoothK;
    if (grad != nullptr) {
      const float inv = clamped ? 0.0f : 1.0f / S;
      grad[m * 3] = nn[0] * inv; grad[m * 3 + 1] = nn[1] * inv; grad[m * 3 + 2] = nn[2] * inv;
    }
  }
}

__global__ void __launch_bounds__(kSphThreads)
k_sphere_set_bwd(int n, const float* __restrict__ centers, const float* __restrict__ radii, const float* __restrict__ tfs,
                 const float* __restrict__ p, int64_t K, const float* __restrict__ g_value, const float* __restrict__ g_grad,
                 float* __restrict__ g_centers, float* __restrict__ g_radii, float* __restrict__ g_tfs) {
  extern __shared__ float sm[];
  float* sS = sm;                 // [n][13] parameters
  float* sG = sm + n * 13;        // [n][13] gradient accumulators: [g_c(3) | g_r | g_T(9)]
  load_spheres(sS, n, centers, radii, tfs);
  for (int i = threadIdx.x; i < n * 13; i += blockDim.x) sG[i] = 0.0f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  // every warp iterates the same number of times (whole-warp shuffles below)
  const int64_t per_iter = (int64_t)gridDim.x * blockDim.x;
  for (int64_t m0 = (int64_t)blockIdx.x * blockDim.x; m0 < K; m0 += per_iter) {
    const int64_t m = m0 + threadIdx.x;
    const bool valid = m < K;
    float pt[3] = {0.0f, 0.0f, 0.0f}, gs = 0.0f, gn[3] = {0.0f, 0.0f, 0.0f};
    if (valid) {
      pt[0] = p[m * 3]; pt[1] = p[m * 3 + 1]; pt[2] = p[m * 3 + 2];
      if (g_value != nullptr) gs = g_value[m];
      if (g_grad != nullptr) { gn[0] = g_grad[m * 3]; gn[1] = g_grad[m * 3 + 1]; gn[2] = g_grad[m * 3 + 2]; }
    }
    // pass 1: S = sum e_i, abar = sum e_i a_i / S
    float S = 0.0f, ea = 0.0f;
    for (int i = 0; i < n; ++i) {
      const float* sp = sS + i * 13;
      const SphereGeom g = sphere_geom(sp, pt);
      const float e = expf(-kSmoothK * g.d);
      float a = 0.0f;
#pragma unroll
      for (int c = 0; c < 3; ++c) a += (sp[4 + 3 * c] * gn[0] + sp[4 + 3 * c + 1] * gn[1] + sp[4 + 3 * c + 2] * gn[2]) * g.qh[c];   // (T g_n) . qh
      S += e; ea += e * a;
    }
    const bool on = valid && (S >= 1e-4f);
    const float invS = on ? 1.0f / S : 0.0f;
    const float abar = ea * invS;
    // pass 2: contributions of this point to every sphere, reduced over the warp
    for (int i = 0; i < n; ++i) {
      const float* sp = sS + i * 13;
      const SphereGeom g = sphere_geom(sp, pt);
      const float w = expf(-kSmoothK * g.d) * invS;
      float h[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) h[c] = sp[4 + 3 * c] * gn[0] + sp[4 + 3 * c + 1] * gn[1] + sp[4 + 3 * c + 2] * gn[2];
      const float a = h[0] * g.qh[0] + h[1] * g.qh[1] + h[2] * g.qh[2];
      const float G = gs * w - kSmoothK * w * (a - abar);
      const float invlen = g.len > 0.0f ? 1.0f / g.len : 0.0f;
      float Q[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) Q[c] = G * g.qh[c] + w * (h[c] - a * g.qh[c]) * invlen;
      float v[13];
      v[0] = -Q[0]; v[1] = -Q[1]; v[2] = -Q[2]; v[3] = -G;
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int b = 0; b < 3; ++b) v[4 + 3 * c + b] = Q[c] * pt[b] + w * g.qh[c] * gn[b];
#pragma unroll
      for (int k = 0; k < 13; ++k) {
        float x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) atomicAdd(sG + i * 13 + k, x);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n * 13; i += blockDim.x) {
    const int s = i / 13, k = i - s * 13;
    const float x = sG[i];
    if (x == 0.0f) continue;
    if (k < 3) atomicAdd(g_centers + s * 3 + k, x);
    else if (k == 3) atomicAdd(g_radii + s, x);
    else atomicAdd(g_tfs + s * 9 + (k - 4), x);
  }
}

}  // namespace nrt

using namespace nrt;

extern "C" int nrt_sphere_set_forward(int n, const float* centers, const float* radii, const float* tfs, const float* p, int64_t K,
                                      float* value, float* grad, void* stream) {
  NRT_REQUIRE(n >= 1 && n <= 1024, "nrt_sphere_set_forward: 1..1024 spheres");
  NRT_REQUIRE(K >= 0, "negative K");
  if (K == 0) return NRT_OK;
  NRT_REQUIRE(centers && radii && tfs && p && value, "nrt_sphere_set_forward: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = (int)std::min<int64_t>((K + kSphThreads - 1) / kSphThreads, (int64_t)nrt_sm_count() * 8);
  NrtProfScope _ps(TAG_SHADE, st);
  k_sphere_set_fwd<<<grid, kSphThreads, (size_t)n * 13 * sizeof(float), st>>>(n, centers, radii, tfs, p, K, value, grad);
  NRT_CUDA(cudaGetLastError());
  return NRT_OK;
}

extern "C" int nrt_sphere_set_backward(int n, const float* centers, const float* radii, const float* tfs, const float* p, int64_t K,
                                       const float* g_value, const float* g_grad, float* g_centers, float* g_radii, float* g_tfs,
                                       void* stream) {
  NRT_REQUIRE(n >= 1 && n <= 1024, "nrt_sphere_set_backward: 1..1024 spheres");
  NRT_REQUIRE(K >= 0, "negative K");
  if (K == 0) return NRT_OK;
  NRT_REQUIRE(centers && radii && tfs && p && g_centers && g_radii && g_tfs, "nrt_sphere_set_backward: null pointer");
  NRT_REQUIRE(g_value || g_grad, "nrt_sphere_set_backward: no incoming gradient");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = (int)std::min<int64_t>((K + kSphThreads - 1) / kSphThreads, (int64_t)nrt_sm_count() * 4);
  NrtProfScope _ps(TAG_SHADE, st);
  k_sphere_set_bwd<<<grid, kSphThreads, (size_t)n * 26 * sizeof(float), st>>>(n, centers, radii, tfs, p, K, g_value, g_grad, g_centers,
                                                                            g_radii, g_tfs);
  NRT_CUDA(cudaGetLastError());
  return NRT_OK;
}
