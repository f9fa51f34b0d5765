// a8 / a12: shading glue as elementwise kernels (one thread per ray, coalesced).  Arithmetic follows the reference's
// op order with the deterministic transcendentals of nrt_detmath.h.
//   coordinate_system / to_local : pytorch3d/pathtracer/interaction.py:9-27, 38-41
//   param_rusin2                 : pytorch3d/pathtracer/utils.py:233-258 (+ rotate_vector :152, nonzero_eps :43)
#include "nrt_common.cuh"

namespace {

__device__ __forceinline__ void normalize3(float v[3], float eps) {   // F.normalize: v / max(|v|, eps)
  const float n = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  const float d = fmaxf(n, eps);
  v[0] = v[0] / d; v[1] = v[1] / d; v[2] = v[2] / d;
}
__device__ __forceinline__ void cross3(const float a[3], const float b[3], float o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
// frame columns (s, t, n); frame[i*3 + c] = component i of column c  (torch.stack([s,t,n], dim=-1))
__device__ __forceinline__ void coordinate_system(const float nin[3], float fr[9]) {
  float n[3] = {nin[0], nin[1], nin[2]};
  normalize3(n, 1e-7f);
  const float sign = n[2] >= 0.0f ? 1.0f : -1.0f;
  const float sz = sign + n[2];
  const float a = -(1.0f / (fabsf(sz) < 1e-6f ? 1e-6f : sz));
  const float b = n[0] * n[1] * a;
  float s[3] = {n[0] * n[0] * a * sign + 1.0f, b * sign, n[0] * -sign};
  normalize3(s, 1e-7f);
  float t[3];
  cross3(s, n, t);
  normalize3(t, 1e-7f);
  cross3(n, t, s);
  normalize3(s, 1e-7f);
#pragma unroll
  for (int i = 0; i < 3; ++i) { fr[i * 3 + 0] = s[i]; fr[i * 3 + 1] = t[i]; fr[i * 3 + 2] = n[i]; }
}
// to_local: normalize(mean over xyz of frame * wo)
__device__ __forceinline__ void to_local(const float fr[9], const float w[3], float o[3]) {
#pragma unroll
  for (int c = 0; c < 3; ++c) o[c] = (fr[0 * 3 + c] * w[0] + fr[1 * 3 + c] * w[1] + fr[2 * 3 + c] * w[2]) / 3.0f;
  normalize3(o, 1e-7f);
}
__device__ __forceinline__ float nonzero_eps(float v) { return fabsf(v) < 1e-7f ? 1e-7f : v; }
// Rodrigues rotation with given cosine / sine
__device__ __forceinline__ void rotate_vector(const float v[3], const float axis[3], float c, float s, float o[3]) {
  const float d = v[0] * axis[0] + v[1] * axis[1] + v[2] * axis[2];
  float cr[3];
  cross3(axis, v, cr);
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i] = v[i] * c + axis[i] * d * (1.0f - c) + cr[i] * s;
}

__global__ void k_shading_frame(const float* __restrict__ normals, const float* __restrict__ rays, int64_t R,
                                float* __restrict__ frame, float* __restrict__ wi) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float n[3] = {normals[r * 3], normals[r * 3 + 1], normals[r * 3 + 2]};
  float fr[9];
  coordinate_system(n, fr);
#pragma unroll
  for (int i = 0; i < 9; ++i) frame[r * 9 + i] = fr[i];
  if (wi != nullptr) {
    const float d[3] = {-rays[r * 6 + 3], -rays[r * 6 + 4], -rays[r * 6 + 5]};   // wi = to_local(-r_d), sdfs.py:159
    float o[3];
    to_local(fr, d, o);
    wi[r * 3] = o[0]; wi[r * 3 + 1] = o[1]; wi[r * 3 + 2] = o[2];
  }
}

__global__ void k_to_local(const float* __restrict__ frame, const float* __restrict__ v, int64_t R, float* __restrict__ out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  float fr[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) fr[i] = frame[r * 9 + i];
  const float w[3] = {v[r * 3], v[r * 3 + 1], v[r * 3 + 2]};
  float o[3];
  to_local(fr, w, o);
  out[r * 3] = o[0]; out[r * 3 + 1] = o[1]; out[r * 3 + 2] = o[2];
}

__global__ void k_param_rusin2(const float* __restrict__ a, const float* __restrict__ b, int64_t R, float* __restrict__ out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  float wo[3] = {a[r * 3], a[r * 3 + 1], a[r * 3 + 2]};
  float wi[3] = {b[r * 3], b[r * 3 + 1], b[r * 3 + 2]};
  normalize3(wo, 1e-12f);
  normalize3(wi, 1e-12f);
  float H[3] = {wo[0] + wi[0], wo[1] + wi[1], wo[2] + wi[2]};
  normalize3(H, 1e-12f);
  const float e1[3] = {0.0f, 1.0f, 0.0f}, e2[3] = {0.0f, 0.0f, 1.0f};
  const float hy = nonzero_eps(H[1]), hx = nonzero_eps(H[0]);
  const float rr = fmaxf(sqrtf(hy * hy + hx * hx), 1e-6f);                 // hypot(...).clamp(min=1e-6)
  float tmp[3], diff[3];
  rotate_vector(wi, e2, H[0] / rr, -(H[1] / rr), tmp);
  normalize3(tmp, 1e-12f);
  const float s = -sqrtf(fmaxf(1.0f - H[2], 1e-6f));                      // quirk: 1 - H_z, not 1 - H_z^2
  rotate_vector(tmp, e1, H[2], s, diff);
  normalize3(diff, 1e-12f);
  out[r * 3 + 0] = nrt_cosf(nrt_atan2f(nonzero_eps(diff[1]), nonzero_eps(diff[0])));
  out[r * 3 + 1] = H[2];
  out[r * 3 + 2] = diff[2];
}

}  // namespace

extern "C" int nrt_shading_frame(const float* normals, const float* rays, int64_t R, float* frame, float* wi, void* stream) {
  NRT_REQUIRE(R >= 0, "nrt_shading_frame: negative R");
  if (R == 0) return NRT_OK;
  NRT_REQUIRE(normals && frame && (wi == nullptr || rays != nullptr), "nrt_shading_frame: null pointer");
  NrtProfScope _ps(TAG_SHADE, (cudaStream_t)stream);
  k_shading_frame<<<nrt_cdiv(R, 256), 256, 0, (cudaStream_t)stream>>>(normals, rays, R, frame, wi);
  NRT_CUDA(cudaGetLastError());
  return NRT_OK;
}
extern "C" int nrt_to_local(const float* frame, const float* v, int64_t R, float* out, void* stream) {
  NRT_REQUIRE(R >= 0, "nrt_to_local: negative R");
  if (R == 0) return NRT_OK;
  NRT_REQUIRE(frame && v && out, "nrt_to_local: null pointer");
  NrtProfScope _ps(TAG_SHADE, (cudaStream_t)stream);
  k_to_local<<<nrt_cdiv(R, 256), 256, 0, (cudaStream_t)stream>>>(frame, v, R, out);
  NRT_CUDA(cudaGetLastError());
  return NRT_OK;
}
extern "C" int nrt_param_rusin2(const float* a, const float* b, int64_t R, float* out, void* stream) {
  NRT_REQUIRE(R >= 0, "nrt_param_rusin2: negative R");
  if (R == 0) return NRT_OK;
  NRT_REQUIRE(a && b && out, "nrt_param_rusin2: null pointer");
  NrtProfScope _ps(TAG_SHADE, (cudaStream_t)stream);
  k_param_rusin2<<<nrt_cdiv(R, 256), 256, 0, (cudaStream_t)stream>>>(a, b, R, out);
  NRT_CUDA(cudaGetLastError());
  return NRT_OK;
}
