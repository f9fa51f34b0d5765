// Pixel -> ray arithmetic shared by k_camera_rays (nrt_camera.cu) and the camera-fed IO policies of the NeRF kernels
// (nrt_tc.cu).  Every operation is an explicit round-to-nearest intrinsic, so the two translation units (one built
// with -fmad=false, one with contraction on) produce bit-identical rays, in the reference's operation order:
//   NeRFCamera.sample_positions             pytorch3d/pathtracer/cameras/cameras.py:23-54
//   DTUCamera.sample_positions, lift        pytorch3d/pathtracer/cameras/cameras.py:132-147, 156-192
//   FoVPerspectiveCameras.sample_positions  pytorch3d/renderer/cameras.py:539-575
#pragma once
#include "nrt_common.cuh"

namespace nrtcam {

// same counter hash as the stratified sample distances (nrt_render.cu; oracle/port.py restates it)
__device__ __forceinline__ float hash_u01(uint64_t seed, uint32_t a, uint32_t b) {
  uint32_t h = (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0x9E3779B1u);
  h = (h ^ a) * 0x85EBCA77u;
  h = (h ^ b) * 0xC2B2AE3Du;
  h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
  return (float)(h >> 8) * (1.0f / 16777216.0f);
}

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float dvd(float a, float b) { return __fdiv_rn(a, b); }

__device__ __forceinline__ void normalize(float x, float y, float z, float* d) {
  // F.normalize: v / max(||v||, 1e-12)
  const float n = fmaxf(__fsqrt_rn(add(add(mul(x, x), mul(y, y)), mul(z, z))), 1e-12f);
  d[0] = dvd(x, n); d[1] = dvd(y, n); d[2] = dvd(z, n);
}

struct CamDev {      // nrt_camera_t by value (device pointers stay device pointers)
  int kind, n_views;
  const float* a; const float* b;
  int avs, ars, bvs, brs;
  float focal, size;
  int x0, y0, nx, ny, bundle, ppp;
  const float* positions;
  float jitter; uint64_t seed;
};

static inline CamDev make_cam_dev(const nrt_camera_t* cam) {
  return CamDev{cam->kind, cam->n_views, cam->a, cam->b, cam->a_view_stride, cam->a_row_stride, cam->b_view_stride,
                cam->b_row_stride, cam->focal, cam->size, cam->x0, cam->y0, cam->nx, cam->ny, cam->bundle,
                cam->pos_per_pixel, cam->positions, cam->jitter, cam->jitter_seed};
}

// ray r of the [n_views, nx, ny, bundle] block -> origin o[3], unit direction d[3], view index
__device__ __forceinline__ void cam_ray(const CamDev& c, int64_t r, float* o, float* d, int* view_out) {
  // 32-bit index arithmetic when the block allows it (every frame size in use): the 64-bit divisions cost ~50 instructions each
  int b, i, j, view;
  if (r < 0x7fffffffLL) {
    const uint32_t r32 = (uint32_t)r;
    const uint32_t p = c.bundle == 1 ? r32 : r32 / (uint32_t)c.bundle;
    b = c.bundle == 1 ? 0 : (int)(r32 - p * (uint32_t)c.bundle);
    const uint32_t q = p / (uint32_t)c.ny;
    j = (int)(p - q * (uint32_t)c.ny);
    view = c.n_views == 1 ? 0 : (int)(q / (uint32_t)c.nx);
    i = (int)(q - (uint32_t)view * (uint32_t)c.nx);
  } else {
    b = (int)(r % c.bundle);
    const int64_t p = r / c.bundle;
    j = (int)(p % c.ny);
    const int64_t q = p / c.ny;
    i = (int)(q % c.nx);
    view = (int)(q / c.nx);
  }
  *view_out = view;
  float u, v;
  if (c.positions) {
    const int64_t k = ((int64_t)i * c.ny + j) * c.ppp + (c.ppp > 1 ? b : 0);
    u = __ldg(c.positions + 2 * k); v = __ldg(c.positions + 2 * k + 1);
  } else {
    u = (float)(c.y0 + j); v = (float)(c.x0 + i);          // main.py:74: positions = stack([grid_y, grid_x])
    if (c.jitter > 0.0f) {
      const uint32_t pix = (uint32_t)(r & 0xffffffffu);
      u = add(u, mul(sub(hash_u01(c.seed, pix, 0x75u), 0.5f), c.jitter));
      v = add(v, mul(sub(hash_u01(c.seed, pix, 0x76u), 0.5f), c.jitter));
    }
  }
  const float* A = c.a + (int64_t)view * c.avs;
  if (c.kind == NRT_CAM_NERF) {
    // cameras.py:39-53
    const float half = mul(c.size, 0.5f);
    const float dx = dvd(sub(u, half), c.focal);
    const float dy = dvd(-sub(v, half), c.focal);
    float w[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float* row = A + k * c.ars;
      w[k] = add(add(mul(dx, __ldg(row)), mul(dy, __ldg(row + 1))), -__ldg(row + 2));     // dz = -1
      o[k] = __ldg(row + 3);
    }
    normalize(w[0], w[1], w[2], d);
  } else if (c.kind == NRT_CAM_DTU) {
    // cameras.py:171-192 with lift (:132-147) at z = 1; the 1600 x 1200 normalisation is the reference's (:177)
    const float* K = c.b + (int64_t)view * c.bvs;
    const float x = mul(u, dvd(1600.0f, c.size)), y = mul(v, dvd(1200.0f, c.size));
    const float fx = __ldg(K), sk = __ldg(K + 1), cx = __ldg(K + 2);
    const float fy = __ldg(K + c.brs + 1), cy = __ldg(K + c.brs + 2);
    const float xl = dvd(sub(add(sub(x, cx), dvd(mul(cy, sk), fy)), dvd(mul(sk, y), fy)), fx);
    const float yl = dvd(sub(y, cy), fy);
    float w[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float* row = A + k * c.ars;
      o[k] = __ldg(row + 3);
      w[k] = sub(add(add(add(mul(__ldg(row), xl), mul(__ldg(row + 1), yl)), __ldg(row + 2)), o[k]), o[k]);
    }
    normalize(w[0], w[1], w[2], d);
  } else {
    // renderer/cameras.py:557-575: NDC point (1 - 2 p / size, z = 1) through the inverse full projection (row vectors);
    // the direction is the normalised unprojected point itself, as in the reference
    const float p0 = add(mul(-2.0f, dvd(u, c.size)), 1.0f), p1 = add(mul(-2.0f, dvd(v, c.size)), 1.0f);
    float h[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      h[k] = add(add(add(mul(p0, __ldg(A + k)), mul(p1, __ldg(A + c.ars + k))), __ldg(A + 2 * c.ars + k)), __ldg(A + 3 * c.ars + k));
    normalize(dvd(h[0], h[3]), dvd(h[1], h[3]), dvd(h[2], h[3]), d);
    const float* C = c.b + (int64_t)view * c.bvs;
    o[0] = __ldg(C); o[1] = __ldg(C + 1); o[2] = __ldg(C + 2);
  }
}

}  // namespace nrtcam
