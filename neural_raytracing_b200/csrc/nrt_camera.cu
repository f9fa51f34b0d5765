// Ray generators on the device: pixel position -> [origin | unit direction] for the three cameras
// the scripts use, written straight into the reference's [n_views, nx, ny, bundle, 6] block.  One thread per ray,
// 24 contiguous bytes per thread (HBM-bound: 24 B per ray written, nothing read but the camera matrices, which every
// thread reads from the same few cache lines).  Built with -fmad=false and written in the reference's operation
// order, so a ray differs from the torch expression by rounding of the 3- and 4-term sums only.
//   NeRFCamera.sample_positions             pytorch3d/pathtracer/cameras/cameras.py:23-54
//   DTUCamera.sample_positions, lift        pytorch3d/pathtracer/cameras/cameras.py:132-147, 156-192
//   FoVPerspectiveCameras.sample_positions  pytorch3d/renderer/cameras.py:539-575
#include "nrt_common.cuh"

// same counter hash as the stratified sample distances (nrt_render.cu; oracle/port.py restates it)
__device__ __forceinline__ float cam_hash_u01(uint64_t seed, uint32_t a, uint32_t b) {
  uint32_t h = (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0x9E3779B1u);
  h = (h ^ a) * 0x85EBCA77u;
  h = (h ^ b) * 0xC2B2AE3Du;
  h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
  return (float)(h >> 8) * (1.0f / 16777216.0f);
}

__device__ __forceinline__ void cam_normalize(float x, float y, float z, float* d) {
  // F.normalize: v / max(||v||, 1e-12)
  const float n = fmaxf(sqrtf(x * x + y * y + z * z), 1e-12f);
  d[0] = x / n; d[1] = y / n; d[2] = z / n;
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

__device__ __forceinline__ void cam_ray(const CamDev& c, int64_t r, float* o, float* d, int* view_out) {
  const int b = (int)(r % c.bundle);
  const int64_t p = r / c.bundle;
  const int j = (int)(p % c.ny);
  const int64_t q = p / c.ny;
  const int i = (int)(q % c.nx);
  const int view = (int)(q / c.nx);
  *view_out = view;
  float u, v;
  if (c.positions) {
    const int64_t k = ((int64_t)i * c.ny + j) * c.ppp + (c.ppp > 1 ? b : 0);
    u = __ldg(c.positions + 2 * k); v = __ldg(c.positions + 2 * k + 1);
  } else {
    u = (float)(c.y0 + j); v = (float)(c.x0 + i);          // main.py:74: positions = stack([grid_y, grid_x])
    if (c.jitter > 0.0f) {
      const uint32_t pix = (uint32_t)(r & 0xffffffffu);
      u = u + (cam_hash_u01(c.seed, pix, 0x75u) - 0.5f) * c.jitter;
      v = v + (cam_hash_u01(c.seed, pix, 0x76u) - 0.5f) * c.jitter;
    }
  }
  const float* A = c.a + (int64_t)view * c.avs;
  if (c.kind == NRT_CAM_NERF) {
    // cameras.py:39-53
    const float half = c.size * 0.5f;
    const float dx = (u - half) / c.focal;
    const float dy = -(v - half) / c.focal;
    const float dz = -1.0f;
    float w[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float* row = A + k * c.ars;
      w[k] = dx * __ldg(row) + dy * __ldg(row + 1) + dz * __ldg(row + 2);
      o[k] = __ldg(row + 3);
    }
    cam_normalize(w[0], w[1], w[2], d);
  } else if (c.kind == NRT_CAM_DTU) {
    // cameras.py:171-192 with lift (:132-147) at z = 1; the 1600 x 1200 normalisation is the reference's (:177)
    const float* K = c.b + (int64_t)view * c.bvs;
    const float x = u * (1600.0f / c.size), y = v * (1200.0f / c.size);
    const float fx = __ldg(K), sk = __ldg(K + 1), cx = __ldg(K + 2);
    const float fy = __ldg(K + c.brs + 1), cy = __ldg(K + c.brs + 2);
    const float xl = (x - cx + cy * sk / fy - sk * y / fy) / fx;
    const float yl = (y - cy) / fy;
    float w[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float* row = A + k * c.ars;
      o[k] = __ldg(row + 3);
      w[k] = (__ldg(row) * xl + __ldg(row + 1) * yl + __ldg(row + 2) + o[k]) - o[k];
    }
    cam_normalize(w[0], w[1], w[2], d);
  } else {
    // renderer/cameras.py:557-575: NDC point (1 - 2 p / size, z = 1) through the inverse full projection (row vectors);
    // the direction is the normalised unprojected point itself, as in the reference
    const float p0 = -2.0f * (u / c.size) + 1.0f, p1 = -2.0f * (v / c.size) + 1.0f;
    float h[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      h[k] = p0 * __ldg(A + k) + p1 * __ldg(A + c.ars + k) + __ldg(A + 2 * c.ars + k) + __ldg(A + 3 * c.ars + k);
    cam_normalize(h[0] / h[3], h[1] / h[3], h[2] / h[3], d);
    const float* C = c.b + (int64_t)view * c.bvs;
    o[0] = __ldg(C); o[1] = __ldg(C + 1); o[2] = __ldg(C + 2);
  }
}

__global__ void __launch_bounds__(256) k_camera_rays(CamDev c, int64_t r0, int64_t n, float* __restrict__ out,
                                                     int32_t* __restrict__ out_view) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  float o[3], d[3];
  int view;
  cam_ray(c, r0 + t, o, d, &view);
  float2* dst = reinterpret_cast<float2*>(out + t * 6);      // 24-byte records, 8-byte aligned
  dst[0] = make_float2(o[0], o[1]);
  dst[1] = make_float2(o[2], d[0]);
  dst[2] = make_float2(d[1], d[2]);
  if (out_view) out_view[t] = view;
}

int nrt_check_camera(const nrt_camera_t* cam, int64_t* total) {
  NRT_REQUIRE(cam != nullptr, "camera descriptor is NULL");
  NRT_REQUIRE(cam->kind == NRT_CAM_NERF || cam->kind == NRT_CAM_DTU || cam->kind == NRT_CAM_FOV,
              "unknown camera kind %d", cam->kind);
  NRT_REQUIRE(cam->n_views >= 1 && cam->nx >= 0 && cam->ny >= 0 && cam->bundle >= 1,
              "camera: bad block %d x %d x %d x %d", cam->n_views, cam->nx, cam->ny, cam->bundle);
  NRT_REQUIRE(cam->a != nullptr, "camera: matrix pointer a is NULL");
  NRT_REQUIRE(cam->kind == NRT_CAM_NERF || cam->b != nullptr, "camera: matrix pointer b is NULL (intrinsics / centres)");
  NRT_REQUIRE(cam->size > 0.0f, "camera: size must be positive");
  NRT_REQUIRE(cam->kind != NRT_CAM_NERF || cam->focal != 0.0f, "NeRF camera: focal is zero");
  NRT_REQUIRE(cam->kind != NRT_CAM_NERF || cam->bundle == 1, "NeRF camera: one ray per pixel (cameras.py:50)");
  NRT_REQUIRE(cam->positions == nullptr || cam->pos_per_pixel == 1 || cam->pos_per_pixel == cam->bundle,
              "camera: pos_per_pixel must be 1 or bundle");
  // strides must cover the elements the kernel reads: 3 rows x 4 columns of a (4 x 4 for the FoV inverse), the first two
  // rows / three columns of the DTU intrinsics, three floats of a FoV centre
  const int a_rows = cam->kind == NRT_CAM_FOV ? 4 : 3;
  NRT_REQUIRE(cam->a_row_stride >= 4 && cam->a_view_stride >= (a_rows - 1) * cam->a_row_stride + 4,
              "camera: matrix a needs %d rows of 4 columns per view (strides %d / %d)", a_rows, cam->a_view_stride,
              cam->a_row_stride);
  NRT_REQUIRE(cam->kind != NRT_CAM_DTU || (cam->b_row_stride >= 3 && cam->b_view_stride >= cam->b_row_stride + 3),
              "DTU camera: intrinsics need two rows of three columns per view (strides %d / %d)", cam->b_view_stride,
              cam->b_row_stride);
  NRT_REQUIRE(cam->kind != NRT_CAM_FOV || cam->b_view_stride >= 3, "FoV camera: centres need three floats per view");
  *total = (int64_t)cam->n_views * cam->nx * cam->ny * cam->bundle;
  return NRT_OK;
}

int nrt_camera_rays_dev(const nrt_camera_t* cam, int64_t r0, int64_t n, float* out_rays, int32_t* out_view,
                        cudaStream_t st) {
  CamDev c{cam->kind, cam->n_views, cam->a, cam->b, cam->a_view_stride, cam->a_row_stride, cam->b_view_stride,
           cam->b_row_stride, cam->focal, cam->size, cam->x0, cam->y0, cam->nx, cam->ny, cam->bundle,
           cam->pos_per_pixel, cam->positions, cam->jitter, cam->jitter_seed};
  NrtProfScope _ps(TAG_CAMERA_RAYS, st);
  k_camera_rays<<<nrt_cdiv(n, 256), 256, 0, st>>>(c, r0, n, out_rays, out_view);
  NRT_CUDA(cudaGetLastError());
  return NRT_OK;
}

extern "C" int nrt_camera_rays(const nrt_camera_t* cam, int64_t r0, int64_t n, float* out_rays, int32_t* out_view,
                               void* stream) {
  int64_t total = 0;
  int rc = nrt_check_camera(cam, &total);
  if (rc != NRT_OK) return rc;
  NRT_REQUIRE(r0 >= 0 && n >= 0 && r0 + n <= total, "nrt_camera_rays: rays %lld..%lld outside the block of %lld",
              (long long)r0, (long long)(r0 + n), (long long)total);
  if (n == 0) return NRT_OK;
  NRT_REQUIRE(out_rays != nullptr, "nrt_camera_rays: out_rays is NULL");
  return nrt_camera_rays_dev(cam, r0, n, out_rays, out_view, (cudaStream_t)stream);
}
