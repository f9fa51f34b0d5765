// Ray generators on the device: pixel position -> [origin | unit direction] for the three cameras
// the scripts use, written straight into the reference's [n_views, nx, ny, bundle, 6] block.  One thread per ray,
// 24 contiguous bytes per thread (HBM-bound: 24 B per ray written, nothing read but the camera matrices, which every
// thread reads from the same few cache lines).  The arithmetic (camera_math.cuh) follows the reference's operation
// order with explicit round-to-nearest operations, so a ray differs from the torch expression by the rounding of the
// 3- and 4-term sums only, and the camera-fed NeRF kernels (nrt_tc.cu) compute bit-identical rays.
//   NeRFCamera.sample_positions             pytorch3d/pathtracer/cameras/cameras.py:23-54
//   DTUCamera.sample_positions, lift        pytorch3d/pathtracer/cameras/cameras.py:132-147, 156-192
//   FoVPerspectiveCameras.sample_positions  pytorch3d/renderer/cameras.py:539-575
#include "nrt_common.cuh"
#include "camera_math.cuh"

using nrtcam::CamDev;
using nrtcam::cam_ray;

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
  const CamDev c = nrtcam::make_cam_dev(cam);
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
