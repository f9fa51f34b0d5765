// tcgen05 / TMEM tensor-core path (placeholder until the fused kernels land in this file).
#include "nrt_common.cuh"

extern "C" int64_t nrt_mlp_tc_blob_bytes(const nrt_mlp_t* m, int prec) {
  (void)m; (void)prec;
  nrt_set_error("tensor-core path not built yet");
  return NRT_E_UNSUPPORTED;
}
extern "C" int nrt_mlp_pack_tc(const nrt_mlp_t* m, int prec, void* blob_out, void* stream) {
  (void)m; (void)prec; (void)blob_out; (void)stream;
  nrt_set_error("tensor-core path not built yet");
  return NRT_E_UNSUPPORTED;
}
int nrt_mlp_forward_tc(const nrt_mlp_t*, int, int, const float*, const float*, int64_t, float*, cudaStream_t) {
  nrt_set_error("tensor-core path not built yet");
  return NRT_E_UNSUPPORTED;
}
int nrt_sdf_eval_tc(const nrt_sphere_sdf_t*, int, const float*, int64_t, float*, cudaStream_t) {
  nrt_set_error("tensor-core path not built yet");
  return NRT_E_UNSUPPORTED;
}
int nrt_nerfle_pass_tc(const nrt_mlp_t*, const nrt_mlp_t*, int, const float*, int64_t, const float*, const float*,
                       int, const float*, int, const int32_t*, int, float*, float*, float*, void*, size_t,
                       cudaStream_t) {
  nrt_set_error("tensor-core path not built yet");
  return NRT_E_UNSUPPORTED;
}
size_t nrt_nerfle_pass_tc_workspace(const nrt_mlp_t*, const nrt_mlp_t*, int64_t, int) { return 0; }
