// Entry points declared in include/nrt_b200.h whose kernels are not written yet.
#include "nrt_common.cuh"
#define NRT_TODO(name) nrt_set_error(name ": not implemented yet"); return NRT_E_UNSUPPORTED;
extern "C" int nrt_mlp_backward(const nrt_mlp_t*, int, const float*, const float*, int64_t, const float*,
                                const float*, const float*, float*, float*, float*, void*) { NRT_TODO("nrt_mlp_backward") }
extern "C" int nrt_sdf_value_grad(const nrt_sphere_sdf_t*, const float*, int64_t, float*, float*, void*) { NRT_TODO("nrt_sdf_value_grad") }
extern "C" int nrt_shading_frame(const float*, const float*, int64_t, float*, float*, void*) { NRT_TODO("nrt_shading_frame") }
extern "C" int nrt_to_local(const float*, const float*, int64_t, float*, void*) { NRT_TODO("nrt_to_local") }
extern "C" int nrt_param_rusin2(const float*, const float*, int64_t, float*, void*) { NRT_TODO("nrt_param_rusin2") }
