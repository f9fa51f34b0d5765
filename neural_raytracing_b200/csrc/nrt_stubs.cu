// Entry points declared in include/nrt_b200.h whose kernels are not written yet.
#include "nrt_common.cuh"
#define NRT_TODO(name) nrt_set_error(name ": not implemented yet"); return NRT_E_UNSUPPORTED;
extern "C" int nrt_mlp_backward(const nrt_mlp_t*, int, const float*, const float*, int64_t, const float*,
                                const float*, const float*, float*, float*, float*, void*) { NRT_TODO("nrt_mlp_backward") }
