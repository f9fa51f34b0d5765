// Test harness (CPU): evaluates the scalar functions of include/nrt_shade_math.h -- the text the CUDA shading kernels
// compile -- for float and for dual numbers, so that values and Jacobians can be checked against torch autograd of the
// mirror's torch expressions (tests/test_shade_math_cpu.py).  TEST INFRASTRUCTURE: nothing in the product loads this.
#include "nrt_shade_math.h"

using namespace nrt;

template <int NIN, class F>
static void run(const float* x, float* y, float* J, int n_out, F f) {
  Dual<NIN> in[NIN], out[8];
  for (int i = 0; i < NIN; ++i) in[i] = dvar<NIN>(x[i], i);
  f(in, out);
  for (int o = 0; o < n_out; ++o) {
    y[o] = out[o].v;
    for (int i = 0; i < NIN; ++i) J[o * NIN + i] = out[o].d[i];
  }
}

// fn: 0 normalize(1e-6) [3->3]   1 to_local_n(n, w) [6->3]   2 param_rusin2(a, b) [6->3]   3 dir_to_elev_azim [3->2]
//     4 fresnel_conductor(cos, eta) [2->1]   5 point light (p; L = c[0..2]) [3->4: d, dist]   6 light field (v) [3->4: d, |v|]
//     7 stage_geom(raw_n; r_d = c[0..2]) [3->6: n, wi]   8 point_light_denominator(dist; c,l,q = c[0..2]) [1->1]
extern "C" int shade_eval(int fn, const float* x, const float* c, float* y, float* J, float* y_float) {
  switch (fn) {
    case 0: {
      run<3>(x, y, J, 3, [](Dual<3>* in, Dual<3>* out) { Dual<3> v[3] = {in[0], in[1], in[2]}; normalize_eps(v, 1e-6f); out[0] = v[0]; out[1] = v[1]; out[2] = v[2]; });
      float v[3] = {x[0], x[1], x[2]}; normalize_eps(v, 1e-6f); y_float[0] = v[0]; y_float[1] = v[1]; y_float[2] = v[2];
      return 3; }
    case 1: {
      run<6>(x, y, J, 3, [](Dual<6>* in, Dual<6>* out) { to_local_n(in, in + 3, out); });
      to_local_n(x, x + 3, y_float);
      return 3; }
    case 2: {
      run<6>(x, y, J, 3, [](Dual<6>* in, Dual<6>* out) { param_rusin2(in, in + 3, out); });
      param_rusin2(x, x + 3, y_float);
      return 3; }
    case 3: {
      run<3>(x, y, J, 2, [](Dual<3>* in, Dual<3>* out) { dir_to_elev_azim(in, out); });
      dir_to_elev_azim(x, y_float);
      return 2; }
    case 4: {
      run<2>(x, y, J, 1, [](Dual<2>* in, Dual<2>* out) { out[0] = fresnel_conductor(in[0], in[1], 0.0f); });
      y_float[0] = fresnel_conductor(x[0], x[1], 0.0f);
      return 1; }
    case 5: {
      run<3>(x, y, J, 4, [c](Dual<3>* in, Dual<3>* out) { stage_point_light(in, c, out, out + 3); });
      stage_point_light(x, c, y_float, y_float + 3);
      return 4; }
    case 6: {
      run<3>(x, y, J, 4, [](Dual<3>* in, Dual<3>* out) { stage_light_field(in, out, out + 3); });
      stage_light_field(x, y_float, y_float + 3);
      return 4; }
    case 7: {
      run<3>(x, y, J, 6, [c](Dual<3>* in, Dual<3>* out) { stage_geom(in, c, out, out + 3); });
      stage_geom(x, c, y_float, y_float + 3);
      return 6; }
    case 8: {
      run<1>(x, y, J, 1, [c](Dual<1>* in, Dual<1>* out) { out[0] = point_light_denominator(in[0], c[0], c[1], c[2]); });
      y_float[0] = point_light_denominator(x[0], c[0], c[1], c[2]);
      return 1; }
  }
  return -1;
}
