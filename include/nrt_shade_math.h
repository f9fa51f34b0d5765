/* Shading glue of the Direct integrator as scalar functions, written ONCE for two number types:
 *   float            -- the forward kernels (deterministic transcendentals of nrt_detmath.h, the reference's op order)
 *   nrt::Dual<N>     -- forward-mode dual numbers with N tangent directions: the backward kernels evaluate the same
 *                       function on Dual inputs, which yields its Jacobian, and contract it with the incoming gradient
 *                       (the per-ray functions have <= 6 inputs and a few hundred flops: 7x that is nothing next to
 *                       the MLPs, and the derivative code cannot drift from the forward code).
 *
 * Reference semantics (pytorch3d/pathtracer/):
 *   normalize_eps          F.normalize(x, eps)                      = x / max(|x|, eps)
 *   coordinate_system      interaction.py:9-27   (1e-6 / 1e-7 guards, three re-normalisations)
 *   to_local               interaction.py:38-41  (frame^T w / 3, re-normalised)
 *   param_rusin2           utils.py:233-258      (+ rotate_vector :152, nonzero_eps :43; the `1 - H_z` quirk)
 *   dir_to_elev_azim       utils.py:490-494
 *   point_light_*          lights/lights.py:89-110
 *   light_field_*          lights/lights.py:175-195 (direction components clamped to [1e-6, 1])
 *   fresnel_conductor      bsdf/bsdfs.py:327-341
 * Host-compilable (plain C++), so the same text is checked on the CPU against torch autograd of the mirror's torch
 * expressions (tests/test_shade_math_cpu.py) before it ever runs on a GPU.  */
#ifndef NRT_SHADE_MATH_H_
#define NRT_SHADE_MATH_H_

#include <math.h>

#include "nrt_detmath.h"

namespace nrt {

template <int N>
struct Dual {
  float v;
  float d[N];
};

/* ---- construction ---- */
template <int N> NRT_HD Dual<N> dconst(float v) {
  Dual<N> r; r.v = v;
  for (int i = 0; i < N; ++i) r.d[i] = 0.0f;
  return r;
}
template <int N> NRT_HD Dual<N> dvar(float v, int k) {
  Dual<N> r = dconst<N>(v);
  r.d[k] = 1.0f;
  return r;
}
NRT_HD float val(float a) { return a; }
template <int N> NRT_HD float val(const Dual<N>& a) { return a.v; }
/* a constant of the number type T */
#ifdef __CUDACC__
#define NRT_HDM __host__ __device__ __forceinline__ static
#else
#define NRT_HDM static inline
#endif
template <class T> struct Cst { NRT_HDM T of(float v) { return v; } };
template <int N> struct Cst<Dual<N> > { NRT_HDM Dual<N> of(float v) { return dconst<N>(v); } };

/* ---- arithmetic ---- */
template <int N> NRT_HD Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; r.v = a.v + b.v;
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i];
  return r;
}
template <int N> NRT_HD Dual<N> operator+(const Dual<N>& a, float b) { Dual<N> r = a; r.v = a.v + b; return r; }
template <int N> NRT_HD Dual<N> operator+(float a, const Dual<N>& b) { return b + a; }
template <int N> NRT_HD Dual<N> operator-(const Dual<N>& a) {
  Dual<N> r; r.v = -a.v;
  for (int i = 0; i < N; ++i) r.d[i] = -a.d[i];
  return r;
}
template <int N> NRT_HD Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; r.v = a.v - b.v;
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i];
  return r;
}
template <int N> NRT_HD Dual<N> operator-(const Dual<N>& a, float b) { Dual<N> r = a; r.v = a.v - b; return r; }
template <int N> NRT_HD Dual<N> operator-(float a, const Dual<N>& b) { return (-b) + a; }
template <int N> NRT_HD Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; r.v = a.v * b.v;
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
  return r;
}
template <int N> NRT_HD Dual<N> operator*(const Dual<N>& a, float b) {
  Dual<N> r; r.v = a.v * b;
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b;
  return r;
}
template <int N> NRT_HD Dual<N> operator*(float a, const Dual<N>& b) { return b * a; }
template <int N> NRT_HD Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; r.v = a.v / b.v;
  const float inv = 1.0f / b.v;
  for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
  return r;
}
template <int N> NRT_HD Dual<N> operator/(const Dual<N>& a, float b) {
  Dual<N> r; r.v = a.v / b;              /* the value exactly as the float path computes it */
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] / b;
  return r;
}
template <int N> NRT_HD Dual<N> operator/(float a, const Dual<N>& b) { return dconst<N>(a) / b; }

/* ---- elementary functions (value: the deterministic fp32 routines; derivative: the textbook one) ---- */
NRT_HD float nsqrt(float a) { return sqrtf(a); }
template <int N> NRT_HD Dual<N> nsqrt(const Dual<N>& a) {
  Dual<N> r; r.v = sqrtf(a.v);
  const float g = a.v > 0.0f ? 0.5f / r.v : 0.0f;
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * g;
  return r;
}
NRT_HD float nabs(float a) { return fabsf(a); }
template <int N> NRT_HD Dual<N> nabs(const Dual<N>& a) { return a.v < 0.0f ? -a : a; }
/* clamp(min=lo): torch passes the gradient where x >= lo ... strictly, where the input is not clamped */
NRT_HD float clamp_min(float a, float lo) { return fmaxf(a, lo); }
template <int N> NRT_HD Dual<N> clamp_min(const Dual<N>& a, float lo) { return a.v < lo ? dconst<N>(lo) : a; }
NRT_HD float clamp_max(float a, float hi) { return fminf(a, hi); }
template <int N> NRT_HD Dual<N> clamp_max(const Dual<N>& a, float hi) { return a.v > hi ? dconst<N>(hi) : a; }
NRT_HD float ncos(float a) { return nrt_cosf(a); }
template <int N> NRT_HD Dual<N> ncos(const Dual<N>& a) {
  float s, c;
  nrt_sincosf(a.v, &s, &c);
  Dual<N> r; r.v = c;
  for (int i = 0; i < N; ++i) r.d[i] = -s * a.d[i];
  return r;
}
NRT_HD float natan2(float y, float x) { return nrt_atan2f(y, x); }
template <int N> NRT_HD Dual<N> natan2(const Dual<N>& y, const Dual<N>& x) {
  Dual<N> r; r.v = nrt_atan2f(y.v, x.v);
  const float den = x.v * x.v + y.v * y.v;
  const float inv = den > 0.0f ? 1.0f / den : 0.0f;
  for (int i = 0; i < N; ++i) r.d[i] = (x.v * y.d[i] - y.v * x.d[i]) * inv;
  return r;
}
NRT_HD float nasin(float a) { return nrt_asinf(a); }
template <int N> NRT_HD Dual<N> nasin(const Dual<N>& a) {
  Dual<N> r; r.v = nrt_asinf(a.v);
  const float g = 1.0f / sqrtf(fmaxf(1.0f - a.v * a.v, 1e-30f));
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * g;
  return r;
}
NRT_HD float nexp(float a) { return nrt_expf(a); }
template <int N> NRT_HD Dual<N> nexp(const Dual<N>& a) {
  Dual<N> r; r.v = nrt_expf(a.v);
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * r.v;
  return r;
}
NRT_HD float nsigmoid(float a) { return nrt_sigmoidf(a); }
template <int N> NRT_HD Dual<N> nsigmoid(const Dual<N>& a) {
  Dual<N> r; r.v = nrt_sigmoidf(a.v);
  const float g = r.v * (1.0f - r.v);
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * g;
  return r;
}
NRT_HD float nsoftplus(float a) { return nrt_softplusf(a); }
template <int N> NRT_HD Dual<N> nsoftplus(const Dual<N>& a) {
  Dual<N> r; r.v = nrt_softplusf(a.v);
  const float g = a.v > 20.0f ? 1.0f : nrt_sigmoidf(a.v);
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * g;
  return r;
}

/* ---- small vector helpers ---- */
template <class T> NRT_HD T dot3(const T a[3], const T b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <class T> NRT_HD void cross3(const T a[3], const T b[3], T o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
/* F.normalize(v, eps): v / max(|v|, eps); returns |v| through `norm` if wanted.  Below eps the divisor is the constant */
template <class T> NRT_HD void normalize_eps(T v[3], float eps, T* norm = nullptr) {
  const T n = nsqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  const T d = clamp_min(n, eps);
  if (norm) *norm = n;
  v[0] = v[0] / d; v[1] = v[1] / d; v[2] = v[2] / d;
}

/* interaction.py:9-27: frame columns (s, t, n) around the (re-normalised) normal */
template <class T> NRT_HD void coordinate_system(const T nin[3], T s[3], T t[3], T n[3]) {
  n[0] = nin[0]; n[1] = nin[1]; n[2] = nin[2];
  normalize_eps(n, 1e-7f);
  const float sign = val(n[2]) >= 0.0f ? 1.0f : -1.0f;
  const T sz = n[2] + sign;
  T a;
  if (fabsf(val(sz)) < 1e-6f) a = (sz - sz) - (1.0f / 1e-6f);      /* constant -1/1e-6 (keeps the number type) */
  else a = -(1.0f / sz);
  const T b = n[0] * n[1] * a;
  s[0] = n[0] * n[0] * a * sign + 1.0f; s[1] = b * sign; s[2] = n[0] * (-sign);
  normalize_eps(s, 1e-7f);
  cross3(s, n, t);
  normalize_eps(t, 1e-7f);
  cross3(n, t, s);
  normalize_eps(s, 1e-7f);
}
/* interaction.py:38-41 on the frame of normal n: normalize((frame^T w) / 3) */
template <class T> NRT_HD void to_local_n(const T n_in[3], const T w[3], T o[3]) {
  T s[3], t[3], n[3];
  coordinate_system(n_in, s, t, n);
  o[0] = dot3(s, w) / 3.0f; o[1] = dot3(t, w) / 3.0f; o[2] = dot3(n, w) / 3.0f;
  normalize_eps(o, 1e-7f);
}
template <class T> NRT_HD T nonzero_eps(const T& v) { return fabsf(val(v)) < 1e-7f ? (v - v) + 1e-7f : v; }
template <class T> NRT_HD void rotate_vector(const T v[3], const float axis[3], const T& c, const T& s, T o[3]) {
  const T d = v[0] * axis[0] + v[1] * axis[1] + v[2] * axis[2];
  T cr[3];
  /* cross(axis, v) with a constant axis */
  cr[0] = v[2] * axis[1] - v[1] * axis[2];
  cr[1] = v[0] * axis[2] - v[2] * axis[0];
  cr[2] = v[1] * axis[0] - v[0] * axis[1];
  for (int i = 0; i < 3; ++i) o[i] = v[i] * c + d * (1.0f - c) * axis[i] + cr[i] * s;
}
/* utils.py:233-258: param_rusin2(wo_arg, wi_arg) -> (cos phi_d, cos theta_h, cos theta_d) */
template <class T> NRT_HD void param_rusin2(const T a_in[3], const T b_in[3], T out[3]) {
  T wo[3] = {a_in[0], a_in[1], a_in[2]};
  T wi[3] = {b_in[0], b_in[1], b_in[2]};
  normalize_eps(wo, 1e-12f);
  normalize_eps(wi, 1e-12f);
  T H[3] = {wo[0] + wi[0], wo[1] + wi[1], wo[2] + wi[2]};
  normalize_eps(H, 1e-12f);
  const float e1[3] = {0.0f, 1.0f, 0.0f}, e2[3] = {0.0f, 0.0f, 1.0f};
  const T hy = nonzero_eps(H[1]), hx = nonzero_eps(H[0]);
  const T rr = clamp_min(nsqrt(hy * hy + hx * hx), 1e-6f);
  T tmp[3], diff[3];
  rotate_vector(wi, e2, H[0] / rr, -(H[1] / rr), tmp);
  normalize_eps(tmp, 1e-12f);
  const T s = -nsqrt(clamp_min(1.0f - H[2], 1e-6f));          /* quirk: 1 - H_z, not 1 - H_z^2 */
  rotate_vector(tmp, e1, H[2], s, diff);
  normalize_eps(diff, 1e-12f);
  out[0] = ncos(natan2(nonzero_eps(diff[1]), nonzero_eps(diff[0])));
  out[1] = H[2];
  out[2] = diff[2];
}
/* utils.py:490-494 */
template <class T> NRT_HD void dir_to_elev_azim(const T d_in[3], T out[2]) {
  T d[3] = {d_in[0], d_in[1], d_in[2]};
  normalize_eps(d, 1e-12f);
  for (int i = 0; i < 3; ++i) d[i] = clamp_max(clamp_min(d[i], -1.0f + 1e-7f), 1.0f - 1e-7f);
  out[0] = nasin(d[2]);
  out[1] = natan2(d[0], nsqrt(clamp_min(1.0f - d[0] * d[0] - d[2] * d[2], 1e-10f)));
}
/* bsdfs.py:327-341 */
template <class T, class U> NRT_HD T fresnel_conductor(const T& cos_t, const U& eta_r, float eta_i) {
  const T ct2 = cos_t * cos_t;
  const T st2 = clamp_min(1.0f - ct2, 1e-10f);
  const T st4 = st2 * st2;
  const T tmp = (eta_r * eta_r - eta_i * eta_i) - st2;
  const T a2b2 = nsqrt(clamp_min(tmp * tmp + (eta_r * eta_r) * (4.0f * eta_i * eta_i), 1e-10f));
  const T a = nsqrt(clamp_min((a2b2 + tmp) * 0.5f, 1e-10f));
  const T t1 = a2b2 + ct2;
  const T t2 = cos_t * a * 2.0f;
  const T r_s = (t1 - t2) / (t1 + t2);
  const T t3 = a2b2 * ct2 + st4;
  const T t4 = t2 * st2;
  const T r_p = r_s * (t3 - t4) / (t3 + t4);
  return (r_s + r_p) * 0.5f;
}

/* ---- the per-hit stages of the fused Direct integrator (integrators.py:156-206) ---------------------------------- */
/* sdfs.py:156-159: n = normalize(raw_n, 1e-6); wi = to_local(frame(n), -r_d).  (p += 5 eps n is linear: done by the caller) */
template <class T> NRT_HD void stage_geom(const T raw_n[3], const float r_d[3], T n[3], T wi[3]) {
  n[0] = raw_n[0]; n[1] = raw_n[1]; n[2] = raw_n[2];
  normalize_eps(n, 1e-6f);
  const T md[3] = {Cst<T>::of(-r_d[0]), Cst<T>::of(-r_d[1]), Cst<T>::of(-r_d[2])};
  to_local_n(n, md, wi);
}
/* lights.py:98-102: d = L - p, dist = |d|, d = normalize(d, 1e-6) */
template <class T> NRT_HD void stage_point_light(const T p[3], const float L[3], T d[3], T* dist) {
  d[0] = L[0] - p[0]; d[1] = L[1] - p[1]; d[2] = L[2] - p[2];
  normalize_eps(d, 1e-6f, dist);
}
/* lights.py:189-193: d = clamp(normalize(v, 1e-6), 1e-6, 1), magnitude |v| */
template <class T> NRT_HD void stage_light_field(const T v[3], T d[3], T* magn) {
  d[0] = v[0]; d[1] = v[1]; d[2] = v[2];
  normalize_eps(d, 1e-6f, magn);
  for (int i = 0; i < 3; ++i) d[i] = clamp_max(clamp_min(d[i], 1e-6f), 1.0f);
}
/* lights.py:103-108: denominator of the point-light falloff (coefficients already clamped at 1e-6) */
template <class T> NRT_HD T point_light_denominator(const T& dist, float c, float l, float q) {
  return clamp_min(dist * l + dist * dist * q + c, 1e-6f);
}

}  /* namespace nrt */

#endif /* NRT_SHADE_MATH_H_ */
