/* nrt_detmath.h -- deterministic fp32 elementary functions.
 *
 * Part of the public contract of the fp32 ("exact") path of libnrt_b200: every
 * transcendental the hot path needs (sin, cos, exp, log, softplus, sigmoid, tanh, atan2,
 * asin) is written here in terms of IEEE-754 binary32 add / mul / fma / div / sqrt /
 * rint only, so a host C compiler (gcc -O2 -ffp-contract=off) and nvcc
 * (-fmad=false, default -prec-div/-prec-sqrt) produce bit-identical results.  That is what
 * lets the sphere-trace hit mask of the CUDA kernels be compared bit-for-bit with the CPU
 * oracle (oracle/c/nrt_oracle.c), which includes this same header as its arithmetic spec.
 *
 * The reference (eager PyTorch, e.g. pytorch3d/pathtracer/utils.py:37-40 sin/cos,
 * :385-387 exp/log, neural_blocks.py:26 / sdfs.py:29 leaky_relu/softplus) uses the
 * platform libm / SLEEF; these functions agree with those to a few ulp, which is the
 * tolerance the reference-vs-oracle golden tests state.
 *
 * Rules for code that wants bit-exactness across host/device:
 *   - every multiply-add that is meant to be fused is written nrt_fma(a,b,c);
 *   - nothing else may be contracted (build flags above);
 *   - no libm calls besides sqrtf / rintf / fabsf / fmaf.
 */
#ifndef NRT_DETMATH_H_
#define NRT_DETMATH_H_

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define NRT_HD __host__ __device__ __forceinline__
#else
#define NRT_HD static inline
#endif

NRT_HD float nrt_fma(float a, float b, float c) { return fmaf(a, b, c); }

NRT_HD uint32_t nrt_f2u(float x) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(x);
#else
  uint32_t u; memcpy(&u, &x, 4); return u;
#endif
}
NRT_HD float nrt_u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float x; memcpy(&x, &u, 4); return x;
#endif
}

/* ---- sin / cos -------------------------------------------------------------------- */
/* Cody-Waite reduction by pi/2 with a 3-term split and fma, then minimax polynomials on
 * [-pi/4, pi/4].  Good to ~1.5 ulp for |x| < ~5e4, far beyond the Fourier-feature
 * arguments of the hot path (|x.B| up to a few hundred radians, SURVEY.md hard part 2). */
NRT_HD void nrt_sincosf(float x, float* s_out, float* c_out) {
  const float two_over_pi = 0.636619772367581343f;
  const float p1 = 1.57079637050628662109375f;      /* fl(pi/2)            */
  const float p2 = -4.37113900018624283e-8f;        /* fl(pi/2 - p1)       */
  const float p3 = -1.71512449885542256e-15f;       /* fl(pi/2 - p1 - p2)  */
  float kf = rintf(x * two_over_pi);
  float r = nrt_fma(-kf, p1, x);
  r = nrt_fma(-kf, p2, r);
  r = nrt_fma(-kf, p3, r);
  int q = (int)kf;
  float r2 = r * r;
  /* sin(r) = r + r^3 * S(r^2) */
  float sp = nrt_fma(r2, -1.9515295891e-4f, 8.3321608736e-3f);
  sp = nrt_fma(sp, r2, -1.6666654611e-1f);
  float sr = nrt_fma(sp * r2, r, r);
  /* cos(r) = 1 - r^2/2 + r^4 * C(r^2) */
  float cp = nrt_fma(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
  cp = nrt_fma(cp, r2, 4.166664568298827e-2f);
  float cr = nrt_fma(cp * r2, r2, nrt_fma(-0.5f, r2, 1.0f));
  float s = (q & 1) ? cr : sr;
  float c = (q & 1) ? sr : cr;
  if (q & 2) s = -s;
  if ((q + 1) & 2) c = -c;
  *s_out = s;
  *c_out = c;
}
NRT_HD float nrt_sinf(float x) { float s, c; nrt_sincosf(x, &s, &c); return s; }
NRT_HD float nrt_cosf(float x) { float s, c; nrt_sincosf(x, &s, &c); return c; }

/* ---- exp -------------------------------------------------------------------------- */
/* exp(x) = 2^k * exp(r), r in [-ln2/2, ln2/2].  Results below 2^-126 flush to 0 and
 * x > 88.72 returns +inf; the hot path only feeds exp with bounded arguments. */
NRT_HD float nrt_expf(float x) {
  if (x < -87.33654f) return 0.0f;
  if (x > 88.72283f) return nrt_u2f(0x7f800000u);
  const float log2e = 1.44269504088896341f;
  const float ln2_hi = 0.693145751953125f;          /* 0x3f317200 */
  const float ln2_lo = 1.42860682030941723e-6f;     /* ln2 - ln2_hi */
  float kf = rintf(x * log2e);
  float r = nrt_fma(-kf, ln2_hi, x);
  r = nrt_fma(-kf, ln2_lo, r);
  float p = 1.9875691500e-4f;
  p = nrt_fma(p, r, 1.3981999507e-3f);
  p = nrt_fma(p, r, 8.3334519073e-3f);
  p = nrt_fma(p, r, 4.1665795894e-2f);
  p = nrt_fma(p, r, 1.6666665459e-1f);
  p = nrt_fma(p, r, 5.0000001201e-1f);
  float e = nrt_fma(p * r, r, r) + 1.0f;
  int k = (int)kf;
  /* scale by 2^k in two steps so that k = -126..128 never needs a denormal factor */
  int k1 = k / 2, k2 = k - k1;
  float f1 = nrt_u2f((uint32_t)(k1 + 127) << 23);
  float f2 = nrt_u2f((uint32_t)(k2 + 127) << 23);
  return (e * f1) * f2;
}

/* ---- log -------------------------------------------------------------------------- */
/* log(x) for finite x > 0 (normal or denormal).  Cephes-style: x = m * 2^e with
 * m in [sqrt(1/2), sqrt(2)), log(m) by a degree-8 polynomial in f = m-1. */
NRT_HD float nrt_logf(float x) {
  if (!(x > 0.0f)) return (x == 0.0f) ? nrt_u2f(0xff800000u) : nrt_u2f(0x7fc00000u);
  uint32_t u = nrt_f2u(x);
  int e = 0;
  if (u < 0x00800000u) { x = x * 8388608.0f; u = nrt_f2u(x); e = -23; }
  e += (int)(u >> 23) - 126;
  float m = nrt_u2f((u & 0x007fffffu) | 0x3f000000u);   /* [0.5, 1) */
  if (m < 0.707106781186547524f) { e -= 1; m = m + m; }
  float f = m - 1.0f;
  float z = f * f;
  float p = 7.0376836292e-2f;
  p = nrt_fma(p, f, -1.1514610310e-1f);
  p = nrt_fma(p, f, 1.1676998740e-1f);
  p = nrt_fma(p, f, -1.2420140846e-1f);
  p = nrt_fma(p, f, 1.4249322787e-1f);
  p = nrt_fma(p, f, -1.6668057665e-1f);
  p = nrt_fma(p, f, 2.0000714765e-1f);
  p = nrt_fma(p, f, -2.4999993993e-1f);
  p = nrt_fma(p, f, 3.3333331174e-1f);
  float y = (f * z) * p;
  float fe = (float)e;
  y = nrt_fma(fe, -2.12194440e-4f, y);
  y = nrt_fma(-0.5f, z, y);
  float r = f + y;
  return nrt_fma(fe, 0.693359375f, r);
}

/* log1p(y) for y >= 0 via the compensated u = 1+y trick. */
NRT_HD float nrt_log1pf(float y) {
  float u = 1.0f + y;
  if (u == 1.0f) return y;
  float l = nrt_logf(u);
  return l * (y / (u - 1.0f));
}

/* torch.nn.functional.softplus(beta=1, threshold=20): x if x > 20 else log1p(exp(x)). */
NRT_HD float nrt_softplusf(float x) {
  if (x > 20.0f) return x;
  return nrt_log1pf(nrt_expf(x));
}
/* 1 / (1 + exp(-x)) */
NRT_HD float nrt_sigmoidf(float x) { return 1.0f / (1.0f + nrt_expf(-x)); }

NRT_HD float nrt_tanhf(float x) {
  float ax = fabsf(x);
  if (ax > 9.02f) return x < 0.0f ? -1.0f : 1.0f;
  float t;
  if (ax < 0.5f) {
    /* odd polynomial: tanh(x) = x + x^3 * P(x^2), avoids cancellation near 0 */
    float z = ax * ax;
    float p = -5.70498872745e-3f;
    p = nrt_fma(p, z, 2.06390887954e-2f);
    p = nrt_fma(p, z, -5.37397155531e-2f);
    p = nrt_fma(p, z, 1.33314422036e-1f);
    p = nrt_fma(p, z, -3.33332819422e-1f);
    t = nrt_fma(p * z, ax, ax);
  } else {
    float e = nrt_expf(2.0f * ax);
    t = 1.0f - 2.0f / (e + 1.0f);
  }
  return x < 0.0f ? -t : t;
}

/* ---- atan / atan2 / asin ------------------------------------------------------------ */
NRT_HD float nrt_atanf_pos(float x) { /* x >= 0 */
  float y0 = 0.0f;
  float z = x;
  if (x > 2.414213562373095f) { y0 = 1.57079632679489662f; z = -1.0f / x; }
  else if (x > 0.4142135623730950f) { y0 = 0.785398163397448310f; z = (x - 1.0f) / (x + 1.0f); }
  float z2 = z * z;
  float p = 8.05374449538e-2f;
  p = nrt_fma(p, z2, -1.38776856032e-1f);
  p = nrt_fma(p, z2, 1.99777106478e-1f);
  p = nrt_fma(p, z2, -3.33329491539e-1f);
  return y0 + nrt_fma(p * z2, z, z);
}
NRT_HD float nrt_atan2f(float y, float x) {
  const float pi = 3.14159265358979324f;
  if (x == 0.0f && y == 0.0f) return 0.0f;
  float ax = fabsf(x), ay = fabsf(y);
  float a;
  if (ax == 0.0f) a = 1.57079632679489662f;
  else a = nrt_atanf_pos(ay / ax);
  if (x < 0.0f) a = pi - a;
  return y < 0.0f ? -a : a;
}
NRT_HD float nrt_asinf(float x) {
  float ax = fabsf(x);
  if (ax > 1.0f) ax = 1.0f;
  float r = nrt_atan2f(ax, sqrtf((1.0f - ax) * (1.0f + ax)));
  return x < 0.0f ? -r : r;
}

#endif /* NRT_DETMATH_H_ */
