// Fused shading glue of the Direct integrator on the K COMPACTED hit rays: everything between
// the MLP evaluations of integrators.py:156-206 as three elementwise stages, forward and backward, one thread per hit.
//
//   geom   : raw normal -> n, offset hit point, local incoming direction (+ frame)       sdfs.py:152-159, interaction.py:9-41
//   light  : light sample -> world / local direction, Rusinkiewicz coords, emitter spectrum, occlusion-MLP input
//                                                                                         lights.py:89-110, 175-195; utils.py:233-258, 490-494
//   blend  : sigmoid(sp_var logits) . [NeuralBSDF | Diffuse | Conductor] spectra x emitter bsdfs.py:108-118, 364-388, 515-536, 634-637
//
// The scalar math lives in include/nrt_shade_math.h, written once for float and for dual numbers: the backward kernels
// evaluate the SAME functions on nrt::Dual<N> inputs (N <= 6), which yields their Jacobians, and contract them with the
// incoming gradients in reverse stage order.  Reductions over the hits (light / BSDF parameters) are block sums + one
// atomicAdd per block.  HBM-bound: ~100-250 B per hit and kernel.
#include "nrt_common.cuh"
#include "nrt_shade_math.h"

using namespace nrt;

namespace {

constexpr int kThreads = 128;

__device__ __forceinline__ void ld3(const float* p, int64_t k, float v[3]) { v[0] = p[k * 3]; v[1] = p[k * 3 + 1]; v[2] = p[k * 3 + 2]; }
__device__ __forceinline__ void st3(float* p, int64_t k, const float v[3]) { p[k * 3] = v[0]; p[k * 3 + 1] = v[1]; p[k * 3 + 2] = v[2]; }

// sum of `v` over the block -> one atomicAdd (all threads of the block must call)
__device__ __forceinline__ void block_atomic_add(float* dst, float v) {
  __shared__ float s_part[kThreads / 32];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int w = 0; w < kThreads / 32; ++w) t += s_part[w];
    if (t != 0.0f) atomicAdd(dst, t);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// geom
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_shade_geom_fwd(const float* __restrict__ raw_n, const float* __restrict__ p_hit, const float* __restrict__ rays, int64_t K,
                 float eps5, float* __restrict__ n_out, float* __restrict__ p_off, float* __restrict__ wi_out,
                 float* __restrict__ frame) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float rn[3], p[3], n[3], wi[3];
  ld3(raw_n, k, rn); ld3(p_hit, k, p);
  const float rd[3] = {rays[k * 6 + 3], rays[k * 6 + 4], rays[k * 6 + 5]};
  stage_geom(rn, rd, n, wi);
  st3(n_out, k, n); st3(wi_out, k, wi);
  const float po[3] = {p[0] + n[0] * eps5, p[1] + n[1] * eps5, p[2] + n[2] * eps5};   // sdfs.py:157
  st3(p_off, k, po);
  if (frame != nullptr) {
    float s[3], t[3], nn[3];
    coordinate_system(n, s, t, nn);
    for (int i = 0; i < 3; ++i) { frame[k * 9 + i * 3] = s[i]; frame[k * 9 + i * 3 + 1] = t[i]; frame[k * 9 + i * 3 + 2] = nn[i]; }
  }
}

__global__ void __launch_bounds__(kThreads)
k_shade_geom_bwd(const float* __restrict__ raw_n, const float* __restrict__ rays, int64_t K, float eps5,
                 const float* __restrict__ g_n, const float* __restrict__ g_p_off, const float* __restrict__ g_wi,
                 float* __restrict__ g_raw_n) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float rn[3];
  ld3(raw_n, k, rn);
  const float rd[3] = {rays[k * 6 + 3], rays[k * 6 + 4], rays[k * 6 + 5]};
  Dual<3> in[3] = {dvar<3>(rn[0], 0), dvar<3>(rn[1], 1), dvar<3>(rn[2], 2)}, n[3], wi[3];
  stage_geom(in, rd, n, wi);
  float gn[3] = {0, 0, 0}, gw[3] = {0, 0, 0};
  if (g_n) ld3(g_n, k, gn);
  if (g_p_off) { float gp[3]; ld3(g_p_off, k, gp); for (int i = 0; i < 3; ++i) gn[i] += eps5 * gp[i]; }
  if (g_wi) ld3(g_wi, k, gw);
  float g[3];
  for (int j = 0; j < 3; ++j) {
    float a = 0.0f;
    for (int i = 0; i < 3; ++i) a += gn[i] * n[i].d[j] + gw[i] * wi[i].d[j];
    g[j] = a;
  }
  st3(g_raw_n, k, g);
}

// ------------------------------------------------------------------------------------------------------------------
// light
// ------------------------------------------------------------------------------------------------------------------
struct LightDev {
  int mode;                      // 0: point lights, 1: light field
  const float* location;         // [n_views,3]
  const float* amp;              // [n_views,3] = scale * normalize(intensity)
  const float* coef;             // [3] const, linear, square, each already clamped at 1e-6
  const int32_t* view_of_hit;    // [K] or null (single view)
  const float* v;                // [K,3] light-field MLP output
  const float* sig_color;        // [3] sigmoid(color)
};

__global__ void __launch_bounds__(kThreads)
k_shade_light_fwd(LightDev L, const float* __restrict__ n_in, const float* __restrict__ wi_in, const float* __restrict__ p_off,
                  int64_t K, float* __restrict__ d_out, float* __restrict__ dist_out, float* __restrict__ wo_out,
                  float* __restrict__ rusin_out, float* __restrict__ e_out, float* __restrict__ elaz_out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float n[3], wi[3], d[3], dist, e[3];
  ld3(n_in, k, n); ld3(wi_in, k, wi);
  if (L.mode == 0) {
    float p[3];
    ld3(p_off, k, p);
    const int view = L.view_of_hit ? L.view_of_hit[k] : 0;
    const float loc[3] = {L.location[view * 3], L.location[view * 3 + 1], L.location[view * 3 + 2]};
    stage_point_light(p, loc, d, &dist);
    const float den = point_light_denominator(dist, L.coef[0], L.coef[1], L.coef[2]);
    for (int c = 0; c < 3; ++c) e[c] = L.amp[view * 3 + c] / den;
  } else {
    float v[3];
    ld3(L.v, k, v);
    stage_light_field(v, d, &dist);
    for (int c = 0; c < 3; ++c) e[c] = dist * L.sig_color[c];
  }
  float wo[3], ru[3];
  to_local_n(n, d, wo);
  param_rusin2(wi, wo, ru);          // bsdfs.py:635: param_rusin2(it.wi, wo)
  st3(d_out, k, d); dist_out[k] = dist; st3(wo_out, k, wo); st3(rusin_out, k, ru); st3(e_out, k, e);
  if (elaz_out != nullptr) {
    float ea[2];
    dir_to_elev_azim(d, ea);
    elaz_out[k * 2] = ea[0]; elaz_out[k * 2 + 1] = ea[1];
  }
}

__global__ void __launch_bounds__(kThreads)
k_shade_light_bwd(LightDev L, const float* __restrict__ n_in, const float* __restrict__ wi_in, const float* __restrict__ p_off,
                  int64_t K, const float* __restrict__ g_wo_in, const float* __restrict__ g_rusin, const float* __restrict__ g_e,
                  const float* __restrict__ g_elaz, float* __restrict__ g_n, float* __restrict__ g_wi, float* __restrict__ g_pv,
                  float* __restrict__ g_amp, float* __restrict__ g_coef, float* __restrict__ g_sig_color, int n_views) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = k < K;
  float ga[3] = {0, 0, 0}, gc[3] = {0, 0, 0}, gs[3] = {0, 0, 0};
  int view = 0;
  if (live) {
    float n[3], wi[3], x[3];
    ld3(n_in, k, n); ld3(wi_in, k, wi);
    view = (L.mode == 0 && L.view_of_hit) ? L.view_of_hit[k] : 0;
    float loc[3] = {0, 0, 0};
    if (L.mode == 0) { ld3(p_off, k, x); loc[0] = L.location[view * 3]; loc[1] = L.location[view * 3 + 1]; loc[2] = L.location[view * 3 + 2]; }
    else ld3(L.v, k, x);
    // (i) light sample: x (p_off or v) -> d, dist, with its Jacobian
    Dual<3> xi[3] = {dvar<3>(x[0], 0), dvar<3>(x[1], 1), dvar<3>(x[2], 2)}, dd[3], ddist;
    if (L.mode == 0) stage_point_light(xi, loc, dd, &ddist); else stage_light_field(xi, dd, &ddist);
    const float d[3] = {dd[0].v, dd[1].v, dd[2].v};
    // (ii) wo = to_local(frame(n), d) with its Jacobian w.r.t. (n, d)
    Dual<6> nd[6] = {dvar<6>(n[0], 0), dvar<6>(n[1], 1), dvar<6>(n[2], 2), dvar<6>(d[0], 3), dvar<6>(d[1], 4), dvar<6>(d[2], 5)}, dwo[3];
    to_local_n(nd, nd + 3, dwo);
    const float wo[3] = {dwo[0].v, dwo[1].v, dwo[2].v};
    // (iii) rusin = param_rusin2(wi, wo) with its Jacobian
    float gwo[3] = {0, 0, 0}, gwi[3] = {0, 0, 0};
    if (g_wo_in) ld3(g_wo_in, k, gwo);
    if (g_rusin) {
      float gr[3];
      ld3(g_rusin, k, gr);
      Dual<6> ab[6] = {dvar<6>(wi[0], 0), dvar<6>(wi[1], 1), dvar<6>(wi[2], 2), dvar<6>(wo[0], 3), dvar<6>(wo[1], 4), dvar<6>(wo[2], 5)}, dr[3];
      param_rusin2(ab, ab + 3, dr);
      for (int j = 0; j < 3; ++j) {
        float a = 0.0f, b = 0.0f;
        for (int i = 0; i < 3; ++i) { a += gr[i] * dr[i].d[j]; b += gr[i] * dr[i].d[3 + j]; }
        gwi[j] += a; gwo[j] += b;
      }
    }
    // back through (ii)
    float gn[3], gd[3];
    for (int j = 0; j < 3; ++j) {
      float a = 0.0f, b = 0.0f;
      for (int i = 0; i < 3; ++i) { a += gwo[i] * dwo[i].d[j]; b += gwo[i] * dwo[i].d[3 + j]; }
      gn[j] = a; gd[j] = b;
    }
    // occlusion-MLP input (elev, azim)(d)
    if (g_elaz) {
      Dual<3> di[3] = {dvar<3>(d[0], 0), dvar<3>(d[1], 1), dvar<3>(d[2], 2)}, ea[2];
      dir_to_elev_azim(di, ea);
      const float g0 = g_elaz[k * 2], g1 = g_elaz[k * 2 + 1];
      for (int j = 0; j < 3; ++j) gd[j] += g0 * ea[0].d[j] + g1 * ea[1].d[j];
    }
    // emitter spectrum -> g_dist and the light parameters
    float gdist = 0.0f;
    if (g_e) {
      float ge[3];
      ld3(g_e, k, ge);
      if (L.mode == 0) {
        Dual<1> dist1 = dvar<1>(ddist.v, 0);
        const Dual<1> den = point_light_denominator(dist1, L.coef[0], L.coef[1], L.coef[2]);
        const float inv = 1.0f / den.v;
        float g_den = 0.0f;
        for (int c = 0; c < 3; ++c) {
          const float amp = L.amp[view * 3 + c];
          ga[c] = ge[c] * inv;
          g_den -= ge[c] * amp * inv * inv;
        }
        gdist = g_den * den.d[0];
        if (den.v > 1e-6f) {                       // not clamped: den = c + l dist + q dist^2
          gc[0] = g_den; gc[1] = g_den * ddist.v; gc[2] = g_den * ddist.v * ddist.v;
        }
      } else {
        for (int c = 0; c < 3; ++c) { gdist += ge[c] * L.sig_color[c]; gs[c] = ge[c] * ddist.v; }
      }
    }
    // back through (i)
    float gx[3];
    for (int j = 0; j < 3; ++j) {
      float a = gdist * ddist.d[j];
      for (int i = 0; i < 3; ++i) a += gd[i] * dd[i].d[j];
      gx[j] = a;
    }
    st3(g_n, k, gn); st3(g_wi, k, gwi); st3(g_pv, k, gx);
  }
  if (L.mode == 0) {
    // per-view amplitude gradient: views are few; threads of a block mostly share one view
    for (int vw = 0; vw < n_views; ++vw)
      for (int c = 0; c < 3; ++c) block_atomic_add(g_amp + vw * 3 + c, (live && view == vw) ? ga[c] : 0.0f);
    for (int c = 0; c < 3; ++c) block_atomic_add(g_coef + c, gc[c]);
  } else {
    for (int c = 0; c < 3; ++c) block_atomic_add(g_sig_color + c, gs[c]);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// blend
// ------------------------------------------------------------------------------------------------------------------
struct BlendDev {
  int nb;
  int kind[NRT_MAX_BSDFS];   // 0 neural, 1 diffuse, 2 conductor
  int slot[NRT_MAX_BSDFS];   // index within its kind
  int neural_act;            // 0 sigmoid, 1 softplus, 2 identity
  int diffuse_pre;           // 0 identity, 1 / pi, 2 softplus, 3 sigmoid
  int n_neural, n_diffuse;
};

__device__ __forceinline__ float act_apply(int id, float x, float* deriv) {
  if (id == 0) { const float s = nrt_sigmoidf(x); *deriv = s * (1.0f - s); return s; }
  if (id == 1) { *deriv = x > 20.0f ? 1.0f : nrt_sigmoidf(x); return nrt_softplusf(x); }
  *deriv = 1.0f;
  return x;
}
__device__ __forceinline__ float pre_apply(int id, float x, float* deriv) {
  if (id == 1) { *deriv = 1.0f / 3.14159265358979323846f; return x / 3.14159265358979323846f; }
  if (id == 2) { *deriv = x > 20.0f ? 1.0f : nrt_sigmoidf(x); return nrt_softplusf(x); }
  if (id == 3) { const float s = nrt_sigmoidf(x); *deriv = s * (1.0f - s); return s; }
  *deriv = 1.0f;
  return x;
}

// spectrum of child b (and, if BWD, what is needed to differentiate it) for one hit
template <bool BWD>
__global__ void __launch_bounds__(kThreads)
k_shade_blend(BlendDev B, const float* __restrict__ logits, const float* __restrict__ neural_raw, const float* __restrict__ wi_in,
              const float* __restrict__ wo_in, const float* __restrict__ e_in, const float* __restrict__ refl,
              const float* __restrict__ cond_spec, const float* __restrict__ cond_eta, float inv_samples, int64_t K,
              float* __restrict__ out,
              // backward only
              const float* __restrict__ g_out, float* __restrict__ g_logits, float* __restrict__ g_neural, float* __restrict__ g_wi,
              float* __restrict__ g_wo, float* __restrict__ g_e, float* __restrict__ g_refl, float* __restrict__ g_cond_spec,
              float* __restrict__ g_cond_eta) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = k < K;
  float g_rf[NRT_MAX_BSDFS][3];          // per-thread parameter gradients (diffuse children), reduced at the end
  float g_cs[3] = {0, 0, 0}, g_ce = 0.0f;
  if (BWD) {
    for (int b = 0; b < NRT_MAX_BSDFS; ++b) g_rf[b][0] = g_rf[b][1] = g_rf[b][2] = 0.0f;
  }
  if (live) {
    float wi[3], wo[3], e[3];
    ld3(wi_in, k, wi); ld3(wo_in, k, wo); ld3(e_in, k, e);
    float r[3] = {0, 0, 0};
    float go[3] = {0, 0, 0}, gr[3] = {0, 0, 0}, gwi[3] = {0, 0, 0}, gwo[3] = {0, 0, 0};
    if (BWD) {
      ld3(g_out, k, go);
      for (int c = 0; c < 3; ++c) gr[c] = go[c] * e[c] * inv_samples;
    }
    for (int b = 0; b < B.nb; ++b) {
      const float lg = logits[k * B.nb + b];
      const float kb = nrt_sigmoidf(lg);                         // bsdfs.py:534-536: sigmoid, not softmax
      float s[3] = {0, 0, 0}, ds[3] = {0, 0, 0};
      const int kind = B.kind[b], slot = B.slot[b];
      float fres = 0.0f;
      bool lobe = false;
      if (kind == 0) {
        const float* raw = neural_raw + ((int64_t)slot * K + k) * 3;
        for (int c = 0; c < 3; ++c) s[c] = act_apply(B.neural_act, raw[c], &ds[c]);
      } else if (kind == 1) {
        for (int c = 0; c < 3; ++c) s[c] = pre_apply(B.diffuse_pre, wo[2] * refl[slot * 3 + c], &ds[c]);   // bsdfs.py:116
      } else {
        // bsdfs.py:364-388: mirror lobe gated by dot(reflect(wi), wo) > 0.94, Fresnel(cos theta_i, softplus(eta), 0)
        lobe = (-wi[0] * wo[0] - wi[1] * wo[1] + wi[2] * wo[2]) > 0.94f;
        if (lobe) {
          fres = fresnel_conductor(wi[2], cond_eta[0], 0.0f);
          for (int c = 0; c < 3; ++c) s[c] = fres * cond_spec[c];
        }
      }
      for (int c = 0; c < 3; ++c) r[c] += kb * s[c];
      if (BWD) {
        float gk = 0.0f;
        for (int c = 0; c < 3; ++c) gk += gr[c] * s[c];
        g_logits[k * B.nb + b] = gk * kb * (1.0f - kb);
        if (kind == 0) {
          float* gn = g_neural + ((int64_t)slot * K + k) * 3;
          for (int c = 0; c < 3; ++c) gn[c] = gr[c] * kb * ds[c];
        } else if (kind == 1) {
          for (int c = 0; c < 3; ++c) {
            const float gz = gr[c] * kb * ds[c];                // gradient w.r.t. wo_z * refl_c
            g_rf[slot][c] += gz * wo[2];
            gwo[2] += gz * refl[slot * 3 + c];
          }
        } else if (lobe) {
          Dual<2> ce[2] = {dvar<2>(wi[2], 0), dvar<2>(cond_eta[0], 1)};
          const Dual<2> f = fresnel_conductor(ce[0], ce[1], 0.0f);
          float gf = 0.0f;
          for (int c = 0; c < 3; ++c) { gf += gr[c] * kb * cond_spec[c]; g_cs[c] += gr[c] * kb * fres; }
          gwi[2] += gf * f.d[0];
          g_ce += gf * f.d[1];
        }
      }
    }
    if (!BWD) {
      const float o[3] = {r[0] * e[0] * inv_samples, r[1] * e[1] * inv_samples, r[2] * e[2] * inv_samples};   // integrators.py:183-187
      st3(out, k, o);
    } else {
      const float ge[3] = {go[0] * r[0] * inv_samples, go[1] * r[1] * inv_samples, go[2] * r[2] * inv_samples};
      st3(g_e, k, ge); st3(g_wi, k, gwi); st3(g_wo, k, gwo);
    }
  }
  if (BWD) {
    for (int b = 0; b < B.nb; ++b)
      if (B.kind[b] == 1)
        for (int c = 0; c < 3; ++c) block_atomic_add(g_refl + B.slot[b] * 3 + c, g_rf[B.slot[b]][c]);
    bool has_cond = false;
    for (int b = 0; b < B.nb; ++b) has_cond |= B.kind[b] == 2;
    if (has_cond) {
      for (int c = 0; c < 3; ++c) block_atomic_add(g_cond_spec + c, g_cs[c]);
      block_atomic_add(g_cond_eta, g_ce);
    }
  }
}

int build_light(const nrt_light_t* l, LightDev* d) {
  NRT_REQUIRE(l != nullptr, "light descriptor is NULL");
  NRT_REQUIRE(l->mode == NRT_LIGHT_POINT || l->mode == NRT_LIGHT_FIELD, "light.mode %d unknown", l->mode);
  if (l->mode == NRT_LIGHT_POINT) NRT_REQUIRE(l->location && l->amp && l->coef && l->n_views >= 1, "point light: location / amp / coef are NULL");
  else NRT_REQUIRE(l->v && l->sig_color, "light field: v / sig_color are NULL");
  d->mode = l->mode; d->location = l->location; d->amp = l->amp; d->coef = l->coef; d->view_of_hit = l->view_of_hit;
  d->v = l->v; d->sig_color = l->sig_color;
  return NRT_OK;
}
int build_blend(const nrt_blend_t* b, BlendDev* d) {
  NRT_REQUIRE(b != nullptr, "blend descriptor is NULL");
  NRT_REQUIRE(b->nb >= 1 && b->nb <= NRT_MAX_BSDFS, "blend.nb %d out of range [1,%d]", b->nb, NRT_MAX_BSDFS);
  d->nb = b->nb; d->neural_act = b->neural_act; d->diffuse_pre = b->diffuse_pre; d->n_neural = 0; d->n_diffuse = 0;
  int n_cond = 0;
  for (int i = 0; i < b->nb; ++i) {
    NRT_REQUIRE(b->kind[i] >= 0 && b->kind[i] <= 2, "blend.kind[%d] = %d unknown", i, b->kind[i]);
    d->kind[i] = b->kind[i];
    if (b->kind[i] == NRT_BSDF_NEURAL) d->slot[i] = d->n_neural++;
    else if (b->kind[i] == NRT_BSDF_DIFFUSE) d->slot[i] = d->n_diffuse++;
    else { d->slot[i] = 0; ++n_cond; }
  }
  NRT_REQUIRE(n_cond <= 1, "at most one Conductor child is supported by the fused blend");
  return NRT_OK;
}

}  // namespace

extern "C" int nrt_shade_geom_forward(const float* raw_n, const float* p_hit, const float* rays_hit, int64_t K, float eps5,
                                      float* n, float* p_off, float* wi, float* frame, void* stream) {
  NRT_REQUIRE(K >= 0, "nrt_shade_geom_forward: negative K");
  if (K == 0) return NRT_OK;
  NRT_REQUIRE(raw_n && p_hit && rays_hit && n && p_off && wi, "nrt_shade_geom_forward: null pointer");
  NrtProfScope _ps(TAG_SHADE, (cudaStream_t)stream);
  k_shade_geom_fwd<<<nrt_cdiv(K, kThreads), kThreads, 0, (cudaStream_t)stream>>>(raw_n, p_hit, rays_hit, K, eps5, n, p_off, wi, frame);
  NRT_CUDA(cudaGetLastError());
  return NRT_OK;
}
extern "C" int nrt_shade_geom_backward(const float* raw_n, const float* rays_hit, int64_t K, float eps5, const float* g_n,
                                       const float* g_p_off, const float* g_wi, float* g_raw_n, void* stream) {
  NRT_REQUIRE(K >= 0, "nrt_shade_geom_backward: negative K");
  if (K == 0) return NRT_OK;
  NRT_REQUIRE(raw_n && rays_hit && g_raw_n, "nrt_shade_geom_backward: null pointer");
  NrtProfScope _ps(TAG_SHADE, (cudaStream_t)stream);
  k_shade_geom_bwd<<<nrt_cdiv(K, kThreads), kThreads, 0, (cudaStream_t)stream>>>(raw_n, rays_hit, K, eps5, g_n, g_p_off, g_wi, g_raw_n);
  NRT_CUDA(cudaGetLastError());
  return NRT_OK;
}

extern "C" int nrt_shade_light_forward(const nrt_light_t* light, const float* n, const float* wi, const float* p_off, int64_t K,
                                       float* d, float* dist, float* wo, float* rusin, float* e, float* elaz, void* stream) {
  LightDev L;
  int rc = build_light(light, &L);
  if (rc != NRT_OK) return rc;
  NRT_REQUIRE(K >= 0, "nrt_shade_light_forward: negative K");
  if (K == 0) return NRT_OK;
  NRT_REQUIRE(n && wi && d && dist && wo && rusin && e && (L.mode == NRT_LIGHT_FIELD || p_off), "nrt_shade_light_forward: null pointer");
  NrtProfScope _ps(TAG_SHADE, (cudaStream_t)stream);
  k_shade_light_fwd<<<nrt_cdiv(K, kThreads), kThreads, 0, (cudaStream_t)stream>>>(L, n, wi, p_off, K, d, dist, wo, rusin, e, elaz);
  NRT_CUDA(cudaGetLastError());
  return NRT_OK;
}
extern "C" int nrt_shade_light_backward(const nrt_light_t* light, const float* n, const float* wi, const float* p_off, int64_t K,
                                        const float* g_wo, const float* g_rusin, const float* g_e, const float* g_elaz,
                                        float* g_n, float* g_wi, float* g_pv, float* g_amp, float* g_coef, float* g_sig_color,
                                        void* stream) {
  LightDev L;
  int rc = build_light(light, &L);
  if (rc != NRT_OK) return rc;
  NRT_REQUIRE(K >= 0, "nrt_shade_light_backward: negative K");
  if (K == 0) return NRT_OK;
  NRT_REQUIRE(n && wi && g_n && g_wi && g_pv, "nrt_shade_light_backward: null pointer");
  NRT_REQUIRE(L.mode == NRT_LIGHT_POINT ? (g_amp && g_coef && p_off) : (g_sig_color != nullptr), "nrt_shade_light_backward: parameter gradients are NULL");
  NrtProfScope _ps(TAG_SHADE, (cudaStream_t)stream);
  k_shade_light_bwd<<<nrt_cdiv(K, kThreads), kThreads, 0, (cudaStream_t)stream>>>(L, n, wi, p_off, K, g_wo, g_rusin, g_e, g_elaz, g_n, g_wi,
                                                                                  g_pv, g_amp, g_coef, g_sig_color, light->n_views);
  NRT_CUDA(cudaGetLastError());
  return NRT_OK;
}

extern "C" int nrt_shade_blend_forward(const nrt_blend_t* cfg, const float* logits, const float* neural_raw, const float* wi,
                                       const float* wo, const float* e, const float* refl, const float* cond_spec,
                                       const float* cond_eta, float inv_samples, int64_t K, float* out, void* stream) {
  BlendDev B;
  int rc = build_blend(cfg, &B);
  if (rc != NRT_OK) return rc;
  NRT_REQUIRE(K >= 0, "nrt_shade_blend_forward: negative K");
  if (K == 0) return NRT_OK;
  NRT_REQUIRE(logits && wi && wo && e && out, "nrt_shade_blend_forward: null pointer");
  NRT_REQUIRE((B.n_neural == 0 || neural_raw) && (B.n_diffuse == 0 || refl), "nrt_shade_blend_forward: child inputs are NULL");
  for (int i = 0; i < B.nb; ++i) NRT_REQUIRE(B.kind[i] != 2 || (cond_spec && cond_eta), "nrt_shade_blend_forward: conductor parameters are NULL");
  NrtProfScope _ps(TAG_SHADE, (cudaStream_t)stream);
  k_shade_blend<false><<<nrt_cdiv(K, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
      B, logits, neural_raw, wi, wo, e, refl, cond_spec, cond_eta, inv_samples, K, out, nullptr, nullptr, nullptr, nullptr, nullptr,
      nullptr, nullptr, nullptr, nullptr);
  NRT_CUDA(cudaGetLastError());
  return NRT_OK;
}
extern "C" int nrt_shade_blend_backward(const nrt_blend_t* cfg, const float* logits, const float* neural_raw, const float* wi,
                                        const float* wo, const float* e, const float* refl, const float* cond_spec,
                                        const float* cond_eta, float inv_samples, int64_t K, const float* g_out, float* g_logits,
                                        float* g_neural, float* g_wi, float* g_wo, float* g_e, float* g_refl, float* g_cond_spec,
                                        float* g_cond_eta, void* stream) {
  BlendDev B;
  int rc = build_blend(cfg, &B);
  if (rc != NRT_OK) return rc;
  NRT_REQUIRE(K >= 0, "nrt_shade_blend_backward: negative K");
  if (K == 0) return NRT_OK;
  NRT_REQUIRE(logits && wi && wo && e && g_out && g_logits && g_wi && g_wo && g_e, "nrt_shade_blend_backward: null pointer");
  NRT_REQUIRE((B.n_neural == 0 || (neural_raw && g_neural)) && (B.n_diffuse == 0 || (refl && g_refl)), "nrt_shade_blend_backward: child buffers are NULL");
  for (int i = 0; i < B.nb; ++i)
    NRT_REQUIRE(B.kind[i] != 2 || (cond_spec && cond_eta && g_cond_spec && g_cond_eta), "nrt_shade_blend_backward: conductor buffers are NULL");
  NrtProfScope _ps(TAG_SHADE, (cudaStream_t)stream);
  k_shade_blend<true><<<nrt_cdiv(K, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
      B, logits, neural_raw, wi, wo, e, refl, cond_spec, cond_eta, inv_samples, K, nullptr, g_out, g_logits, g_neural, g_wi, g_wo, g_e,
      g_refl, g_cond_spec, g_cond_eta);
  NRT_CUDA(cudaGetLastError());
  return NRT_OK;
}
