/* nrt_oracle.c -- CPU restatement of the reference's per-ray hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.  The product
 * (libnrt_b200.so) never links or calls it and has no CPU fallback.
 *
 * What it restates (reference = /root/reference/pytorch3d/pathtracer, eager PyTorch fp32):
 *   oracle_mlp_forward      neural_blocks.py:75-86 (SkipConnMLP.forward) + utils.py:37-40 (fourier2)
 *   oracle_sdf_eval         shapes/sdfs.py:37-46 (SphereSDF) + utils.py:385-387 (smooth_min)
 *   oracle_sphere_trace     shapes/sdfs.py:111-131 (SDF.intersect march loop)
 *   oracle_shadow_test      shapes/sdfs.py:162-181 (SDF.intersect_test)
 *   oracle_min_scan         shapes/sdfs.py:232-249 (SDF.throughput)
 *   oracle_sdf_value_grad   shapes/sdfs.py:184-197 (autograd_diff), written as the forward-mode
 *                           Jacobian of the same network
 *   oracle_nerfle_render    shapes/nerf.py:175-214 (NeRFLE.forward incl. compositing quirks)
 *   oracle_composite        shapes/nerf.py:206-213
 *
 * Arithmetic: IEEE binary32, every dot product accumulated as
 *   acc = bias; for k in 0..K-1: acc = fmaf(a[k], W[k][n], acc)
 * with the transcendentals of include/nrt_detmath.h, compiled with -ffp-contract=off.  That
 * fixed order is the contract the CUDA fp32 path follows, so sphere-trace hit masks can be
 * compared bit for bit.  Against the reference itself (MKL/SLEEF summation order) it agrees
 * to ~1e-5; that is pinned by tests/golden/ (generated from the unmodified reference by
 * tests/golden/make_golden.py) in tests/test_oracle_vs_golden.py.
 *
 * The loops are lock-step over all rays exactly like the reference (no early exit, no
 * compaction): the reference's cost model is part of what the CPU baseline reports.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "nrt_b200.h"
#include "nrt_detmath.h"

/* ---- minimal static-chunk parallel_for on pthreads (no OpenMP runtime in the image) ---- */
#include <pthread.h>
#include <unistd.h>
typedef void (*range_fn)(void* ctx, int64_t lo, int64_t hi);
typedef struct { range_fn fn; void* ctx; int64_t lo, hi; } pf_task_t;
static void* pf_entry(void* a) { pf_task_t* t = (pf_task_t*)a; t->fn(t->ctx, t->lo, t->hi); return NULL; }
static int g_threads = 0;
void oracle_set_threads(int n) { g_threads = n; }
int oracle_get_threads(void) {
  if (g_threads > 0) return g_threads;
  const char* e = getenv("NRT_ORACLE_THREADS");
  int n = e ? atoi(e) : (int)sysconf(_SC_NPROCESSORS_ONLN);
  if (n < 1) n = 1;
  if (n > 256) n = 256;
  return n;
}
static void parallel_for(int64_t n, range_fn fn, void* ctx) {
  int T = oracle_get_threads();
  if (n < 2 * T) T = (int)(n > 0 ? (n + 1) / 2 : 1);
  if (T <= 1) { fn(ctx, 0, n); return; }
  pthread_t th[256]; pf_task_t tk[256];
  /* interleave small blocks so that per-ray cost differences balance out */
  int64_t chunk = (n + T - 1) / T;
  for (int i = 0; i < T; ++i) {
    tk[i].fn = fn; tk[i].ctx = ctx; tk[i].lo = i * chunk; tk[i].hi = (i + 1) * chunk < n ? (i + 1) * chunk : n;
    if (tk[i].lo > n) tk[i].lo = n;
    pthread_create(&th[i], NULL, pf_entry, &tk[i]);
  }
  for (int i = 0; i < T; ++i) pthread_join(th[i], NULL);
}

#define MAXW 1024 /* max(dim_p, hidden, out) supported by the scratch arrays */

static float act_apply(int act, float x) {
  if (act == NRT_ACT_SOFTPLUS) return nrt_softplusf(x);
  return x > 0.0f ? x : x * 0.01f;
}
static float act_grad(int act, float z) { /* derivative w.r.t. the pre-activation */
  if (act == NRT_ACT_SOFTPLUS) return z > 20.0f ? 1.0f : nrt_sigmoidf(z);
  return z > 0.0f ? 1.0f : 0.01f;
}
static float out_act_apply(int oa, float x) {
  switch (oa) {
    case NRT_OUT_SIGMOID: return nrt_sigmoidf(x);
    case NRT_OUT_SOFTPLUS: return nrt_softplusf(x);
    case NRT_OUT_TANH: return nrt_tanhf(x);
    default: return x;
  }
}

typedef struct {
  int dim_p, n_lin;
  int K[NRT_MAX_LAYERS + 2], N[NRT_MAX_LAYERS + 2], w_off[NRT_MAX_LAYERS + 2], b_off[NRT_MAX_LAYERS + 2];
  int skip_layer[NRT_MAX_LAYERS];
} shape_t;

static int resolve(const nrt_mlp_t* m, shape_t* s) {
  s->dim_p = m->in_size + 2 * m->freqs + m->latent_size;
  s->n_lin = m->num_layers + 2;
  int off = 0;
  for (int li = 0; li < s->n_lin; ++li) {
    int K, N;
    if (li == 0) { K = s->dim_p; N = m->hidden; }
    else if (li == s->n_lin - 1) { K = m->hidden; N = m->out_size; }
    else {
      int i = li - 1;
      int sk = (i % m->skip) == 0 && i != m->num_layers - 1; /* neural_blocks.py:45-49, :82 */
      s->skip_layer[i] = sk;
      K = m->hidden + (sk ? s->dim_p : 0); N = m->hidden;
    }
    s->K[li] = K; s->N[li] = N;
    s->w_off[li] = off; off += K * N;
    s->b_off[li] = off; off += N;
  }
  if (s->dim_p > MAXW || m->hidden > MAXW || m->out_size > MAXW) return -1;
  return off;
}

int64_t oracle_mlp_param_count(const nrt_mlp_t* m) { shape_t s; return resolve(m, &s); }

/* y[n] = b[n] + sum_k a[k] W[k][n], k ascending, one fma per term */
static void linear(const float* W, const float* b, const float* a, int K, int N, float* y) {
  for (int n = 0; n < N; ++n) y[n] = b[n];
  for (int k = 0; k < K; ++k) {
    const float ak = a[k];
    const float* w = W + (size_t)k * N;
    for (int n = 0; n < N; ++n) y[n] = fmaf(ak, w[n], y[n]);
  }
}

/* utils.py:37-40: [x, sin(x@B), cos(x@B)] (+ latent appended, neural_blocks.py:78-79) */
static void encode(const nrt_mlp_t* m, const float* x, const float* latent, float* enc) {
  const int I = m->in_size, F = m->freqs;
  for (int j = 0; j < I; ++j) enc[j] = x[j];
  for (int f = 0; f < F; ++f) {
    float arg = 0.0f;
    for (int j = 0; j < I; ++j) arg = fmaf(x[j], m->basis[j * F + f], arg);
    nrt_sincosf(arg, &enc[I + f], &enc[I + F + f]);
  }
  for (int j = 0; j < m->latent_size; ++j) enc[I + 2 * F + j] = latent[j];
}

/* One sample through the network.  out: pre-output-activation values. */
static void mlp_one(const nrt_mlp_t* m, const shape_t* s, const float* x, const float* latent, float* out) {
  float enc[MAXW], enc_act[MAXW], h[MAXW], in[2 * MAXW], z[MAXW];
  const int H = m->hidden, DP = s->dim_p;
  encode(m, x, latent, enc);
  for (int j = 0; j < DP; ++j) enc_act[j] = act_apply(m->act, enc[j]);
  linear(m->params + s->w_off[0], m->params + s->b_off[0], enc, DP, H, z);
  for (int j = 0; j < H; ++j) h[j] = act_apply(m->act, z[j]);   /* activation(x) of the next consumer */
  for (int i = 0; i < m->num_layers; ++i) {
    const int li = 1 + i;
    memcpy(in, h, sizeof(float) * H);
    if (s->skip_layer[i]) memcpy(in + H, enc_act, sizeof(float) * DP);  /* act(cat([x, init])) */
    linear(m->params + s->w_off[li], m->params + s->b_off[li], in, s->K[li], H, z);
    for (int j = 0; j < H; ++j) h[j] = act_apply(m->act, z[j]);
  }
  const int lo = s->n_lin - 1;
  linear(m->params + s->w_off[lo], m->params + s->b_off[lo], h, H, m->out_size, out);
}

typedef struct { const nrt_mlp_t* m; const shape_t* s; int out_act; const float* x; const float* latent; float* out; } mlpfwd_ctx;
static void w_mlp_forward(void* vc, int64_t lo, int64_t hi) {
  mlpfwd_ctx* c = (mlpfwd_ctx*)vc;
  const nrt_mlp_t* m = c->m;
  for (int64_t i = lo; i < hi; ++i) {
    float o[MAXW];
    mlp_one(m, c->s, c->x + i * m->in_size, c->latent ? c->latent + i * m->latent_size : NULL, o);
    for (int n = 0; n < m->out_size; ++n) c->out[i * m->out_size + n] = out_act_apply(c->out_act, o[n]);
  }
}

int oracle_mlp_forward(const nrt_mlp_t* m, int out_act, const float* x, const float* latent, int64_t M,
                       float* out) {
  shape_t s;
  if (resolve(m, &s) < 0) return -1;
  mlpfwd_ctx c = {m, &s, out_act, x, latent, out};
  parallel_for(M, w_mlp_forward, &c);
  return 0;
}

/* ---- SphereSDF ---------------------------------------------------------------------- */
static float sphere_smin(const nrt_sphere_sdf_t* s, const float* p) {
  float sum = 0.0f;
  for (int i = 0; i < s->n; ++i) {
    const float* T = s->tfs + i * 9;
    float q[3];
    for (int j = 0; j < 3; ++j) {
      /* (tfs + I) p : einsum "ijk,ibk->ibj" (sdfs.py:38-40) */
      float t0 = T[j * 3 + 0] + (j == 0 ? 1.0f : 0.0f);
      float t1 = T[j * 3 + 1] + (j == 1 ? 1.0f : 0.0f);
      float t2 = T[j * 3 + 2] + (j == 2 ? 1.0f : 0.0f);
      float a = t0 * p[0];
      a = fmaf(t1, p[1], a);
      a = fmaf(t2, p[2], a);
      q[j] = a - s->centers[i * 3 + j];
    }
    float n2 = q[0] * q[0];
    n2 = fmaf(q[1], q[1], n2);
    n2 = fmaf(q[2], q[2], n2);
    float d = sqrtf(n2) - s->radii[i];
    sum = sum + nrt_expf(-32.0f * d);
  }
  sum = fmaxf(sum, 1e-4f);            /* .clamp(min=1e-4) utils.py:387 */
  return -nrt_logf(sum) / 32.0f;
}

static float sdf_one(const nrt_sphere_sdf_t* s, const shape_t* sh, const float* p) {
  float o[MAXW];
  float a = sphere_smin(s, p);
  mlp_one(&s->shift, sh, p, NULL, o);
  return a + o[0];                    /* out + self.shift(p) sdfs.py:46 */
}

/* sdf at p[i] (rays == NULL) or at rays[i].o + rays[i].d * depth[i] */
typedef struct { const nrt_sphere_sdf_t* s; const shape_t* sh; const float* p; const float* rays; const float* depth; float* out; } sdfpts_ctx;
static void w_sdf_points(void* vc, int64_t lo, int64_t hi) {
  sdfpts_ctx* c = (sdfpts_ctx*)vc;
  for (int64_t r = lo; r < hi; ++r) {
    if (c->rays) {
      const float* o = c->rays + r * 6; const float* d = o + 3;
      float p[3];
      for (int j = 0; j < 3; ++j) p[j] = o[j] + d[j] * c->depth[r];   /* r_o + r_d * depths */
      c->out[r] = sdf_one(c->s, c->sh, p);
    } else {
      c->out[r] = sdf_one(c->s, c->sh, c->p + r * 3);
    }
  }
}

int oracle_sdf_eval(const nrt_sphere_sdf_t* s, const float* p, int64_t M, float* out) {
  shape_t sh;
  if (resolve(&s->shift, &sh) < 0) return -1;
  sdfpts_ctx c = {s, &sh, p, NULL, NULL, out};
  parallel_for(M, w_sdf_points, &c);
  return 0;
}

/* ---- SDF.intersect march loop (sdfs.py:111-131), lock-step over all rays -------------- */
int oracle_sphere_trace(const nrt_sphere_sdf_t* s, const float* rays, int64_t R, float eps, int max_steps,
                        float max_t, float* depth, uint8_t* hit, unsigned long long* evals) {
  shape_t sh;
  if (resolve(&s->shift, &sh) < 0) return -1;
  uint8_t* remaining = (uint8_t*)malloc(R);
  float* dists = (float*)malloc(sizeof(float) * R);
  for (int64_t r = 0; r < R; ++r) { depth[r] = 0.0f; remaining[r] = 1; hit[r] = 0; }
  for (int it = 0; it < max_steps; ++it) {
    sdfpts_ctx c = {s, &sh, NULL, rays, depth, dists};   /* evaluated for every ray, like the reference */
    parallel_for(R, w_sdf_points, &c);
    for (int64_t r = 0; r < R; ++r) {
      remaining[r] = remaining[r] && (depth[r] < max_t);
      uint8_t h = remaining[r] && (dists[r] <= eps);
      hit[r] = hit[r] || h;
      remaining[r] = remaining[r] && !h;
      if (remaining[r]) depth[r] = depth[r] + dists[r];
    }
  }
  if (evals) *evals = (unsigned long long)R * (unsigned long long)max_steps;
  free(remaining); free(dists);
  return 0;
}

/* ---- SDF.intersect_test (sdfs.py:162-181) ------------------------------------------------ */
int oracle_shadow_test(const nrt_sphere_sdf_t* s, const float* rays, const float* max_t, int64_t R, float eps,
                       int max_steps, uint8_t* not_blocked) {
  shape_t sh;
  if (resolve(&s->shift, &sh) < 0) return -1;
  uint8_t* remaining = (uint8_t*)malloc(R);
  float* depth = (float*)malloc(sizeof(float) * R);
  float* dists = (float*)malloc(sizeof(float) * R);
  const float t0 = (float)(1e2 * (double)eps);
  for (int64_t r = 0; r < R; ++r) { depth[r] = t0; remaining[r] = 1; }
  for (int it = 0; it < max_steps; ++it) {
    sdfpts_ctx c = {s, &sh, NULL, rays, depth, dists};
    parallel_for(R, w_sdf_points, &c);
    for (int64_t r = 0; r < R; ++r) {
      uint8_t h = remaining[r] && (dists[r] < eps);
      if (remaining[r]) depth[r] = depth[r] + dists[r];
      remaining[r] = remaining[r] && !h;
    }
  }
  for (int64_t r = 0; r < R; ++r) not_blocked[r] = (depth[r] >= max_t[r]) || remaining[r];
  free(remaining); free(depth); free(dists);
  return 0;
}

/* ---- SDF.throughput (sdfs.py:232-249) ---------------------------------------------------- */
typedef struct { const nrt_sphere_sdf_t* s; const shape_t* shp; const float* rays; double step; int n_steps; int32_t* best_idx; float* best_pos; float* min_val; } scan_ctx;
static void w_min_scan(void* vc, int64_t lo, int64_t hi) {
  scan_ctx* c = (scan_ctx*)vc;
  const nrt_sphere_sdf_t* s = c->s; const shape_t sh = *c->shp; const float* rays = c->rays;
  const double step = c->step; const int n_steps = c->n_steps;
  int32_t* best_idx = c->best_idx; float* best_pos = c->best_pos; float* min_val = c->min_val;
  for (int64_t r = lo; r < hi; ++r) {
    const float* o = rays + r * 6; const float* d = o + 3;
    float cur = sdf_one(s, &sh, o);
    int idx = 0;
    for (int i = 0; i < n_steps; ++i) {
      const float t = (float)(step * (double)(i + 1));   /* python float, cast when it scales d */
      float p[3];
      for (int j = 0; j < 3; ++j) p[j] = o[j] + t * d[j];
      float v = sdf_one(s, &sh, p);
      if (v < cur) idx = i + 1;
      cur = fminf(cur, v);
    }
    best_idx[r] = idx;
    if (min_val) min_val[r] = cur;
    const float tb = (float)idx * (float)step;            /* idxs(long) * step -> fp32 product */
    for (int j = 0; j < 3; ++j) best_pos[r * 3 + j] = o[j] + tb * d[j];
  }
}

int oracle_min_scan(const nrt_sphere_sdf_t* s, const float* rays, int64_t R, double step, int n_steps,
                    int32_t* best_idx, float* best_pos, float* min_val) {
  shape_t sh;
  if (resolve(&s->shift, &sh) < 0) return -1;
  scan_ctx c = {s, &sh, rays, step, n_steps, best_idx, best_pos, min_val};
  parallel_for(R, w_min_scan, &c);
  return 0;
}

/* ---- value + d/dp (forward-mode restatement of autograd_diff, sdfs.py:184-197) -------- */
static void sdf_value_grad_one(const nrt_sphere_sdf_t* s, const shape_t* sh, const float* p, float* value,
                               float* grad) {
  const nrt_mlp_t* m = &s->shift;
  const int I = 3, F = m->freqs, H = m->hidden, DP = sh->dim_p;
  /* sphere part */
  float sum = 0.0f, gs[3] = {0, 0, 0};
  for (int i = 0; i < s->n; ++i) {
    const float* T = s->tfs + i * 9;
    float A[9], q[3];
    for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) A[j * 3 + k] = T[j * 3 + k] + (j == k ? 1.0f : 0.0f);
    for (int j = 0; j < 3; ++j) {
      float a = A[j * 3] * p[0]; a = fmaf(A[j * 3 + 1], p[1], a); a = fmaf(A[j * 3 + 2], p[2], a);
      q[j] = a - s->centers[i * 3 + j];
    }
    float n2 = q[0] * q[0]; n2 = fmaf(q[1], q[1], n2); n2 = fmaf(q[2], q[2], n2);
    float nq = sqrtf(n2);
    float e = nrt_expf(-32.0f * (nq - s->radii[i]));
    sum += e;
    if (nq > 0.0f) for (int k = 0; k < 3; ++k) {
      float dn = (A[0 * 3 + k] * q[0] + A[1 * 3 + k] * q[1] + A[2 * 3 + k] * q[2]) / nq;  /* d|q|/dp_k */
      gs[k] += e * dn;
    }
  }
  float val = -nrt_logf(fmaxf(sum, 1e-4f)) / 32.0f;
  float gsm[3];
  for (int k = 0; k < 3; ++k) gsm[k] = (sum >= 1e-4f) ? gs[k] / sum : 0.0f;  /* -1/32 * (-32 e dn)/sum */
  /* MLP: primal + 3 tangents */
  float enc[MAXW], denc[3][MAXW];
  encode(m, p, NULL, enc);
  for (int c = 0; c < 3; ++c) {
    for (int j = 0; j < I; ++j) denc[c][j] = (j == c) ? 1.0f : 0.0f;
    for (int f = 0; f < F; ++f) {
      float b = m->basis[c * F + f];
      denc[c][I + f] = enc[I + F + f] * b;       /* d sin = cos * B */
      denc[c][I + F + f] = -enc[I + f] * b;      /* d cos = -sin * B */
    }
  }
  float z[MAXW], dz[3][MAXW], h[MAXW], dh[3][MAXW], in[2 * MAXW], din[3][2 * MAXW], zero[MAXW];
  memset(zero, 0, sizeof(zero));
  linear(m->params + sh->w_off[0], m->params + sh->b_off[0], enc, DP, H, z);
  for (int c = 0; c < 3; ++c) linear(m->params + sh->w_off[0], zero, denc[c], DP, H, dz[c]);
  for (int i = 0; i <= m->num_layers; ++i) {
    /* activation of the consumer of z */
    for (int j = 0; j < H; ++j) {
      float g = act_grad(m->act, z[j]);
      h[j] = act_apply(m->act, z[j]);
      for (int c = 0; c < 3; ++c) dh[c][j] = g * dz[c][j];
    }
    if (i == m->num_layers) break;
    const int li = 1 + i;
    memcpy(in, h, sizeof(float) * H);
    for (int c = 0; c < 3; ++c) memcpy(din[c], dh[c], sizeof(float) * H);
    if (sh->skip_layer[i]) {
      for (int j = 0; j < DP; ++j) {
        float g = act_grad(m->act, enc[j]);
        in[H + j] = act_apply(m->act, enc[j]);
        for (int c = 0; c < 3; ++c) din[c][H + j] = g * denc[c][j];
      }
    }
    linear(m->params + sh->w_off[li], m->params + sh->b_off[li], in, sh->K[li], H, z);
    for (int c = 0; c < 3; ++c) linear(m->params + sh->w_off[li], zero, din[c], sh->K[li], H, dz[c]);
  }
  const int lo = sh->n_lin - 1;
  float o[4], d_o[4];
  linear(m->params + sh->w_off[lo], m->params + sh->b_off[lo], h, H, 1, o);
  *value = val + o[0];
  for (int c = 0; c < 3; ++c) {
    linear(m->params + sh->w_off[lo], zero, dh[c], H, 1, d_o);
    grad[c] = gsm[c] + d_o[0];
  }
}

typedef struct { const nrt_sphere_sdf_t* s; const shape_t* sh; const float* p; float* value; float* grad; } vg_ctx;
static void w_value_grad(void* vc, int64_t lo, int64_t hi) {
  vg_ctx* c = (vg_ctx*)vc;
  for (int64_t i = lo; i < hi; ++i) sdf_value_grad_one(c->s, c->sh, c->p + i * 3, c->value + i, c->grad + i * 3);
}

int oracle_sdf_value_grad(const nrt_sphere_sdf_t* s, const float* p, int64_t M, float* value, float* grad) {
  shape_t sh;
  if (resolve(&s->shift, &sh) < 0) return -1;
  vg_ctx c = {s, &sh, p, value, grad};
  parallel_for(M, w_value_grad, &c);
  return 0;
}

/* ---- compositing (nerf.py:206-213) ---------------------------------------------------------
 * sigma_raw [S,R], rgb [S,R,3] sample-major, ts [S].  cumprod order = s ascending. */
typedef struct { const float* sigma_raw; const float* rgb; const float* ts; int S; int64_t R; float* out; } comp_ctx;
static void w_composite(void* vc, int64_t lo, int64_t hi) {
  comp_ctx* cc = (comp_ctx*)vc;
  const float* sigma_raw = cc->sigma_raw; const float* rgb = cc->rgb; const float* ts = cc->ts;
  const int S = cc->S; const int64_t R = cc->R; float* out = cc->out;
  for (int64_t r = lo; r < hi; ++r) {
    float cp = 1.0f, acc[3] = {0, 0, 0}, a0 = 0.0f, c0[3] = {0, 0, 0};
    for (int s = 0; s < S; ++s) {
      float sigma = fmaxf(sigma_raw[(int64_t)s * R + r], 0.0f);         /* F.relu */
      float alpha = 1.0f - nrt_expf(-sigma * ts[s]);                    /* absolute t (quirk) */
      const float* c = rgb + ((int64_t)s * R + r) * 3;
      if (s == 0) { a0 = alpha; c0[0] = c[0]; c0[1] = c[1]; c0[2] = c[2]; }
      else {
        /* after roll(1): weights[s] = alpha_s * cp_{s-1}; cp[-1] = 1 overrides the last one */
        float w = alpha * ((s == S - 1) ? 1.0f : cp);
        for (int j = 0; j < 3; ++j) acc[j] = acc[j] + w * c[j];
      }
      cp = cp * fmaxf(1.0f - alpha, 1e-10f);
    }
    float w0 = a0 * ((S == 1) ? 1.0f : cp);  /* sample 0 receives the rolled-around total product */
    for (int j = 0; j < 3; ++j) out[r * 3 + j] = acc[j] + w0 * c0[j];
  }
}
int oracle_composite(const float* sigma_raw, const float* rgb, const float* ts, int S, int64_t R, float* out) {
  comp_ctx c = {sigma_raw, rgb, ts, S, R, out};
  parallel_for(R, w_composite, &c);
  return 0;
}

/* ---- NeRFLE.forward (nerf.py:175-214) ------------------------------------------------------
 * rays [R,6]; ts [S] shared, or ts_per_ray [R,S]; light_code [n_views, light_dim];
 * writes rgb [R,3]; optionally per-sample sigma_raw [R,S] and rgb_s [R,S,3] (ray-major). */
typedef struct {
  const nrt_mlp_t* first; const nrt_mlp_t* second; const shape_t* s1; const shape_t* s2; const float* rays;
  const float* ts; const float* ts_per_ray; int S; const float* light_code; int light_dim;
  const int32_t* view_of_ray; int second_out_act; float* out_rgb; float* out_sigma; float* out_srgb;
} nerf_ctx;
static void w_nerfle(void* vc, int64_t lo, int64_t hi) {
  nerf_ctx* k = (nerf_ctx*)vc;
  const nrt_mlp_t* first = k->first; const nrt_mlp_t* second = k->second;
  const shape_t s1 = *k->s1, s2 = *k->s2;
  const float* rays = k->rays; const float* ts = k->ts; const float* ts_per_ray = k->ts_per_ray;
  const int S = k->S; const float* light_code = k->light_code; const int light_dim = k->light_dim;
  const int32_t* view_of_ray = k->view_of_ray; const int second_out_act = k->second_out_act;
  float* out_rgb = k->out_rgb; float* out_sigma = k->out_sigma; float* out_srgb = k->out_srgb;
  const int nlat = first->out_size - 1;
  for (int64_t r = lo; r < hi; ++r) {
    const float* o = rays + r * 6; const float* d = o + 3;
    const int view = view_of_ray ? view_of_ray[r] : 0;
    float cp = 1.0f, acc[3] = {0, 0, 0}, a0 = 0.0f, c0[3] = {0, 0, 0};
    for (int s = 0; s < S; ++s) {
      const float t = ts_per_ray ? ts_per_ray[r * S + s] : ts[s];
      float p[3], f[MAXW], x2[MAXW], c[MAXW];
      for (int j = 0; j < 3; ++j) p[j] = o[j] + t * d[j];               /* r_o + ts (x) r_d */
      mlp_one(first, &s1, p, NULL, f);
      for (int j = 0; j < nlat; ++j) x2[j] = f[1 + j];                  /* latent = first_out[..., 1:] */
      for (int j = 0; j < 3; ++j) x2[nlat + j] = d[j];
      for (int j = 0; j < light_dim; ++j) x2[nlat + 3 + j] = light_code[(int64_t)view * light_dim + j];
      mlp_one(second, &s2, x2, NULL, c);
      for (int j = 0; j < 3; ++j) c[j] = out_act_apply(second_out_act, c[j]);
      if (out_sigma) {
        out_sigma[r * S + s] = f[0];
        for (int j = 0; j < 3; ++j) out_srgb[(r * S + s) * 3 + j] = c[j];
      }
      float sigma = fmaxf(f[0], 0.0f);
      float alpha = 1.0f - nrt_expf(-sigma * t);
      if (s == 0) { a0 = alpha; c0[0] = c[0]; c0[1] = c[1]; c0[2] = c[2]; }
      else {
        float w = alpha * ((s == S - 1) ? 1.0f : cp);
        for (int j = 0; j < 3; ++j) acc[j] = acc[j] + w * c[j];
      }
      cp = cp * fmaxf(1.0f - alpha, 1e-10f);
    }
    float w0 = a0 * ((S == 1) ? 1.0f : cp);
    if (out_rgb) for (int j = 0; j < 3; ++j) out_rgb[r * 3 + j] = acc[j] + w0 * c0[j];
  }
}
int oracle_nerfle_render(const nrt_mlp_t* first, const nrt_mlp_t* second, const float* rays, int64_t R,
                         const float* ts, const float* ts_per_ray, int S, const float* light_code,
                         int light_dim, const int32_t* view_of_ray, int second_out_act, float* out_rgb,
                         float* out_sigma, float* out_srgb) {
  shape_t s1, s2;
  if (resolve(first, &s1) < 0 || resolve(second, &s2) < 0) return -1;
  if (second->in_size != first->out_size - 1 + 3 + light_dim) return -2;
  nerf_ctx c = {first, second, &s1, &s2, rays, ts, ts_per_ray, S, light_code, light_dim, view_of_ray,
                second_out_act, out_rgb, out_sigma, out_srgb};
  parallel_for(R, w_nerfle, &c);
  return 0;
}
