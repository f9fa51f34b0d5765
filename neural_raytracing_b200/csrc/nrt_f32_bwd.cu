// Fused fp32 backward of SkipConnMLP (reverse mode of neural_blocks.py:75-86 + utils.py:37-40).
//
// One 256-thread CTA owns a tile of TM samples and walks the layers back to front with everything in shared
// memory: the post-activation hidden states saved by the forward (acts, [layer][k][M]) are re-loaded per
// layer, the encoding is recomputed from x, and per layer
//   wgrad  dW[k][n] += sum_m in[k][m] * gz[n][m]      (register tile 4k x 8n, reduced with red.global.add)
//   dgrad  gin[k][m]  = sum_n W[k][n] * gz[n][m]      (same tiled GEMM as the forward, fed by nn.Linear's
//                                                       native [N][K] weight layout so rows stream over n)
// Activation derivatives are taken from the activation OUTPUT (leaky: sign; softplus: 1 - exp(-a)), so the
// forward only has to save post-activation values.
#include <algorithm>

#include "mlp_tile_f32.cuh"

namespace nrt {

__device__ __forceinline__ float out_act_grad_from_out(int out_act, float o) {
  switch (out_act) {
    case NRT_OUT_SIGMOID: return o * (1.0f - o);
    case NRT_OUT_SOFTPLUS: return 1.0f - nrt_expf(-o);
    case NRT_OUT_TANH: return 1.0f - o * o;
    default: return 1.0f;
  }
}

struct BwdArgs {
  const float* x; const float* latent; const float* out; const float* acts; const float* g_out;
  const float* w_nk;     // per layer W [N][K] (torch layout), same layer order as the packed blob, no biases
  float* g_params;       // packed-f32 layout (W^T [K][N] + bias per layer), accumulated into
  float* g_x; float* g_latent;
  int64_t M; int out_act;
  int wnk_off[NRT_NLIN];   // float offset of layer li inside w_nk
};

// rows are padded to TM + 4 floats and the n-rows a thread owns are interleaved (n = tn + NG*j) so that the
// LDS.128 of a quarter warp hit 32 distinct banks
template <int TM> struct Pad { static constexpr int S = TM + 4; };

// out[k][m] (k < KO) = sum_n W_nk[n][k0 + k] * g[n][m], n < N; generic (slow) path for the encoding rows
template <int TM>
__device__ void dgrad_generic(const float* __restrict__ Wnk, int Kfull, int k0, int KO, int N,
                              const float* __restrict__ g, float* __restrict__ outp, bool accumulate_scaled,
                              const float* __restrict__ scale_from_out, int act) {
  constexpr int S = Pad<TM>::S;
  for (int idx = threadIdx.x; idx < KO * TM; idx += kThreads) {
    const int k = idx / TM, mm = idx - k * TM;
    float acc = 0.0f;
    for (int n = 0; n < N; ++n) acc = nrt_fma(__ldg(Wnk + (size_t)n * Kfull + k0 + k), g[n * S + mm], acc);
    if (accumulate_scaled) outp[k * S + mm] += acc * act_grad_from_out(act, scale_from_out[k * S + mm]);
    else outp[k * S + mm] += acc;
  }
}

// Tiled dgrad for the H hidden rows: gin[k][m] = sum_n W_nk[n][k] g[n][m], k < H, n < N (= H).
// Weight rows n stream through wbuf in chunks of kKC (each row holds the first H of Kfull entries).
template <int H, int TM>
__device__ void dgrad_hidden(const float* __restrict__ Wnk, int Kfull, const float* __restrict__ g,
                             float* __restrict__ outp, float* __restrict__ wbuf) {
  using C = TileCfg<H, TM>;
  constexpr int S = Pad<TM>::S;
  const int tid = threadIdx.x;
  const int tn = tid % C::NG, tmg = tid / C::NG;
  const int m0 = tmg * C::RM;
  float acc[C::RM][C::RN];
#pragma unroll
  for (int r = 0; r < C::RM; ++r)
#pragma unroll
    for (int j = 0; j < C::RN; ++j) acc[r][j] = 0.0f;
  constexpr int nchunks = H / kKC;
  auto prefetch = [&](int c) {
    float* dst = wbuf + (c & 1) * (kKC * H);
    if ((Kfull & 3) == 0) {
      for (int i = tid; i < kKC * H / 4; i += kThreads) {
        const int row = i / (H / 4), col = (i - row * (H / 4)) * 4;
        cp_async16(dst + row * H + col, Wnk + (size_t)(c * kKC + row) * Kfull + col);
      }
    } else {   // rows of the torch-layout weight are not 16-byte aligned (odd fan-in): plain loads
      for (int i = tid; i < kKC * H; i += kThreads) {
        const int row = i / H, col = i - row * H;
        dst[i] = __ldg(Wnk + (size_t)(c * kKC + row) * Kfull + col);
      }
    }
    cp_async_commit();
  };
  prefetch(0);
  for (int c = 0; c < nchunks; ++c) {
    if (c + 1 < nchunks) { prefetch(c + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    const float* wb = wbuf + (c & 1) * (kKC * H);
#pragma unroll 4
    for (int nn = 0; nn < kKC; ++nn) {
      const float4 a4 = *reinterpret_cast<const float4*>(g + (c * kKC + nn) * S + m0);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      float w[C::RN];
#pragma unroll
      for (int v = 0; v < C::NV; ++v) {
        const float* wp = wb + nn * H + v * (C::NG * C::VEC) + tn * C::VEC;
        if (C::VEC == 4) {
          const float4 t = *reinterpret_cast<const float4*>(wp);
          w[v * 4] = t.x; w[v * 4 + 1] = t.y; w[v * 4 + 2] = t.z; w[v * 4 + 3] = t.w;
        } else {
          const float2 t = *reinterpret_cast<const float2*>(wp);
          w[v * 2] = t.x; w[v * 2 + 1] = t.y;
        }
      }
#pragma unroll
      for (int r = 0; r < C::RM; ++r)
#pragma unroll
        for (int j = 0; j < C::RN; ++j) acc[r][j] = nrt_fma(a[r], w[j], acc[r][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int v = 0; v < C::NV; ++v)
#pragma unroll
    for (int e = 0; e < C::VEC; ++e) {
      const int k = v * (C::NG * C::VEC) + tn * C::VEC + e;
      const int j = v * C::VEC + e;
      *reinterpret_cast<float4*>(outp + k * S + m0) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
    }
  __syncthreads();
}

// wgrad: gW[k][n] += sum_m in[k][m] g[n][m] for k < K (rows from in0 then in1), n < N; gb[n] += sum_m g[n][m].
// Register tile 4 (k) x 8 (n, interleaved); partial sums go to global memory with red.add.
template <int TM, bool JAC = false>
__device__ void wgrad_tile(const float* __restrict__ in0, int K0, const float* __restrict__ in1, int K1,
                           const float* __restrict__ g, int N, float* __restrict__ gW, float* __restrict__ gb) {
  constexpr int S = Pad<TM>::S;
  const int K = K0 + K1;
  // n interleave: group gi covers n = (gi % NGI) + NGI * j + (gi / NGI) * 8 * NGI, j = 0..7 (blocks of 8*NGI columns;
  // the last block may be partial, its out-of-range slots are masked)
  const int NGI = N >= 128 ? 16 : (N >= 64 ? 8 : (N >= 32 ? 4 : (N >= 16 ? 2 : 1)));
  const int n_groups = ((N + 8 * NGI - 1) / (8 * NGI)) * NGI;
  const int k_groups = (K + 3) / 4;
  for (int t = threadIdx.x; t < n_groups * k_groups; t += kThreads) {
    const int gi = t % n_groups, kg = t / n_groups;
    const int nb = (gi % NGI) + (gi / NGI) * 8 * NGI;
    int nidx[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) nidx[j] = nb + NGI * j;
    const int k0 = kg * 4;
    const float* arow[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int k = min(k0 + r, K - 1);
      arow[r] = (k < K0) ? (in0 + k * S) : (in1 + (k - K0) * S);
    }
    float acc[4][8];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[r][j] = 0.0f;
    for (int mm = 0; mm < TM; mm += 4) {
      float4 a[4], b[8];
#pragma unroll
      for (int r = 0; r < 4; ++r) a[r] = *reinterpret_cast<const float4*>(arow[r] + mm);
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = *reinterpret_cast<const float4*>(g + min(nidx[j], N - 1) * S + mm);
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[r][j] = nrt_fma(a[r].x, b[j].x, acc[r][j]);
          acc[r][j] = nrt_fma(a[r].y, b[j].y, acc[r][j]);
          acc[r][j] = nrt_fma(a[r].z, b[j].z, acc[r][j]);
          acc[r][j] = nrt_fma(a[r].w, b[j].w, acc[r][j]);
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (k0 + r < K && nidx[j] < N) atomicAdd(gW + (size_t)(k0 + r) * N + nidx[j], acc[r][j]);
  }
  for (int n = threadIdx.x; n < N; n += kThreads) {
    float s = 0.0f;
    for (int mm = 0; mm < TM; mm += (JAC ? 4 : 1)) s += g[n * S + mm];   // JAC: tangent columns see no bias
    atomicAdd(gb + n, s);
  }
}

// GX = false: no input / latent gradient is wanted, so the encoding-gradient buffer and every data-gradient product
// into the encoding disappear; the shared memory that frees lets the 256-wide nets run 32-sample tiles instead of 16
// (half the weight-gradient red.adds and half the weight re-streaming per sample; measured 3 % on a DTU-style step).
template <int H, int TM, bool GX = true>
__global__ void __launch_bounds__(kThreads, 1)
k_mlp_bwd(MlpDev m, BwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  constexpr int S = Pad<TM>::S;
  const int DP = m.dim_p, OUT = m.out;
  // region R1 is used twice: first by the forward-style (unpadded) buffers of the encoding recompute, then by
  // the three padded [H][S] hidden-state buffers
  TileSmem ts;
  float* p = smem;
  const int r1 = max(2 * DP * TM, 3 * H * S);
  ts.enc_raw = p;
  ts.enc_act = p + DP * TM;
  float* hin = p;                   // post-activation input of the current layer
  float* gz = p + H * S;            // gradient w.r.t. the current layer's pre-activation output
  float* gin = p + 2 * H * S;       // dgrad result (gradient w.r.t. hin)
  p += r1;
  ts.h0 = ts.h1 = nullptr; ts.outb = nullptr;
  ts.wbuf = p; p += 2 * kKC * H;
  float* encr = p; p += DP * S;     // raw encoding (padded copy)
  float* enca = p; p += DP * S;     // activated encoding
  float* genc = p; if (GX) p += DP * S;     // gradient w.r.t. the raw encoding (GX only)
  float* go = p; p += OUT * S;      // gradient w.r.t. the pre-output-activation result
  const int tid = threadIdx.x;
  const int lat0 = m.in_size + 2 * m.freqs;
  const int64_t ntiles = (a.M + TM - 1) / TM;

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t mb = tile * TM;
    const int valid = (int)min((int64_t)TM, a.M - mb);
    // ---- tiles whose incoming gradient is exactly zero contribute nothing: skip them ----
    // (the shading masks its MLP outputs with where(active, ., 0), bsdfs.py:521-525: for a DTU-style crop 60-80 % of
    //  the rays are misses, spatially coherent, and their g_out is exactly 0)
    {
      int nz = 0;
      for (int idx = tid; idx < OUT * valid; idx += kThreads) nz |= (a.g_out[mb * OUT + idx] != 0.0f) ? 1 : 0;
      if (!__syncthreads_or(nz)) {
        if (a.g_x != nullptr)
          for (int idx = tid; idx < valid * m.in_size; idx += kThreads) a.g_x[mb * m.in_size + idx] = 0.0f;
        if (a.g_latent != nullptr)
          for (int idx = tid; idx < valid * m.latent; idx += kThreads) a.g_latent[mb * m.latent + idx] = 0.0f;
        continue;
      }
    }
    // ---- recompute the encoding ----
    for (int idx = tid; idx < TM * m.in_size; idx += kThreads) {
      const int mm = idx / m.in_size, j = idx - mm * m.in_size;
      ts.enc_raw[j * TM + mm] = (mm < valid) ? a.x[(mb + mm) * m.in_size + j] : 0.0f;
    }
    for (int idx = tid; idx < TM * m.latent; idx += kThreads) {
      const int mm = idx / m.latent, j = idx - mm * m.latent;
      ts.enc_raw[(lat0 + j) * TM + mm] = (mm < valid) ? a.latent[(mb + mm) * m.latent + j] : 0.0f;
    }
    __syncthreads();
    encode_tile<TM>(m, ts);
    for (int idx = tid; idx < DP * TM; idx += kThreads) {
      const int k = idx / TM, mm = idx - k * TM;
      encr[k * S + mm] = ts.enc_raw[idx];
      enca[k * S + mm] = ts.enc_act[idx];
      if (GX) genc[k * S + mm] = 0.0f;
    }
    __syncthreads();     // the unpadded encoding buffers alias hin / gz / gin
    // ---- gradient at the output (through the output activation) ----
    for (int idx = tid; idx < OUT * TM; idx += kThreads) {
      const int mm = idx / OUT, n = idx - mm * OUT;
      float g = 0.0f;
      if (mm < valid) g = a.g_out[(mb + mm) * OUT + n] * out_act_grad_from_out(a.out_act, a.out[(mb + mm) * OUT + n]);
      go[n * S + mm] = g;
    }
    auto load_acts = [&](int l) {   // hin <- post-activation hidden state l (0 = init output)
      for (int idx = tid; idx < H * TM; idx += kThreads) {
        const int k = idx / TM, mm = idx - k * TM;
        hin[k * S + mm] = (mm < valid) ? a.acts[((size_t)l * H + k) * a.M + mb + mm] : 0.0f;
      }
    };
    load_acts(m.L);
    __syncthreads();
    // ---- out layer ----
    {
      const int li = m.n_lin - 1;
      wgrad_tile<TM>(hin, H, nullptr, 0, go, OUT, a.g_params + m.w_off[li], a.g_params + m.b_off[li]);
    }
    // dgrad of the out layer: gz[k][m] = act'(hin) * sum_n W_out[k][n] go[n][m]  (W_out^T [H][OUT] is k-major in params)
    {
      const int li = m.n_lin - 1;
      const float* Wkn = m.params + m.w_off[li];
      for (int idx = tid; idx < H * TM; idx += kThreads) {
        const int k = idx / TM, mm = idx - k * TM;
        float acc = 0.0f;
        for (int n = 0; n < OUT; ++n) acc = nrt_fma(__ldg(Wkn + k * OUT + n), go[n * S + mm], acc);
        gz[k * S + mm] = acc * act_grad_from_out(m.act, hin[k * S + mm]);
      }
    }
    __syncthreads();
    // ---- hidden layers, back to front ----
    for (int i = m.L - 1; i >= 0; --i) {
      const int li = 1 + i;
      const bool sk = (m.skip_mask >> i) & 1u;
      const int Kfull = H + (sk ? DP : 0);
      load_acts(i);              // input of layer i = post-activation hidden state i
      __syncthreads();
      wgrad_tile<TM>(hin, H, enca, sk ? DP : 0, gz, H, a.g_params + m.w_off[li], a.g_params + m.b_off[li]);
      __syncthreads();
      const float* Wl = a.w_nk + a.wnk_off[li];      // nn.Linear layout [H][Kfull] of this layer
      dgrad_hidden<H, TM>(Wl, Kfull, gz, gin, ts.wbuf);
      if (GX && sk) dgrad_generic<TM>(Wl, Kfull, H, DP, H, gz, genc, true, enca, m.act);
      __syncthreads();
      // gz <- gin * act'(hin)   (hin = act(z_i))
      for (int idx = tid; idx < H * TM; idx += kThreads) {
        const int k = idx / TM, mm = idx - k * TM;
        gz[k * S + mm] = gin[k * S + mm] * act_grad_from_out(m.act, hin[k * S + mm]);
      }
      __syncthreads();
    }
    // ---- init layer: input = raw encoding ----
    wgrad_tile<TM>(encr, DP, nullptr, 0, gz, H, a.g_params + m.w_off[0], a.g_params + m.b_off[0]);
    if (GX && (a.g_x != nullptr || a.g_latent != nullptr)) {
      dgrad_generic<TM>(a.w_nk, DP, 0, DP, H, gz, genc, false, nullptr, m.act);
      __syncthreads();
      // d enc / d x: [x, sin(xB), cos(xB)]
      if (a.g_x != nullptr) {
        const int I = m.in_size, F = m.freqs;
        for (int idx = tid; idx < I * TM; idx += kThreads) {
          const int j = idx / TM, mm = idx - j * TM;
          float acc = genc[j * S + mm];
          for (int f = 0; f < F; ++f) {
            const float b = __ldg(m.basis + j * F + f);
            acc += b * (encr[(I + F + f) * S + mm] * genc[(I + f) * S + mm] - encr[(I + f) * S + mm] * genc[(I + F + f) * S + mm]);
          }
          if (mm < valid) a.g_x[(mb + mm) * I + j] = acc;
        }
      }
      if (a.g_latent != nullptr) {
        for (int idx = tid; idx < m.latent * TM; idx += kThreads) {
          const int j = idx / TM, mm = idx - j * TM;
          if (mm < valid) a.g_latent[(mb + mm) * m.latent + j] = genc[(lat0 + j) * S + mm];
        }
      }
    }
    __syncthreads();
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Reverse pass of the forward-mode (value, Jacobian) evaluation of k_mlp_value_jac (nrt_sdf_grad.cu): what
// loss.backward() does through SDF.autograd_diff's create_graph normals in the reference (sdfs.py:184-197 feeding
// eikonal_loss utils.py:294 and the shading frame), i.e. a double backward, as ONE first-order pass over the
// four-column network.  Every point owns four tile columns [value, d/dp0, d/dp1, d/dp2]:
//   z_c = W a_c (+ b for the value column),   a_v = act(z_v),   a_t = act'(z_v) * z_t
// so the Linear layers are ordinary dgrad / wgrad over 4M columns (bias gradient from the value columns only) and
// the activation couples the columns of a point:
//   g_z_t = act'(z_v) g_a_t,    g_z_v = act'(z_v) g_a_v + act''(z_v) sum_t g_a_t z_t.
// With s = act'(z_v) taken from the saved OUTPUT a_v (softplus: 1 - exp(-a_v); leaky: sign) and z_t = a_t / s the
// second term is (1 - s) sum_t g_a_t a_t for softplus (act'' = s (1 - s)) and 0 for leaky_relu.
// The Fourier basis is not trained and p carries no gradient in the reference (the march is no_grad), so nothing
// flows into the encoding.
struct JacBwdArgs {
  const float* p; const float* acts; const float* g_value; const float* g_jac;
  const float* w_nk; float* g_params; int64_t M;
  int wnk_off[NRT_NLIN];
};

template <int TM>
__device__ __forceinline__ void jac_act_backward(int act, const float* __restrict__ gin, const float* __restrict__ hin,
                                                 float* __restrict__ gz, int H) {
  constexpr int S = Pad<TM>::S;
  for (int idx = threadIdx.x; idx < H * (TM / 4); idx += kThreads) {
    const int k = idx / (TM / 4), q = idx - k * (TM / 4);
    const float4 a = *reinterpret_cast<const float4*>(hin + k * S + 4 * q);
    const float4 g = *reinterpret_cast<const float4*>(gin + k * S + 4 * q);
    const float s = act_grad_from_out(act, a.x);
    float gv = s * g.x;
    if (act == NRT_ACT_SOFTPLUS) gv += (1.0f - s) * (g.y * a.y + g.z * a.z + g.w * a.w);
    *reinterpret_cast<float4*>(gz + k * S + 4 * q) = make_float4(gv, s * g.y, s * g.z, s * g.w);
  }
}

template <int H, int TM>
__global__ void __launch_bounds__(kThreads, 1)
k_mlp_jac_bwd(MlpDev m, JacBwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  constexpr int S = Pad<TM>::S;
  constexpr int PTS = TM / 4;
  const int DP = m.dim_p, OUT = m.out;
  TileSmem ts;
  float* p = smem;
  const int r1 = max(2 * DP * TM, 3 * H * S);
  ts.enc_raw = p;
  ts.enc_act = p + DP * TM;
  float* hin = p;
  float* gz = p + H * S;
  float* gin = p + 2 * H * S;
  p += r1;
  ts.h0 = ts.h1 = nullptr; ts.outb = nullptr;
  ts.wbuf = p; p += 2 * kKC * H;
  float* encr = p; p += DP * S;
  float* enca = p; p += DP * S;
  float* go = p; p += OUT * S;
  const int tid = threadIdx.x;
  const int64_t Mc = a.M * 4;                       // columns of the saved activations
  const int64_t ntiles = (a.M + PTS - 1) / PTS;

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t pb = tile * PTS;
    const int valid = (int)min((int64_t)PTS, a.M - pb) * 4;     // valid columns
    // ---- recompute the four-column encoding (value + tangents) ----
    for (int idx = tid; idx < PTS * 3; idx += kThreads) {
      const int q = idx / 3, j = idx - q * 3;
      ts.enc_raw[j * TM + 4 * q] = (pb + q < a.M) ? a.p[(pb + q) * 3 + j] : 0.0f;
    }
    __syncthreads();
    encode_tile<TM, true>(m, ts);
    for (int idx = tid; idx < DP * TM; idx += kThreads) {
      const int k = idx / TM, mm = idx - k * TM;
      encr[k * S + mm] = ts.enc_raw[idx];
      enca[k * S + mm] = ts.enc_act[idx];
    }
    __syncthreads();     // the unpadded encoding buffers alias hin / gz / gin
    for (int idx = tid; idx < OUT * TM; idx += kThreads) {
      const int n = idx / TM, mm = idx - n * TM;
      const int64_t pt = pb + (mm >> 2);
      const int c = mm & 3;
      float g = 0.0f;
      if (mm < valid) g = (c == 0) ? a.g_value[pt * OUT + n] : a.g_jac[(pt * OUT + n) * 3 + c - 1];
      go[n * S + mm] = g;
    }
    auto load_acts = [&](int l) {
      for (int idx = tid; idx < H * TM; idx += kThreads) {
        const int k = idx / TM, mm = idx - k * TM;
        hin[k * S + mm] = (mm < valid) ? a.acts[((size_t)l * H + k) * Mc + pb * 4 + mm] : 0.0f;
      }
    };
    load_acts(m.L);
    __syncthreads();
    {
      const int li = m.n_lin - 1;
      wgrad_tile<TM, true>(hin, H, nullptr, 0, go, OUT, a.g_params + m.w_off[li], a.g_params + m.b_off[li]);
      const float* Wkn = m.params + m.w_off[li];
      for (int idx = tid; idx < H * TM; idx += kThreads) {
        const int k = idx / TM, mm = idx - k * TM;
        float acc = 0.0f;
        for (int n = 0; n < OUT; ++n) acc = nrt_fma(__ldg(Wkn + k * OUT + n), go[n * S + mm], acc);
        gin[k * S + mm] = acc;
      }
    }
    __syncthreads();
    jac_act_backward<TM>(m.act, gin, hin, gz, H);
    __syncthreads();
    for (int i = m.L - 1; i >= 0; --i) {
      const int li = 1 + i;
      const bool sk = (m.skip_mask >> i) & 1u;
      const int Kfull = H + (sk ? DP : 0);
      load_acts(i);
      __syncthreads();
      wgrad_tile<TM, true>(hin, H, enca, sk ? DP : 0, gz, H, a.g_params + m.w_off[li], a.g_params + m.b_off[li]);
      __syncthreads();
      dgrad_hidden<H, TM>(a.w_nk + a.wnk_off[li], Kfull, gz, gin, ts.wbuf);
      jac_act_backward<TM>(m.act, gin, hin, gz, H);
      __syncthreads();
    }
    wgrad_tile<TM, true>(encr, DP, nullptr, 0, gz, H, a.g_params + m.w_off[0], a.g_params + m.b_off[0]);
    __syncthreads();
  }
}

}  // namespace nrt
using namespace nrt;

extern "C" int nrt_mlp_backward(const nrt_mlp_t* mm, int out_act, const float* x, const float* latent, int64_t M,
                                const float* out, const float* acts, const float* g_out, const float* params_nk,
                                float* g_params, float* g_x, float* g_latent, void* stream) {
  MlpDev d;
  int rc = nrt_build_mlp_dev(mm, &d);
  if (rc != NRT_OK) return rc;
  NRT_REQUIRE(M >= 0, "nrt_mlp_backward: negative M");
  if (M == 0) return NRT_OK;
  NRT_REQUIRE(x && out && acts && g_out && params_nk && g_params, "nrt_mlp_backward: null pointer");
  NRT_REQUIRE(d.latent == 0 || latent != nullptr, "nrt_mlp_backward: latent is NULL");
  NRT_REQUIRE(((uintptr_t)params_nk & 15) == 0, "params_nk must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  BwdArgs a{x, latent, out, acts, g_out, params_nk, g_params, g_x, g_latent, M, out_act, {0}};
  {
    int off = 0;
    for (int li = 0; li < d.n_lin; ++li) { a.wnk_off[li] = off; off += d.K[li] * d.N[li]; }
  }
#define NRT_BWD_CASE(HV, TMV)                                                                                      \
  if (d.hidden == HV) {                                                                                            \
    const size_t fl = std::max<size_t>((size_t)2 * d.dim_p * TMV, (size_t)3 * HV * (TMV + 4)) + 2 * kKC * HV +     \
                      (size_t)(3 * d.dim_p + d.out) * (TMV + 4);                                                    \
    const size_t bytes = fl * sizeof(float);                                                                       \
    NRT_REQUIRE(bytes <= 227 * 1024, "nrt_mlp_backward: %zu bytes of shared memory needed", bytes);                \
    NRT_CUDA(cudaFuncSetAttribute(k_mlp_bwd<HV, TMV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));   \
    const int64_t ntiles = (M + TMV - 1) / TMV;                                                                    \
    const int grid = (int)std::min<int64_t>(ntiles, (int64_t)nrt_sm_count() * 4);                                  \
    NrtProfScope _ps(TAG_MLP_BWD_F32, st);                                                                         \
    k_mlp_bwd<HV, TMV><<<grid, kThreads, bytes, st>>>(d, a);                                                       \
    NRT_CUDA(cudaGetLastError());                                                                                  \
    return NRT_OK;                                                                                                 \
  }
  if (d.hidden == 256 && g_x == nullptr && g_latent == nullptr) {
    // no input gradient (sp_var on it.p, LightField on the hit points): 32-sample tiles
    constexpr int HV = 256, TMV = 32;
    const size_t fl = std::max<size_t>((size_t)2 * d.dim_p * TMV, (size_t)3 * HV * (TMV + 4)) + 2 * kKC * HV +
                      (size_t)(2 * d.dim_p + d.out) * (TMV + 4);
    const size_t bytes = fl * sizeof(float);
    if (bytes <= 227 * 1024) {
      NRT_CUDA(cudaFuncSetAttribute(k_mlp_bwd<HV, TMV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
      const int64_t ntiles = (M + TMV - 1) / TMV;
      const int grid = (int)std::min<int64_t>(ntiles, (int64_t)nrt_sm_count() * 4);
      NrtProfScope _ps(TAG_MLP_BWD_F32, st);
      k_mlp_bwd<HV, TMV, false><<<grid, kThreads, bytes, st>>>(d, a);
      NRT_CUDA(cudaGetLastError());
      return NRT_OK;
    }
  }
  NRT_BWD_CASE(32, 64)
  NRT_BWD_CASE(64, 64)
  NRT_BWD_CASE(96, 64)
  NRT_BWD_CASE(128, 64)
  NRT_BWD_CASE(256, 16)
#undef NRT_BWD_CASE
  nrt_set_error("nrt_mlp_backward: unsupported hidden size %d", d.hidden);
  return NRT_E_UNSUPPORTED;
}

extern "C" int nrt_mlp_value_jac_backward(const nrt_mlp_t* mm, const float* p, int64_t M, const float* acts,
                                          const float* g_value, const float* g_jac, const float* params_nk,
                                          float* g_params, void* stream) {
  MlpDev d;
  int rc = nrt_build_mlp_dev(mm, &d);
  if (rc != NRT_OK) return rc;
  NRT_REQUIRE(M >= 0, "nrt_mlp_value_jac_backward: negative M");
  NRT_REQUIRE(d.in_size == 3 && d.latent == 0, "nrt_mlp_value_jac_backward: needs in_size 3 and no latent");
  if (M == 0) return NRT_OK;
  NRT_REQUIRE(p && acts && g_value && g_jac && params_nk && g_params, "nrt_mlp_value_jac_backward: null pointer");
  NRT_REQUIRE(((uintptr_t)params_nk & 15) == 0, "params_nk must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  JacBwdArgs a{p, acts, g_value, g_jac, params_nk, g_params, M, {0}};
  {
    int off = 0;
    for (int li = 0; li < d.n_lin; ++li) { a.wnk_off[li] = off; off += d.K[li] * d.N[li]; }
  }
#define NRT_JBWD_CASE(HV, TMV)                                                                                     \
  if (d.hidden == HV) {                                                                                            \
    const size_t fl = std::max<size_t>((size_t)2 * d.dim_p * TMV, (size_t)3 * HV * (TMV + 4)) + 2 * kKC * HV +     \
                      (size_t)(2 * d.dim_p + d.out) * (TMV + 4);                                                    \
    const size_t bytes = fl * sizeof(float);                                                                       \
    NRT_REQUIRE(bytes <= 227 * 1024, "nrt_mlp_value_jac_backward: %zu bytes of shared memory needed", bytes);      \
    NRT_CUDA(cudaFuncSetAttribute(k_mlp_jac_bwd<HV, TMV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes)); \
    const int64_t ntiles = (M + TMV / 4 - 1) / (TMV / 4);                                                          \
    const int grid = (int)std::min<int64_t>(ntiles, (int64_t)nrt_sm_count() * 4);                                  \
    NrtProfScope _ps(TAG_MLP_BWD_F32, st);                                                                         \
    k_mlp_jac_bwd<HV, TMV><<<grid, kThreads, bytes, st>>>(d, a);                                                   \
    NRT_CUDA(cudaGetLastError());                                                                                  \
    return NRT_OK;                                                                                                 \
  }
  NRT_JBWD_CASE(32, 64)
  NRT_JBWD_CASE(64, 64)
  NRT_JBWD_CASE(128, 64)
#undef NRT_JBWD_CASE
  nrt_set_error("nrt_mlp_value_jac_backward: unsupported hidden size %d", d.hidden);
  return NRT_E_UNSUPPORTED;
}
