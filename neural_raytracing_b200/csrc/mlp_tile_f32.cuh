// fp32 "exact" fused SkipConnMLP tile evaluator (device code).
//
// One CTA of 256 threads evaluates the whole MLP for a tile of TM samples with every
// activation resident in shared memory (no HBM round trip between layers); weights are
// streamed from L2 in 16-row chunks with cp.async double buffering.  Each output is
// accumulated as  acc = bias; for k in 0..K-1: acc = fma(a[k], W[k][n], acc)  -- the fixed
// k-sequential order that oracle/c/nrt_oracle.c restates, so results are bit-identical
// to the CPU oracle (which is what the bit-exact hit-mask claim rests on).
//
// Reference semantics: pytorch3d/pathtracer/neural_blocks.py:75-86 (forward),
// utils.py:37-40 (fourier2).  Quirk kept: on skip layers the activation is applied to the
// concatenated [x | encoding] tensor (neural_blocks.py:82-84).
#pragma once
#include "nrt_common.cuh"

namespace nrt {

constexpr int kThreads = 256;
constexpr int kKC = 16;  // weight rows per streamed chunk

template <int H, int TM>
struct TileCfg {
  static constexpr int RM = 4;
  static constexpr int MG = TM / RM;
  static constexpr int NG = kThreads / MG;
  static constexpr int RN = H / NG;
  static constexpr int VEC = (RN % 4 == 0) ? 4 : 2;
  static constexpr int NV = RN / VEC;
  static_assert(MG * NG == kThreads, "tile/thread mismatch");
  static_assert(RN * NG == H && RN >= 2, "hidden size must split evenly over the n-groups");
};

// Shared-memory carve-up for one MLP evaluation context.
struct TileSmem {
  float* enc_raw;  // [dim_p][TM]   raw encoding  [x | sin | cos | latent]
  float* enc_act;  // [dim_p][TM]   act(encoding) (consumed by skip layers)
  float* h0;       // [H][TM]
  float* h1;       // [H][TM]
  float* wbuf;     // [2][kKC*H]
  float* outb;     // [out][TM]
};

__host__ __device__ inline size_t tile_smem_floats(int dim_p, int H, int out, int TM) {
  return (size_t)2 * dim_p * TM + (size_t)2 * H * TM + (size_t)2 * kKC * H + (size_t)out * TM;
}

__device__ inline float* carve_tile(TileSmem& s, float* base, int dim_p, int H, int out, int TM) {
  s.enc_raw = base; base += dim_p * TM;
  s.enc_act = base; base += dim_p * TM;
  s.h0 = base; base += H * TM;
  s.h1 = base; base += H * TM;
  s.wbuf = base; base += 2 * kKC * H;
  s.outb = base; base += out * TM;
  return base;
}

__device__ __forceinline__ float act_apply(int act, float x) {
  if (act == NRT_ACT_SOFTPLUS) return nrt_softplusf(x);
  return x > 0.0f ? x : x * 0.01f;
}
// derivative of the activation expressed through its OUTPUT a = act(z)
__device__ __forceinline__ float act_grad_from_out(int act, float a) {
  if (act == NRT_ACT_SOFTPLUS) {
    // a = log(1+e^z) => sigmoid(z) = 1 - e^{-a}; (z > 20: a = z, derivative 1 - e^-a ~ 1)
    return 1.0f - nrt_expf(-a);
  }
  return a > 0.0f ? 1.0f : 0.01f;
}
// derivative of the activation expressed through its INPUT z (torch semantics: softplus' = sigmoid,
// 1 beyond the threshold; leaky_relu' = 1 for z > 0 else the slope)
__device__ __forceinline__ float act_grad_from_in(int act, float z) {
  if (act == NRT_ACT_SOFTPLUS) return z > 20.0f ? 1.0f : nrt_sigmoidf(z);
  return z > 0.0f ? 1.0f : 0.01f;
}
__device__ __forceinline__ float out_act_apply(int out_act, float x) {
  switch (out_act) {
    case NRT_OUT_SIGMOID: return nrt_sigmoidf(x);
    case NRT_OUT_SOFTPLUS: return nrt_softplusf(x);
    case NRT_OUT_TANH: return nrt_tanhf(x);
    default: return x;
  }
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// Fourier features of the in_size raw inputs already sitting in enc_raw rows [0,in_size):
// rows [in, in+f) = sin(x.B), rows [in+f, in+2f) = cos(x.B)  (utils.py:37-40).
// Then enc_act = act(enc_raw) for all dim_p rows.  Ends with __syncthreads().
// JAC = true: forward-mode Jacobian.  Tile columns come in groups of four [value, d/dp0, d/dp1, d/dp2] of the
// same point; the caller wrote the point into the value column of rows [0,3) (tangent columns are filled here).
template <int TM, bool JAC = false>
__device__ void encode_tile(const MlpDev& m, const TileSmem& s) {
  const int tid = threadIdx.x;
  const int F = m.freqs, I = m.in_size;
  if (JAC) {
    for (int idx = tid; idx < I * TM; idx += kThreads) {
      const int j = idx / TM, mm = idx - j * TM;
      if (mm & 3) s.enc_raw[idx] = ((mm & 3) - 1 == j) ? 1.0f : 0.0f;     // d p_j / d p_c
    }
    __syncthreads();
  }
  for (int idx = tid; idx < F * TM; idx += kThreads) {
    const int f = idx / TM, mm = idx - f * TM;
    if (JAC && (mm & 3)) continue;
    float arg = 0.0f;
    for (int j = 0; j < I; ++j) arg = nrt_fma(s.enc_raw[j * TM + mm], __ldg(m.basis + j * F + f), arg);
    float sn, cs;
    nrt_sincosf(arg, &sn, &cs);
    s.enc_raw[(I + f) * TM + mm] = sn;
    s.enc_raw[(I + F + f) * TM + mm] = cs;
    if (JAC) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float b = __ldg(m.basis + c * F + f);
        s.enc_raw[(I + f) * TM + mm + 1 + c] = cs * b;        // d sin = cos * B
        s.enc_raw[(I + F + f) * TM + mm + 1 + c] = -sn * b;   // d cos = -sin * B
      }
    }
  }
  __syncthreads();
  for (int idx = tid; idx < m.dim_p * TM; idx += kThreads) {
    if (!JAC) { s.enc_act[idx] = act_apply(m.act, s.enc_raw[idx]); continue; }
    const int mm = idx % TM;
    if (mm & 3) continue;
    const float z = s.enc_raw[idx];
    const float g = act_grad_from_in(m.act, z);
    s.enc_act[idx] = act_apply(m.act, z);
    s.enc_act[idx + 1] = g * s.enc_raw[idx + 1];
    s.enc_act[idx + 2] = g * s.enc_raw[idx + 2];
    s.enc_act[idx + 3] = g * s.enc_raw[idx + 3];
  }
  __syncthreads();
}

// One Linear(K0+K1 -> H) over the tile: input rows come from in0 ([K0][TM]) followed by
// in1 ([K1][TM]); writes act(z) (or z if !apply_act) to outp [H][TM].  Ends with a barrier.
template <int H, int TM, bool JAC = false>
__device__ void hidden_layer(const MlpDev& m, int li, const float* __restrict__ in0, int K0,
                             const float* __restrict__ in1, int K1, float* __restrict__ outp,
                             float* __restrict__ wbuf, bool apply_act) {
  using C = TileCfg<H, TM>;
  const int tid = threadIdx.x;
  const int tn = tid % C::NG, tmg = tid / C::NG;
  const int m0 = tmg * C::RM;
  const float* __restrict__ Wg = m.params + m.w_off[li];
  const float* __restrict__ bg = m.params + m.b_off[li];
  const int K = K0 + K1;

  float acc[C::RM][C::RN];
#pragma unroll
  for (int v = 0; v < C::NV; ++v)
#pragma unroll
    for (int e = 0; e < C::VEC; ++e) {
      const float b = __ldg(bg + v * (C::NG * C::VEC) + tn * C::VEC + e);
#pragma unroll
      for (int r = 0; r < C::RM; ++r) acc[r][v * C::VEC + e] = (JAC && r > 0) ? 0.0f : b;   // tangents carry no bias
    }

  const int nchunks = (K + kKC - 1) / kKC;
  auto prefetch = [&](int c) {
    const int kb = c * kKC;
    const int rows = min(kKC, K - kb);
    const int n16 = rows * H / 4;  // 16-byte packets
    const float* src = Wg + (size_t)kb * H;
    float* dst = wbuf + (c & 1) * (kKC * H);
    for (int i = tid; i < n16; i += kThreads) cp_async16(dst + i * 4, src + i * 4);
    cp_async_commit();
  };
  prefetch(0);
  for (int c = 0; c < nchunks; ++c) {
    if (c + 1 < nchunks) {
      prefetch(c + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* wb = wbuf + (c & 1) * (kKC * H);
    const int kb = c * kKC;
    const int kn = min(kKC, K - kb);
#pragma unroll 4
    for (int kk = 0; kk < kn; ++kk) {
      const int k = kb + kk;
      const float* arow = (k < K0) ? (in0 + k * TM) : (in1 + (k - K0) * TM);
      const float4 a4 = *reinterpret_cast<const float4*>(arow + m0);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      float w[C::RN];
#pragma unroll
      for (int v = 0; v < C::NV; ++v) {
        const float* wp = wb + kk * H + v * (C::NG * C::VEC) + tn * C::VEC;
        if (C::VEC == 4) {
          const float4 t = *reinterpret_cast<const float4*>(wp);
          w[v * 4 + 0] = t.x; w[v * 4 + 1] = t.y; w[v * 4 + 2] = t.z; w[v * 4 + 3] = t.w;
        } else {
          const float2 t = *reinterpret_cast<const float2*>(wp);
          w[v * 2 + 0] = t.x; w[v * 2 + 1] = t.y;
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
      const int n = v * (C::NG * C::VEC) + tn * C::VEC + e;
      float4 o;
      const int j = v * C::VEC + e;
      if (apply_act && JAC) {
        // the four columns of this thread are (value, three tangents) of one point
        const float g = act_grad_from_in(m.act, acc[0][j]);
        o.x = act_apply(m.act, acc[0][j]); o.y = g * acc[1][j]; o.z = g * acc[2][j]; o.w = g * acc[3][j];
      } else if (apply_act) {
        o.x = act_apply(m.act, acc[0][j]); o.y = act_apply(m.act, acc[1][j]);
        o.z = act_apply(m.act, acc[2][j]); o.w = act_apply(m.act, acc[3][j]);
      } else {
        o.x = acc[0][j]; o.y = acc[1][j]; o.z = acc[2][j]; o.w = acc[3][j];
      }
      *reinterpret_cast<float4*>(outp + n * TM + m0) = o;
    }
  __syncthreads();
}

// Final Linear(H -> out) (small fan-out): one (n, m) output per thread iteration, weights
// through the read-only path.  Writes raw (pre output-activation) values to outb [out][TM].
template <int TM, bool JAC = false>
__device__ void out_layer(const MlpDev& m, const float* __restrict__ hin, float* __restrict__ outb) {
  const int li = m.n_lin - 1;
  const float* __restrict__ Wg = m.params + m.w_off[li];
  const float* __restrict__ bg = m.params + m.b_off[li];
  const int NO = m.out, K = m.hidden;
  for (int idx = threadIdx.x; idx < NO * TM; idx += kThreads) {
    const int n = idx / TM, mm = idx - n * TM;
    float acc = (JAC && (mm & 3)) ? 0.0f : __ldg(bg + n);
    for (int k = 0; k < K; ++k) acc = nrt_fma(hin[k * TM + mm], __ldg(Wg + k * NO + n), acc);
    outb[n * TM + mm] = acc;
  }
  __syncthreads();
}

// Whole MLP for one tile.  Precondition: raw inputs in enc_raw rows [0,in) and latent in
// rows [in+2f, dim_p); all threads have passed a barrier since those writes.
// Postcondition: outb [out][TM] holds the pre-output-activation result.
// If acts_g != nullptr the post-activation hidden states are stored for the backward pass:
// acts_g[(l*H + k)*M_total + m_base + mm], l = 0..L.
template <int H, int TM, bool JAC = false>
__device__ void mlp_tile_forward(const MlpDev& m, const TileSmem& s, float* acts_g, int64_t M_total,
                                 int64_t m_base, int valid) {
  encode_tile<TM, JAC>(m, s);
  float* cur = s.h0;
  float* nxt = s.h1;
  auto save = [&](int l, const float* h) {
    if (acts_g == nullptr) return;
    for (int idx = threadIdx.x; idx < H * TM; idx += kThreads) {
      const int k = idx / TM, mm = idx - k * TM;
      if (mm < valid) acts_g[((size_t)l * H + k) * M_total + m_base + mm] = h[idx];
    }
  };
  hidden_layer<H, TM, JAC>(m, 0, s.enc_raw, m.dim_p, nullptr, 0, cur, s.wbuf, true);
  save(0, cur);
  for (int i = 0; i < m.L; ++i) {
    const bool sk = (m.skip_mask >> i) & 1u;
    hidden_layer<H, TM, JAC>(m, 1 + i, cur, H, s.enc_act, sk ? m.dim_p : 0, nxt, s.wbuf, true);
    save(1 + i, nxt);
    float* t = cur; cur = nxt; nxt = t;
  }
  out_layer<TM, JAC>(m, cur, s.outb);
}

}  // namespace nrt
