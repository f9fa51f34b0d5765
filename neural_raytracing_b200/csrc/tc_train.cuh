// Shared pieces of the tensor-core TRAINING path (nrt_tc_train.cu: networks whose weights are resident in shared memory;
// nrt_tc_train_wide.cu: the 256-wide networks with streamed weights): the per-batch workspace of saved tiles, the
// gradient IO policies with the device-side loss scale, and the weight-gradient kernel.  Internal header of libnrt_b200.
#pragma once
#include "tc_wide.cuh"

namespace tc {

// ---------------------------------------------------------------------------------------------
// dgrad blob: the weights with the roles of K and N swapped, UMMA canonical K-major, backward op order
//   op 0        : output layer   B'[n' = hidden unit][k' = output]
//   op 1+j      : hidden layer l = L-1-j   B'[n' = input of the layer (h | enc if NEEDX and skip)][k' = unit]
//   op L+1 (X)  : init layer     B'[n' = encoding column][k' = unit]
//   op L+2 (X)  : Fourier basis  B'[n' = input j][k' = frequency f] = basis[j][f]
// ---------------------------------------------------------------------------------------------
struct DLayout {
  int n_ops;
  int opN[kMaxOps], opK[kMaxOps], op_off[kMaxOps];
  int w_elems, bytes;
};
__host__ __device__ constexpr DLayout make_dlayout(int in, int lat, int f, int h, int L, int skip, int out, bool needx) {
  const Layout y = make_layout(in, lat, f, h, L, skip, out);
  DLayout d{};
  d.n_ops = 1 + L + (needx ? 2 : 0);
  int off = 0;
  for (int o = 0; o < d.n_ops; ++o) {
    int N = 0, K = 0;
    if (o == 0) { N = h; K = y.NOP; }
    else if (o <= L) { const int l = L - o; N = h + ((needx && is_skip(l, skip, L)) ? y.KE : 0); K = h; }
    else if (o == L + 1) { N = y.KE; K = h; }
    else { N = y.XR; K = y.FP; }
    d.opN[o] = N; d.opK[o] = K; d.op_off[o] = off; off += N * K;
  }
  d.w_elems = off;
  d.bytes = off * 2;
  return d;
}

// ---------------------------------------------------------------------------------------------
// training workspace (one per MLP and batch): all sections 256-byte aligned
// ---------------------------------------------------------------------------------------------
struct TrainWs {
  uint16_t* acts;      // [L+1][ntiles][(H+16) x 128]
  uint16_t* enc_raw;   // [ntiles][(KE+16) x 128]
  uint16_t* enc_act;   // [ntiles][(KE+16) x 128]
  uint16_t* dz;        // [L+1][ntiles][H x 128]
  uint16_t* gout;      // [ntiles][NOP x 128]
  uint32_t* masks;     // [L+1][H/32][ntiles*128] sign bits of a_l
  float* scale;        // [0]: max |g_out * out_act'| as float bits (atomicMax), [1]: loss scale S, [2]: 1/S
  int64_t ntiles;
  size_t bytes;
};
static TrainWs carve_ws(const Layout& y, int h, int L, int64_t M, void* base) {
  TrainWs w{};
  w.ntiles = (M + 127) / 128;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off += (n + 255) / 256 * 256; return o; };
  const size_t nt = (size_t)w.ntiles;
  const size_t o_acts = take((size_t)(L + 1) * nt * (h + kTileRowsExtra) * 128 * 2);
  const size_t o_er = take(nt * (y.KE + kTileRowsExtra) * 128 * 2);
  const size_t o_ea = take(nt * (y.KE + kTileRowsExtra) * 128 * 2);
  const size_t o_dz = take((size_t)(L + 1) * nt * h * 128 * 2);
  const size_t o_go = take(nt * y.NOP * 128 * 2);
  const size_t o_mk = take((size_t)(L + 1) * (h / 32) * nt * 128 * 4);
  const size_t o_sc = take(256);
  w.bytes = off;
  uint8_t* b = reinterpret_cast<uint8_t*>(base);
  if (b) {
    w.acts = (uint16_t*)(b + o_acts); w.enc_raw = (uint16_t*)(b + o_er); w.enc_act = (uint16_t*)(b + o_ea);
    w.dz = (uint16_t*)(b + o_dz); w.gout = (uint16_t*)(b + o_go);
    w.scale = (float*)(b + o_sc); w.masks = (uint32_t*)(b + o_mk);
  }
  return w;
}

// ---------------------------------------------------------------------------------------------
// training-forward IO: materialised x in, activated output out
// ---------------------------------------------------------------------------------------------
template <int IN, int OUT>
struct IoTrainFwd {
  const float* x; float* out; int out_act;
  __device__ __forceinline__ void load(int64_t m, float* v) const {
#pragma unroll
    for (int j = 0; j < IN; ++j) v[j] = __ldg(x + m * IN + j);
  }
  __device__ __forceinline__ void store(int64_t m, const float* o) const {
#pragma unroll
    for (int j = 0; j < OUT; ++j) {
      float v = o[j];
      if (out_act == NRT_OUT_SIGMOID) v = 1.0f / (1.0f + __expf(-v));
      else if (out_act == NRT_OUT_SOFTPLUS) v = v > 20.0f ? v : __logf(1.0f + __expf(v));
      else if (out_act == NRT_OUT_TANH) v = tanhf(v);
      out[m * OUT + j] = v;
    }
  }
};

// gradient IO: g_out (w.r.t. the ACTIVATED output `out`) in, g_x out
template <int IN, int OUT>
struct IoGrad {
  const float* out; const float* g_out; float* g_x; int out_act; const float* scale;
  // raw gradient w.r.t. the pre-activation output (no loss scale)
  __device__ __forceinline__ float g_pre(int64_t m, int j) const {
    float v = __ldg(g_out + m * OUT + j);
    if (out_act != NRT_OUT_NONE) {
      const float y = __ldg(out + m * OUT + j);
      if (out_act == NRT_OUT_SIGMOID) v *= y * (1.0f - y);
      else if (out_act == NRT_OUT_SOFTPLUS) v *= 1.0f - __expf(-y);
      else if (out_act == NRT_OUT_TANH) v *= 1.0f - y * y;
    }
    return v;
  }
  __device__ __forceinline__ void load_g(int64_t m, float* g) const {
    const float S = scale[1];
#pragma unroll
    for (int j = 0; j < OUT; ++j) g[j] = g_pre(m, j) * S;
  }
  __device__ __forceinline__ void store_gx1(int64_t m, int j, float v) const { g_x[m * IN + j] = v * scale[2]; }
};

// loss scale: S = 2^(5 - ceil(log2(max|g|))) so that the largest scaled gradient entering the chain is in [16, 32):
// 2000x head-room below fp16's 65504 for layers that amplify the gradient, 5e5x above its smallest normal number
// (IO policies whose gradient source is column-interleaved declare kColMajor: consecutive threads then take consecutive
//  samples of one output column, which is the coalesced order for them)
template <class IO, class = void> struct GradColMajor : std::false_type {};
template <class IO> struct GradColMajor<IO, std::void_t<decltype(IO::kColMajor)>> : std::true_type {};
template <class IO, int OUT>
__global__ void k_grad_absmax(IO io, int64_t M, float* __restrict__ scale) {
  float mx = 0.0f;
  constexpr bool COL = GradColMajor<IO>::value;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < M * OUT; idx += (int64_t)gridDim.x * blockDim.x) {
    const float v = fabsf(COL ? io.g_pre(idx % M, (int)(idx / M)) : io.g_pre(idx / OUT, (int)(idx % OUT)));
    if (v < 3.0e38f) mx = fmaxf(mx, v);   // ignores inf / nan
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0 && mx > 0.0f) atomicMax(reinterpret_cast<unsigned int*>(scale), __float_as_uint(mx));
}
static __global__ void k_grad_scale(float* __restrict__ scale) {
  const float mx = scale[0];
  int e = 0;
  if (mx > 0.0f) { frexpf(mx, &e); }          // mx = f * 2^e, f in [0.5, 1)
  const int k = mx > 0.0f ? 5 - e : 0;
  scale[1] = ldexpf(1.0f, k);
  scale[2] = ldexpf(1.0f, -k);
}

// 32 fp32 gradient columns x leaky_relu'(sign bits in `mask`) x c (a power of two: the per-layer rescale of the deep
// 256-wide chains, 1 elsewhere) -> 16 packed 16-bit pairs
template <int FMT>
__device__ __forceinline__ void dconvert32(const uint32_t* __restrict__ acc, uint32_t mask, uint32_t* __restrict__ pk, float c = 1.0f) {
  const float c_neg = 0.01f * c;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      // saturate instead of overflowing to inf (fp16 operands: a network whose weights amplify the gradient by more
      // than the 2000x head-room of the loss scale loses the largest entries, not the whole step)
      const float d = __uint_as_float(acc[8 * g + i]) * (((mask >> (8 * g + i)) & 1u) ? c_neg : c);
      v[i] = fminf(fmaxf(d, -60000.0f), 60000.0f);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) pk[4 * g + i] = Elem<FMT>::pack(v[2 * i], v[2 * i + 1]);
  }
}

// ---------------------------------------------------------------------------------------------
// wgrad: D'[n_out (lanes)][cols] = sum over samples dZ[s][n_out] * src[s][col]
// ---------------------------------------------------------------------------------------------
struct WgradJob {
  const uint16_t* a_tiles; int a_rows;     // dZ (or g_out) tiles: the M = 128 operand (rows staged in shared memory)
  int a_tile_rows;                          // rows of a whole dZ tile in HBM (> a_rows: this job stages rows unit0 .. unit0 + a_rows - 1)
  const uint16_t* s0_tiles; int s0_rows;   // first source (activation tile or raw encoding), with the ones row
  const uint16_t* s1_tiles; int s1_rows;   // second source (activated encoding of a skip layer) or null
  int n_valid;                              // valid output units (lanes)
  int N;                                    // fan-out of the linear layer (row stride of W^T [K][N])
  int w_off, b_off;                         // float offsets in the packed-f32 gradient blob
  int s0_kind;                              // 0: hidden activations (col c -> k = c), 1: encoding (col c -> enc_ref_index)
  int s0_valid;                             // columns of source 0 before the ones row
  int k_base1;                              // k offset of source 1 rows (hidden width)
  int unit0;                                // first output unit of this job (256-wide layers: one job per 128 units)
  int s0_kbase;                             // k offset of source 0 rows in W^T (jobs whose only source is the encoding of a skip layer)
  int lexp_from;                            // dZ of this job carries the per-layer rescales lexp[lexp_from .. n_lexp-1] (see WgradJobs)
  int split0, nsplit;                       // CTAs split0 .. split0 + nsplit - 1 share this job's sample range (wgrad_assign_splits)
};
constexpr int kMaxJobs = 64;
// lexp (optional): per-layer power-of-two exponents of the dgrad chain's rescale (256-wide nets): the dZ tiles of Linear li
// hold S * 2^(lexp[li] + ... + lexp[n_lexp-1]) * dZ, which the flush divides out
struct WgradJobs { WgradJob j[kMaxJobs]; int n; Layout y; int in_size; const int* lexp; int n_lexp; };

// One CTA per SM (512 TMEM columns, ~200 KB of stages).  The jobs share EXACTLY one or two waves of CTAs (round 1: the same
// number of splits for every job on a ceil(2 x SMs / jobs) x jobs grid = 301 CTAs for 7 jobs on 148 SMs: a third wave of 5
// CTAs, 6.3 instead of 5.0 ms for the 65,536-ray NeRFLE step).  A CTA's time per tile is a fixed part (barrier round trips of
// the 2-stage ring) plus the bytes it streams (dZ + sources): splits in proportion to  fixed + rows.  Measured on B200
// (tools/train_scaling.py, wgrad ms at 4,096 / 16,384 / 65,536 rays; profiles/r02_wgrad_sweep.log):
//   bytes only, one wave 0.49 / 1.78 / 6.96   fixed = 256 rows, one wave 0.39 / 1.34 / 5.16   equal splits, two waves 0.45 / 1.34 / 4.97
static inline int wgrad_assign_splits(WgradJobs& jb, int64_t nt, int ctas) {
  double w[kMaxJobs], tot = 0.0;
  const bool large = nt >= 8192;                 // >= ~28 tiles per CTA and job in two waves
  static const char* env_c = getenv("NRT_WGRAD_C");          // development knobs
  static const char* env_w = getenv("NRT_WGRAD_WAVES");
  const double kFixed = env_c ? atof(env_c) : (large ? 1.0e5 : 256.0);
  const double kWaves = env_w ? atof(env_w) : (large ? 2.0 : 1.0);
  ctas = (int)(ctas * kWaves);
  for (int i = 0; i < jb.n; ++i) {
    const WgradJob& j = jb.j[i];
    w[i] = kFixed + (double)j.a_rows + j.s0_rows + (j.s1_tiles ? j.s1_rows : 0);
    tot += w[i];
  }
  int used = 0;
  for (int i = 0; i < jb.n; ++i) {
    int k = (int)(ctas * w[i] / tot);            // floor; the remainder goes to the largest jobs below
    k = (int)std::max<int64_t>(1, std::min<int64_t>(nt, k));
    jb.j[i].nsplit = k; used += k;
  }
  for (bool more = true; more && used < ctas;) {  // hand out what the floors left, heaviest tiles-per-CTA first
    more = false;
    int best = -1; double load = 0.0;
    for (int i = 0; i < jb.n; ++i)
      if (jb.j[i].nsplit < nt && w[i] / jb.j[i].nsplit > load) { load = w[i] / jb.j[i].nsplit; best = i; }
    if (best >= 0) { jb.j[best].nsplit++; used++; more = true; }
  }
  int at = 0;
  for (int i = 0; i < jb.n; ++i) { jb.j[i].split0 = at; at += jb.j[i].nsplit; }
  return at;
}

template <int FMT>
__global__ void __launch_bounds__(160, 1)
k_mlp_wgrad_tc(const __grid_constant__ WgradJobs jobs_g, int64_t ntiles, int stage_bytes, int s0_off, int s1_off,
               float* __restrict__ g_params, const float* __restrict__ scale) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_full[2];
  __shared__ __align__(8) uint64_t bar_empty[2];
  __shared__ __align__(8) uint64_t bar_done;
  __shared__ uint32_t tmem_base_s;
  __shared__ WgradJob job;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    int ji = 0;
    while (ji + 1 < jobs_g.n && (int)blockIdx.x >= jobs_g.j[ji + 1].split0) ++ji;
    job = jobs_g.j[ji];
    mbar_init(&bar_full[0], 1); mbar_init(&bar_full[1], 1);
    mbar_init(&bar_empty[0], 1); mbar_init(&bar_empty[1], 1);
    mbar_init(&bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const int64_t sp = (int64_t)blockIdx.x - job.split0;
  const int64_t t_begin = ntiles * sp / job.nsplit, t_end = ntiles * (sp + 1) / job.nsplit;
  const int n = (int)(t_end - t_begin);
  const uint32_t a_bytes = (uint32_t)job.a_rows * 256u, s0_bytes = (uint32_t)job.s0_rows * 256u;
  const uint32_t s1_bytes = job.s1_tiles ? (uint32_t)job.s1_rows * 256u : 0u;

  if (warp == 4) {
    // producer + MMA issuer (whole warp convergent, one elected lane acts)
    // both operands are MN-major tiles (feature-contiguous core-matrix rows): a_major (bit 15) = b_major (bit 16) = 1
    constexpr uint32_t kMajorMN = (1u << 15) | (1u << 16);
    const uint32_t idesc_base = (1u << 4) | ((uint32_t)FMT << 7) | ((uint32_t)FMT << 10) | kMajorMN | ((uint32_t)(128 >> 4) << 24);
    // the job lives in shared memory and every tcgen05.mma is a volatile asm with a memory clobber: keep what the issue
    // loop needs in registers
    const uint32_t lbo_a = (uint32_t)job.a_rows * 16u, lbo0 = (uint32_t)job.s0_rows * 16u, lbo1 = (uint32_t)job.s1_rows * 16u;
    uint32_t e_col[4], e_id[4], e_src[4];
    uint64_t e_boff[4];
    int n_ent = 0;
#pragma unroll
    for (int src = 0; src < 2; ++src) {
      const int rows = src == 0 ? job.s0_rows : (s1_bytes ? job.s1_rows : 0);
      const uint32_t col0 = src == 0 ? 0u : (uint32_t)job.s0_rows;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int n0 = 256 * c;
        if (n0 < rows) {
          const uint32_t nn = (uint32_t)min(256, rows - n0);
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (e == n_ent) {
              e_col[e] = col0 + (uint32_t)n0; e_id[e] = idesc_base | ((nn >> 3) << 17); e_src[e] = (uint32_t)src;
              e_boff[e] = (uint64_t)(((uint32_t)(n0 / 8) * 128u) >> 4);
            }
          ++n_ent;
        }
      }
    }
    for (int i = 0; i <= n; ++i) {
      if (i < n) {
        const int slot = i & 1;
        if (i >= 2) mbar_wait(&bar_empty[slot], ((i >> 1) - 1) & 1);
        if (elect_one()) {
          uint8_t* sb = smem + (size_t)slot * stage_bytes;
          const int64_t t = t_begin + i;
          mbar_expect_tx(&bar_full[slot], a_bytes + s0_bytes + s1_bytes);
          const uint16_t* at = job.a_tiles + t * (int64_t)(job.a_tile_rows * 128);
          if (job.a_tile_rows == job.a_rows) {
            bulk_g2s(sb, at, a_bytes, &bar_full[slot]);
          } else {
            // a 128-unit half of a 256-row tile: per 8-sample group, a_rows features x 8 samples are contiguous
            for (int g = 0; g < 16; ++g)
              bulk_g2s(sb + (size_t)g * job.a_rows * 16, at + (size_t)g * job.a_tile_rows * 8 + (size_t)job.unit0 * 8,
                       (uint32_t)job.a_rows * 16u, &bar_full[slot]);
          }
          for (uint32_t off = 0; off < s0_bytes; off += 32768u)
            bulk_g2s(sb + s0_off + off, reinterpret_cast<const uint8_t*>(job.s0_tiles + t * (int64_t)(job.s0_rows * 128)) + off,
                     min(32768u, s0_bytes - off), &bar_full[slot]);
          for (uint32_t off = 0; off < s1_bytes; off += 32768u)
            bulk_g2s(sb + s1_off + off, reinterpret_cast<const uint8_t*>(job.s1_tiles + t * (int64_t)(job.s1_rows * 128)) + off,
                     min(32768u, s1_bytes - off), &bar_full[slot]);
        }
        __syncwarp();
      }
      if (i >= 1) {
        const int j = i - 1, slot = j & 1;
        mbar_wait(&bar_full[slot], (j >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sb = smem_u32(smem + (size_t)slot * stage_bytes);
          const uint64_t ad = make_desc(sb, lbo_a, 128), b0 = make_desc(sb + (uint32_t)s0_off, lbo0, 128), b1 = make_desc(sb + (uint32_t)s1_off, lbo1, 128);
#pragma unroll 1
          for (int kc = 0; kc < 8; ++kc) {
            const uint32_t acc = (j > 0 || kc > 0) ? 1u : 0u;
            const uint64_t adk = ad + (uint64_t)((kc * 2 * lbo_a) >> 4);
            // D'[tmem] (+)= A''[smem] * B''[smem]^T; a source wider than the MMA's N limit (256) goes in two N-chunks
            // (the instruction operands were prepared once per job: e_col / e_id / e_boff / e_src, at most 4 entries)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (e < n_ent) {
                const uint64_t bd = (e_src[e] ? b1 + (uint64_t)((kc * 2 * lbo1) >> 4) : b0 + (uint64_t)((kc * 2 * lbo0) >> 4)) + e_boff[e];
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                             ::"r"(tmem + e_col[e]), "l"(adk), "l"(bd), "r"(e_id[e]), "r"(acc) : "memory");
              }
            }
          }
          tc_commit(&bar_empty[slot]);
          if (j == n - 1) tc_commit(&bar_done);
        }
        __syncwarp();
      }
    }
  } else if (n > 0) {
    // epilogue: lane = output unit; coalesced float atomics into the packed-f32 gradient blob
    mbar_wait(&bar_done, 0);
    tc_fence_after();
    const int unit = tid;   // 0..127
    const uint32_t trow = tmem + (((uint32_t)(warp * 32)) << 16);
    const Layout& y = jobs_g.y;
    const int in_size = jobs_g.in_size;
    const int ncols = job.s0_rows + (job.s1_tiles ? job.s1_rows : 0);
    float inv_s = scale[2];
    if (jobs_g.lexp != nullptr) {
      int e = 0;
      for (int l = job.lexp_from; l < jobs_g.n_lexp; ++l) e += jobs_g.lexp[l];
      inv_s = ldexpf(inv_s, -e);
    }
    for (int c0 = 0; c0 < ncols; c0 += 16) {
      uint32_t v[16];
      TmemIO<16>::ld(trow + c0, v);
      tc_wait_ld();
      if (unit < job.n_valid) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int c = c0 + j;
          int dst = -1;
          if (c < job.s0_rows) {
            if (c < job.s0_valid) {
              const int k = job.s0_kind == 0 ? c : enc_ref_index(y, in_size, c);
              if (k >= 0) dst = job.w_off + (job.s0_kbase + k) * job.N + job.unit0 + unit;
            } else if (c == job.s0_valid && job.b_off >= 0) {
              dst = job.b_off + job.unit0 + unit;
            }
          } else {
            const int k = enc_ref_index(y, in_size, c - job.s0_rows);
            if (k >= 0 && c - job.s0_rows < y.KE) dst = job.w_off + (job.k_base1 + k) * job.N + job.unit0 + unit;
          }
          if (dst >= 0) atomicAdd(g_params + dst, __uint_as_float(v[j]) * inv_s);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}


// per-layer power-of-two rescale of a deep dgrad chain (lexp[l], l = 0..L: dZ_l is multiplied by 2^lexp[l] when it leaves
// the accumulator, and the weight-gradient flush divides the accumulated product out): 2^-lexp[l] ~ the gain of the step
// dZ_{l+1} -> dZ_l estimated from the weights, |dA| ~ |dZ| * ||W||_F / sqrt(hidden rows), times ~0.7 for the activation's
// derivative.  Without it a chain whose layers shrink the gradient (default-initialised 256-wide layers: 0.41 per layer; a
// freshly initialised SDF residual net: 0.1) leaves fp16's range after a few layers, whatever the loss scale.
static __global__ void k_dgrad_layer_scales(MlpDev m, int* __restrict__ lexp) {
  // block l: Linear li = l + 1 (the one that consumes a_l); hidden rows only (the encoding rows carry no gradient here)
  const int l = blockIdx.x, li = l + 1, h = m.hidden;
  const int n_out = li == m.n_lin - 1 ? m.out : h;
  const float* w = m.params + m.w_off[li];
  float ss = 0.0f;
  for (int i = threadIdx.x; i < h * n_out; i += blockDim.x) { const float v = w[i]; ss = fmaf(v, v, ss); }
  __shared__ float red[32];
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    const float gain = 0.7071f * sqrtf(t / (float)h);
    int e = 0;
    if (gain > 0.0f && gain < 3.0e38f) e = -(int)lrintf(log2f(gain));
    lexp[l] = max(-8, min(8, e));
  }
}

// blob: chunks in consumption order.  op 0 (output layer): halves nh = 0, 1 of B'[n = hidden unit][k = output];
// op o = 1..L (hidden layer l = L - o, Linear li = l + 1): quarters q = 2 * nh + kh of B'[n = input unit][k = output unit];
// every chunk UMMA canonical K-major with N = 128: element (n, k) at ((k / 8) * 128 + n) * 8 + k % 8

// weight gradients of a network whose layers are at most 128 wide: one job per Linear (skip layers: hidden activations and
// the activated encoding as two sources of one job)
template <class NET, int FMT>
static int launch_wgrad_std(const MlpDev& d, const TrainWs& ws, float* g_params, cudaStream_t st, const int* lexp = nullptr) {
  // ---- weight gradients: one job per linear layer ----
  constexpr int H = NET::H, L = NET::L, KE = NET::KE, NOP = NET::NOP;
  constexpr int FRA = H + kTileRowsExtra, FRE = KE + kTileRowsExtra;
  WgradJobs jobs{};
  jobs.y = NET::Y; jobs.in_size = NET::IN; jobs.n = L + 2;
  jobs.lexp = lexp; jobs.n_lexp = lexp ? L + 1 : 0;
  const int64_t nt = ws.ntiles;
  for (int li = 0; li <= L + 1; ++li) {
    WgradJob& j = jobs.j[li];
    j.N = d.N[li]; j.w_off = d.w_off[li]; j.b_off = d.b_off[li]; j.k_base1 = H;
    j.a_tile_rows = (li == L + 1) ? NOP : H;
    j.lexp_from = li;
    if (li == L + 1) {
      j.a_tiles = ws.gout; j.a_rows = NOP; j.n_valid = NET::OUT;
      j.s0_tiles = ws.acts + (int64_t)L * nt * FRA * 128; j.s0_rows = FRA; j.s0_kind = 0; j.s0_valid = H;
    } else if (li == 0) {
      j.a_tiles = ws.dz; j.a_rows = H; j.n_valid = H;
      j.s0_tiles = ws.enc_raw; j.s0_rows = FRE; j.s0_kind = 1; j.s0_valid = KE;
    } else {
      j.a_tiles = ws.dz + (int64_t)li * nt * H * 128; j.a_rows = H; j.n_valid = H;
      j.s0_tiles = ws.acts + (int64_t)(li - 1) * nt * FRA * 128; j.s0_rows = FRA; j.s0_kind = 0; j.s0_valid = H;
      if (is_skip(li - 1, NET::SKIP, L)) { j.s1_tiles = ws.enc_act; j.s1_rows = FRE; }
    }
  }
  // stage: [A'' | source 0 | source 1].  The M = 128 operand reads 128 rows per sample group whatever a_rows is, so its
  // region spans (15 * a_rows + 128) * 16 bytes; the garbage rows only reach accumulator lanes >= a_rows (ignored)
  constexpr int A_ROWS = H > NOP ? H : NOP;
  constexpr int S0_OFF = ((15 * A_ROWS + 128) * 16 + 1023) / 1024 * 1024;
  constexpr int S0_BYTES = (FRA > FRE ? FRA : FRE) * 256;
  constexpr int S1_OFF = S0_OFF + S0_BYTES;
  const int stage_bytes = S1_OFF + FRE * 256;
  const size_t bytes = 2 * (size_t)stage_bytes + 8192;
  static_assert(2 * (S1_OFF + FRE * 256) + 8192 + 1024 <= 227 * 1024, "wgrad stages do not fit in shared memory");
  static_assert(FRA + FRE <= 512 && FRA <= 256 && FRE <= 256, "wgrad accumulator / MMA N limits");
  auto kern = k_mlp_wgrad_tc<FMT>;
  NRT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  const int ctas = wgrad_assign_splits(jobs, nt, nrt_sm_count());
  {
    NrtProfScope _ps(TAG_TC_WGRAD, st);
    kern<<<ctas, 160, bytes, st>>>(jobs, nt, stage_bytes, S0_OFF, S1_OFF, g_params, ws.scale);   // job table by value (kernel parameters): no copy, graph-capturable
  }
  NRT_CUDA(cudaGetLastError());
  return NRT_OK;
}

}  // namespace tc
