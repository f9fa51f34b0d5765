// Tensor-core TRAINING path of SphereSDF.shift (8 x 128, softplus, 32 frequencies; shapes/sdfs.py:23-31):
//   * first order: the differentiable sdf(best_pos) of SDF.throughput (sdfs.py:249) -- rows = points;
//   * VALUE + JACOBIAN: what SDF.autograd_diff (sdfs.py:184-197) and the double backward of loss.backward() through it
//     compute for the shading normals and eikonal_loss.  Forward mode over FOUR rows per point,
//         row 4i = value, rows 4i + 1..3 = d / dp_0..2:
//     the Linear layers see all four (tangent rows without bias), the activation couples them,
//         a_v = softplus(z_v),   a_t = sigmoid(z_v) * z_t,
//     and the reverse pass of that four-row network is an ordinary dgrad / wgrad over the 4K rows with
//         g_z_t = s * g_a_t,   g_z_v = s * g_a_v + (1 - s) * sum_t g_a_t a_t,   s = sigmoid(z_v) = 1 - exp(-a_v)
//     (softplus'' z_t = s (1 - s) z_t = (1 - s) a_t).  The four rows of a point are four lanes of one warp, so the
//     coupling is a shuffle.  The hit points carry no gradient (the march is no_grad, sdfs.py:119-131).
// Kernels: one 128-row tile per CTA, TWO CTAs per SM (256 TMEM columns and < 100 KB of shared memory each: the second
// CTA's MMAs run under the first one's epilogue), weights streamed from L2 (the 333 KB of 16-bit weights do not fit in
// shared memory), the raw and the activated encoding as shared-memory A operands, hidden activations as the TMEM A
// operand.  The weight gradients come from the shared k_mlp_wgrad_tc (tc_train.cuh).
#include "tc_train.cuh"

namespace tc {

using NetSdfShift = Net<3, 0, 32, 128, 8, 3, 1, NRT_ACT_SOFTPLUS>;

constexpr int kSdfKC = 64;        // K elements per streamed forward chunk
constexpr int kSdfNB = 3;         // forward ring depth

template <class NET>
struct SdfT {
  static constexpr Layout Y = NET::Y;
  static constexpr int H = NET::H, L = NET::L, KE = NET::KE, F = NET::F, IN = NET::IN, XR = NET::XR, NOP = NET::NOP;
  static_assert(H == 128 && NET::ENC_CUDA && NET::ACT == NRT_ACT_SOFTPLUS && NET::LAT == 0 && NET::FUSE_OUT && NET::OUT == 1 &&
                    NET::SKIP == 3 && XR == 16 && IN == 3, "SDF training kernels: the 8x128 softplus residual net");
  static constexpr int N_OPS = L + 1;                       // init + L hidden layers (blob ops 1 .. L+1)
  static constexpr int CH_BYTES = H * kSdfKC * 2;           // 16 KB
  static constexpr int ENC_BYTES = KE * 128 * 2;            // one encoding operand (raw / activated)
  static constexpr int BIAS_BYTES = Y.bias_floats * 4;
  static constexpr int SMEM_FWD = 2 * ENC_BYTES + kSdfNB * CH_BYTES + BIAS_BYTES;
  static constexpr int op_k(int o) { return Y.opK[1 + o]; }
  static constexpr int op_chunks(int o) { return (op_k(o) + kSdfKC - 1) / kSdfKC; }
  static constexpr int chunks_per_tile() { int n = 0; for (int o = 0; o < N_OPS; ++o) n += op_chunks(o); return n; }
  static constexpr int CPT = chunks_per_tile();
  // dgrad: one chunk per op (the whole transposed layer), ring of 2
  static constexpr DLayout DY = make_dlayout(NET::IN, 0, NET::F, NET::H, NET::L, NET::SKIP, NET::OUT, false);
  static constexpr int DCH_BYTES = H * H * 2;               // 32 KB
  static constexpr int SMEM_BWD = 2 * DCH_BYTES;
  static_assert(2 * (SMEM_FWD + 2048) <= 227 * 1024 && 2 * (SMEM_BWD + 2048) <= 227 * 1024, "two CTAs per SM");
};

// softplus(v) and sigmoid(v) from one ex2: e = exp(-|v|), u = 1 + e
__device__ __forceinline__ void softplus_sigmoid(float v, float* sp, float* sg) {
  const float e = ex2_approx(-1.4426950408889634f * fabsf(v));
  const float u = 1.0f + e;
  const float r = __fdividef(1.0f, u);
  *sp = fmaf(0.6931471805599453f, lg2_approx(u), fmaxf(v, 0.0f));
  *sg = v >= 0.0f ? r : e * r;
}

template <int N>
__device__ __forceinline__ void preload_bias_if(uint32_t dD, const float* __restrict__ bias, bool on) {
  uint32_t r[N];
#pragma unroll
  for (int j = 0; j < N / 4; ++j) {
    const float4 b = *reinterpret_cast<const float4*>(bias + 4 * j);
    r[4 * j] = on ? __float_as_uint(b.x) : 0u; r[4 * j + 1] = on ? __float_as_uint(b.y) : 0u;
    r[4 * j + 2] = on ? __float_as_uint(b.z) : 0u; r[4 * j + 3] = on ? __float_as_uint(b.w) : 0u;
  }
  tmem_store<N>(dD, r);
}

// ---------------------------------------------------------------------------------------------
// forward: p [K,3] -> value (+ Jacobian), saving the activation tiles
// ---------------------------------------------------------------------------------------------
template <class NET, int FMT, bool JAC>
__global__ void __launch_bounds__(160, 2)
k_sdf_train_fwd_tc(const uint8_t* __restrict__ blob, const float* __restrict__ p, int64_t M, float* __restrict__ value,
                   float* __restrict__ jac, int out_act, SaveTiles sv) {
  using T = SdfT<NET>;
  using E = Elem<FMT>;
  constexpr Layout Y = NET::Y;
  constexpr int H = 128, L = NET::L, KE = NET::KE, F = NET::F, IN = 3, XR = 16;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sEncRaw = smem;
  uint8_t* sEncAct = smem + T::ENC_BYTES;
  uint8_t* sRing = smem + 2 * T::ENC_BYTES;
  const float* sBias = reinterpret_cast<const float*>(smem + 2 * T::ENC_BYTES + kSdfNB * T::CH_BYTES);
  __shared__ __align__(8) uint64_t bar_full[kSdfNB], bar_empty[kSdfNB], bar_ready, bar_done, bar_bias;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int64_t ntiles = (M + 127) / 128;
  const int64_t my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (tid == 0) {
    for (int i = 0; i < kSdfNB; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
    mbar_init(&bar_ready, 128); mbar_init(&bar_done, 1); mbar_init(&bar_bias, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    if ((tid & 31) == 0) {
      mbar_expect_tx(&bar_bias, (uint32_t)T::BIAS_BYTES);
      bulk_g2s(const_cast<float*>(sBias), blob + (size_t)Y.w_elems * 2, (uint32_t)T::BIAS_BYTES, &bar_bias);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t dD = tmem, aU = tmem + 128;     // accumulator 128 columns, hidden operand 64 columns
  mbar_wait(&bar_bias, 0);

  if (warp == 4) {
    // ===================== producer + MMA issuer =====================
    const int64_t total_chunks = my_tiles * T::CPT;
    int p_op = 0, p_chunk = 0;
    int64_t p_i = 0, c_i = 0;
    uint32_t n_ready = 0;
    const uint32_t ring_addr = smem_u32(sRing), raw_addr = smem_u32(sEncRaw), act_addr = smem_u32(sEncAct);
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)FMT << 7) | ((uint32_t)FMT << 10) | ((uint32_t)(H >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    auto k_of = [](int o) { return o == 0 ? KE : H + (is_skip(o - 1, 3, L) ? KE : 0); };
    auto top_up = [&]() {
      while (p_i < total_chunks && p_i - c_i < kSdfNB) {
        const int b = (int)(p_i % kSdfNB);
        if (p_i >= kSdfNB) mbar_wait(&bar_empty[b], (uint32_t)((p_i / kSdfNB - 1) & 1));
        const int K = k_of(p_op);
        if (elect_one()) {
          const int k0 = p_chunk * kSdfKC;
          const uint32_t bytes = (uint32_t)(min(kSdfKC, K - k0) * H * 2);
          const uint8_t* src = blob + (size_t)Y.op_off[1 + p_op] * 2 + (size_t)(k0 / 8) * H * 16;
          mbar_expect_tx(&bar_full[b], bytes);
          bulk_g2s(sRing + (size_t)b * T::CH_BYTES, src, bytes, &bar_full[b]);
        }
        __syncwarp();
        if (++p_chunk == (K + kSdfKC - 1) / kSdfKC) { p_chunk = 0; if (++p_op == T::N_OPS) p_op = 0; }
        ++p_i;
      }
    };
    for (int64_t t = 0; t < my_tiles; ++t) {
      for (int op = 0; op < T::N_OPS; ++op) {
        top_up();
        mbar_wait(&bar_ready, n_ready & 1); n_ready++;
        tc_fence_after();
        const int K = k_of(op);
        const int nch = (K + kSdfKC - 1) / kSdfKC;
        for (int c = 0; c < nch; ++c) {
          top_up();
          const int b = (int)(c_i % kSdfNB);
          mbar_wait(&bar_full[b], (uint32_t)((c_i / kSdfNB) & 1));
          tc_fence_after();
          if (elect_one()) {
            const int k0 = c * kSdfKC, kc = min(kSdfKC, K - k0);
            const uint64_t bd0 = make_desc(ring_addr + (uint32_t)b * T::CH_BYTES, (uint32_t)H * 16u, 128);
            for (int j = 0; j < kc / 16; ++j) {
              const int k = k0 + 16 * j;                  // [hidden 128 | encoding KE]; init layer: encoding only
              const uint64_t bd = bd0 + (uint64_t)((j * 2 * H * 16) >> 4);
              const int ke = (op == 0) ? k : k - H;
              // the accumulator was pre-loaded with the bias (value rows) or zero (tangent rows): always accumulate
              if (ke >= 0) {
                const uint64_t ad = make_desc((op == 0 ? raw_addr : act_addr) + (uint32_t)(ke / 8) * 2048u, 2048u, 128u);
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                             ::"r"(dD), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
              } else {
                mma_ts(dD, aU + (uint32_t)(k / 16) * 8u, bd, idesc, 1u);
              }
            }
            tc_commit(&bar_empty[b]);
            if (c == nch - 1) tc_commit(&bar_done);
          }
          __syncwarp();
          ++c_i;
        }
      }
    }
  } else {
    // ===================== epilogue warpgroup: thread = row =====================
    const int row = tid, lane = tid & 31;
    const int role = JAC ? (row & 3) : 0;          // 0: value row, 1..3: d / dp_{role-1}
    const int vl = lane & ~3;                      // lane that holds this point's value row
    const uint32_t lane_off = ((uint32_t)(warp * 32)) << 16;
    const uint32_t tD = dD + lane_off, tU = aU + lane_off;
    uint32_t n_done = 0;
    uint4* raw_row = reinterpret_cast<uint4*>(sEncRaw) + row;     // 16-byte slot of k-group g at [g * 128 + row]
    uint4* act_row = reinterpret_cast<uint4*>(sEncAct) + row;
    const float* sB = sBias + Y.basis_f32_off;
    const float* wout = sBias + Y.wout_f32_off;
    for (int64_t t = 0; t < my_tiles; ++t) {
      const int64_t tile = (int64_t)blockIdx.x + t * gridDim.x;
      const int64_t m = tile * 128 + row;
      const bool valid = m < M;
      const int64_t pt = JAC ? (m >> 2) : m;
      uint16_t* sv_raw = tile_row_ptr(sv.enc_raw, tile, KE + kTileRowsExtra, row);
      uint16_t* sv_act = tile_row_ptr(sv.enc_act, tile, KE + kTileRowsExtra, row);
      // ---- encoding of this row: value row [x | sin | cos], tangent row j [e_j | cos * B_j | -sin * B_j]; and what the
      //      skip layers see: softplus(enc_v) resp. sigmoid(enc_v) * enc_t ----
      {
        float x[IN];
#pragma unroll
        for (int j = 0; j < IN; ++j) x[j] = valid ? __ldg(p + pt * IN + j) : 0.0f;
        uint16_t rv[8], av[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { rv[j] = 0; av[j] = 0; }
#pragma unroll
        for (int j = 0; j < IN; ++j) {
          float sp, sg;
          softplus_sigmoid(x[j], &sp, &sg);
          const float r = role == 0 ? x[j] : (role == j + 1 ? 1.0f : 0.0f);
          const float a = role == 0 ? sp : (role == j + 1 ? sg : 0.0f);
          const uint16_t rh = E::cvt(r), ah = E::cvt(a);
          rv[j] = rh; rv[IN + j] = E::cvt(r - E::back(rh));
          av[j] = ah; av[IN + j] = E::cvt(a - E::back(ah));
        }
        const uint4 q_r = make_uint4((uint32_t)rv[0] | ((uint32_t)rv[1] << 16), (uint32_t)rv[2] | ((uint32_t)rv[3] << 16),
                                     (uint32_t)rv[4] | ((uint32_t)rv[5] << 16), (uint32_t)rv[6] | ((uint32_t)rv[7] << 16));
        const uint4 q_a = make_uint4((uint32_t)av[0] | ((uint32_t)av[1] << 16), (uint32_t)av[2] | ((uint32_t)av[3] << 16),
                                     (uint32_t)av[4] | ((uint32_t)av[5] << 16), (uint32_t)av[6] | ((uint32_t)av[7] << 16));
        const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
        raw_row[0] = q_r; raw_row[128] = zero4;
        act_row[0] = q_a; act_row[128] = zero4;
        *reinterpret_cast<uint4*>(sv_raw) = q_r; *reinterpret_cast<uint4*>(sv_raw + 64) = zero4;
        *reinterpret_cast<uint4*>(sv_act) = q_a; *reinterpret_cast<uint4*>(sv_act + 64) = zero4;
#pragma unroll 1
        for (int g = 0; g < F / 8; ++g) {
          uint32_t rs[4], rc[4], as[4], ac[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float r_s[2], r_c[2], a_s[2], a_c[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int f = 8 * g + 2 * e + h;
              float ph = x[0] * sB[f];
#pragma unroll
              for (int j = 1; j < IN; ++j) ph = fmaf(x[j], sB[j * NET::FP + f], ph);
              float s, c;
              sincos_fast(ph, &s, &c);
              float sps, sgs, spc, sgc;
              softplus_sigmoid(s, &sps, &sgs);
              softplus_sigmoid(c, &spc, &sgc);
              const float bj = role == 0 ? 0.0f : sB[(role - 1) * NET::FP + f];
              r_s[h] = role == 0 ? s : c * bj;
              r_c[h] = role == 0 ? c : -s * bj;
              a_s[h] = role == 0 ? sps : sgs * r_s[h];
              a_c[h] = role == 0 ? spc : sgc * r_c[h];
            }
            rs[e] = E::pack(r_s[0], r_s[1]); rc[e] = E::pack(r_c[0], r_c[1]);
            as[e] = E::pack(a_s[0], a_s[1]); ac[e] = E::pack(a_c[0], a_c[1]);
          }
          const uint4 v_rs = make_uint4(rs[0], rs[1], rs[2], rs[3]), v_rc = make_uint4(rc[0], rc[1], rc[2], rc[3]);
          const uint4 v_as = make_uint4(as[0], as[1], as[2], as[3]), v_ac = make_uint4(ac[0], ac[1], ac[2], ac[3]);
          raw_row[(XR / 8 + g) * 128] = v_rs; raw_row[(XR / 8 + F / 8 + g) * 128] = v_rc;
          act_row[(XR / 8 + g) * 128] = v_as; act_row[(XR / 8 + F / 8 + g) * 128] = v_ac;
          *reinterpret_cast<uint4*>(sv_raw + (XR / 8 + g) * 64) = v_rs; *reinterpret_cast<uint4*>(sv_raw + (XR / 8 + F / 8 + g) * 64) = v_rc;
          *reinterpret_cast<uint4*>(sv_act + (XR / 8 + g) * 64) = v_as; *reinterpret_cast<uint4*>(sv_act + (XR / 8 + F / 8 + g) * 64) = v_ac;
        }
        // the constant-1 row (bias gradient): value rows only
        sv_raw[tile_elem(KE)] = role == 0 ? one16<FMT>() : (uint16_t)0;
        sv_act[tile_elem(KE)] = role == 0 ? one16<FMT>() : (uint16_t)0;
        preload_bias_if<H>(tD, sBias + Y.bias_off[1], role == 0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(&bar_ready);
      }
      // ---- hidden layers (+ the fp32 output layer in the last one) ----
      float o = role == 0 ? sBias[Y.bias_off[NET::STAGES - 1]] : 0.0f;
#pragma unroll 1
      for (int st = 0; st <= L; ++st) {
        mbar_wait(&bar_done, n_done & 1); n_done++;
        tc_fence_after();
        uint16_t* save_row = tile_row_ptr(sv.acts, (int64_t)st * sv.ntiles + tile, H + kTileRowsExtra, row);
        uint32_t buf[2][32];
        TmemIO<32>::ld(tD, buf[0]);
        tc_wait_ld();
#pragma unroll
        for (int c = 0; c < H / 32; ++c) {
          if (c + 1 < H / 32) TmemIO<32>::ld(tD + 32 * (c + 1), buf[(c + 1) & 1]);
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float a2[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const float z = __uint_as_float(buf[c & 1][2 * i + h]);
              const float zv = JAC ? __shfl_sync(0xffffffffu, z, vl) : z;
              float sp, sg;
              softplus_sigmoid(zv, &sp, &sg);
              a2[h] = role == 0 ? sp : sg * z;
              if (st == L) o = fmaf(a2[h], wout[(32 * c + 2 * i + h) * 4], o);
            }
            pk[i] = E::pack(a2[0], a2[1]);
          }
          if (st < L) TmemIO<16>::st(tU + 16 * c, pk);
          save_cols<16>(save_row, 32 * c, pk);
          if (c + 1 < H / 32) tc_wait_ld();
        }
        save_row[tile_elem(H)] = role == 0 ? one16<FMT>() : (uint16_t)0;
        if (st < L) {
          preload_bias_if<H>(tD, sBias + Y.bias_off[2 + st], role == 0);
          tc_wait_st();
          tc_fence_before();
          mbar_arrive(&bar_ready);
        } else {
          tc_fence_before();
        }
      }
      if (valid) {
        if (JAC) {
          if (role == 0) value[pt] = o;
          else jac[pt * 3 + role - 1] = o;
        } else {
          float v = o;
          if (out_act == NRT_OUT_SIGMOID) v = 1.0f / (1.0f + __expf(-v));
          else if (out_act == NRT_OUT_SOFTPLUS) v = v > 20.0f ? v : __logf(1.0f + __expf(v));
          else if (out_act == NRT_OUT_TANH) v = tanhf(v);
          value[m] = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

// ---------------------------------------------------------------------------------------------
// dgrad: g -> dZ_L -> ... -> dZ_0 through the (coupled) softplus derivative
// ---------------------------------------------------------------------------------------------
struct IoGradJac {       // row 4i: g w.r.t. the value (may be absent), rows 4i + 1..3: g w.r.t. the Jacobian
  const float* g_value; const float* g_jac; const float* scale;
  __device__ __forceinline__ float g_pre(int64_t m, int) const {
    const int role = (int)(m & 3);
    const int64_t pt = m >> 2;
    if (role == 0) return g_value ? __ldg(g_value + pt) : 0.0f;
    return __ldg(g_jac + pt * 3 + role - 1);
  }
  __device__ __forceinline__ void load_g(int64_t m, float* g) const { g[0] = g_pre(m, 0) * scale[1]; }
};

template <class NET, class IO, int FMT, bool JAC>
__global__ void __launch_bounds__(160, 2)
k_sdf_dgrad_tc(const uint8_t* __restrict__ blob, IO io, int64_t M, TrainWs ws) {
  using T = SdfT<NET>;
  using E = Elem<FMT>;
  constexpr DLayout D = T::DY;
  constexpr int H = 128, L = NET::L, NOP = NET::NOP;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_full[2], bar_empty[2], bar_ready, bar_done;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_cl[L + 1];                     // per-layer rescale 2^lexp[l] (the blob's tail, k_dgrad_layer_scales)

  const int tid = threadIdx.x, warp = tid >> 5;
  const int64_t ntiles = (M + 127) / 128;
  const int64_t my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (tid <= L) s_cl[tid] = ldexpf(1.0f, reinterpret_cast<const int*>(blob + D.bytes)[tid]);
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
    mbar_init(&bar_ready, 128); mbar_init(&bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t dM = tmem, aA = tmem + 128;

  if (warp == 4) {
    // ===================== producer + MMA issuer: one chunk per op =====================
    const int64_t total = my_tiles * (L + 1);
    int64_t p_i = 0, c_i = 0;
    int p_op = 0;
    uint32_t n_ready = 0;
    const uint32_t ring_addr = smem_u32(smem);
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)FMT << 7) | ((uint32_t)FMT << 10) | ((uint32_t)(H >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    auto top_up = [&]() {
      while (p_i < total && p_i - c_i < 2) {
        const int b = (int)(p_i & 1);
        if (p_i >= 2) mbar_wait(&bar_empty[b], (uint32_t)((p_i / 2 - 1) & 1));
        if (elect_one()) {
          const uint32_t bytes = (uint32_t)(H * (p_op == 0 ? NOP : H) * 2);
          mbar_expect_tx(&bar_full[b], bytes);
          bulk_g2s(smem + (size_t)b * T::DCH_BYTES, blob + (size_t)D.op_off[p_op == 0 ? 0 : 1] * 2 + (p_op == 0 ? 0 : (size_t)(p_op - 1) * H * H * 2),
                   bytes, &bar_full[b]);
        }
        __syncwarp();
        if (++p_op == L + 1) p_op = 0;
        ++p_i;
      }
    };
    for (int64_t t = 0; t < my_tiles; ++t) {
      for (int op = 0; op <= L; ++op) {
        top_up();
        mbar_wait(&bar_ready, n_ready & 1); n_ready++;
        tc_fence_after();
        const int b = (int)(c_i & 1);
        mbar_wait(&bar_full[b], (uint32_t)((c_i / 2) & 1));
        tc_fence_after();
        if (elect_one()) {
          const uint64_t bd0 = make_desc(ring_addr + (uint32_t)b * T::DCH_BYTES, (uint32_t)H * 16u, 128);
          const int ksteps = (op == 0 ? NOP : H) / 16;
          for (int j = 0; j < ksteps; ++j)
            mma_ts(dM, aA + 8u * (uint32_t)j, bd0 + (uint64_t)((j * 2 * H * 16) >> 4), idesc, j > 0 ? 1u : 0u);
          tc_commit(&bar_empty[b]);
          tc_commit(&bar_done);
        }
        __syncwarp();
        ++c_i;
        top_up();
      }
    }
  } else {
    // ===================== epilogue warpgroup =====================
    const int row = tid, lane = tid & 31;
    const int role = JAC ? (row & 3) : 0;
    const int vl = lane & ~3;
    const uint32_t lane_off = ((uint32_t)(warp * 32)) << 16;
    const uint32_t tM = dM + lane_off, tA = aA + lane_off;
    uint32_t n_done = 0;
    for (int64_t t = 0; t < my_tiles; ++t) {
      const int64_t tile = (int64_t)blockIdx.x + t * gridDim.x;
      const int64_t m = tile * 128 + row;
      const bool valid = m < M;
      {
        float g[1] = {0.0f};
        if (valid) io.load_g(m, g);
        uint32_t pk[NOP / 2];
#pragma unroll
        for (int j = 0; j < NOP / 2; ++j) pk[j] = 0u;
        pk[0] = E::pack(g[0], 0.0f);
        tmem_store<NOP / 2>(tA, pk);
        save_cols<NOP / 2>(tile_row_ptr(ws.gout, tile, NOP, row), 0, pk);
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(&bar_ready);
      }
#pragma unroll 1
      for (int i = 0; i <= L; ++i) {
        const int l = L - i;
        // this row of the saved activations a_l (what softplus' needs), loaded under the layer's MMAs
        const uint16_t* arow = tile_row_ptr(ws.acts, (int64_t)l * ntiles + tile, H + kTileRowsExtra, row);
        uint4 av[H / 8];
#pragma unroll
        for (int g = 0; g < H / 8; ++g) av[g] = __ldg(reinterpret_cast<const uint4*>(arow + g * 64));
        mbar_wait(&bar_done, n_done & 1); n_done++;
        tc_fence_after();
        uint16_t* zrow = tile_row_ptr(ws.dz, (int64_t)l * ntiles + tile, H, row);
        const float cl = s_cl[l];
        uint32_t buf[2][32];
        TmemIO<32>::ld(tM, buf[0]);
        tc_wait_ld();
#pragma unroll
        for (int c = 0; c < H / 32; ++c) {
          if (c + 1 < H / 32) TmemIO<32>::ld(tM + 32 * (c + 1), buf[(c + 1) & 1]);
          uint32_t pk[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const uint4 w4 = av[4 * c + (q >> 2)];
            const uint32_t aw = (q & 3) == 0 ? w4.x : (q & 3) == 1 ? w4.y : (q & 3) == 2 ? w4.z : w4.w;
            float gz[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const float a = E::back((uint16_t)(aw >> (16 * h)));
              const float ga = __uint_as_float(buf[c & 1][2 * q + h]) * cl;
              const float a_v = JAC ? __shfl_sync(0xffffffffu, a, vl) : a;
              const float s = 1.0f - ex2_approx(-1.4426950408889634f * a_v);      // sigmoid(z_v) = 1 - exp(-softplus(z_v))
              float v = s * ga;
              if (JAC) {
                float pr = role == 0 ? 0.0f : ga * a;
                pr += __shfl_xor_sync(0xffffffffu, pr, 1);
                pr += __shfl_xor_sync(0xffffffffu, pr, 2);
                if (role == 0) v = fmaf(1.0f - s, pr, v);
              }
              gz[h] = fminf(fmaxf(v, -60000.0f), 60000.0f);
            }
            pk[q] = E::pack(gz[0], gz[1]);
          }
          if (i < L) TmemIO<16>::st(tA + 16 * c, pk);
          save_cols<16>(zrow, 32 * c, pk);
          if (c + 1 < H / 32) tc_wait_ld();
        }
        if (i < L) {
          tc_wait_st();
          tc_fence_before();
          mbar_arrive(&bar_ready);
        } else {
          tc_fence_before();
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

template <int FMT, bool JAC>
static int sdf_forward(const nrt_mlp_t* m, const float* p, int64_t M, float* value, float* jac, int out_act, const TrainWs& ws,
                       cudaStream_t st) {
  using NET = NetSdfShift;
  using T = SdfT<NET>;
  SaveTiles sv{ws.acts, ws.enc_raw, ws.enc_act, ws.masks, ws.ntiles};
  const size_t bytes = (size_t)T::SMEM_FWD + 1024;
  auto kern = k_sdf_train_fwd_tc<NET, FMT, JAC>;
  NRT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  const int grid = (int)std::min<int64_t>(ws.ntiles, (int64_t)2 * nrt_sm_count());
  NrtProfScope _ps(TAG_TC_TRAIN_FWD, st);
  kern<<<grid, 160, bytes, st>>>(reinterpret_cast<const uint8_t*>(m->params_tc), p, M, value, jac, out_act, sv);
  NRT_CUDA(cudaGetLastError());
  return NRT_OK;
}

template <int FMT, bool JAC, class IO>
static int sdf_backward(const MlpDev& d, IO io, int64_t M, const void* dblob, const TrainWs& ws, float* g_params, cudaStream_t st) {
  using NET = NetSdfShift;
  using T = SdfT<NET>;
  io.scale = ws.scale;
  {
    NrtProfScope _ps(TAG_TC_DGRAD, st);
    NRT_CUDA(cudaMemsetAsync(ws.scale, 0, 16, st));
    k_grad_absmax<IO, 1><<<(int)std::min<int64_t>((M + 255) / 256, 148 * 8), 256, 0, st>>>(io, M, ws.scale);
    k_grad_scale<<<1, 1, 0, st>>>(ws.scale);
    NRT_CUDA(cudaGetLastError());
  }
  {
    const size_t bytes = (size_t)T::SMEM_BWD + 1024;
    auto kern = k_sdf_dgrad_tc<NET, IO, FMT, JAC>;
    NRT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    const int grid = (int)std::min<int64_t>(ws.ntiles, (int64_t)2 * nrt_sm_count());
    NrtProfScope _ps(TAG_TC_DGRAD, st);
    kern<<<grid, 160, bytes, st>>>(reinterpret_cast<const uint8_t*>(dblob), io, M, ws);
    NRT_CUDA(cudaGetLastError());
  }
  return launch_wgrad_std<NET, FMT>(d, ws, g_params, st, reinterpret_cast<const int*>(reinterpret_cast<const uint8_t*>(dblob) + T::DY.bytes));
}

template <class NET>
static bool matches_sdf(const MlpDev& d) {
  return d.in_size == NET::IN && d.latent == NET::LAT && d.freqs == NET::F && d.hidden == NET::H && d.L == NET::L &&
         d.skip == NET::SKIP && d.out == NET::OUT && d.act == NET::ACT;
}

}  // namespace tc

using namespace tc;

// ---- internal interface used by the first-order C entry points in nrt_tc_train.cu ----
bool nrt_train_is_sdf(const MlpDev& d) { return matches_sdf<NetSdfShift>(d); }

// the dgrad blob of this net ends with the per-layer rescale exponents (L + 1 int32, padded to 16 bytes)
int64_t nrt_train_sdf_dgrad_tail_bytes(const MlpDev& d) { return (d.L + 1 + 3) / 4 * 16; }
int nrt_train_sdf_pack_tail(const MlpDev& d, void* tail, cudaStream_t st) {
  k_dgrad_layer_scales<<<d.L + 1, 256, 0, st>>>(d, reinterpret_cast<int*>(tail));
  NRT_CUDA(cudaGetLastError());
  return NRT_OK;
}

int nrt_train_sdf_forward(const nrt_mlp_t* m, int prec, int out_act, const float* x, int64_t M, float* out, const TrainWs& ws,
                          cudaStream_t st) {
  return prec == NRT_PREC_F16 ? sdf_forward<0, false>(m, x, M, out, nullptr, out_act, ws, st)
                              : sdf_forward<1, false>(m, x, M, out, nullptr, out_act, ws, st);
}

int nrt_train_sdf_backward(const MlpDev& d, int prec, int out_act, int64_t M, const float* out, const float* g_out,
                           const void* dblob, const TrainWs& ws, float* g_params, cudaStream_t st) {
  IoGrad<3, 1> io{out, g_out, nullptr, out_act, ws.scale};
  return prec == NRT_PREC_F16 ? sdf_backward<0, false>(d, io, M, dblob, ws, g_params, st)
                              : sdf_backward<1, false>(d, io, M, dblob, ws, g_params, st);
}

// ---- value + Jacobian on the tensor cores (replaces nrt_mlp_value_jac_forward / _backward under a 16-bit train precision;
//      reference: SDF.autograd_diff, shapes/sdfs.py:184-197, and autograd's double backward through it) ----
extern "C" int64_t nrt_mlp_value_jac_tc_workspace_bytes(const nrt_mlp_t* m, int64_t K) {
  MlpDev d;
  int rc = nrt_build_mlp_dev(m, &d);
  if (rc != NRT_OK) return rc;
  NRT_REQUIRE(nrt_train_is_sdf(d), "tensor-core value + Jacobian: instantiated for SphereSDF.shift (3 -> 1, 8 x 128, softplus, 32 frequencies)");
  const Layout y = make_layout(d.in_size, d.latent, d.freqs, d.hidden, d.L, d.skip, d.out);
  return (int64_t)carve_ws(y, d.hidden, d.L, 4 * K, nullptr).bytes;
}

extern "C" int nrt_mlp_value_jac_forward_tc(const nrt_mlp_t* m, int prec, const float* p, int64_t K, float* value, float* jac,
                                            void* workspace, size_t workspace_bytes, void* stream) {
  MlpDev d;
  int rc = nrt_build_mlp_dev(m, &d);
  if (rc != NRT_OK) return rc;
  NRT_REQUIRE(K >= 0, "negative K");
  if (K == 0) return NRT_OK;
  NRT_REQUIRE(p && value && jac && workspace, "nrt_mlp_value_jac_forward_tc: null pointer");
  NRT_REQUIRE(prec == NRT_PREC_F16 || prec == NRT_PREC_BF16, "tensor-core training path: prec must be F16 or BF16");
  NRT_REQUIRE(nrt_train_is_sdf(d), "tensor-core value + Jacobian: instantiated for SphereSDF.shift (3 -> 1, 8 x 128, softplus, 32 frequencies)");
  NRT_REQUIRE(m->params_tc != nullptr, "mlp.params_tc is NULL: call nrt_mlp_pack_tc (same prec) first");
  const Layout y = make_layout(d.in_size, d.latent, d.freqs, d.hidden, d.L, d.skip, d.out);
  const TrainWs ws = carve_ws(y, d.hidden, d.L, 4 * K, workspace);
  NRT_REQUIRE(workspace_bytes >= ws.bytes && ((uintptr_t)workspace & 255) == 0, "training workspace too small or not 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  return prec == NRT_PREC_F16 ? sdf_forward<0, true>(m, p, 4 * K, value, jac, NRT_OUT_NONE, ws, st)
                              : sdf_forward<1, true>(m, p, 4 * K, value, jac, NRT_OUT_NONE, ws, st);
}

extern "C" int nrt_mlp_value_jac_backward_tc(const nrt_mlp_t* m, int prec, int64_t K, const float* g_value, const float* g_jac,
                                             const void* dgrad_blob, void* workspace, size_t workspace_bytes, float* g_params,
                                             void* stream) {
  MlpDev d;
  int rc = nrt_build_mlp_dev(m, &d);
  if (rc != NRT_OK) return rc;
  NRT_REQUIRE(K >= 0, "negative K");
  if (K == 0) return NRT_OK;
  NRT_REQUIRE(g_jac && dgrad_blob && workspace && g_params, "nrt_mlp_value_jac_backward_tc: null pointer");
  NRT_REQUIRE(prec == NRT_PREC_F16 || prec == NRT_PREC_BF16, "tensor-core training path: prec must be F16 or BF16");
  NRT_REQUIRE(nrt_train_is_sdf(d), "tensor-core value + Jacobian: instantiated for SphereSDF.shift");
  const Layout y = make_layout(d.in_size, d.latent, d.freqs, d.hidden, d.L, d.skip, d.out);
  const TrainWs ws = carve_ws(y, d.hidden, d.L, 4 * K, workspace);
  NRT_REQUIRE(workspace_bytes >= ws.bytes, "training workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  IoGradJac io{g_value, g_jac, ws.scale};
  return prec == NRT_PREC_F16 ? sdf_backward<0, true>(d, io, 4 * K, dgrad_blob, ws, g_params, st)
                              : sdf_backward<1, true>(d, io, 4 * K, dgrad_blob, ws, g_params, st);
}
