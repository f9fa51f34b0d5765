// tcgen05 kernel of the 256-wide SkipConnMLPs (weights streamed in K-chunks): shared by the inference entry points
// (nrt_tc_wide.cu) and the training path (nrt_tc_train.cu).  Internal header of libnrt_b200.
#pragma once
#include "tc_core.cuh"

namespace tc {

constexpr int kWideKC = 64;       // K elements per streamed chunk
constexpr int kWideNB = 4;        // ring depth

template <class NET>
struct Wide {
  static constexpr Layout Y = NET::Y;
  static constexpr int H = NET::H, L = NET::L, KE = NET::KE, F = NET::F, IN = NET::IN;
  static_assert(H == 256 && NET::ENC_CUDA && NET::ACT == NRT_ACT_LEAKY_RELU && NET::LAT == 0,
                "wide kernel: 256 hidden units, 3-D input, leaky_relu");
  // init + L hidden layers (blob ops 1 .. L+1); out <= 4: the output layer runs in the last epilogue (fp32), else it is
  // one more streamed op with N = NOP
  static constexpr int N_OPS = L + 1 + (NET::FUSE_OUT ? 0 : 1);
  static constexpr int NOP = NET::NOP;
  static constexpr int CH_BYTES = H * kWideKC * 2;          // 32 KB
  static constexpr int ENC_BYTES = KE * 128 * 2;            // encoding as an A operand
  static constexpr int BIAS_BYTES = Y.bias_floats * 4;
  static constexpr int SMEM_BYTES = ENC_BYTES + kWideNB * CH_BYTES + BIAS_BYTES;
  static_assert(SMEM_BYTES + 1024 <= 227 * 1024, "wide kernel: shared memory");
  static constexpr int op_k(int o) { return Y.opK[1 + o]; }
  static constexpr int op_chunks(int o) { return (op_k(o) + kWideKC - 1) / kWideKC; }
  static constexpr int chunks_per_tile() { int n = 0; for (int o = 0; o < N_OPS; ++o) n += op_chunks(o); return n; }
  static constexpr int CPT = chunks_per_tile();
};

// walks the (op, chunk) sequence of a tile
template <class W>
struct ChunkCursor {
  int op = 0, chunk = 0;
  __device__ __forceinline__ void advance() {
    if (++chunk == chunks_of(op)) { chunk = 0; if (++op == W::N_OPS) op = 0; }
  }
  __device__ __forceinline__ static int k_of(int o) {          // runtime op -> K (init / plain / skip layer / out)
    if (o == 0) return W::KE;
    if (o == W::L + 1) return W::H;
    return W::H + (is_skip(o - 1, 3, W::L) ? W::KE : 0);
  }
  __device__ __forceinline__ static int n_of(int o) { return o == W::L + 1 ? W::NOP : W::H; }
  __device__ __forceinline__ static int chunks_of(int o) { return (k_of(o) + kWideKC - 1) / kWideKC; }
};

// (convert_row_savef32 / convert_row_out_savef32: tc_core.cuh)
// SV: NoSave (inference), SaveF32 (fp32 activations for the fused fp32 backward) or SaveTiles (16-bit MN-major tiles +
// leaky_relu' sign masks for the tensor-core backward, see tc_core.cuh)
template <class NET, class IO, int FMT, class SV = NoSave>
__global__ void __launch_bounds__(160, 1)
k_mlp_wide_tc(const uint8_t* __restrict__ blob, IO io, int64_t M, SV sv = SV{}) {
  constexpr bool SAVEF32 = SV::kF32;
  constexpr bool SAVET = SV::kOn;
  using W = Wide<NET>;
  using E = Elem<FMT>;
  constexpr Layout Y = NET::Y;
  constexpr int H = 256, L = NET::L, KE = NET::KE, F = NET::F, IN = NET::IN, XR = NET::XR;
  static_assert(NET::SKIP == 3, "skip period");
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sEnc = smem;                                        // [KE/8][128][8] 16-bit
  uint8_t* sRing = smem + W::ENC_BYTES;
  const float* sBias = reinterpret_cast<const float*>(smem + W::ENC_BYTES + kWideNB * W::CH_BYTES);
  __shared__ __align__(8) uint64_t bar_full[kWideNB], bar_empty[kWideNB], bar_ready, bar_done, bar_bias;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int64_t ntiles = (M + 127) / 128;
  const int64_t my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (tid == 0) {
    for (int i = 0; i < kWideNB; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
    mbar_init(&bar_ready, 128); mbar_init(&bar_done, 1); mbar_init(&bar_bias, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    if ((tid & 31) == 0) {
      mbar_expect_tx(&bar_bias, (uint32_t)W::BIAS_BYTES);
      bulk_g2s(const_cast<float*>(sBias), blob + (size_t)Y.w_elems * 2, (uint32_t)W::BIAS_BYTES, &bar_bias);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tmem_base_s != 0u) __trap();
  constexpr uint32_t dD = 0, aU = 256;       // TMEM: accumulator columns 0..255, hidden operand columns 256..383
  mbar_wait(&bar_bias, 0);

  if (warp == 4) {
    // ===================== producer + MMA issuer =====================
    const int64_t total_chunks = my_tiles * W::CPT;
    ChunkCursor<W> pc;                        // producer cursor
    int64_t p_i = 0, c_i = 0;
    uint32_t n_ready = 0;
    const uint32_t ring_addr = smem_u32(sRing), enc_addr = smem_u32(sEnc);
    constexpr uint32_t idesc_h = (1u << 4) | ((uint32_t)FMT << 7) | ((uint32_t)FMT << 10) | ((uint32_t)(H >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t idesc_o = (1u << 4) | ((uint32_t)FMT << 7) | ((uint32_t)FMT << 10) | ((uint32_t)(W::NOP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    auto top_up = [&]() {
      while (p_i < total_chunks && p_i - c_i < kWideNB) {
        const int b = (int)(p_i % kWideNB);
        if (p_i >= kWideNB) mbar_wait(&bar_empty[b], (uint32_t)((p_i / kWideNB - 1) & 1));
        if (elect_one()) {
          const int K = ChunkCursor<W>::k_of(pc.op);
          const int k0 = pc.chunk * kWideKC;
          const int N = ChunkCursor<W>::n_of(pc.op);
          const uint32_t bytes = (uint32_t)(min(kWideKC, K - k0) * N * 2);
          const uint8_t* src = blob + (size_t)Y.op_off[1 + pc.op] * 2 + (size_t)(k0 / 8) * N * 16;
          mbar_expect_tx(&bar_full[b], bytes);
          bulk_g2s(sRing + (size_t)b * W::CH_BYTES, src, bytes, &bar_full[b]);
        }
        __syncwarp();
        pc.advance();
        ++p_i;
      }
    };
    for (int64_t t = 0; t < my_tiles; ++t) {
      for (int op = 0; op < W::N_OPS; ++op) {
        top_up();                                        // keep the ring full while the epilogue runs
        mbar_wait(&bar_ready, n_ready & 1); n_ready++;
        tc_fence_after();
        const int K = ChunkCursor<W>::k_of(op);
        const int N = ChunkCursor<W>::n_of(op);
        const uint32_t idesc = (op == L + 1) ? idesc_o : idesc_h;
        const int nch = (K + kWideKC - 1) / kWideKC;
        for (int c = 0; c < nch; ++c) {
          top_up();
          const int b = (int)(c_i % kWideNB);
          mbar_wait(&bar_full[b], (uint32_t)((c_i / kWideNB) & 1));
          tc_fence_after();
          if (elect_one()) {
            const int k0 = c * kWideKC, kc = min(kWideKC, K - k0);
            const uint64_t bd0 = make_desc(ring_addr + (uint32_t)b * W::CH_BYTES, (uint32_t)N * 16u, 128);
            for (int j = 0; j < kc / 16; ++j) {
              const int k = k0 + 16 * j;                  // K index inside the op: [hidden 256 | encoding KE], init: encoding only
              const uint64_t bd = bd0 + (uint64_t)((j * 2 * N * 16) >> 4);
              const int ke = (op == 0) ? k : k - H;       // index into the encoding (>= 0: SS MMA, A from shared memory)
              // every layer's accumulator was pre-loaded with its bias by the epilogue: always accumulate
              if (ke >= 0) {
                const uint64_t ad = make_desc(enc_addr + (uint32_t)(ke / 8) * 2048u, 2048u, 128u);
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
    // ===================== epilogue warpgroup: thread = row = sample =====================
    const int row = tid;                      // 0..127
    const uint32_t lane_off = ((uint32_t)(warp * 32)) << 16;
    const uint32_t tD = dD + lane_off, tU = aU + lane_off;
    uint32_t n_done = 0;
    // this row's 16-byte slots in the encoding operand: k-group g at sEnc + (g*128 + row)*16
    uint4* enc_row = reinterpret_cast<uint4*>(sEnc) + row;
    for (int64_t t = 0; t < my_tiles; ++t) {
      const int64_t tile = (int64_t)blockIdx.x + t * gridDim.x;
      const int64_t m = tile * 128 + row;
      const bool valid = m < M;
      // ---- encoding (raw) -> shared memory, init bias -> accumulator ----
      {
        float x[IN];
        if (valid) io.load(m, x);
        else {
#pragma unroll
          for (int j = 0; j < IN; ++j) x[j] = 0.0f;
        }
        uint16_t v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0;
#pragma unroll
        for (int j = 0; j < IN; ++j) {
          const uint16_t hi = E::cvt(x[j]);
          v[j] = hi; v[IN + j] = E::cvt(x[j] - E::back(hi));
        }
        enc_row[0 * 128] = make_uint4((uint32_t)v[0] | ((uint32_t)v[1] << 16), (uint32_t)v[2] | ((uint32_t)v[3] << 16),
                                      (uint32_t)v[4] | ((uint32_t)v[5] << 16), (uint32_t)v[6] | ((uint32_t)v[7] << 16));
        enc_row[1 * 128] = make_uint4(0u, 0u, 0u, 0u);
        static_assert(XR == 16 && IN <= 4, "raw-x segment layout");
        const float* sB = sBias + Y.basis_f32_off;
        // sin block then cos block, 8 frequencies (one 16-byte slot) at a time
#pragma unroll 1
        for (int g = 0; g < F / 8; ++g) {
          uint32_t sp[4], cp[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float s2[2], c2[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int f = 8 * g + 2 * e + h;
              float p = x[0] * sB[f];
#pragma unroll
              for (int j = 1; j < IN; ++j) p = fmaf(x[j], sB[j * NET::FP + f], p);
              sincos_fast(p, &s2[h], &c2[h]);
            }
            sp[e] = E::pack(s2[0], s2[1]); cp[e] = E::pack(c2[0], c2[1]);
          }
          enc_row[(XR / 8 + g) * 128] = make_uint4(sp[0], sp[1], sp[2], sp[3]);
          enc_row[(XR / 8 + F / 8 + g) * 128] = make_uint4(cp[0], cp[1], cp[2], cp[3]);
        }
        if constexpr (SAVET) {
          // the raw encoding as the init layer sees it, one 16-byte feature group at a time (+ the ones row)
          uint16_t* rr = tile_row_ptr(sv.enc_raw, tile, KE + kTileRowsExtra, row);
#pragma unroll 1
          for (int g = 0; g < KE / 8; ++g) *reinterpret_cast<uint4*>(rr + g * 64) = enc_row[g * 128];
          rr[tile_elem(KE)] = one16<FMT>();
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) preload_bias<64>(tD + 64 * c, sBias + Y.bias_off[1] + 64 * c);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // st.shared -> visible to the tensor core
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(&bar_ready);
      }
      // ---- hidden layers ----
      constexpr int NHID = NET::FUSE_OUT ? L : L + 1;      // epilogues that feed another MMA op
#pragma unroll 1
      for (int st = 0; st < NHID; ++st) {
        mbar_wait(&bar_done, n_done & 1); n_done++;
        tc_fence_after();
        if constexpr (SAVEF32)
          convert_row_savef32<NET::ACT, FMT, H>(tD, tU, valid ? sv.acts + ((int64_t)st * H) * M + m : nullptr, M);
        else if constexpr (SAVET)
          convert_row<NET::ACT, FMT, H, true>(tD, tU, tile_row_ptr(sv.acts, (int64_t)st * sv.ntiles + tile, H + kTileRowsExtra, row), 0, H,
                                               sv.masks + (int64_t)st * (H / 32) * (sv.ntiles * 128) + tile * 128 + row, sv.ntiles * 128);
        else
          convert_row<NET::ACT, FMT, H>(tD, tU);
        if (st < L) {
#pragma unroll
          for (int c = 0; c < 4; ++c) preload_bias<64>(tD + 64 * c, sBias + Y.bias_off[2 + st] + 64 * c);
        } else {
          preload_bias<NET::NOP>(tD, sBias + Y.bias_off[NET::STAGES - 1]);
        }
        if (st == 0) {
          // the init layer has consumed the raw encoding: activate it in place for the skip layers
#pragma unroll 1
          for (int g = 1; g < KE / 8; ++g) {
            uint4 q = enc_row[g * 128];
            q.x = leaky_packed<FMT>(q.x); q.y = leaky_packed<FMT>(q.y); q.z = leaky_packed<FMT>(q.z); q.w = leaky_packed<FMT>(q.w);
            enc_row[g * 128] = q;
          }
          {
            // raw-x group [x_hi | x_lo | 0]: act(x_hi + x_lo), split again (act(x_hi) + act(x_lo) would be wrong when
            // the two halves differ in sign)
            const uint4 q = enc_row[0];
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
            uint16_t h[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) h[j] = (uint16_t)(w[j >> 1] >> (16 * (j & 1)));
#pragma unroll
            for (int j = 0; j < IN; ++j) {
              const float a = act_fast<NET::ACT>(E::back(h[j]) + E::back(h[IN + j]));
              const uint16_t ahi = E::cvt(a);
              h[j] = ahi; h[IN + j] = E::cvt(a - E::back(ahi));
            }
            enc_row[0] = make_uint4((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16),
                                    (uint32_t)h[4] | ((uint32_t)h[5] << 16), (uint32_t)h[6] | ((uint32_t)h[7] << 16));
          }
          if constexpr (SAVET) {
            uint16_t* ra = tile_row_ptr(sv.enc_act, tile, KE + kTileRowsExtra, row);
#pragma unroll 1
            for (int g = 0; g < KE / 8; ++g) *reinterpret_cast<uint4*>(ra + g * 64) = enc_row[g * 128];
            ra[tile_elem(KE)] = one16<FMT>();
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(&bar_ready);
      }
      // ---- output ----
      {
        mbar_wait(&bar_done, n_done & 1); n_done++;
        tc_fence_after();
        float o[NET::OUT];
        if constexpr (NET::FUSE_OUT) {
          // last hidden activations + output layer (fp32, CUDA cores)
#pragma unroll
          for (int j = 0; j < NET::OUT; ++j) o[j] = sBias[Y.bias_off[NET::STAGES - 1] + j];
          if constexpr (SAVEF32)
            convert_row_out_savef32<NET::ACT, FMT, H, NET::OUT>(tD, sBias + Y.wout_f32_off, o,
                                                                  valid ? sv.acts + ((int64_t)L * H) * M + m : nullptr, M);
          else if constexpr (SAVET)
            convert_row_out<NET::ACT, FMT, H, NET::OUT, true>(tD, sBias + Y.wout_f32_off, o,
                tile_row_ptr(sv.acts, (int64_t)L * sv.ntiles + tile, H + kTileRowsExtra, row),
                sv.masks + (int64_t)L * (H / 32) * (sv.ntiles * 128) + tile * 128 + row, sv.ntiles * 128);
          else
            convert_row_out<NET::ACT, FMT, H, NET::OUT, false>(tD, sBias + Y.wout_f32_off, o);
        } else {
          constexpr int OC = (NET::OUT + 7) / 8 * 8;
          uint32_t acc[OC];
          tmem_load<OC>(tD, acc);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < NET::OUT; ++j) o[j] = __uint_as_float(acc[j]);   // bias pre-loaded into the accumulator
        }
        if (valid) io.store(m, o);
        tc_fence_before();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(0u) : "memory");
}

template <int IN, int OUT>
struct IoPlainWide {
  const float* x; float* out; int out_act;
  // number of output columns of the network that is evaluated, when it is narrower than the instantiation's OUT (a
  // ComposeSpatialVarying with, say, 10 bases runs on the 16-output instantiation: the blob layout of out = 5..16 is the same,
  // the pack kernel zero-fills the missing rows); 0 = OUT
  int out_cols;
  __device__ __forceinline__ void load(int64_t m, float* v) const {
#pragma unroll
    for (int j = 0; j < IN; ++j) v[j] = __ldg(x + m * IN + j);
  }
  __device__ __forceinline__ void store(int64_t m, const float* o) const {
    const int nc = out_cols > 0 ? out_cols : OUT;
#pragma unroll
    for (int j = 0; j < OUT; ++j) {
      float v = o[j];
      if (out_act == NRT_OUT_SIGMOID) v = 1.0f / (1.0f + __expf(-v));
      else if (out_act == NRT_OUT_SOFTPLUS) v = v > 20.0f ? v : __logf(1.0f + __expf(v));
      else if (out_act == NRT_OUT_TANH) v = tanhf(v);
      if (j < nc) out[m * nc + j] = v;
    }
  }
};


// ComposeSpatialVarying.sp_var_fn with 4 bases (colocate.py:70-78), 8 (nerf_synthetic.py:68-75) and 16 (dtu.py:101-106);
// LightField.light_field_approx
using NetSpVar4 = Net<3, 0, 128, 256, 16, 3, 4, NRT_ACT_LEAKY_RELU>;
using NetSpVar8 = Net<3, 0, 128, 256, 16, 3, 8, NRT_ACT_LEAKY_RELU>;
using NetSpVar16 = Net<3, 0, 128, 256, 16, 3, 16, NRT_ACT_LEAKY_RELU>;
using NetLightField = Net<3, 0, 16, 256, 10, 3, 3, NRT_ACT_LEAKY_RELU>;

}  // namespace tc
