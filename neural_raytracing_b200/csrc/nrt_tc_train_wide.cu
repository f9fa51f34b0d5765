// Tensor-core TRAINING path of the 256-wide SkipConnMLPs (ComposeSpatialVarying.sp_var_fn 16x256, bsdfs.py:487-496;
// LightField.light_field_approx 10x256, lights.py:159-164): what loss.backward() does to them in the reference's
// training loops (training_utils.py:211-260), as
//   forward  k_mlp_wide_tc<.., SaveTiles> (tc_wide.cuh): the streamed-weight forward, which also writes a_l, the raw and
//            the activated encoding as MN-major UMMA tiles + the leaky_relu' sign masks;
//   dgrad    k_mlp_wide_dgrad_tc (here): g_out -> dZ_L -> ... -> dZ_0.  These nets are evaluated at the hit points,
//            which carry no gradient (the march is no_grad, sdfs.py:119-131), so there is no input gradient and the
//            encoding rows of the skip layers drop out of the chain: every op is [128 x 256] . [256 x 256].
//            The transposed weights are streamed from L2 as 128 x 128 QUARTERS (32 KB, 4-deep ring); the fp32
//            accumulator (256 TMEM columns) is split into two N-halves with their own `done` barriers and the 16-bit
//            dZ operand is double buffered (2 x 128 columns: 512 columns in all), so that the epilogue of one half
//            (leaky_relu' mask, pack, tcgen05.st, tile store) runs under the MMAs of the other half and the first
//            quarter of the next layer starts as soon as the first half of its operand exists;
//   wgrad    k_mlp_wgrad_tc (tc_train.cuh): one job per Linear, 128-unit half and source (hidden activations / encoding).
#include "tc_train.cuh"

namespace tc {

template <class NET>
struct WideD {
  static constexpr int H = 256, L = NET::L, NOP = NET::NOP;
  static constexpr int NB = 4;                            // ring depth
  static constexpr int CH_BYTES = 128 * 128 * 2;          // one quarter: B'[128 inputs][128 units]
  static constexpr int C0_BYTES = 128 * NOP * 2;          // output layer: B'[128 hidden units][NOP outputs]
  static constexpr int CPT = 2 + 4 * L;                   // chunks per tile
  static constexpr int SMEM_BYTES = NB * CH_BYTES;
  static_assert(NET::H == 256 && NET::ACT == NRT_ACT_LEAKY_RELU && NET::LAT == 0, "wide dgrad: 256 hidden units, leaky_relu");
};
__host__ __device__ constexpr int wide_dgrad_elems(int L, int nop) { return 256 * nop + L * 65536; }
// the blob ends with L + 1 int32 exponents lexp[l] (l = 0..L): dZ_l is multiplied by 2^lexp[l] when it leaves the
// accumulator, so that a deep chain whose layers shrink (or amplify) the gradient stays inside fp16's normal range.
// 2^-lexp[l] ~ the gain of the step dZ_{l+1} -> dZ_l estimated from the weights: |dA| ~ |dZ| * ||W||_F / sqrt(fan_in rows
// used), times sqrt((1 + 0.01^2) / 2) for the leaky_relu' mask (default-initialised 256-wide layers: 0.41 per layer,
// 6e-7 over 16 layers; the loss scale alone leaves 5e5x below its target)
__host__ __device__ constexpr int wide_dgrad_bytes(int L, int nop) { return wide_dgrad_elems(L, nop) * 2 + (L + 1 + 3) / 4 * 16; }

template <int FMT>
__global__ void k_pack_dgrad_wide(MlpDev m, int nop, uint8_t* __restrict__ blob) {
  uint16_t* w = reinterpret_cast<uint16_t*>(blob);
  const int L = m.L, h = m.hidden;
  const int total = wide_dgrad_elems(L, nop);
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    float v = 0.0f;
    if (idx < 256 * nop) {
      const int nh = idx / (128 * nop), e = idx - nh * 128 * nop;
      const int n = nh * 128 + (e >> 3) % 128, k = (e >> 3) / 128 * 8 + (e & 7);
      if (k < m.out) v = m.params[m.w_off[m.n_lin - 1] + n * m.out + k];
    } else {
      const int r = idx - 256 * nop;
      const int o = 1 + r / 65536, q = (r % 65536) / 16384, e = r % 16384;
      const int n = (q >> 1) * 128 + (e >> 3) % 128, k = (q & 1) * 128 + (e >> 3) / 128 * 8 + (e & 7);
      const int li = 1 + (L - o);
      v = m.params[m.w_off[li] + n * h + k];
    }
    w[idx] = Elem<FMT>::cvt(v);
  }
}

template <class NET, class IO, int FMT>
__global__ void __launch_bounds__(160, 1)
k_mlp_wide_dgrad_tc(const uint8_t* __restrict__ blob, IO io, int64_t M, TrainWs ws) {
  using E = Elem<FMT>;
  using WD = WideD<NET>;
  constexpr int H = 256, L = NET::L, NOP = NET::NOP, NB = WD::NB;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_full[NB], bar_empty[NB], bar_ready[2], bar_done[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_cl[L + 1];                     // per-layer rescale 2^lexp[l]

  const int tid = threadIdx.x, warp = tid >> 5;
  const int64_t ntiles = (M + 127) / 128;
  const int64_t my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (tid <= L) s_cl[tid] = ldexpf(1.0f, reinterpret_cast<const int*>(blob + (size_t)wide_dgrad_elems(L, NOP) * 2)[tid]);
  if (tid == 0) {
    for (int i = 0; i < NB; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_ready[i], 128); mbar_init(&bar_done[i], 1); }
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
  if (tmem_base_s != 0u) __trap();
  // TMEM: accumulator halves at columns 0 / 128; 16-bit operand buffers at 256 / 384 (op o reads buffer o & 1)
  constexpr uint32_t dD = 0, aA = 256;

  if (warp == 4) {
    // ===================== producer + MMA issuer =====================
    const int64_t total_chunks = my_tiles * WD::CPT;
    int64_t p_i = 0, c_i = 0;
    int p_pos = 0;                                   // producer position inside the tile's chunk sequence
    uint32_t n_r0 = 0, n_r1 = 0;
    const uint32_t ring_addr = smem_u32(smem);
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)FMT << 7) | ((uint32_t)FMT << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    auto top_up = [&]() {
      while (p_i < total_chunks && p_i - c_i < NB) {
        const int b = (int)(p_i % NB);
        if (p_i >= NB) mbar_wait(&bar_empty[b], (uint32_t)((p_i / NB - 1) & 1));
        if (elect_one()) {
          const uint32_t bytes = p_pos < 2 ? (uint32_t)WD::C0_BYTES : (uint32_t)WD::CH_BYTES;
          const size_t off = p_pos < 2 ? (size_t)p_pos * WD::C0_BYTES : (size_t)2 * WD::C0_BYTES + (size_t)(p_pos - 2) * WD::CH_BYTES;
          mbar_expect_tx(&bar_full[b], bytes);
          bulk_g2s(smem + (size_t)b * WD::CH_BYTES, blob + off, bytes, &bar_full[b]);
        }
        __syncwarp();
        if (++p_pos == WD::CPT) p_pos = 0;
        ++p_i;
      }
    };
    // one chunk: KS K-steps of 16 into accumulator half `nh`, A operand columns from `a_col`
    auto issue = [&](int nh, uint32_t a_col, int ksteps, bool first_acc, uint64_t* done) {
      top_up();
      const int b = (int)(c_i % NB);
      mbar_wait(&bar_full[b], (uint32_t)((c_i / NB) & 1));
      tc_fence_after();
      if (elect_one()) {
        const uint64_t bd0 = make_desc(ring_addr + (uint32_t)b * WD::CH_BYTES, 128u * 16u, 128);
        for (int j = 0; j < ksteps; ++j)
          mma_ts(dD + 128u * (uint32_t)nh, a_col + 8u * (uint32_t)j, bd0 + (uint64_t)((j * 2 * 128 * 16) >> 4), idesc,
                 (first_acc && j == 0) ? 0u : 1u);
        tc_commit(&bar_empty[b]);
        if (done) tc_commit(done);
      }
      __syncwarp();
      ++c_i;
    };
    for (int64_t t = 0; t < my_tiles; ++t) {
      for (int op = 0; op <= L; ++op) {
        const uint32_t aRd = aA + 128u * (uint32_t)(op & 1);
        top_up();
        if (op == 0) {
          mbar_wait(&bar_ready[0], n_r0 & 1); n_r0++;
          mbar_wait(&bar_ready[1], n_r1 & 1); n_r1++;
          tc_fence_after();
          issue(0, aRd, NOP / 16, true, &bar_done[0]);
          issue(1, aRd, NOP / 16, true, &bar_done[1]);
        } else {
          mbar_wait(&bar_ready[0], n_r0 & 1); n_r0++;      // first half of dZ written, accumulator half 0 drained
          tc_fence_after();
          issue(0, aRd, 8, true, nullptr);
          mbar_wait(&bar_ready[1], n_r1 & 1); n_r1++;      // second half written, accumulator half 1 drained
          tc_fence_after();
          issue(0, aRd + 64u, 8, false, &bar_done[0]);
          issue(1, aRd, 8, true, nullptr);
          issue(1, aRd + 64u, 8, false, &bar_done[1]);
        }
      }
    }
  } else {
    // ===================== epilogue warpgroup: thread = row = sample =====================
    const int row = tid;
    const uint32_t lane_off = ((uint32_t)(warp * 32)) << 16;
    uint32_t n_d0 = 0, n_d1 = 0;
    const int64_t mpad = ntiles * 128;
    for (int64_t t = 0; t < my_tiles; ++t) {
      const int64_t tile = (int64_t)blockIdx.x + t * gridDim.x;
      const int64_t m = tile * 128 + row;
      const bool valid = m < M;
      {
        float g[NET::OUT];
        if (valid) io.load_g(m, g);
        else {
#pragma unroll
          for (int j = 0; j < NET::OUT; ++j) g[j] = 0.0f;
        }
        uint32_t pk[NOP / 2];
#pragma unroll
        for (int j = 0; j < NOP / 2; ++j)
          pk[j] = E::pack(2 * j < NET::OUT ? g[2 * j] : 0.0f, 2 * j + 1 < NET::OUT ? g[2 * j + 1] : 0.0f);
        tmem_store<NOP / 2>(aA + lane_off, pk);
        save_cols<NOP / 2>(tile_row_ptr(ws.gout, tile, NOP, row), 0, pk);
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(&bar_ready[0]);
        mbar_arrive(&bar_ready[1]);
      }
#pragma unroll 1
      for (int i = 0; i <= L; ++i) {
        const int l = L - i;
        const uint32_t aWr = aA + 128u * (uint32_t)((i + 1) & 1) + lane_off;
        uint32_t mask[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) mask[c] = __ldg(ws.masks + ((int64_t)(l * 8 + c)) * mpad + m);
        uint16_t* zrow = tile_row_ptr(ws.dz, (int64_t)l * ntiles + tile, H, row);
        const float cl = s_cl[l];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (half == 0) { mbar_wait(&bar_done[0], n_d0 & 1); n_d0++; }
          else { mbar_wait(&bar_done[1], n_d1 & 1); n_d1++; }
          tc_fence_after();
          const uint32_t src = dD + 128u * (uint32_t)half + lane_off;
          uint32_t buf[2][32];
          TmemIO<32>::ld(src, buf[0]);
          tc_wait_ld();
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (c + 1 < 4) TmemIO<32>::ld(src + 32 * (c + 1), buf[(c + 1) & 1]);
            uint32_t pk[16];
            dconvert32<FMT>(buf[c & 1], mask[4 * half + c], pk, cl);
            if (i < L) TmemIO<16>::st(aWr + 64u * (uint32_t)half + 16u * (uint32_t)c, pk);
            save_cols<16>(zrow, 128 * half + 32 * c, pk);
            if (c + 1 < 4) tc_wait_ld();
          }
          if (i < L) {
            tc_wait_st();
            tc_fence_before();
            mbar_arrive(&bar_ready[half]);
          } else {
            tc_fence_before();
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(0u) : "memory");
}

// ---------------------------------------------------------------------------------------------
// input gradient (the hit points carry one through the 5-epsilon offset along the normal, sdfs.py:196): the encoding
// rows that the chain above leaves out,
//     dEnc = act'(enc) * sum over skip layers l of dZ_{l+1} . W_l[enc rows]^T  +  dZ_0 . W_init^T,
//     g_x  = dEnc[x] + (cos * dEnc[sin] - sin * dEnc[cos]) . basis^T                       (utils.py:37-40)
// as one more kernel over the saved dZ tiles: a [128 x KE] fp32 accumulator in TMEM, A = a whole dZ tile (64 KB, one
// bulk copy, double buffered; the MN-major tile read as a K-major operand: LBO = 128 B between feature groups, SBO = the
// tile's sample-group stride), B = the encoding rows of the transposed weights streamed in K-chunks of 32 units.
// The sources carry different per-layer rescales 2^E_li; the pack kernel folds 2^(E_c - E_li) <= 1 into the weights
// (E_c = the smallest exponent among the sources: what underflows there is negligible against the E_c term).
// ---------------------------------------------------------------------------------------------
template <class NET>
struct WideE {
  static constexpr int H = 256, L = NET::L, KE = NET::KE, IN = NET::IN, F = NET::F, XR = NET::XR;
  static constexpr int KC = 32, NB = 4;
  static constexpr int CH_BYTES = KE * KC * 2;
  static constexpr int A_BYTES = H * 128 * 2;
  static constexpr int n_src() { int n = 1; for (int l = 0; l < L; ++l) if (is_skip(l, NET::SKIP, L)) ++n; return n; }
  static constexpr int NSRC = n_src();                 // skip layers (ascending l), then the init layer
  static constexpr int NA = KE > 256 ? (KE / 2 + 15) / 16 * 16 : KE;   // N of the first MMA of a K-step, rest in a second one
  static constexpr int NBK = KE - NA;
  static constexpr int BASIS_BYTES = IN * F * 4;
  static constexpr int SMEM_BYTES = 2 * A_BYTES + NB * CH_BYTES + BASIS_BYTES;
  static_assert(NSRC >= 2 && KE <= 512 && NA % 16 == 0 && NBK % 16 == 0 && NA <= 256 && NBK <= 256, "dEnc kernel shape");
  static_assert(SMEM_BYTES + 2048 <= 227 * 1024, "dEnc kernel: shared memory");
  static_assert(XR == 16 && IN <= 4, "raw-x segment layout");
};
__host__ __device__ constexpr int wide_n_src(int L, int skip) { int n = 1; for (int l = 0; l < L; ++l) if (is_skip(l, skip, L)) ++n; return n; }
// Linear index of source i (skip layers ascending, init last)
__host__ __device__ constexpr int wide_src_li(int i, int L, int skip) {
  int n = 0;
  for (int l = 0; l < L; ++l) if (is_skip(l, skip, L)) { if (n == i) return l + 1; ++n; }
  return 0;
}
__host__ __device__ constexpr int wide_denc_bytes(int L, int skip, int ke) { return wide_n_src(L, skip) * ke * 256 * 2 + 16; }

// blob: per source the canonical K-major operand B'[n = encoding column][k = unit] (N = KE, K = 256), then one float:
// 2^-E_c (what the kernel's output is multiplied with, next to 1 / loss scale)
template <int FMT>
__global__ void k_pack_denc_wide(MlpDev m, Layout y, const int* __restrict__ lexp, uint8_t* __restrict__ blob) {
  uint16_t* w = reinterpret_cast<uint16_t*>(blob);
  const int L = m.L, h = m.hidden, KE = y.KE;
  const int nsrc = wide_n_src(L, m.skip);
  // E_li = lexp[li] + ... + lexp[L]; E_c = min over the sources
  int E[NRT_MAX_LAYERS + 2];
  int run = 0;
  for (int l = L; l >= 0; --l) { run += lexp[l]; E[l] = run; }
  int Ec = E[0];
  for (int i = 0; i + 1 < nsrc; ++i) Ec = min(Ec, E[wide_src_li(i, L, m.skip)]);
  const int total = nsrc * KE * 256;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int i = idx / (KE * 256), e = idx - i * KE * 256;
    const int n = (e >> 3) % KE, k = (e >> 3) / KE * 8 + (e & 7);
    const int li = wide_src_li(i, L, m.skip);
    const int r = enc_ref_index(y, m.in_size, n);
    float v = 0.0f;
    if (r >= 0) v = m.params[m.w_off[li] + ((li == 0 ? 0 : h) + r) * h + k];
    w[idx] = Elem<FMT>::cvt(ldexpf(v, Ec - E[li]));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *reinterpret_cast<float*>(blob + (size_t)total * 2) = ldexpf(1.0f, -Ec);
}

template <class NET, int FMT>
__global__ void __launch_bounds__(192, 1)
k_mlp_wide_denc_tc(const uint8_t* __restrict__ blob, const float* __restrict__ basis, int64_t M, TrainWs ws, float* __restrict__ g_x) {
  using E = Elem<FMT>;
  using WE = WideE<NET>;
  constexpr int H = 256, L = NET::L, KE = NET::KE, IN = NET::IN, F = NET::F, XR = NET::XR, NB = WE::NB, NSRC = WE::NSRC;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + 2 * WE::A_BYTES;
  float* sBasis = reinterpret_cast<float*>(smem + 2 * WE::A_BYTES + NB * WE::CH_BYTES);   // [IN][F]
  __shared__ __align__(8) uint64_t a_full[2], a_empty[2], b_full[NB], b_empty[NB], bar_done_skip, bar_ready_init, bar_done_all, bar_acc_free;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int64_t ntiles = (M + 127) / 128;
  const int64_t my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  for (int i = tid; i < IN * F; i += blockDim.x) sBasis[i] = basis[i];
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < NB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    mbar_init(&bar_done_skip, 1); mbar_init(&bar_done_all, 1);
    mbar_init(&bar_ready_init, 128); mbar_init(&bar_acc_free, 128);
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
  if (tmem_base_s != 0u) __trap();
  constexpr uint32_t dE = 0;
  const float unscale_c = *reinterpret_cast<const float*>(blob + (size_t)NSRC * KE * 256 * 2);

  if (warp == 5) {
    // ===================== producer: dZ tiles and weight chunks, in consumption order =====================
    int64_t a_i = 0, b_i = 0;
    for (int64_t t = 0; t < my_tiles; ++t) {
      const int64_t tile = (int64_t)blockIdx.x + t * gridDim.x;
      for (int i = 0; i < NSRC; ++i) {
        const int li = wide_src_li(i, L, NET::SKIP);
        const int ab = (int)(a_i & 1);
        if (a_i >= 2) mbar_wait(&a_empty[ab], (uint32_t)((a_i / 2 - 1) & 1));
        if (elect_one()) {
          const uint8_t* src = reinterpret_cast<const uint8_t*>(ws.dz + ((int64_t)li * ntiles + tile) * (H * 128));
          mbar_expect_tx(&a_full[ab], (uint32_t)WE::A_BYTES);
          for (uint32_t off = 0; off < (uint32_t)WE::A_BYTES; off += 32768u)
            bulk_g2s(sA + (size_t)ab * WE::A_BYTES + off, src + off, 32768u, &a_full[ab]);
        }
        __syncwarp();
        ++a_i;
        for (int c = 0; c < H / WE::KC; ++c) {
          const int b = (int)(b_i % NB);
          if (b_i >= NB) mbar_wait(&b_empty[b], (uint32_t)((b_i / NB - 1) & 1));
          if (elect_one()) {
            mbar_expect_tx(&b_full[b], (uint32_t)WE::CH_BYTES);
            bulk_g2s(sB + (size_t)b * WE::CH_BYTES, blob + ((size_t)i * (H / WE::KC) + c) * WE::CH_BYTES, (uint32_t)WE::CH_BYTES, &b_full[b]);
          }
          __syncwarp();
          ++b_i;
        }
      }
    }
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    int64_t a_i = 0, b_i = 0;
    const uint32_t sA_addr = smem_u32(sA), sB_addr = smem_u32(sB);
    constexpr uint32_t idesc_a = (1u << 4) | ((uint32_t)FMT << 7) | ((uint32_t)FMT << 10) | ((uint32_t)(WE::NA >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t idesc_b = (1u << 4) | ((uint32_t)FMT << 7) | ((uint32_t)FMT << 10) | ((uint32_t)((WE::NBK > 0 ? WE::NBK : 16) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int64_t t = 0; t < my_tiles; ++t) {
      for (int i = 0; i < NSRC; ++i) {
        if (i == 0 && t > 0) { mbar_wait(&bar_acc_free, (uint32_t)((t - 1) & 1)); tc_fence_after(); }
        if (i == NSRC - 1) { mbar_wait(&bar_ready_init, (uint32_t)(t & 1)); tc_fence_after(); }
        const int ab = (int)(a_i & 1);
        mbar_wait(&a_full[ab], (uint32_t)((a_i / 2) & 1));
        tc_fence_after();
        for (int c = 0; c < H / WE::KC; ++c) {
          const int b = (int)(b_i % NB);
          mbar_wait(&b_full[b], (uint32_t)((b_i / NB) & 1));
          tc_fence_after();
          if (elect_one()) {
            // A: K-major view of the MN-major dZ tile: core matrix (sample group, feature group) at sg * H*16 + fg * 128
            const uint64_t ad0 = make_desc(sA_addr + (uint32_t)ab * WE::A_BYTES + (uint32_t)(c * WE::KC / 8) * 128u, 128u, (uint32_t)H * 16u);
            const uint64_t bd0 = make_desc(sB_addr + (uint32_t)b * WE::CH_BYTES, (uint32_t)KE * 16u, 128);
#pragma unroll
            for (int j = 0; j < WE::KC / 16; ++j) {
              const uint32_t acc = (i == 0 && c == 0 && j == 0) ? 0u : 1u;
              const uint64_t ad = ad0 + (uint64_t)((j * 2 * 128) >> 4);
              const uint64_t bd = bd0 + (uint64_t)((j * 2 * KE * 16) >> 4);
              asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                           "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                           ::"r"(dE), "l"(ad), "l"(bd), "r"(idesc_a), "r"(acc) : "memory");
              if constexpr (WE::NBK > 0)
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                             ::"r"(dE + (uint32_t)WE::NA), "l"(ad), "l"(bd + (uint64_t)((WE::NA / 8 * 128) >> 4)), "r"(idesc_b), "r"(acc) : "memory");
            }
            tc_commit(&b_empty[b]);
            if (c == H / WE::KC - 1) {
              tc_commit(&a_empty[ab]);
              if (i == NSRC - 2) tc_commit(&bar_done_skip);
              if (i == NSRC - 1) tc_commit(&bar_done_all);
            }
          }
          __syncwarp();
          ++b_i;
        }
        ++a_i;
      }
    }
  } else {
    // ===================== epilogue warpgroup: thread = row = sample =====================
    const int row = tid;
    const uint32_t lane_off = ((uint32_t)(warp * 32)) << 16;
    const uint32_t tE = dE + lane_off;
    for (int64_t t = 0; t < my_tiles; ++t) {
      const int64_t tile = (int64_t)blockIdx.x + t * gridDim.x;
      const int64_t m = tile * 128 + row;
      const uint16_t* er = tile_row_ptr(ws.enc_raw, tile, KE + kTileRowsExtra, row);
      // ---- sum over the skip layers -> through act'(enc) (sign of the raw encoding), in place ----
      mbar_wait(&bar_done_skip, (uint32_t)(t & 1));
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < KE / 16; ++c) {
        uint32_t e[16];
        TmemIO<16>::ld(tE + 16 * c, e);
        const uint4 r0 = *reinterpret_cast<const uint4*>(er + (2 * c) * 64), r1 = *reinterpret_cast<const uint4*>(er + (2 * c + 1) * 64);
        const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const uint32_t sign = (rw[j >> 1] >> (16 * (j & 1) + 15)) & 1u;
          if (sign) e[j] = __float_as_uint(0.01f * __uint_as_float(e[j]));
        }
        TmemIO<16>::st(tE + 16 * c, e);
      }
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&bar_ready_init);
      // ---- + init layer -> g_x ----
      mbar_wait(&bar_done_all, (uint32_t)(t & 1));
      tc_fence_after();
      float g[IN];
      {
        uint32_t a[8];
        TmemIO<8>::ld(tE, a);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < IN; ++j) g[j] = __uint_as_float(a[j]);
      }
#pragma unroll 1
      for (int c = 0; c < F / 8; ++c) {
        uint32_t ds[8], dc[8];
        TmemIO<8>::ld(tE + XR + 8 * c, ds);
        TmemIO<8>::ld(tE + XR + F + 8 * c, dc);
        const uint4 sv = *reinterpret_cast<const uint4*>(er + (XR / 8 + c) * 64), cv = *reinterpret_cast<const uint4*>(er + ((XR + F) / 8 + c) * 64);
        const uint32_t sw[4] = {sv.x, sv.y, sv.z, sv.w}, cw[4] = {cv.x, cv.y, cv.z, cv.w};
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float sn = E::back((uint16_t)(sw[j >> 1] >> (16 * (j & 1)))), cs = E::back((uint16_t)(cw[j >> 1] >> (16 * (j & 1))));
          const float q = cs * __uint_as_float(ds[j]) - sn * __uint_as_float(dc[j]);
#pragma unroll
          for (int i = 0; i < IN; ++i) g[i] = fmaf(q, sBasis[i * F + 8 * c + j], g[i]);
        }
      }
      tc_fence_before();
      mbar_arrive(&bar_acc_free);
      if (m < M) {
        const float us = ws.scale[2] * unscale_c;
#pragma unroll
        for (int j = 0; j < IN; ++j) g_x[m * IN + j] = g[j] * us;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(0u) : "memory");
}

template <class NET>
static bool matches_wt(const MlpDev& d) {
  return d.in_size == NET::IN && d.latent == NET::LAT && d.freqs == NET::F && d.hidden == NET::H && d.L == NET::L &&
         d.skip == NET::SKIP && d.out == NET::OUT && d.act == NET::ACT;
}

template <class NET, int FMT>
static int wide_train_forward(const nrt_mlp_t* m, int out_act, const float* x, int64_t M, float* out, const TrainWs& ws, cudaStream_t st) {
  using W = Wide<NET>;
  IoPlainWide<NET::IN, NET::OUT> io{x, out, out_act, 0};
  SaveTiles sv{ws.acts, ws.enc_raw, ws.enc_act, ws.masks, ws.ntiles};
  const size_t bytes = (size_t)W::SMEM_BYTES + 1024;
  const int grid = (int)std::min<int64_t>(ws.ntiles, (int64_t)nrt_sm_count());
  auto kern = k_mlp_wide_tc<NET, decltype(io), FMT, SaveTiles>;
  NRT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  NrtProfScope _ps(TAG_TC_TRAIN_FWD, st);
  kern<<<grid, 160, bytes, st>>>(reinterpret_cast<const uint8_t*>(m->params_tc), io, M, sv);
  NRT_CUDA(cudaGetLastError());
  return NRT_OK;
}

template <class NET, int FMT>
static int wide_train_backward(const MlpDev& d, int out_act, int64_t M, const float* out, const float* g_out, const void* dblob,
                               const TrainWs& ws, float* g_params, float* g_x, cudaStream_t st) {
  using WD = WideD<NET>;
  IoGrad<NET::IN, NET::OUT> io{out, g_out, nullptr, out_act, ws.scale};
  {
    NrtProfScope _ps(TAG_TC_DGRAD, st);
    NRT_CUDA(cudaMemsetAsync(ws.scale, 0, 16, st));
    k_grad_absmax<decltype(io), NET::OUT><<<(int)std::min<int64_t>((M * NET::OUT + 255) / 256, 148 * 8), 256, 0, st>>>(io, M, ws.scale);
    k_grad_scale<<<1, 1, 0, st>>>(ws.scale);
    NRT_CUDA(cudaGetLastError());
  }
  {
    const size_t bytes = (size_t)WD::SMEM_BYTES + 1024;
    auto kern = k_mlp_wide_dgrad_tc<NET, decltype(io), FMT>;
    NRT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    const int grid = (int)std::min<int64_t>(ws.ntiles, (int64_t)nrt_sm_count());
    NrtProfScope _ps(TAG_TC_DGRAD, st);
    kern<<<grid, 160, bytes, st>>>(reinterpret_cast<const uint8_t*>(dblob), io, M, ws);
    NRT_CUDA(cudaGetLastError());
  }
  constexpr int H = 256, L = NET::L, KE = NET::KE, NOP = NET::NOP;
  if (g_x != nullptr) {
    using WE = WideE<NET>;
    const size_t bytes = (size_t)WE::SMEM_BYTES + 1024;
    auto kern = k_mlp_wide_denc_tc<NET, FMT>;
    NRT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    const int grid = (int)std::min<int64_t>(ws.ntiles, (int64_t)nrt_sm_count());
    NrtProfScope _ps(TAG_TC_DGRAD, st);
    kern<<<grid, 192, bytes, st>>>(reinterpret_cast<const uint8_t*>(dblob) + wide_dgrad_bytes(L, NOP), d.basis, M, ws, g_x);
    NRT_CUDA(cudaGetLastError());
  }
  // ---- weight gradients: per Linear one job per 128-unit half and source ----
  constexpr int FRA = H + kTileRowsExtra, FRE = KE + kTileRowsExtra;
  WgradJobs jb{};
  jb.y = NET::Y; jb.in_size = NET::IN; jb.n = 0;
  const int64_t nt = ws.ntiles;
  auto add = [&](const uint16_t* a, int a_rows, int a_tile_rows, int unit0, int n_valid, const uint16_t* s0, int s0_rows, int kind,
                 int s0_valid, int kbase, int li, bool bias) {
    WgradJob& j = jb.j[jb.n++];
    j.a_tiles = a; j.a_rows = a_rows; j.a_tile_rows = a_tile_rows; j.unit0 = unit0; j.n_valid = n_valid;
    j.s0_tiles = s0; j.s0_rows = s0_rows; j.s0_kind = kind; j.s0_valid = s0_valid; j.s0_kbase = kbase;
    j.s1_tiles = nullptr; j.s1_rows = 0; j.k_base1 = H;
    j.N = d.N[li]; j.w_off = d.w_off[li]; j.b_off = bias ? d.b_off[li] : -1;
    j.lexp_from = li;           // dZ of Linear li = dz[li] carries 2^(lexp[li] + .. + lexp[L]); g_out (li = L + 1) none
  };
  jb.lexp = reinterpret_cast<const int*>(reinterpret_cast<const uint8_t*>(dblob) + (size_t)wide_dgrad_elems(L, NOP) * 2);
  jb.n_lexp = L + 1;
  for (int li = 0; li <= L + 1; ++li) {
    if (li == L + 1) {
      add(ws.gout, NOP, NOP, 0, NET::OUT, ws.acts + (int64_t)L * nt * FRA * 128, FRA, 0, H, 0, li, true);
    } else if (li == 0) {
      for (int u = 0; u < H; u += 128) add(ws.dz, 128, H, u, 128, ws.enc_raw, FRE, 1, KE, 0, li, true);
    } else {
      const uint16_t* dz = ws.dz + (int64_t)li * nt * H * 128;
      for (int u = 0; u < H; u += 128) add(dz, 128, H, u, 128, ws.acts + (int64_t)(li - 1) * nt * FRA * 128, FRA, 0, H, 0, li, true);
      if (is_skip(li - 1, NET::SKIP, L))
        for (int u = 0; u < H; u += 128) add(dz, 128, H, u, 128, ws.enc_act, FRE, 1, KE, H, li, false);
    }
  }
  static_assert(3 + 2 * L + 2 * ((L + 2) / 3) <= kMaxJobs, "job table");
  // stage: [A'' (128 rows) | source 0]
  constexpr int S0_OFF = 128 * 256;
  constexpr int S0_BYTES = (FRA > FRE ? FRA : FRE) * 256;
  const int stage_bytes = S0_OFF + S0_BYTES;
  const size_t bytes = 2 * (size_t)stage_bytes + 8192;
  static_assert(2 * (S0_OFF + S0_BYTES) + 8192 + 1024 <= 227 * 1024, "wgrad stages do not fit in shared memory");
  static_assert(FRA <= 512 && FRE <= 512, "wgrad accumulator columns");
  auto kern = k_mlp_wgrad_tc<FMT>;
  NRT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  const int ctas = wgrad_assign_splits(jb, nt, nrt_sm_count());
  {
    NrtProfScope _ps(TAG_TC_WGRAD, st);
    kern<<<ctas, 160, bytes, st>>>(jb, nt, stage_bytes, S0_OFF, S0_OFF + S0_BYTES, g_params, ws.scale);
  }
  NRT_CUDA(cudaGetLastError());
  return NRT_OK;
}

}  // namespace tc

using namespace tc;

// ---- internal interface used by the C entry points in nrt_tc_train.cu ----
int nrt_train_wide_id(const MlpDev& d) {
  if (matches_wt<NetSpVar4>(d)) return 1;
  if (matches_wt<NetSpVar8>(d)) return 2;
  if (matches_wt<NetSpVar16>(d)) return 3;
  if (matches_wt<NetLightField>(d)) return 4;
  return 0;
}

int64_t nrt_train_wide_dgrad_blob_bytes(const MlpDev& d, bool need_x) {
  const Layout y = make_layout(d.in_size, d.latent, d.freqs, d.hidden, d.L, d.skip, d.out);
  return (int64_t)wide_dgrad_bytes(d.L, c16(d.out)) + (need_x ? wide_denc_bytes(d.L, d.skip, y.KE) : 0);
}

int nrt_train_wide_pack_dgrad(const MlpDev& d, int prec, bool need_x, void* blob_out, cudaStream_t st) {
  const int total = wide_dgrad_elems(d.L, c16(d.out));
  const int grid = std::min(nrt_cdiv(total, 256), 1184);
  NrtProfScope _ps(TAG_TC_PACK, st);
  if (prec == NRT_PREC_F16) k_pack_dgrad_wide<0><<<grid, 256, 0, st>>>(d, c16(d.out), (uint8_t*)blob_out);
  else k_pack_dgrad_wide<1><<<grid, 256, 0, st>>>(d, c16(d.out), (uint8_t*)blob_out);
  int* lexp = reinterpret_cast<int*>((uint8_t*)blob_out + (size_t)total * 2);
  k_dgrad_layer_scales<<<d.L + 1, 256, 0, st>>>(d, lexp);
  if (need_x) {
    const Layout y = make_layout(d.in_size, d.latent, d.freqs, d.hidden, d.L, d.skip, d.out);
    uint8_t* eb = (uint8_t*)blob_out + wide_dgrad_bytes(d.L, c16(d.out));
    const int etotal = wide_n_src(d.L, d.skip) * y.KE * 256;
    const int egrid = std::min(nrt_cdiv(etotal, 256), 1184);
    if (prec == NRT_PREC_F16) k_pack_denc_wide<0><<<egrid, 256, 0, st>>>(d, y, lexp, eb);
    else k_pack_denc_wide<1><<<egrid, 256, 0, st>>>(d, y, lexp, eb);
  }
  NRT_CUDA(cudaGetLastError());
  return NRT_OK;
}

#define NRT_WIDE_DISPATCH(CALL_F16, CALL_BF16) (prec == NRT_PREC_F16 ? (CALL_F16) : (CALL_BF16))

int nrt_train_wide_forward(const nrt_mlp_t* m, const MlpDev& d, int prec, int out_act, const float* x, int64_t M, float* out,
                           const TrainWs& ws, cudaStream_t st) {
  switch (nrt_train_wide_id(d)) {
    case 1: return NRT_WIDE_DISPATCH((wide_train_forward<NetSpVar4, 0>(m, out_act, x, M, out, ws, st)), (wide_train_forward<NetSpVar4, 1>(m, out_act, x, M, out, ws, st)));
    case 2: return NRT_WIDE_DISPATCH((wide_train_forward<NetSpVar8, 0>(m, out_act, x, M, out, ws, st)), (wide_train_forward<NetSpVar8, 1>(m, out_act, x, M, out, ws, st)));
    case 3: return NRT_WIDE_DISPATCH((wide_train_forward<NetSpVar16, 0>(m, out_act, x, M, out, ws, st)), (wide_train_forward<NetSpVar16, 1>(m, out_act, x, M, out, ws, st)));
    case 4: return NRT_WIDE_DISPATCH((wide_train_forward<NetLightField, 0>(m, out_act, x, M, out, ws, st)), (wide_train_forward<NetLightField, 1>(m, out_act, x, M, out, ws, st)));
  }
  nrt_set_error("tensor-core training path: this MLP shape is not instantiated");
  return NRT_E_UNSUPPORTED;
}

int nrt_train_wide_backward(const MlpDev& d, int prec, int out_act, int64_t M, const float* out, const float* g_out,
                            const void* dblob, const TrainWs& ws, float* g_params, float* g_x, cudaStream_t st) {
  switch (nrt_train_wide_id(d)) {
    case 1: return NRT_WIDE_DISPATCH((wide_train_backward<NetSpVar4, 0>(d, out_act, M, out, g_out, dblob, ws, g_params, g_x, st)), (wide_train_backward<NetSpVar4, 1>(d, out_act, M, out, g_out, dblob, ws, g_params, g_x, st)));
    case 2: return NRT_WIDE_DISPATCH((wide_train_backward<NetSpVar8, 0>(d, out_act, M, out, g_out, dblob, ws, g_params, g_x, st)), (wide_train_backward<NetSpVar8, 1>(d, out_act, M, out, g_out, dblob, ws, g_params, g_x, st)));
    case 3: return NRT_WIDE_DISPATCH((wide_train_backward<NetSpVar16, 0>(d, out_act, M, out, g_out, dblob, ws, g_params, g_x, st)), (wide_train_backward<NetSpVar16, 1>(d, out_act, M, out, g_out, dblob, ws, g_params, g_x, st)));
    case 4: return NRT_WIDE_DISPATCH((wide_train_backward<NetLightField, 0>(d, out_act, M, out, g_out, dblob, ws, g_params, g_x, st)), (wide_train_backward<NetLightField, 1>(d, out_act, M, out, g_out, dblob, ws, g_params, g_x, st)));
  }
  nrt_set_error("tensor-core training path: this MLP shape is not instantiated");
  return NRT_E_UNSUPPORTED;
}
