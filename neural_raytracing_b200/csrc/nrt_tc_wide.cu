// tcgen05 forward of the 256-wide SkipConnMLPs (ComposeSpatialVarying.sp_var_fn 16x256 with 128 frequencies,
// bsdfs.py:487-496; LightField.light_field_approx 10x256, lights.py:159-164).
//
// These networks fit neither in shared memory (2.9 MB / 1.4 MB of 16-bit weights; a single 515 x 256 operand is 270 KB)
// nor, with two tiles in flight, in TMEM.  So: ONE 128-sample tile per CTA, accumulator [128 x 256] fp32 (256 TMEM
// columns) + hidden activations as the 16-bit TMEM A operand (128 columns), and
//   * the weights are STREAMED from L2 in K-chunks of 64 (32 KB, the canonical K-major tiles of the blob are
//     contiguous in K) through a 4-deep shared-memory ring: one warp is producer (cp.async.bulk on `full` mbarriers)
//     and MMA issuer (tcgen05.commit on `empty` mbarriers frees a buffer); it runs ahead across layer and tile
//     boundaries, so the next layer's first chunks land while the epilogue of the current layer runs;
//   * the (activated) Fourier encoding, which every third layer re-reads, lives in SHARED memory in the canonical A
//     layout ([k/8][128 rows][8]): its K-steps are SS MMAs, the hidden K-steps TS MMAs (A from TMEM);
//   * phases, sin / cos (3 x F FMAs per sample) and the 4-wide output layer run on the CUDA cores in fp32.
// L2 -> SM traffic is the whole weight set per tile (22.7 KB per sample for sp_var): ~36k cycles per tile at ~80 B/clk,
// below the ~52k cycles of tensor time, so the kernel is tensor-bound with the stream hidden.
#include "tc_wide.cuh"

namespace tc {

template <class NET>
static bool matches_w(const MlpDev& d) {
  return d.in_size == NET::IN && d.latent == NET::LAT && d.freqs == NET::F && d.hidden == NET::H && d.L == NET::L &&
         d.skip == NET::SKIP && d.out == NET::OUT && d.act == NET::ACT;
}

// every field but the output width (the sp_var family: one instantiation serves every basis count up to its own)
template <class NET>
static bool matches_w_upto(const MlpDev& d, int out_lo) {
  return d.in_size == NET::IN && d.latent == NET::LAT && d.freqs == NET::F && d.hidden == NET::H && d.L == NET::L &&
         d.skip == NET::SKIP && d.out >= out_lo && d.out <= NET::OUT && d.act == NET::ACT;
}

template <class NET, int FMT, bool SAVE>
static int launch_wide(const nrt_mlp_t* m, int out_act, const float* x, int64_t M, float* out, float* acts, cudaStream_t st) {
  using SVP = typename std::conditional<SAVE, SaveF32, NoSave>::type;
  using W = Wide<NET>;
  IoPlainWide<NET::IN, NET::OUT> io{x, out, out_act, m->out_size == NET::OUT ? 0 : m->out_size};
  const size_t bytes = (size_t)W::SMEM_BYTES + 1024;
  const int64_t ntiles = (M + 127) / 128;
  const int grid = (int)std::min<int64_t>(ntiles, (int64_t)nrt_sm_count());
  auto kern = k_mlp_wide_tc<NET, decltype(io), FMT, SVP>;
  SVP sv{};
  if constexpr (SAVE) { sv.acts = acts; sv.M = M; }
  NRT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  NrtProfScope _ps(TAG_TC_MLP_WIDE, st);
  kern<<<grid, 160, bytes, st>>>(reinterpret_cast<const uint8_t*>(m->params_tc), io, M, sv);
  NRT_CUDA(cudaGetLastError());
  return NRT_OK;
}

template <class NET>
static int forward_wide(const nrt_mlp_t* m, int prec, int out_act, const float* x, int64_t M, float* out, float* acts,
                        cudaStream_t st) {
  if (prec == NRT_PREC_BF16)
    return acts ? launch_wide<NET, 1, true>(m, out_act, x, M, out, acts, st) : launch_wide<NET, 1, false>(m, out_act, x, M, out, nullptr, st);
  return acts ? launch_wide<NET, 0, true>(m, out_act, x, M, out, acts, st) : launch_wide<NET, 0, false>(m, out_act, x, M, out, nullptr, st);
}

}  // namespace tc

using namespace tc;

// returns NRT_E_UNSUPPORTED (without setting an error) when the shape is not one of the wide networks
// (acts, optional: the post-activation layer inputs for nrt_mlp_backward, [(num_layers + 1) * 256][M] floats)
int nrt_mlp_forward_tc_wide(const nrt_mlp_t* m, const MlpDev& d, int prec, int out_act, const float* x, int64_t M, float* out,
                            float* acts, cudaStream_t st, bool* handled) {
  *handled = true;
  if (matches_w<NetSpVar4>(d)) return forward_wide<NetSpVar4>(m, prec, out_act, x, M, out, acts, st);
  if (matches_w<NetSpVar8>(d)) return forward_wide<NetSpVar8>(m, prec, out_act, x, M, out, acts, st);
  if (matches_w<NetSpVar16>(d)) return forward_wide<NetSpVar16>(m, prec, out_act, x, M, out, acts, st);
  if (matches_w<NetLightField>(d)) return forward_wide<NetLightField>(m, prec, out_act, x, M, out, acts, st);
  // any other basis count of ComposeSpatialVarying.sp_var_fn up to 16: the next wider instantiation (same blob layout within
  // 1..4 and within 5..16; the missing output rows are zero in the blob and are not stored)
  if (matches_w_upto<NetSpVar4>(d, 1)) return forward_wide<NetSpVar4>(m, prec, out_act, x, M, out, acts, st);
  if (matches_w_upto<NetSpVar16>(d, 5)) return forward_wide<NetSpVar16>(m, prec, out_act, x, M, out, acts, st);
  *handled = false;
  return NRT_OK;
}
