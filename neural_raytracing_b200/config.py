"""Process-wide knobs of the native backend."""

# arithmetic of the fused MLP kernels where no gradient is required:
#   "f32"  : fp32 FMA in a fixed order, bit-exact against the CPU oracle (default)
#   "f16"  : tcgen05 tensor cores, fp16 operands / fp32 accumulate (networks that fit in smem)
#   "bf16" : same with bf16 operands
precision = "f32"

# networks the tensor-core path instantiates: (in, latent, freqs, hidden, layers, skip, out, act)
TC_NETS = {
    (3, 0, 16, 128, 5, 3, 65, 0),   # NeRFLE.first
    (70, 0, 16, 64, 8, 3, 3, 0),    # NeRFLE.second (point light)
    (115, 0, 16, 64, 8, 3, 3, 0),   # NeRFLE.second (environment-light code)
    (3, 0, 64, 96, 6, 3, 3, 0),     # NeuralBSDF.mlp
    (5, 0, 16, 64, 8, 3, 1, 0),     # occlusion MLP
    (3, 0, 32, 128, 8, 3, 1, 1),    # SphereSDF.shift (softplus; weights streamed through shared memory)
    (3, 0, 128, 256, 16, 3, 4, 0),  # ComposeSpatialVarying.sp_var_fn, 4 bases (K-chunk streamed, nrt_tc_wide.cu)
    (3, 0, 128, 256, 16, 3, 8, 0),  # ... 8 bases (nerf_synthetic.py)
    (3, 0, 128, 256, 16, 3, 16, 0), # ... 16 bases (dtu.py)
    (3, 0, 16, 256, 10, 3, 3, 0),   # LightField.light_field_approx (K-chunk streamed)
    (3, 32, 16, 32, 5, 3, 33, 0),   # PlainNeRF.first (per-image latent as part of the encoding operand)
    (2, 64, 16, 32, 5, 3, 3, 0),    # PlainNeRF.second
}

def tc_instantiated(key):
    """Whether the tensor-core (gradient-free) path evaluates a network of this shape: the shapes above, plus
    ComposeSpatialVarying.sp_var_fn for ANY number of bases up to 16 (the 4- and the 16-output instantiations serve the narrower
    ones; training kernels exist for 4 / 8 / 16 only, other counts train on the fp32 kernels)."""
    if key in TC_NETS:
        return True
    return tuple(key[:6]) == (3, 0, 128, 256, 16, 3) and 1 <= key[6] <= 16 and key[7] == 0


# arithmetic of the differentiable (training) MLP evaluations:
#   "f32"  : fused fp32 forward/backward kernels (default; gradients agree with float64 autograd to ~1e-6)
#   "f16"  : tcgen05 forward-with-saved-tiles + fused dgrad chain + wgrad kernel, fp16 operands with an automatic
#            power-of-two loss scale; "bf16": same with bf16 operands
train_precision = "f32"
# networks the tensor-core TRAINING path instantiates
TRAIN_TC_NETS = {
    (3, 0, 16, 128, 5, 3, 65, 0),   # NeRFLE.first
    (70, 0, 16, 64, 8, 3, 3, 0),    # NeRFLE.second (point light)
    (115, 0, 16, 64, 8, 3, 3, 0),   # NeRFLE.second (environment-light code)
    (3, 0, 64, 96, 6, 3, 3, 0),     # NeuralBSDF.mlp
    (5, 0, 16, 64, 8, 3, 1, 0),     # occlusion MLP
    # the 256-wide nets (nrt_tc_train_wide.cu; weights streamed)
    (3, 0, 128, 256, 16, 3, 4, 0), (3, 0, 128, 256, 16, 3, 8, 0), (3, 0, 128, 256, 16, 3, 16, 0),   # sp_var_fn
    (3, 0, 16, 256, 10, 3, 3, 0),                                                                      # LightField
    (3, 0, 32, 128, 8, 3, 1, 1),    # SphereSDF.shift (softplus; nrt_tc_train_sdf.cu: first order and value + Jacobian; no g_x)
}
# of those, the ones whose tensor-core backward also produces the INPUT gradient for 3..5-D (hi+lo split) inputs
TRAIN_TC_GX_NETS = {
    (3, 0, 64, 96, 6, 3, 3, 0), (5, 0, 16, 64, 8, 3, 1, 0),
    # 256-wide nets: the encoding rows of the chain as one more kernel over the saved dZ tiles (k_mlp_wide_denc_tc)
    (3, 0, 128, 256, 16, 3, 4, 0), (3, 0, 128, 256, 16, 3, 8, 0), (3, 0, 128, 256, 16, 3, 16, 0), (3, 0, 16, 256, 10, 3, 3, 0),
}


# networks whose TRAINING forward runs on the tensor cores while their backward stays the fused fp32 kernel (the forward
# kernel writes the post-activation layer inputs in the fp32 layout nrt_mlp_backward reads): the 256-wide nets
# (nrt_tc_wide.cu), NeuralBSDF.mlp and the occlusion MLP (k_mlp_tc with the SaveF32 policy)
TRAIN_TC_FWD_NETS = {
    (3, 0, 128, 256, 16, 3, 4, 0), (3, 0, 128, 256, 16, 3, 8, 0), (3, 0, 128, 256, 16, 3, 16, 0),   # sp_var_fn
    (3, 0, 16, 256, 10, 3, 3, 0),                                                                      # LightField
    (3, 0, 64, 96, 6, 3, 3, 0),                                                                        # NeuralBSDF.mlp
    (5, 0, 16, 64, 8, 3, 1, 0),                                                                        # occlusion MLP
}


def set_train_precision(p):
    global train_precision
    assert p in ("f32", "f16", "bf16"), p
    train_precision = p


def set_precision(p):
    global precision
    assert p in ("f32", "f16", "bf16"), p
    precision = p


# pathtrace (main.py:13-93) renders gradient-free frames in row blocks of up to this many rays per integrator call instead of
# the caller's chunk_size tiles (the kernels are latency-bound on the reference's default 32 x 32 = 1,024-ray tiles: a 64-step
# march takes 1.8 ms for 4,096 rays and 8.7 ms for 262,144).  0 keeps the caller's tiles.  Rays are independent, so the image
# is the same; 524,288 rays of a Direct-lit SDF frame need < 4 GB of intermediates.
max_tile_rays = 524288


def set_max_tile_rays(n):
    global max_tile_rays
    max_tile_rays = int(n)
