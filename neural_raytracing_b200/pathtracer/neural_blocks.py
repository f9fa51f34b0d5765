"""SkipConnMLP with the reference's constructor, attributes and semantics
(pytorch3d/pathtracer/neural_blocks.py:12-86), evaluated by the fused CUDA kernels of
libnrt_b200 whenever no gradient is required."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import config, ops
from .utils import create_fourier_basis2, fourier2


def _default_activation(x):
    # the reference's default is an in-place leaky_relu (neural_blocks.py:26); out-of-place here,
    # values are identical
    return F.leaky_relu(x)


def _activation_id(fn):
    if fn is _default_activation or fn is F.leaky_relu:
        return ops.ACT_LEAKY_RELU
    if fn is F.softplus:
        return ops.ACT_SOFTPLUS
    name = getattr(fn, "__name__", "")
    qual = getattr(fn, "__qualname__", "")
    if name == "<lambda>" and "SkipConnMLP" in qual:
        return ops.ACT_LEAKY_RELU
    return None


class SkipConnMLP(nn.Module):
    """MLP with Fourier-feature input encoding and re-concatenation of the encoding every `skip`
    layers.  Same parameters / attributes as the reference so that scripts which poke at
    `.init`, `.layers`, `.out`, `.basis_p`, `.activation` keep working."""

    def __init__(self, num_layers=8, hidden_size=64, in_size=3, out=3, skip=3, freqs=16, sigma=2 << 4,
                 device="cuda", activation=_default_activation, latent_size=0, zero_init=False,
                 xavier_init=False):
        super().__init__()
        assert type(freqs) == int
        self.in_size = in_size
        self.basis_p, map_size = create_fourier_basis2(freqs, features=in_size, freq=sigma, device=device)
        self.dim_p = map_size + latent_size
        self.skip = skip
        self.latent_size = latent_size
        widths = []
        for i in range(num_layers):
            concat = (i % skip) == 0 and i != num_layers - 1
            widths.append(hidden_size + self.dim_p if concat else hidden_size)
        self.init = nn.Linear(self.dim_p, hidden_size)
        self.layers = nn.ModuleList([nn.Linear(w, hidden_size) for w in widths])
        self.out = nn.Linear(hidden_size, out)
        # neural_blocks.py:66-71: the two inits are applied one after the other (zero first, then xavier overrides the
        # weights when both flags are set); either one zeroes the biases
        if zero_init:
            for lin in self._linears():
                nn.init.zeros_(lin.weight)
                nn.init.zeros_(lin.bias)
        if xavier_init:
            for lin in self._linears():
                nn.init.xavier_uniform_(lin.weight)
                nn.init.zeros_(lin.bias)
        self.activation = activation
        self._pack_key = None
        self._packed = None

    # ---- helpers -------------------------------------------------------------------------
    def _linears(self):
        return [self.init] + list(self.layers) + [self.out]

    def _flat_params(self):
        ps = []
        for lin in self._linears():
            ps += [lin.weight, lin.bias]
        return ps

    def packed(self) -> "ops.PackedMLP":
        """Device parameters in the C-ABI layout; rebuilt when a weight changed (optimizer step,
        load_state_dict, .to(device), manual assignment).  After training.FlatParameters re-homed the Linears into one
        flat buffer in that very layout, the packed blob IS the parameter storage: nothing is copied, only the derived
        16-bit blobs are dropped when the buffer changed (or while a CUDA graph is being captured, so that the pack
        kernels are part of the graph)."""
        flat = getattr(self, "_flat_view", None)
        if flat is not None:
            key = (flat.data_ptr(), flat._version, self.basis_p.data_ptr(), self.basis_p._version, id(self.activation))
            capturing = flat.is_cuda and torch.cuda.is_current_stream_capturing()
            if self._packed is None or self._pack_key is None or capturing or key[2:] != self._pack_key[2:]:
                act = _activation_id(self.activation)
                if act is None:
                    raise ops.NrtError("SkipConnMLP: unsupported activation %r for the fused kernels" % (self.activation,))
                basis = self.basis_p if self.basis_p.device == flat.device else self.basis_p.to(flat.device)
                self.basis_p = basis
                self._packed = ops.PackedMLP(self.in_size, self.latent_size, basis.shape[-1], self.init.out_features,
                                             len(self.layers), self.skip, self.out.out_features, act,
                                             basis.detach().contiguous().float(), flat.detach())
            elif key != self._pack_key:
                self._packed.tc_blobs.clear(); self._packed.dgrad_blobs.clear(); self._packed._params_nk = None
            self._pack_key = key
            return self._packed
        ps = self._flat_params()
        basis = self.basis_p
        if basis.device != ps[0].device:   # basis_p is a plain attribute: .to() does not move it
            basis = basis.to(ps[0].device)
            self.basis_p = basis
        key = tuple((p.data_ptr(), p._version) for p in ps) + (basis.data_ptr(), basis._version, id(self.activation))
        if key != self._pack_key:
            act = _activation_id(self.activation)
            if act is None:
                raise ops.NrtError("SkipConnMLP: unsupported activation %r for the fused kernels "
                                   "(supported: the default leaky_relu, F.leaky_relu, F.softplus)" % (self.activation,))
            flat = ops.PackedMLP.pack([p for p in ps[0::2]], [p for p in ps[1::2]])
            self._packed = ops.PackedMLP(self.in_size, self.latent_size, (basis.shape[-1]), self.init.out_features,
                                         len(self.layers), self.skip, self.out.out_features, act,
                                         basis.detach().contiguous().float(), flat)
            self._pack_key = key
        return self._packed

    def invalidate_packed(self):
        """Drops the cached device copies of the parameters (packed-f32 blob, tensor-core blobs, transposed weights).
        The cache key is (data_ptr, _version) of every parameter: call this after a weight update that does not bump
        tensor versions -- writes through `.data`, `p.set_()`, replays of a CUDA graph that contains the optimizer
        step (training.GraphedStep does it itself)."""
        self._pack_key = None
        self._packed = None

    def _shape_key(self):
        m = self
        return (m.in_size, m.latent_size, m.basis_p.shape[-1], m.init.out_features, len(m.layers), m.skip,
                m.out.out_features, _activation_id(m.activation))

    def precision(self):
        """Arithmetic used for gradient-free evaluation."""
        if config.precision == "f32":
            return "f32"
        m = self
        key = (m.in_size, m.latent_size, m.basis_p.shape[-1], m.init.out_features, len(m.layers), m.skip,
               m.out.out_features, _activation_id(m.activation))
        return config.precision if config.tc_instantiated(key) else "f32"

    def train_precision(self):
        """Arithmetic of the differentiable evaluation (config.train_precision where the tensor-core training
        kernels are instantiated for this shape, else fp32)."""
        if config.train_precision == "f32" or self.latent_size:
            return "f32"
        m = self
        key = (m.in_size, m.latent_size, m.basis_p.shape[-1], m.init.out_features, len(m.layers), m.skip,
               m.out.out_features, _activation_id(m.activation))
        return config.train_precision if key in config.TRAIN_TC_NETS else "f32"

    def train_forward_precision(self):
        """Arithmetic of the training FORWARD of a network whose backward is the fused fp32 kernel."""
        if config.train_precision == "f32" or self.latent_size:
            return "f32"
        m = self
        key = (m.in_size, m.latent_size, m.basis_p.shape[-1], m.init.out_features, len(m.layers), m.skip,
               m.out.out_features, _activation_id(m.activation))
        return config.train_precision if key in config.TRAIN_TC_FWD_NETS else "f32"

    def _needs_grad(self, *tensors):
        if not torch.is_grad_enabled():
            return False
        if any(t is not None and t.requires_grad for t in tensors):
            return True
        return any(p.requires_grad for p in self.parameters())

    # ---- forward -------------------------------------------------------------------------
    def forward_reference_ops(self, p, latent=None):
        """Differentiable (any order) evaluation with plain torch ops; same op sequence as
        neural_blocks.py:75-86.  Used only where autograd must see inside the network."""
        lead = p.shape[:-1]
        enc = fourier2(p.reshape(-1, self.in_size), self.basis_p)
        if latent is not None:
            enc = torch.cat([enc, latent.reshape(-1, self.latent_size)], dim=-1)
        h = self.init(enc)
        last = len(self.layers) - 1
        for i, lin in enumerate(self.layers):
            if i != last and (i % self.skip) == 0:
                h = torch.cat([h, enc], dim=-1)
            h = lin(self.activation(h))
        return self.out(self.activation(h)).reshape(lead + (self.out.out_features,))

    def forward(self, p, latent=None, out_act=ops.OUT_NONE):
        if p.is_cuda and not self._needs_grad(p, latent):
            return ops.mlp_forward(self.packed(), p.detach().float(), None if latent is None else latent.detach().float(),
                                   out_act=out_act, prec=self.precision())
        if p.is_cuda and _FUSED_BACKWARD[0] and not getattr(self, "_higher_order", False) and \
                _activation_id(self.activation) is not None:
            return _FusedMLP.apply(self, p, latent, out_act, *self._flat_params())
        y = self.forward_reference_ops(p, latent)
        if out_act == ops.OUT_SIGMOID:
            y = y.sigmoid()
        elif out_act == ops.OUT_SOFTPLUS:
            y = F.softplus(y)
        elif out_act == ops.OUT_TANH:
            y = y.tanh()
        return y


# first-order autograd through the fused forward/backward kernels (enabled once nrt_mlp_backward exists)
_FUSED_BACKWARD = [True]


class _FusedMLP(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, p, latent, out_act, *params):
        pk = module.packed()
        x = p.detach().float().contiguous()
        lat = None if latent is None else latent.detach().float().contiguous()
        ctx.pk, ctx.out_act = pk, out_act
        ctx.has_latent = latent is not None
        ctx.need_in = p.requires_grad or (latent is not None and latent.requires_grad)
        ctx.tc_prec = module.train_precision()
        ctx.flat_grad = getattr(module, "_flat_grad", None)
        if ctx.tc_prec != "f32" and (not ctx.need_in or pk.in_size > 5 or module._shape_key() in config.TRAIN_TC_GX_NETS):
            # tensor-core training path: the workspace holds the saved activation tiles
            out, ws = ops.mlp_forward_train_tc(pk, x, out_act=out_act, prec=ctx.tc_prec)
            ctx.lead = p.shape[:-1]
            ctx.save_for_backward(out, ws)
            return out.reshape(p.shape[:-1] + (pk.out_size,))
        ctx.tc_prec = "f32"
        # 256-wide nets under a 16-bit train precision: tensor-core forward that saves its (operand-rounded)
        # activations for the fused fp32 backward; everything else: the exact fp32 forward
        fwd_prec = module.train_forward_precision() if lat is None else "f32"
        out, acts = ops.mlp_forward(pk, x, lat, out_act=out_act, prec=fwd_prec, save_acts=True)
        ctx.save_for_backward(x, lat if lat is not None else x.new_empty(0), out, acts)
        return out

    @staticmethod
    def backward(ctx, g):
        if ctx.tc_prec != "f32":
            out, ws = ctx.saved_tensors
            M = out.shape[0]
            g_params, g_x = ops.mlp_backward_tc(ctx.pk, M, out, g.contiguous().float().reshape(M, -1), ws,
                                                out_act=ctx.out_act, need_input_grad=ctx.need_in, prec=ctx.tc_prec,
                                                g_params=ctx.flat_grad)
            if ctx.flat_grad is not None:    # accumulated straight into the flat gradient buffer (training.FlatParameters)
                gx = None if g_x is None else g_x.reshape(ctx.lead + (ctx.pk.in_size,))
                return (None, gx, None, None) + (None,) * (2 * len(ctx.pk.dims))
            gW, gb = ctx.pk.unpack(g_params)
            flat = []
            for w, b in zip(gW, gb):
                flat += [w, b]
            gx = None if g_x is None else g_x.reshape(ctx.lead + (ctx.pk.in_size,))
            return (None, gx, None, None) + tuple(flat)
        x, lat, out, acts = ctx.saved_tensors
        lat = lat if ctx.has_latent else None
        g_params, g_x, g_lat = ops.mlp_backward(ctx.pk, x, lat, out, acts, g.contiguous().float(), out_act=ctx.out_act,
                                                need_input_grad=ctx.need_in, g_params=ctx.flat_grad)
        if ctx.flat_grad is not None:
            gx = None if g_x is None else g_x.reshape(x.shape)
            g_lat = None if g_lat is None else g_lat.reshape(lat.shape)
            return (None, gx, g_lat, None) + (None,) * (2 * len(ctx.pk.dims))
        gW, gb = ctx.pk.unpack(g_params)
        flat = []
        for w, b in zip(gW, gb):
            flat += [w, b]
        gx = None if g_x is None else g_x.reshape(x.shape)
        if g_lat is not None:
            g_lat = g_lat.reshape(lat.shape)          # the kernels see [M, latent]; autograd wants the caller's shape
        return (None, gx, g_lat, None) + tuple(flat)
