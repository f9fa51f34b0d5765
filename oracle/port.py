"""numpy restatements used as oracles for what the C oracle does not cover.  TEST INFRASTRUCTURE.

1. Hierarchical (coarse -> importance-resampled fine) volumetric rendering.  NOT IN THE REFERENCE
   (nerf.py:178 uses one shared uniform ts); BASELINE.json's config 2 asks for 64 coarse + 128 fine
   samples, so the extension is pinned here by a plain restatement of the standard NeRF inverse-CDF
   resampling combined with the REFERENCE's compositing formula (nerf.py:205-213), and by the
   invariant that n_fine = 0 reproduces the reference exactly.
2. The eager shading glue (interaction.py, utils.py, lights.py, bsdfs.py) lives in the package
   itself as torch ops; its oracle is the golden output of the unmodified reference.
"""
import numpy as np

from . import c_oracle

F32 = np.float32
M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def hash_u01(seed, a, b):
    """32-bit integer hash -> [0,1) with 24 bits (same integer recipe as the kernels: nrt_render.cu, hash_u01)."""
    M = np.uint64(0xFFFFFFFF)
    seed = np.uint64(seed)
    with np.errstate(over="ignore"):
        h = ((seed & M) ^ (((seed >> np.uint64(32)) * np.uint64(0x9E3779B1)) & M)) & M
        h = ((h ^ (a.astype(np.uint64) & M)) * np.uint64(0x85EBCA77)) & M
        h = ((h ^ (b.astype(np.uint64) & M)) * np.uint64(0xC2B2AE3D)) & M
        h ^= h >> np.uint64(16); h = (h * np.uint64(0x7FEB352D)) & M
        h ^= h >> np.uint64(15); h = (h * np.uint64(0x846CA68B)) & M
        h ^= h >> np.uint64(16)
    return (h >> np.uint64(8)).astype(F32) * F32(1.0 / 16777216.0)


def stratified_ts(R, S, t_near, t_far, seed, r_off=0):
    r = np.arange(R)[:, None] + r_off
    s = np.arange(S)[None, :]
    u = hash_u01(seed, np.broadcast_to(r, (R, S)), np.broadcast_to(s, (R, S))) if seed else np.full((R, S), 0.5, F32)
    return (F32(t_near) + (s.astype(F32) + u) / F32(S) * F32(t_far - t_near)).astype(F32)


def alphas(sigma_raw, ts):
    return (F32(1) - np.exp(-np.maximum(sigma_raw, 0).astype(F32) * ts.astype(F32))).astype(F32)


def sample_pdf(sigma_c, ts_c, n_fine, seed=0, r_off=0):
    """Inverse-CDF resampling over the bins between coarse mid-points, weights from the reference's
    compositing (alpha from absolute t).  sigma_c, ts_c: [R,Sc] -> ts_f [R,n_fine] (ascending)."""
    R, Sc = sigma_c.shape
    a = alphas(sigma_c, ts_c)
    x = np.maximum(F32(1) - a, F32(1e-10))
    cp_before = np.concatenate([np.ones((R, 1), F32), np.cumprod(x, axis=1, dtype=F32)[:, :-1]], axis=1)
    w = (a * cp_before + F32(1e-5))[:, 1:Sc - 1]                      # interior samples 1..Sc-2
    pdf = w / w.sum(axis=1, keepdims=True, dtype=F32)
    cdf_hi = np.cumsum(pdf, axis=1, dtype=F32)
    cdf_lo = cdf_hi - pdf
    j = np.arange(n_fine)[None, :]
    jit = hash_u01(np.uint64(seed) ^ np.uint64(0x5bd1e995), np.broadcast_to(np.arange(R)[:, None] + r_off, (R, n_fine)),
                   np.broadcast_to(j, (R, n_fine))) if seed else np.full((R, n_fine), 0.5, F32)
    u = (j.astype(F32) + jit) / F32(n_fine)
    # first interior bin whose upper cdf bound is >= u (last bin otherwise)
    b = (u[:, :, None] > cdf_hi[:, None, :]).sum(axis=2)
    b = np.minimum(b, Sc - 3)
    s = b + 1
    rows = np.arange(R)[:, None]
    lo = F32(0.5) * (ts_c[rows, s - 1] + ts_c[rows, s])
    hi = F32(0.5) * (ts_c[rows, s] + ts_c[rows, s + 1])
    f = np.clip((u - cdf_lo[rows, b]) / pdf[rows, b], 0, 1).astype(F32)
    return (lo + f * (hi - lo)).astype(F32)


def composite_ray_major(sigma_raw, rgb, ts):
    """nerf.py:205-213 on ray-major arrays [R,S], [R,S,3], [R,S] (per-ray distances)."""
    R, S = sigma_raw.shape
    a = alphas(sigma_raw, ts)
    x = np.maximum(F32(1) - a, F32(1e-10))
    cp = np.cumprod(x, axis=1, dtype=F32)
    cp = np.roll(cp, 1, axis=1)         # torch.roll(cp, 1, 0)
    cp[:, -1] = 1                       # cp[-1] = 1
    w = a * cp
    return (w[..., None] * rgb).sum(axis=1, dtype=F32)


def nerfle_render_hierarchical(first, second, rays, light_code, n_coarse, n_fine, t_near, t_far, seed=0,
                               ts_shared=None, view_of_ray=None):
    """coarse pass -> resample -> fine pass -> merged compositing; MLPs through the C oracle."""
    rays = np.ascontiguousarray(rays, F32).reshape(-1, 6)
    R = rays.shape[0]
    if ts_shared is not None and not seed:
        ts_c = np.broadcast_to(np.asarray(ts_shared, F32)[None, :], (R, n_coarse)).copy()
    else:
        ts_c = stratified_ts(R, n_coarse, t_near, t_far, seed)
    _, sig_c, rgb_c = c_oracle.nerfle_render(first, second, rays, ts_per_ray=ts_c, light_code=light_code,
                                             view_of_ray=view_of_ray, store=True)
    if n_fine == 0:
        return composite_ray_major(sig_c, rgb_c, ts_c)
    ts_f = sample_pdf(sig_c, ts_c, n_fine, seed)
    _, sig_f, rgb_f = c_oracle.nerfle_render(first, second, rays, ts_per_ray=ts_f, light_code=light_code,
                                             view_of_ray=view_of_ray, store=True)
    t_all = np.concatenate([ts_c, ts_f], axis=1)
    order = np.argsort(t_all, axis=1, kind="stable")          # coarse first on ties
    rows = np.arange(R)[:, None]
    sig = np.concatenate([sig_c, sig_f], axis=1)[rows, order]
    rgb = np.concatenate([rgb_c, rgb_f], axis=1)[rows, order]
    return composite_ray_major(sig, rgb, t_all[rows, order])


# ---------------------------------------------------------------------------------------------
# torch-CPU restatement of the reference's eager op sequence (what the CPU baseline times).
# The reference IS eager PyTorch (MKL GEMMs + ATen elementwise kernels), so a port that issues the
# same ops on the same shapes reproduces its cost model; the scalar C oracle would under-state it.
# ---------------------------------------------------------------------------------------------
def torch_mlp(w, x, latent=None, act="leaky"):
    """neural_blocks.py:75-86 with synth-format weights (w['W'][i] is [N,K] like nn.Linear)."""
    import torch
    import torch.nn.functional as Fn
    a = Fn.softplus if act == "softplus" else Fn.leaky_relu
    B = torch.as_tensor(w["basis"])
    enc = torch.cat([x, (x @ B).sin(), (x @ B).cos()], dim=-1)       # utils.py:37-40
    if latent is not None:
        enc = torch.cat([enc, latent], dim=-1)
    Ws = [torch.as_tensor(v) for v in w["W"]]
    bs = [torch.as_tensor(v) for v in w["b"]]
    h = Fn.linear(enc, Ws[0], bs[0])
    L = w["num_layers"]
    for i in range(L):
        if i != L - 1 and (i % w["skip"]) == 0:
            h = torch.cat([h, enc], dim=-1)
        h = Fn.linear(a(h), Ws[1 + i], bs[1 + i])
    return Fn.linear(a(h), Ws[L + 1], bs[L + 1])


def torch_nerfle_samples(w1, w2, rays, ts, light_code, chunk=2048):
    """nerf.py:175-203 for per-ray distances ts [R,S]: returns sigma_raw [R,S], sigmoid rgb [R,S,3]."""
    import torch
    R, S = ts.shape
    sig = torch.empty(R, S)
    rgb = torch.empty(R, S, 3)
    lc = torch.as_tensor(light_code).reshape(1, 1, -1)
    with torch.no_grad():
        for r0 in range(0, R, chunk):
            ro, rd, t = rays[r0:r0 + chunk, :3], rays[r0:r0 + chunk, 3:], ts[r0:r0 + chunk]
            n = ro.shape[0]
            pts = ro[:, None, :] + t[..., None] * rd[:, None, :]
            f = torch_mlp(w1, pts.reshape(-1, 3)).reshape(n, S, -1)
            x2 = torch.cat([f[..., 1:], rd[:, None, :].expand(n, S, 3), lc.expand(n, S, lc.shape[-1])], dim=-1)
            c = torch_mlp(w2, x2.reshape(n * S, -1)).sigmoid().reshape(n, S, 3)
            sig[r0:r0 + n] = f[..., 0]
            rgb[r0:r0 + n] = c
    return sig, rgb


def torch_nerfle_render(w1, w2, rays, light_code, n_coarse, n_fine, t_near, t_far, seed=0):
    """Whole config-2 style render on the CPU with torch (MLPs) + the numpy sampling restatement."""
    import torch
    rays = torch.as_tensor(rays).reshape(-1, 6)
    R = rays.shape[0]
    ts_c = stratified_ts(R, n_coarse, t_near, t_far, seed)
    sig_c, rgb_c = torch_nerfle_samples(w1, w2, rays, torch.from_numpy(ts_c), light_code)
    sig_c, rgb_c = sig_c.numpy(), rgb_c.numpy()
    if n_fine == 0:
        return composite_ray_major(sig_c, rgb_c, ts_c)
    ts_f = sample_pdf(sig_c, ts_c, n_fine, seed)
    sig_f, rgb_f = torch_nerfle_samples(w1, w2, rays, torch.from_numpy(ts_f), light_code)
    t_all = np.concatenate([ts_c, ts_f], axis=1)
    order = np.argsort(t_all, axis=1, kind="stable")
    rows = np.arange(R)[:, None]
    sig = np.concatenate([sig_c, sig_f.numpy()], axis=1)[rows, order]
    rgb = np.concatenate([rgb_c, rgb_f.numpy()], axis=1)[rows, order]
    return composite_ray_major(sig, rgb, t_all[rows, order])


def torch_sphere_sdf(w, p):
    """SphereSDF.forward (shapes/sdfs.py:37-46, utils.py:385-387) with synth-format weights: w['centers'] [n,3],
    w['radii'] [n], w['tfs'] [n,3,3], w['shift'] (8x128 softplus SkipConnMLP)."""
    import torch
    c, r, tf = torch.as_tensor(w["centers"]), torch.as_tensor(w["radii"]), torch.as_tensor(w["tfs"])
    tfs = tf + torch.eye(3).unsqueeze(0)
    q = torch.einsum("ijk,ibk->ibj", tfs, p.reshape(1, -1, 3).expand(tfs.shape[0], -1, -1)) - c.unsqueeze(1)
    sd = q.norm(p=2, dim=-1) - r.unsqueeze(-1)
    out = -torch.exp(-32.0 * sd).sum(0).clamp(min=1e-4).log() / 32.0
    return out + torch_mlp(w["shift"], p.reshape(-1, 3), act="softplus").reshape_as(out)


def torch_sdf_march_and_scan(w, rays, eps=1e-3, max_steps=64, max_t=10.0, scan_n=128, scan_dist=2.2):
    """The SDF evaluations of one colocate.py-style frame on the CPU, as the reference issues them: the lock-step
    sphere-trace march (shapes/sdfs.py:111-131: max_steps evaluations of EVERY ray, no early exit) and the
    min-along-ray scan of SDF.throughput (:232-249: scan_n + 1 evaluations of every ray).  Returns (depth [R], hit [R],
    argmin index [R]).  193 of the ~200 network evaluations per ray of the cfg1 pipeline are these."""
    import torch
    rays = torch.as_tensor(rays).reshape(-1, 6)
    r_o, r_d = rays[:, :3], rays[:, 3:]
    with torch.no_grad():
        depths = torch.zeros(rays.shape[0], 1)
        remaining = torch.ones(rays.shape[0], dtype=torch.bool)
        hit = torch.zeros_like(remaining)
        for _ in range(max_steps):
            remaining = remaining & (depths < max_t).squeeze(-1)
            d = torch_sphere_sdf(w, r_o + r_d * depths)
            hits = remaining & (d <= eps)
            hit = hit | hits
            remaining = remaining & ~hits
            depths = torch.where(remaining.unsqueeze(-1), depths + d.unsqueeze(-1), depths)
        step = scan_dist / scan_n
        cur = torch_sphere_sdf(w, r_o)
        idx = torch.zeros_like(cur, dtype=torch.long)
        for i in range(scan_n):
            sd = torch_sphere_sdf(w, r_o + (step * (i + 1)) * r_d)
            idx = torch.where(sd < cur, i + 1, idx)
            cur = torch.minimum(cur, sd)
    return depths.squeeze(-1), hit, idx


class TorchNerfleTrainer:
    """nerfle.py:104-120 on the CPU with torch autograd: NeRFLE.forward (nerf.py:175-214) -> F.mse_loss-style loss ->
    backward -> AdamW(lr 8e-5, wd 0) (nerfle.py:55-57).  Same op sequence as the reference's eager step; used as the
    CPU baseline of the training configs (bench.py) and nowhere in the product."""

    def __init__(self, w1, w2, lr=8e-5):
        import torch
        self.torch = torch
        self.w = []
        for w in (w1, w2):
            self.w.append(dict(w, W=[torch.tensor(v, requires_grad=True) for v in w["W"]],
                               b=[torch.tensor(v, requires_grad=True) for v in w["b"]]))
        self.params = [p for w in self.w for p in w["W"] + w["b"]]
        self.opt = torch.optim.AdamW(self.params, lr=lr, weight_decay=0)

    def forward(self, rays, ts, light_loc):
        """rays [N,R,6], ts [S], light_loc [N,3] -> rgb [N,R,3] (sample-major inside, like the reference)."""
        torch = self.torch
        r_o, r_d = rays[..., :3], rays[..., 3:]
        pts = r_o[None] + ts[:, None, None, None] * r_d[None]                       # [S,N,R,3]
        f = torch_mlp(self.w[0], pts.reshape(-1, 3)).reshape(pts.shape[:-1] + (-1,))
        latent, alpha = f[..., 1:], f[..., 0]
        light = light_loc[None, :, None, :].expand(latent.shape[:-1] + (3,))
        x2 = torch.cat([latent, r_d[None].expand(latent.shape[:-1] + (3,)), light], dim=-1)
        rgb = torch_mlp(self.w[1], x2.reshape(-1, x2.shape[-1])).sigmoid().reshape(latent.shape[:-1] + (3,))
        sigma_a = torch.relu(alpha)
        a = 1 - torch.exp(-sigma_a * ts[:, None, None])
        cp = torch.cumprod((1 - a).clamp(min=1e-10), dim=0)
        cp = torch.roll(cp, 1, 0)
        cp = torch.cat([cp[:-1], torch.ones_like(cp[:1])], dim=0)                   # cp[-1] = 1
        return ((a * cp)[..., None] * rgb).sum(dim=0)

    def step(self, rays, ts, light_loc, target, denom):
        torch = self.torch
        self.opt.zero_grad()
        loss = (self.forward(rays, ts, light_loc) - target).square().sum() / denom
        loss.backward()
        self.opt.step()
        return float(loss.detach())
