"""Training and evaluation loops with the call surface of pytorch3d/pathtracer/training_utils.py (SURVEY.md section 8f,
rank 1): `train_sample` :123, `train_nerf` :211, `train_dtu` :347, `test` :487, `test_nerf` :302, `test_dtu` :436.
These are the callers of the per-ray path in scripts/colocate.py, nerf_synthetic.py and dtu.py.

One engine (`_fit`) serves the three training entry points, which differ only in how a batch of view indices becomes
a camera, in the mask weight of `masked_loss`, in the NaN policy and in a few render arguments; one engine
(`_evaluate`) serves the three test entry points.  What every iteration does is the reference's:

    views  <- LossSampler.sample(N)               (numpy RNG; views with a large last loss are preferred)
    (u, v) <- uv_select(mask[0], crop_size)       (python RNG)
    got, mi <- pathtrace_sample(... crop ...)     -> the fused CUDA kernels
    loss   <- masked_loss(got[..., :3], exp crop, mi.throughput, mask crop, mask_weight) + extra_loss(mi, got, exp, mask)
    loss.backward(); opt.step(); LossSampler.update_idxs(views, loss)   (`train_sample` never updates the sampler: kept)

Differences from the reference, all outside the arithmetic: images are written with PIL instead of matplotlib and only
if the target directory exists; progress goes through tqdm when it is importable; `train_sample` takes the camera
factory that the reference forgot to define (`mk_camera`, `focal` are undefined names at training_utils.py:161; the
default builds `OpenGLPerspectiveCameras(device, R, T)` like its validation branch :198 and nerfle.py:96); a
`step_hook(i, loss)` lets a caller collect statistics.  SSIM is the in-repo restatement (pathtracer/ssim.py).
"""
import os

import numpy as np
import torch
import torch.nn.functional as F

from . import main as _main
from .cameras import DTUCamera, NeRFCamera
from .integrators import NeRFIntegrator
from .ssim import ssim
from .utils import LossSampler, masked_loss, mse2psnr, rand_uv_mask
from ..renderer.cameras import OpenGLPerspectiveCameras


# ---- small host-side helpers -------------------------------------------------------------------------------
def save_image(name, img):
    """training_utils.py:21 (plt.imsave there): clamp to [0,1], 8 bit, skip silently when the directory is absent."""
    d = os.path.dirname(name)
    if d and not os.path.isdir(d):
        return False
    from PIL import Image
    a = (img.detach().float().cpu().clamp(0, 1).numpy() * 255.0 + 0.5).astype(np.uint8)
    if a.ndim == 3 and a.shape[-1] == 1:
        a = a[..., 0]
    if a.ndim == 3 and a.shape[-1] > 4:
        a = a[..., :3]
    Image.fromarray(a).save(name)
    return True


def save_plot(expected, got, name):
    """training_utils.py:22-33: expected and rendered image side by side."""
    e, g = expected.detach().float().cpu(), got.detach().float().cpu()
    c = min(e.shape[-1], g.shape[-1], 3)
    return save_image(name, torch.cat([e[..., :c], g[..., :c]], dim=1))


def no_update(cameras, lights):
    return


def _progress(iters, silent, really_silent=False):
    """(iterator, report(loss, i)): a tqdm bar, or one line per iteration (silent), or one per 1000 (really_silent)."""
    bar = None
    if not silent and not really_silent:
        try:
            from tqdm import trange
            bar = trange(iters)
        except ImportError:
            bar = None
    if bar is not None:
        return bar, lambda loss, i: bar.set_postfix(refresh=False, loss=f"{loss:.05}")
    if really_silent:
        return range(iters), lambda loss, i: print(f"{i:06}: {loss:.05}") if i % 1000 == 0 else None
    return range(iters), lambda loss, i: print(f"{i:06}: {loss:.05}")


def _default_uv_select(mask, crop_size):
    return rand_uv_mask(mask, crop_size)


def _zero_extra(mi, got, exp, mask):
    return 0


# ---- the training engine -------------------------------------------------------------------------------------
def _fit(shape, bsdf, integrator, lights, exp_imgs, exp_masks, opt, size, crop_size, make_cameras, valid_cameras, *,
         N, iters, num_ckpts, save_freq, valid_freq, max_valid_size, extra_loss, save_fn, name_fn, valid_name_fn,
         uv_select, silent, really_silent=False, mask_weight, nan_raises, update_sampler, sample_kwargs, valid_kwargs,
         valid_integrator, before_render=None, step_hook=None):
    device = exp_imgs[0].device
    ckpt_freq = (iters // num_ckpts) - 1
    losses = []
    selector = LossSampler(len(exp_imgs))
    iterator, report = _progress(iters, silent, really_silent)
    for i in iterator:
        idxs = selector.sample(n=N)
        exp = torch.stack([exp_imgs[j] for j in idxs])
        mask = torch.stack([exp_masks[j] for j in idxs])
        cameras = make_cameras(idxs)
        if before_render is not None:
            before_render(cameras, lights)
        opt.zero_grad()
        u, v = uv_select(mask[0], crop_size)
        got, mi = _main.pathtrace_sample(shape, size=size, chunk_size=size, bundle_size=1, crop_size=crop_size, bsdf=bsdf,
                                         integrator=integrator, cameras=cameras, lights=lights, device=device, uv=(u, v),
                                         addition=lambda it: it, squeeze_first=False, silent=True, **sample_kwargs)
        if (i % save_freq) == 0:
            save_image(name_fn(i), got[0])
        u, v = int(u), int(v)
        exp = exp[:, u:u + crop_size, v:v + crop_size]
        mask = mask[:, u:u + crop_size, v:v + crop_size]
        loss = masked_loss(got[..., :3], exp, mi.throughput.squeeze(-1), mask, mask_weight=mask_weight,
                           with_logits=mi.with_logits) + extra_loss(mi, got, exp, mask)
        if loss.isnan():
            if not nan_raises:
                continue                      # train_sample skips the step (training_utils.py:185)
            loss.backward()                   # train_nerf / train_dtu propagate, step and raise (:265-269, :400-404)
            opt.step()
            raise Exception("Unexpected NaN")
        loss.backward()
        opt.step()
        loss = loss.detach().item()
        losses.append(loss)
        if update_sampler:
            selector.update_idxs(idxs, loss)
        report(loss, i)
        if step_hook is not None:
            step_hook(i, loss)
        if ((i % ckpt_freq) == 0) and (i != 0):
            save_fn(i)
        if (i % valid_freq) == 0:
            with torch.no_grad():
                cams = valid_cameras(idxs)
                if before_render is not None:
                    before_render(cams, lights)
                validate, _ = _main.pathtrace(shape, size=size, chunk_size=min(size, max_valid_size), bundle_size=1,
                                              bsdf=bsdf, integrator=valid_integrator, cameras=cams, lights=lights,
                                              device=device, silent=True, **valid_kwargs)
                save_image(valid_name_fn(i), validate)
    return losses


def train_sample(shape, bsdf, integrator, lights, Rs, Ts, exp_imgs, exp_masks, opt, size, crop_size, N=3, iters=50_000,
                 num_ckpts=5, save_freq=50, valid_freq=250, max_valid_size=128, extra_loss=_zero_extra,
                 save_fn=lambda i: None, name_fn=lambda i: f"outputs/train_{i:05}.png",
                 valid_name_fn=lambda i: f"outputs/valid_{i:05}.png", uv_select=_default_uv_select,
                 light_update=no_update, silent=False, really_silent=False, w_isect=False, mk_camera=None,
                 step_hook=None):
    """colocate.py-style training (training_utils.py:123-209): per-view look-at (R, T), the light follows the camera
    through `light_update`, mask weight 15, NaN steps are skipped, the loss sampler is never updated."""
    device = exp_imgs[0].device
    if mk_camera is None:
        def mk_camera(R, T, device):
            return OpenGLPerspectiveCameras(device=device, R=R, T=T)

    def cams(idxs):
        return mk_camera(torch.cat([Rs[j] for j in idxs], dim=0), torch.cat([Ts[j] for j in idxs], dim=0), device)

    def valid_cams(idxs):
        return OpenGLPerspectiveCameras(device=device, R=Rs[idxs[0]][:1], T=Ts[idxs[0]][:1])

    return _fit(shape, bsdf, integrator, lights, exp_imgs, exp_masks, opt, size, crop_size, cams, valid_cams, N=N,
                iters=iters, num_ckpts=num_ckpts, save_freq=save_freq, valid_freq=valid_freq,
                max_valid_size=max_valid_size, extra_loss=extra_loss, save_fn=save_fn, name_fn=name_fn,
                valid_name_fn=valid_name_fn, uv_select=uv_select, silent=silent, really_silent=really_silent,
                mask_weight=15, nan_raises=False, update_sampler=False, sample_kwargs=dict(w_isect=w_isect),
                valid_kwargs=dict(w_isect=w_isect), valid_integrator=NeRFIntegrator(integrator),
                before_render=light_update, step_hook=step_hook)


def train_nerf(shape, bsdf, integrator, lights, cam_to_worlds, focal, exp_imgs, exp_masks, opt, size, crop_size, N=3,
               iters=50_000, num_ckpts=5, save_freq=50, valid_freq=250, max_valid_size=128, extra_loss=_zero_extra,
               save_fn=lambda i: None, name_fn=lambda i: f"outputs/train_{i:05}.png",
               valid_name_fn=lambda i: f"outputs/valid_{i:05}.png", uv_select=_default_uv_select, silent=False,
               step_hook=None):
    """nerf_synthetic.py-style training (training_utils.py:211-300): NeRF cameras, the integrator wrapped in
    NeRFIntegrator (alpha channel from the silhouette logit), background 0, mask weight 15, NaN raises."""
    device = exp_imgs[0].device
    wrapped = NeRFIntegrator(integrator)

    def cams(idxs):
        return NeRFCamera(cam_to_world=torch.stack([cam_to_worlds[j] for j in idxs], dim=0), focal=focal, device=device)

    def valid_cams(idxs):
        return NeRFCamera(cam_to_world=cam_to_worlds[idxs[0]].unsqueeze(0), focal=focal, device=device)

    return _fit(shape, bsdf, wrapped, lights, exp_imgs, exp_masks, opt, size, crop_size, cams, valid_cams, N=N,
                iters=iters, num_ckpts=num_ckpts, save_freq=save_freq, valid_freq=valid_freq,
                max_valid_size=max_valid_size, extra_loss=extra_loss, save_fn=save_fn, name_fn=name_fn,
                valid_name_fn=valid_name_fn, uv_select=uv_select, silent=silent, mask_weight=15, nan_raises=True,
                update_sampler=True, sample_kwargs=dict(background=0), valid_kwargs={}, valid_integrator=wrapped,
                step_hook=step_hook)


def train_dtu(shape, bsdf, integrator, lights, poses, intrinsics, exp_imgs, exp_masks, opt, size, crop_size, N=3,
              iters=50_000, num_ckpts=5, save_freq=50, valid_freq=250, max_valid_size=128, extra_loss=_zero_extra,
              save_fn=lambda i: None, name_fn=lambda i: f"outputs/train_{i:05}.png",
              valid_name_fn=lambda i: f"outputs/valid_{i:05}.png", uv_select=_default_uv_select, silent=False,
              step_hook=None):
    """dtu.py-style training (training_utils.py:347-434): DTU / IDR cameras, mask weight 10, NaN raises.  Kept: the
    validation render pairs the first sampled pose with the FIRST intrinsic matrix of the data set (:425)."""
    device = exp_imgs[0].device
    wrapped = NeRFIntegrator(integrator)

    def cams(idxs):
        return DTUCamera(pose=torch.stack([poses[j] for j in idxs], dim=0),
                         intrinsic=torch.stack([intrinsics[j] for j in idxs], dim=0), device=device)

    def valid_cams(idxs):
        return DTUCamera(pose=poses[idxs[0]][None], intrinsic=intrinsics[0][None], device=device)

    return _fit(shape, bsdf, wrapped, lights, exp_imgs, exp_masks, opt, size, crop_size, cams, valid_cams, N=N,
                iters=iters, num_ckpts=num_ckpts, save_freq=save_freq, valid_freq=valid_freq,
                max_valid_size=max_valid_size, extra_loss=extra_loss, save_fn=save_fn, name_fn=name_fn,
                valid_name_fn=valid_name_fn, uv_select=uv_select, silent=silent, mask_weight=10, nan_raises=True,
                update_sampler=True, sample_kwargs=dict(background=0), valid_kwargs={}, valid_integrator=wrapped,
                step_hook=step_hook)


# ---- the evaluation engine -----------------------------------------------------------------------------------
def _evaluate(density_field, integrator, bsdf, lights, cameras_of, n_views, exp_imgs, size, chunk, name_fn, *,
              masks=None, ssim_stride=1, before_render=None, render_kwargs=None, verbose=True):
    """Renders every view, clamps to [0,1], reports mean l1 / l2 / PSNR and the SSIM of the stacked images; returns
    the numbers as a dict (the reference only prints them)."""
    device = exp_imgs[0].device
    l1s, l2s, psnrs, gots, exps = [], [], [], [], []
    with torch.no_grad():
        for i in range(n_views):
            exp = exp_imgs[i]
            cams = cameras_of(i)
            if before_render is not None:
                before_render(cams, lights)
            got = _main.pathtrace(density_field, size=size, chunk_size=min(size, chunk), bundle_size=1, bsdf=bsdf,
                                  integrator=integrator, cameras=cams, lights=lights, device=device, silent=True,
                                  background=0, **(render_kwargs or {}))[0].clamp(min=0, max=1)
            save_plot(exp, got, name_fn(i))
            if masks is not None:
                m = (masks[i] == 1)[..., None]
                exp, got = exp * m, got * m
            mse = F.mse_loss(exp, got)
            l1s.append(F.l1_loss(exp, got).item())
            l2s.append(mse.item())
            psnrs.append(mse2psnr(mse).item())
            gots.append(got)
            exps.append(exp)
        g = torch.stack(gots[::ssim_stride], dim=0).permute(0, 3, 1, 2)
        e = torch.stack(exps[::ssim_stride], dim=0).permute(0, 3, 1, 2)
        ssim_value = ssim(g, e, data_range=1, size_average=True).item()
    stats = {"l1": float(np.mean(l1s)), "l2": float(np.mean(l2s)), "psnr": float(np.mean(psnrs)), "ssim": ssim_value}
    if verbose:
        print("Avg l1 loss", stats["l1"])
        print("Avg l2 loss", stats["l2"])
        print("Avg PSNR loss", stats["psnr"])
        print("SSIM loss", stats["ssim"])
    return stats


def test(density_field, integrator, bsdf, lights, Rs, Ts, exp_imgs, size, max_chunk_size=128, light_update=no_update,
         name_fn=lambda i: f"outputs/test_{i:03}.png", w_isect=False):
    """training_utils.py:487-536 (colocate-style views; SSIM over every third image, like the reference)."""
    device = exp_imgs[0].device
    return _evaluate(density_field, integrator, bsdf, lights,
                     lambda i: OpenGLPerspectiveCameras(device=device, R=Rs[i], T=Ts[i]), len(Rs), exp_imgs, size,
                     max_chunk_size, name_fn, ssim_stride=3, before_render=light_update,
                     render_kwargs=dict(w_isect=w_isect))


def test_nerf(density_field, integrator, bsdf, lights, cam_to_worlds, focal, exp_imgs, size,
              name_fn=lambda i: f"outputs/test_{i:03}.png"):
    """training_utils.py:302-344."""
    device = exp_imgs[0].device
    return _evaluate(density_field, integrator, bsdf, lights,
                     lambda i: NeRFCamera(cam_to_world=cam_to_worlds[i].unsqueeze(0), focal=focal, device=device),
                     len(cam_to_worlds), exp_imgs, size, 256, name_fn)


def test_dtu(density_field, integrator, bsdf, lights, poses, intrinsics, exp_imgs, exp_masks, size,
             name_fn=lambda i: f"outputs/test_{i:03}.png"):
    """training_utils.py:436-485 (metrics on the masked images)."""
    device = exp_imgs[0].device
    return _evaluate(density_field, integrator, bsdf, lights,
                     lambda i: DTUCamera(pose=poses[i][None, ...], intrinsic=intrinsics[i][None, ...], device=device),
                     len(poses), exp_imgs, size, 128, name_fn, masks=exp_masks)


# keep pytest from collecting the reference-named entry points when this module is imported into a test file
test.__test__ = False
test_nerf.__test__ = False
test_dtu.__test__ = False


# ---- dataset layouts the scripts load their targets from (training_utils.py:538-595) ------------------------------
def test_nerf_resources(directory, size=128, kind="test", device="cuda"):
    """NeRF-synthetic layout (nerf_synthetic.py:125): `directory + transforms_<kind>.json` with `camera_angle_x` and
    per-frame `file_path` / `transform_matrix`.  -> (cam_to_worlds [3,4] each, camera centre pulled onto the unit
    sphere; focal = 0.5 size / tan(fov_x / 2); RGB targets; masks = ceil(alpha - 1e-5))."""
    import json
    from .utils import load_image
    assert kind in ("train", "test")
    with open(directory + "transforms_%s.json" % kind) as f:
        meta = json.load(f)
    focal = 0.5 * size / np.tan(0.5 * float(meta["camera_angle_x"]))
    poses, images, masks = [], [], []
    for frame in meta["frames"]:
        rgba = load_image(os.path.join(directory, frame["file_path"] + ".png"), resize=(size, size)).to(device)
        images.append(rgba[..., :3])
        masks.append((rgba[..., 3] - 1e-5).ceil())
        pose = torch.tensor(frame["transform_matrix"], dtype=torch.float, device=device)[:3, :4]
        pose[:3, 3] = F.normalize(pose[:3, 3], dim=-1)
        poses.append(pose)
    return poses, focal, images, masks


def test_colocate_resources(kind, size=128, dist=1, device="cuda", root="mitsuba_scenes/cbox_relight"):
    """The relighting set of colocate.py:162 / nerfle.py:177: 4 x 4 camera poses (elevation 0..45, azimuth -90..90), each
    with 3 x 3 light positions on the sphere of radius 1.05 dist; images `gt_<kind>_<i>_<j>_<k>_<l>.png` under `root`.
    -> (Rs, Ts, RGB targets, alpha masks, light positions), 144 entries each."""
    from .utils import load_image
    from ..renderer.cameras import look_at_view_transform

    def on_sphere(elev, azim, rad):
        elev, azim = torch.deg2rad(elev), torch.deg2rad(azim)
        return torch.stack([rad * elev.cos() * azim.sin(), rad * elev.cos() * azim.cos(), rad * elev.sin()], dim=0)

    Rs, Ts, images, masks, light_xyz = [], [], [], [], []
    for i, elev in enumerate(torch.linspace(0, 45, 4, device=device)):
        for j, azim in enumerate(torch.linspace(-90, 90, 4, device=device)):
            R, T = look_at_view_transform(dist=dist, elev=elev, azim=azim, device=device)
            for k, light_elev in enumerate(torch.linspace(0, 45, 3, device=device)):
                for l, light_azim in enumerate(torch.linspace(-90, 90, 3, device=device)):
                    rgba = load_image(os.path.join(root, "gt_%s_%03d_%03d_%03d_%03d.png" % (kind, i, j, k, l)),
                                      (size, size)).to(device)
                    Rs.append(R)
                    Ts.append(T)
                    images.append(rgba[..., :3])
                    masks.append(rgba[..., 3])
                    light_xyz.append(on_sphere(light_elev, light_azim, dist * 1.05))
    return Rs, Ts, images, masks, light_xyz
