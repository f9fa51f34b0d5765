"""CPU tests of the dataset layouts the scripts load their targets from (training_utils.test_nerf_resources,
test_colocate_resources; pytorch3d/pathtracer/training_utils.py:538-595) against what the UNMODIFIED reference's loaders
return for the same miniature datasets (tests/golden/dataset_loaders.npz, tiny_datasets.py)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "golden"))

import tiny_datasets  # noqa: E402
from neural_raytracing_b200.pathtracer.training_utils import test_colocate_resources, test_nerf_resources  # noqa: E402

test_colocate_resources.__test__ = False      # library functions with the reference's names, not pytest cases
test_nerf_resources.__test__ = False

G = np.load(os.path.join(HERE, "golden", "dataset_loaders.npz"))


def test_nerf_synthetic_layout(tmp_path):
    d = tiny_datasets.write_nerf_synthetic(str(tmp_path / "lego")) + os.sep
    c2w, focal, imgs, masks = test_nerf_resources(d, size=8, kind="test", device="cpu")
    assert len(c2w) == len(imgs) == len(masks) == 3 and abs(focal - float(G["nerf_focal"])) < 1e-9
    assert np.abs(torch.stack(c2w).numpy() - G["nerf_c2w"]).max() < 1e-6
    assert np.allclose(torch.stack(c2w)[:, :, 3].norm(dim=-1).numpy(), 1.0, atol=1e-6)      # centres on the unit sphere
    assert np.array_equal(torch.stack(imgs).numpy(), G["nerf_imgs"])
    m = torch.stack(masks).numpy()
    assert np.array_equal(m, G["nerf_masks"]) and set(np.unique(m)) == {0.0, 1.0}


def test_colocate_relighting_layout(tmp_path):
    root = tiny_datasets.write_colocate(str(tmp_path / "mitsuba_scenes" / "cbox_relight"), "bunny")
    Rs, Ts, imgs, masks, xyzs = test_colocate_resources("bunny", size=4, dist=1.3, device="cpu", root=root)
    assert len(Rs) == len(Ts) == len(imgs) == len(masks) == len(xyzs) == 144
    assert np.abs(torch.cat(Rs).numpy() - G["col_R"]).max() < 1e-6 and np.abs(torch.cat(Ts).numpy() - G["col_T"]).max() < 1e-6
    assert np.array_equal(torch.stack(imgs).numpy(), G["col_imgs"]) and np.array_equal(torch.stack(masks).numpy(), G["col_masks"])
    xyz = torch.stack(xyzs).numpy()
    assert np.abs(xyz - G["col_xyz"]).max() < 1e-6
    assert np.allclose(np.linalg.norm(xyz, axis=-1), 1.3 * 1.05, atol=1e-5)
