"""Deterministic miniature datasets in the two on-disk layouts the scripts read (NeRF-synthetic transforms json + RGBA
PNGs; the cbox relighting set of colocate.py / nerfle.py), shared by tests/golden/make_golden.py (which runs the
reference's loaders on them) and tests/test_dataset_loaders_cpu.py (which runs this repo's)."""
import json
import os

import numpy as np


def _rgba(seed, n=12):
    rs = np.random.RandomState(seed)
    img = rs.randint(0, 256, size=(n, n, 4)).astype(np.uint8)
    img[: n // 3, :, 3] = 0            # a fully transparent band and a fully opaque one: the mask thresholds matter
    img[-(n // 3):, :, 3] = 255
    return img


def write_nerf_synthetic(directory, n_frames=3):
    from PIL import Image
    os.makedirs(os.path.join(directory, "test"), exist_ok=True)
    frames = []
    for i in range(n_frames):
        Image.fromarray(_rgba(100 + i), "RGBA").save(os.path.join(directory, "test", "r_%d.png" % i))
        # NeRF-convention camera-to-world (-z forward) on a sphere of radius 4 looking at the origin, like the real dataset
        az, el = 0.5 + 0.9 * i, 0.35 + 0.1 * i
        c = np.array([np.cos(el) * np.sin(az), np.sin(el), np.cos(el) * np.cos(az)])
        fwd = -c
        right = np.cross(fwd, [0, 1, 0]); right /= np.linalg.norm(right)
        up = np.cross(right, fwd)
        m = np.eye(4)
        m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = right, up, -fwd, 4.0 * c
        frames.append({"file_path": "./test/r_%d" % i, "transform_matrix": m.tolist()})
    with open(os.path.join(directory, "transforms_test.json"), "w") as f:
        json.dump({"camera_angle_x": 0.6911112070083618, "frames": frames}, f)
    return directory


def write_colocate(root, kind="bunny"):
    from PIL import Image
    os.makedirs(root, exist_ok=True)
    n = 0
    for i in range(4):
        for j in range(4):
            for k in range(3):
                for l in range(3):
                    Image.fromarray(_rgba(1000 + n, 6), "RGBA").save(
                        os.path.join(root, "gt_%s_%03d_%03d_%03d_%03d.png" % (kind, i, j, k, l)))
                    n += 1
    return root
