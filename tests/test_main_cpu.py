"""CPU tests of the host logic of pathtrace's tiling (pathtracer/main.py): when a gradient-free GPU frame is rendered in row
blocks and how many rows a block gets; the tile loop itself is exercised on the GPU (tests/test_gpu_pipeline.py)."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from neural_raytracing_b200 import config  # noqa: E402
from neural_raytracing_b200.pathtracer import main as M  # noqa: E402


def test_row_block_policy():
    prev = config.max_tile_rays
    try:
        config.set_max_tile_rays(524288)
        with torch.no_grad():
            assert M._row_block(800, 800, 1, 32, M.nothing, "cuda") == 655          # 524,288 // 800 rows per call
            assert M._row_block(256, 256, 1, 32, M.nothing, "cuda:1") == 256        # the whole frame fits
            assert M._row_block(256, 256, 4, 32, M.nothing, "cuda") == 256          # bundle of 4: 262,144 rays, still one block
            assert M._row_block(3840, 2160, 1, 128, M.nothing, "cuda") == 242
            assert M._row_block(256, 256, 1, 256, M.nothing, "cuda") == 0           # the caller's tile is already the frame
            assert M._row_block(256, 256, 1, 32, M.nothing, "cpu") == 0             # CPU: the reference's tiles
            assert M._row_block(256, 256, 1, 32, lambda it: it, "cuda") == 0        # the caller reads the last tile's interaction
            assert M._row_block(64, 1 << 20, 1, 32, M.nothing, "cuda") == 0         # a single row would exceed the budget: keep tiles
            config.set_max_tile_rays(0)
            assert M._row_block(800, 800, 1, 32, M.nothing, "cuda") == 0
            config.set_max_tile_rays(100)
            assert M._row_block(16, 16, 1, 4, M.nothing, "cuda") == 6
        config.set_max_tile_rays(524288)
        assert M._row_block(800, 800, 1, 32, M.nothing, "cuda") == 0                # under autograd: the caller's tiles
    finally:
        config.set_max_tile_rays(prev)
