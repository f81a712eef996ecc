"""n2: depth-buffer pyramid — oracle self-checks (CPU) and CUDA parity (GPU)."""
import numpy as np
import pytest

import oracle
from vren_b200 import synthetic


def level_sizes(W, H):
    out, l = [], 0
    while True:
        w, h = max(W >> l, 1), max(H >> l, 1)
        out.append((w, h))
        l += 1
        if w == 1 and h == 1:
            return out


@pytest.mark.parametrize("size", [(1, 1), (2, 2), (3, 5), (64, 64), (100, 37), (640, 360)])
def test_oracle_pyramid_properties(built, size):
    W, H = size
    depth = synthetic.depth_buffer(max(W, 8), max(H, 8), seed=W * 31 + H)[:H, :W].copy()
    pyr, levels = oracle.depth_pyramid(depth)
    sizes = level_sizes(W, H)
    assert levels == len(sizes) == int(np.floor(np.log2(max(W, H)))) + 1      # depth_buffer_pyramid.cpp:18
    off = 0
    prev = None
    for (w, h) in sizes:
        lvl = pyr[off:off + w * h].reshape(h, w)
        if prev is None:
            assert np.array_equal(lvl, depth)                                  # level 0 = copy
        else:
            ph, pw = prev.shape
            for y in range(h):
                for x in range(w):
                    block = prev[2 * y:min(2 * y + 2, ph), 2 * x:min(2 * x + 2, pw)]
                    assert lvl[y, x] == max(0.0, block.max())
        prev = lvl
        off += w * h
    assert off == pyr.size


@pytest.mark.gpu
@pytest.mark.parametrize("size", [(1, 1), (2, 3), (63, 65), (64, 64), (100, 37), (130, 70), (640, 360), (1000, 700), (1920, 1080), (3840, 2160), (4099, 515)])
def test_depth_pyramid_matches_oracle(vren, size):
    import torch

    W, H = size
    depth = synthetic.depth_buffer(max(W, 8), max(H, 8), seed=W + H)[:H, :W].copy()
    want, levels = oracle.depth_pyramid(depth)
    lib = vren.load()
    assert lib.vrenb200_depth_pyramid_level_count(W, H) == levels
    assert lib.vrenb200_depth_pyramid_bytes(W, H) == want.size * 4
    got = vren.depth_pyramid(torch.from_numpy(depth).cuda()).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    off = 0
    for l, (w, h) in enumerate(level_sizes(W, H)):
        assert lib.vrenb200_depth_pyramid_level_offset(W, H, l) == off
        assert lib.vrenb200_depth_pyramid_level_width(W, l) == w and lib.vrenb200_depth_pyramid_level_height(H, l) == h
        off += w * h
