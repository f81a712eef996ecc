"""n4 — vren_demo::visualize_bvh (visualize_bvh.cpp:59-94, show_bvh.comp:62-78): the debug-line consumer of the BVH level
layout.  Oracle known answers on the CPU, bit-exact parity of the CUDA kernel on the GPU."""
import numpy as np
import pytest

import oracle
from test_oracle_primitives import make_leaves

COLORS = [0xff0000, 0xffff00, 0x00ff00, 0x0000ff, 0x00ffff, 0xff00ff, 0xffffff]


def f32bits(*v):
    return np.array(v, np.float32).view(np.uint32)


def test_oracle_single_box_edges():
    nodes = np.zeros(1, dtype=oracle.BVH_NODE)
    nodes["min"][0] = (1, 2, 3)
    nodes["max"][0] = (4, 5, 6)
    nodes["next"][0] = 0xFFFFFFFF
    v = oracle.visualize_bvh(nodes, 0)
    assert v.shape == (24, 4) and (v[:, 3] == COLORS[2]).all()                     # colors[level_count - level + 2]
    pos = v[:, :3].copy().view(np.float32)
    # first macro line: m -> (M.x, m.y, m.z); last: (m.x, m.y, M.z) -> (m.x, M.y, M.z)
    assert pos[0].tolist() == [1, 2, 3] and pos[1].tolist() == [4, 2, 3]
    assert pos[22].tolist() == [1, 2, 6] and pos[23].tolist() == [1, 5, 6]
    # the 12 segments are the 12 edges of the box: each joins two corners that differ in exactly one coordinate
    edges = set()
    for a, b in zip(pos[0::2], pos[1::2]):
        assert int((a != b).sum()) == 1
        edges.add((tuple(a), tuple(b)) if tuple(a) < tuple(b) else (tuple(b), tuple(a)))
    assert len(edges) == 12


def test_oracle_level_layout_and_invalid_nodes():
    # 33 nodes: level 1 = 32 leaves at [0, 32), level 0 = root at 32 (the layout build_bvh produces)
    nodes, padded, length = make_leaves(20, 3)
    assert padded == 32 and length == 33
    built = oracle.build_bvh(nodes, padded)
    v = oracle.visualize_bvh(built, 1)
    assert v.shape == (33 * 24, 4)
    assert (v[: 20 * 24, 3] == COLORS[2]).all() and (v[32 * 24:, 3] == COLORS[3]).all()
    assert not v[20 * 24: 32 * 24].any()                                            # INVALID leaves: degenerate black lines
    root = v[32 * 24:, :3].copy().view(np.float32)
    assert np.allclose(root.min(0), built["min"][32]) and np.allclose(root.max(0), built["max"][32])


@pytest.mark.gpu
@pytest.mark.parametrize("leaf_count", [1, 20, 1000, 40000])
def test_gpu_matches_oracle(leaf_count):
    import torch

    from vren_b200 import lib as vren

    lib = vren.load()
    nodes, padded, length = make_leaves(leaf_count, 11 + leaf_count)
    built = oracle.build_bvh(nodes, padded)
    levels = lib.vrenb200_calc_bvh_level_count(leaf_count)
    want = oracle.visualize_bvh(built, levels)
    assert lib.vrenb200_visualize_bvh_vertex_count(levels) == length * 24
    dev = torch.from_numpy(built.view(np.uint8).copy()).cuda()
    got = vren.visualize_bvh(dev, levels).cpu().numpy().view(np.uint32)
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_gpu_argument_errors():
    import torch

    from vren_b200 import lib as vren

    lib = vren.load()
    buf = torch.zeros(4096, dtype=torch.uint8, device="cuda")
    assert lib.vrenb200_visualize_bvh(None, None, 0, buf.data_ptr()) == vren.EINVAL_ARG
    assert lib.vrenb200_visualize_bvh(None, buf.data_ptr(), 5, buf.data_ptr()) == vren.ELIMIT
    assert lib.vrenb200_visualize_bvh(None, buf.data_ptr() + 4, 0, buf.data_ptr()) == vren.EALIGN
