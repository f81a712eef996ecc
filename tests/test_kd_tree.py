"""n4 — vren::kd_tree (base/kd_tree.cpp:5-129).  TEST(KDTree, NearestNeigborSearch) (vren_test/vren_test/kd_tree.cpp:29):
the tree's nearest neighbour equals a linear scan's.  The host build/search run everywhere; the batched device search
needs a GPU."""
import ctypes as C

import numpy as np
import pytest

from vren_b200 import lib as vren

NODE = np.dtype([("value", np.uint32), ("axis_and_link", np.uint32)])


def build(points, max_leaf=16, stride=None):
    lib = vren.load()
    pts = np.ascontiguousarray(points, dtype=np.float32)
    stride = stride or pts.shape[1]
    idx = np.arange(pts.shape[0], dtype=np.uint32)
    nodes = np.zeros(2 * pts.shape[0] + 1, dtype=NODE)
    count = lib.vrenb200_kd_tree_build(pts.ctypes.data, stride, idx.ctypes.data, pts.shape[0], nodes.ctypes.data, max_leaf)
    assert 0 < count <= 2 * pts.shape[0]
    return pts, idx, nodes[:count].copy()


def search(pts, stride, nodes, sample, filt=None):
    lib = vren.load()
    bp, bd = C.c_uint32(0xFFFFFFFF), C.c_float(np.inf)
    s = np.ascontiguousarray(sample, dtype=np.float32)
    cb = C.CFUNCTYPE(C.c_int, C.c_uint32, C.c_void_p)(filt) if filt else None
    lib.vrenb200_kd_tree_search(pts.ctypes.data, stride, nodes.ctypes.data, s.ctypes.data, C.cast(cb, C.c_void_p) if cb else None, None,
                                C.byref(bp), C.byref(bd))
    return bp.value, bd.value


def linear(pts, sample, mask=None):
    d = ((pts[:, :3].astype(np.float32) - np.asarray(sample, np.float32)) ** 2)
    d2 = (d[:, 0] + d[:, 1]) + d[:, 2]
    if mask is not None:
        d2 = np.where(mask, d2, np.inf)
    i = int(np.argmin(d2))
    return i, float(d2[i])


@pytest.mark.parametrize("n,max_leaf", [(1, 16), (17, 16), (1000, 1), (200000, 32)])
def test_nearest_neighbour_equals_linear_scan(n, max_leaf):
    rng = np.random.default_rng(n)
    pts, idx, nodes = build(rng.uniform(-100, 100, (n, 3)), max_leaf)
    assert sorted(idx.tolist()) == list(range(n))                       # the indices are permuted, none lost
    leaves = nodes[(nodes["axis_and_link"] & 3) == 3]
    assert sorted(leaves["value"].tolist()) == list(range(n))           # every point is in exactly one leaf slot
    for s in rng.uniform(-120, 120, (100, 3)).astype(np.float32):
        got, gd = search(pts, 3, nodes, s)
        want, wd = linear(pts, s)
        assert got == want and np.isclose(gd, wd, rtol=1e-6)


def test_stride_filter_and_degenerate_clouds():
    rng = np.random.default_rng(5)
    raw = rng.uniform(0, 10, (5000, 5)).astype(np.float32)              # xyz + two payload floats: stride 5
    pts, idx, nodes = build(raw, 8, stride=5)
    s = np.array([5, 5, 5], np.float32)
    got, _ = search(pts, 5, nodes, s)
    assert got == linear(raw, s)[0]
    even = lambda p, _user: int(p % 2 == 0)                             # kd_tree_search_filter_t
    got, _ = search(pts, 5, nodes, s, even)
    assert got == linear(raw, s, mask=(np.arange(5000) % 2 == 0))[0]
    # identical points and points on a line: the mean split separates nothing, the build must still terminate
    same = np.ones((3000, 3), np.float32)
    pts, idx, nodes = build(same, 4)
    assert search(pts, 3, nodes, [0, 0, 0])[1] == 3.0
    line = np.zeros((4097, 3), np.float32)
    line[:, 1] = np.arange(4097)
    pts, idx, nodes = build(line, 2)
    assert search(pts, 3, nodes, [0.2, 1000.4, 0])[0] == 1000


@pytest.mark.gpu
def test_batched_device_search_equals_host_search():
    import torch

    lib = vren.load()
    rng = np.random.default_rng(77)
    pts, idx, nodes = build(rng.normal(0, 30, (300000, 3)), 16)
    samples = rng.normal(0, 40, (20000, 3)).astype(np.float32)
    d_pts = torch.from_numpy(pts).cuda()
    d_nodes = torch.from_numpy(nodes.view(np.uint32).reshape(-1, 2).view(np.int32).copy()).cuda()
    d_s = torch.from_numpy(samples).cuda()
    bp = torch.empty(len(samples), dtype=torch.int32, device="cuda")
    bd = torch.empty(len(samples), dtype=torch.float32, device="cuda")
    assert lib.vrenb200_kd_tree_search_batch(None, d_pts.data_ptr(), 3, d_nodes.data_ptr(), d_s.data_ptr(), len(samples), bp.data_ptr(), bd.data_ptr()) == 0
    got_p, got_d = bp.cpu().numpy().view(np.uint32), bd.cpu().numpy()
    for i in range(0, len(samples), 97):
        hp, hd = search(pts, 3, nodes, samples[i])
        assert got_p[i] == hp and got_d[i] == np.float32(hd)
    # and against the brute force on the device for all of them
    d2 = torch.cdist(d_s[:2000].double(), d_pts.double()).min(dim=1)
    assert torch.equal(d2.indices.cpu(), torch.from_numpy(got_p[:2000].astype(np.int64)))
