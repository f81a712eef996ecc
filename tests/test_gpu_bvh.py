"""GPU parity for a4 (bucket sort), a5 (32-ary BVH) and a6 (light BVH chain) vs the oracle."""
import math

import numpy as np
import pytest

import oracle
from conftest import splitmix64
from test_oracle_primitives import make_leaves

pytestmark = pytest.mark.gpu


def rand_u32(seed, n):
    return (splitmix64(seed, n) & np.uint64(0xFFFFFFFF)).astype(np.uint32)


def rand_f32(seed, n, lo, hi):
    u = (splitmix64(seed, n) >> np.uint64(40)).astype(np.float64) / float(1 << 24)
    return (lo + u * (hi - lo)).astype(np.float32)


def to_dev(a):
    import torch

    a = np.ascontiguousarray(a)
    if a.dtype == np.uint32:
        return torch.from_numpy(a.view(np.int32)).cuda()
    return torch.from_numpy(a).cuda()


# ---- a4 -------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 5, 120, 2400, 8192, 8193, 100003, 1 << 20, 4 * 5 * 6 * 4 * 5 * 6 * 4 * 5])
@pytest.mark.parametrize("full_range_x", [False, True])
def test_bucket_sort_matches_oracle(vren, n, full_range_x):
    keys = rand_u32(23, n) if full_range_x else (rand_u32(23, n) & np.uint32(0xFFFF))   # key = rand()%65536, value = i
    pairs = np.stack([keys, np.arange(n, dtype=np.uint32)], axis=1)
    want, want_counters = oracle.bucket_sort(pairs)
    _, got, counters = vren.bucket_sort(to_dev(pairs))
    got = got.cpu().numpy().view(np.uint32)
    assert np.array_equal(got, want)                          # stable == canonical tie-break
    assert np.array_equal(counters.cpu().numpy().view(np.uint32), want_counters)   # bucket END offsets
    # the reference test's own (weaker) checks, vren_test/.../bucket_sort.cpp:154-179
    masked = got[:, 0] & 0xFFFF
    assert np.all(masked[:-1] <= masked[1:])
    assert np.array_equal(np.sort(got[:, 1]), np.arange(n, dtype=np.uint32))


def test_bucket_sort_skewed_and_morton_like(vren):
    n = 300000
    keys = (rand_u32(29, n) % np.uint32(7)).astype(np.uint32) * np.uint32(4681)     # 7 hot buckets
    pairs = np.stack([keys, rand_u32(30, n)], axis=1)
    want, wc = oracle.bucket_sort(pairs)
    _, got, counters = vren.bucket_sort(to_dev(pairs))
    assert np.array_equal(got.cpu().numpy().view(np.uint32), want)
    assert np.array_equal(counters.cpu().numpy().view(np.uint32), wc)


@pytest.mark.parametrize("by_search", [1, 0])
@pytest.mark.parametrize("n", [1, 3, 777, 70001, (1 << 20) + 5])
def test_bucket_sort_end_offsets_both_paths(vren, n, by_search):
    """END offsets by search in the sorted output (large inputs) == by counting (small inputs), forced either way;
    sparse keys leave most buckets empty (an empty bucket repeats its predecessor's END)"""
    keys = (rand_u32(41, n) % np.uint32(1000)) * np.uint32(61) + np.uint32(0xABCD0000)   # 1000 used buckets, high half noise
    pairs = np.stack([keys, np.arange(n, dtype=np.uint32)], axis=1)
    want, wc = oracle.bucket_sort(pairs)
    _, got, counters = vren.bucket_sort(to_dev(pairs), end_offsets=by_search)
    assert np.array_equal(got.cpu().numpy().view(np.uint32), want)
    assert np.array_equal(counters.cpu().numpy().view(np.uint32), wc)


# ---- a5 -------------------------------------------------------------------------------------------------------------
def nodes_equal(got, want):
    """bit-exact where the reference defines the value; INVALID nodes compare on `next` only"""
    assert np.array_equal(got["next"], want["next"])
    valid = want["next"] != 0xFFFFFFFE
    assert np.array_equal(got["min"][valid].view(np.uint32), want["min"][valid].view(np.uint32))
    assert np.array_equal(got["max"][valid].view(np.uint32), want["max"][valid].view(np.uint32))


@pytest.mark.parametrize("leaf_count", [0, 1, 10, 32, 33, 100, 1000, 1024, 1025, 10000, 40000, 1 << 20])
def test_build_bvh_matches_oracle(vren, leaf_count):
    import torch

    nodes, padded, length = make_leaves(leaf_count, 29)
    want = oracle.build_bvh(nodes, padded)
    dev = torch.from_numpy(nodes.view(np.uint8).copy()).cuda()
    vren.build_bvh(dev, padded)
    got = dev.cpu().numpy().view(oracle.BVH_NODE)
    nodes_equal(got, want)
    lib = vren.load()
    assert lib.vrenb200_calc_bvh_padded_leaf_count(leaf_count) == padded
    assert lib.vrenb200_calc_bvh_buffer_length(leaf_count) == length
    assert lib.vrenb200_calc_bvh_root_index(leaf_count) == length - 1


def test_build_bvh_reference_property_and_preconditions(vren):
    """TEST(build_bvh, main): traversal of the GPU-built tree == linear scan (vren_test/.../build_bvh.cpp:175-221)"""
    import torch

    orc = oracle.load()
    for leaf_count, queries in ((1, 16), (10, 16), (100, 8), (1000, 4), (10000, 2)):
        nodes, padded, length = make_leaves(leaf_count, 31 + leaf_count)
        dev = torch.from_numpy(nodes.view(np.uint8).copy()).cuda()
        vren.build_bvh(dev, padded)
        got = np.ascontiguousarray(dev.cpu().numpy().view(oracle.BVH_NODE))
        pts = rand_f32(37, 3 * queries, 0, 100)
        a = np.zeros(padded, np.uint32)
        b = np.zeros(padded, np.uint32)
        for q in range(queries):
            p = np.ascontiguousarray(pts[3 * q:3 * q + 3])
            na = orc.oracle_bvh_traverse_point(got.ctypes.data, length - 1, p, a, padded)
            nb = orc.oracle_bvh_linear_point(got.ctypes.data, padded, p, b, padded)
            assert na == nb and np.array_equal(a[:na], b[:nb])
    lib = vren.load()
    st = torch.cuda.current_stream().cuda_stream
    buf = torch.zeros(64 * 40, dtype=torch.uint8, device="cuda")
    for bad in (0, 16, 33, 64, 1000):      # assert(leaf_count >= 32 && is_power_of(leaf_count, 32)), build_bvh.cpp:45
        assert lib.vrenb200_build_bvh(st, buf.data_ptr(), bad) == 1


# ---- a6 -------------------------------------------------------------------------------------------------------------
def view_matrix(yaw, pitch, pos):
    """camera.cpp:11-38 (glm::rotate / translate), evaluated in float64 then rounded: it is an INPUT of the path"""
    cy, sy, cp, sp = math.cos(yaw), math.sin(yaw), math.cos(-pitch), math.sin(-pitch)
    ry = np.array([[cy, 0, sy, 0], [0, 1, 0, 0], [-sy, 0, cy, 0], [0, 0, 0, 1]])
    rx = np.array([[1, 0, 0, 0], [0, cp, -sp, 0], [0, sp, cp, 0], [0, 0, 0, 1]])
    orient = ry @ rx
    t = np.eye(4)
    t[:3, 3] = -np.asarray(pos, dtype=np.float64)
    v = np.linalg.inv(orient) @ t
    return v.T.astype(np.float32).reshape(-1)       # column-major


def make_lights(L, seed, box=60.0, intensity=(0.5, 2.0)):
    pos = np.zeros((L, 4), np.float32)
    pos[:, 0] = rand_f32(seed, L, -box, box)
    pos[:, 1] = rand_f32(seed + 1, L, -box / 2, box / 2)
    pos[:, 2] = rand_f32(seed + 2, L, 1.0, 2 * box)
    lights = np.zeros((L, 4), np.float32)
    lights[:, :3] = rand_f32(seed + 3, 3 * L, 0, 1).reshape(L, 3)
    lights[:, 3] = rand_f32(seed + 4, L, *intensity)
    return pos, lights


@pytest.mark.parametrize("L", [1, 2, 31, 32, 33, 1000, 1024, 4097, 65536, 100000])
@pytest.mark.parametrize("external_scratch", [True, False])
def test_construct_point_light_bvh_matches_oracle(vren, L, external_scratch):
    pos, lights = make_lights(L, 41 + L)
    view = view_matrix(0.3, -0.1, (1.0, 2.0, -3.0))
    wvp, wnodes, wpairs = oracle.construct_point_light_bvh(pos, lights, view)
    vp, bvh, idx = vren.construct_point_light_bvh(to_dev(pos), to_dev(lights), view.tolist(), external_scratch)
    assert np.array_equal(vp.cpu().numpy().view(np.uint32), wvp.view(np.uint32))                      # K9 bit-exact
    got_pairs = idx[: L * 8].cpu().numpy().view(np.uint32).reshape(L, 2)
    assert np.array_equal(got_pairs, wpairs)                                                            # Morton + sort
    got_nodes = bvh[: wnodes.size * 32].cpu().numpy().view(oracle.BVH_NODE)
    nodes_equal(got_nodes, wnodes)


@pytest.mark.parametrize("L", [1000000, 1 << 20])
def test_construct_point_light_bvh_c4_sizes(vren, L):
    """BASELINE C4 through the whole a6 chain: 10^6 and 2^20 lights (both pad to 2^20 leaves, 4 levels)"""
    pos, lights = make_lights(L, 4100 + (L & 0xFF))
    view = view_matrix(-0.7, 0.2, (5.0, -1.0, 2.0))
    wvp, wnodes, wpairs = oracle.construct_point_light_bvh(pos, lights, view)
    vp, bvh, idx = vren.construct_point_light_bvh(to_dev(pos), to_dev(lights), view.tolist())
    assert np.array_equal(vp.cpu().numpy().view(np.uint32), wvp.view(np.uint32))
    assert np.array_equal(idx[: L * 8].cpu().numpy().view(np.uint32).reshape(L, 2), wpairs)
    nodes_equal(bvh[: wnodes.size * 32].cpu().numpy().view(oracle.BVH_NODE), wnodes)
    assert wnodes.size == 1082401


def test_light_bvh_degenerate_axis(vren):
    """all lights share y: (p - min)/(max - min) = 0/0 -> canonical bin 0"""
    L = 500
    pos, lights = make_lights(L, 77)
    pos[:, 1] = 3.25
    view = np.eye(4, dtype=np.float32).reshape(-1)
    wvp, wnodes, wpairs = oracle.construct_point_light_bvh(pos, lights, view)
    vp, bvh, idx = vren.construct_point_light_bvh(to_dev(pos), to_dev(lights), view.tolist())
    assert np.array_equal(idx[: L * 8].cpu().numpy().view(np.uint32).reshape(L, 2), wpairs)
    nodes_equal(bvh[: wnodes.size * 32].cpu().numpy().view(oracle.BVH_NODE), wnodes)
