"""GPU parity for a7 (unique cluster keys) and a8 (light assignment) vs the oracle, and the a6->a7->a8 chain (a9).

The reference has no test for this pass (parity unpinned upstream); the oracle restates the shaders with the
canonical choices of SURVEY 8c.  The CUDA path emits the canonical tile-major order itself, so keys, per-pixel
references, counts, offsets and light index lists are all compared bit-exactly, plus the order-free properties the
reference guarantees (keys[ref(x,y)] == key(x,y); per-key light sets).
"""
import math

import numpy as np
import pytest

import oracle
from vren_b200 import synthetic

pytestmark = pytest.mark.gpu


def dev(a):
    import torch

    a = np.ascontiguousarray(a)
    if a.dtype == np.uint32:
        return torch.from_numpy(a.view(np.int32)).cuda()
    return torch.from_numpy(a).cuda()


def host_u32(t):
    return t.cpu().numpy().view(np.uint32)


def both_cameras(vren, w, h, **kw):
    oc = oracle.default_camera(w, h, **kw)
    vc = vren.Camera(oc.fov_y, oc.aspect_ratio, oc.near_plane, oc.far_plane)
    return oc, vc


@pytest.mark.parametrize("size", [(32, 32), (64, 48), (256, 144), (640, 360), (1000, 700), (1920, 1080)])
@pytest.mark.parametrize("with_normals", [False, True])
def test_find_unique_clusters_matches_oracle(vren, size, with_normals):
    w, h = size
    depth = synthetic.depth_buffer(w, h, seed=w + h)
    normals = synthetic.normal_buffer(w, h, seed=w) if with_normals else None
    oc, vc = both_cameras(vren, w, h)
    want_keys, want_ref = oracle.find_unique_clusters(depth, normals, oc)
    keys, disp, ref = vren.find_unique_clusters(dev(depth), None if normals is None else dev(normals), vc)
    disp = host_u32(disp)
    assert disp[0] == want_keys.size and disp[1] == 1 and disp[2] == 1 and disp[3] == 0
    got_keys = host_u32(keys)[: disp[0]]
    assert np.array_equal(got_keys, want_keys)                       # canonical tile-major order, ascending per tile
    got_ref = host_u32(ref).reshape(h, w)
    assert np.array_equal(got_ref, want_ref)
    # order-free contract of the reference: the key a pixel references is the pixel's own key
    px_keys = got_keys[got_ref]
    ys, xs = np.mgrid[0:h, 0:w]
    assert np.array_equal(px_keys & 0xFF, (xs >> 5) & 0xFF) and np.array_equal((px_keys >> 8) & 0xFF, (ys >> 5) & 0xFF)
    if not with_normals:
        assert np.all(px_keys >> 26 == 63)                           # zero normal -> all-ones field


def test_find_unique_clusters_background_and_overflow(vren):
    w, h = 256, 128
    depth = np.ones((h, w), np.float32)                              # cleared depth buffer: every pixel at the far plane
    oc, vc = both_cameras(vren, w, h)
    want_keys, want_ref = oracle.find_unique_clusters(depth, None, oc)
    keys, disp, ref = vren.find_unique_clusters(dev(depth), None, vc)
    assert host_u32(disp)[0] == want_keys.size == (w // 32) * (h // 32)
    assert np.array_equal(host_u32(keys)[: want_keys.size], want_keys)
    # overflow of the key list is detected (the reference silently overruns config.hpp:23)
    depth = synthetic.depth_buffer(w, h, seed=9)
    want_keys, _ = oracle.find_unique_clusters(depth, None, oc)
    cap = want_keys.size // 2
    keys, disp, ref = vren.find_unique_clusters(dev(depth), None, vc, max_keys=cap)
    disp = host_u32(disp)
    assert disp[3] == 1 and disp[0] == cap
    assert np.array_equal(host_u32(keys)[:cap], want_keys[:cap])


def run_chain(vren, w, h, L, seed, intensity=(1.0, 1.0), max_keys=1 << 17, max_assigned=1 << 23, yaw=0.0, with_normals=False):
    depth = synthetic.depth_buffer(w, h, seed=seed)
    normals = synthetic.normal_buffer(w, h, seed=seed + 9) if with_normals else None
    pos, lights = synthetic.point_lights(L, seed=seed + 1, aspect=w / h, intensity=intensity)
    view = synthetic.view_matrix(yaw, 0.0, (0.0, 0.0, 0.0))
    oc, vc = both_cameras(vren, w, h)
    # oracle chain
    wvp, wnodes, wpairs = oracle.construct_point_light_bvh(pos, lights, view)
    wkeys, wref = oracle.find_unique_clusters(depth, normals, oc)
    wcounts, woffsets, windices, wtotal = oracle.assign_lights(w, h, oc, wkeys, max_keys, wnodes, L, wpairs, wvp, max_assigned)
    # CUDA chain (a9 order: a6 -> a7 -> a8)
    import torch

    vp, bvh, idx = vren.construct_point_light_bvh(dev(pos), dev(lights), view.tolist())
    dn = None if normals is None else torch.from_numpy(normals.view(np.int16)).cuda().view(torch.float16)
    keys, disp, ref = vren.find_unique_clusters(dev(depth), dn, vc, max_keys=max_keys)
    counts, offsets, indices, status = vren.assign_lights(w, h, vc, keys, disp, bvh, L, idx, vp, max_keys=max_keys, max_assigned=max_assigned)
    return (wkeys, wcounts, woffsets, windices, wtotal), (host_u32(keys), host_u32(counts), host_u32(offsets), host_u32(indices), host_u32(status))


@pytest.mark.parametrize("w,h,L", [(64, 64, 1), (64, 64, 40), (256, 144, 1000), (640, 360, 5000), (1280, 720, 33000), (1920, 1080, 65536)])
def test_assign_lights_chain_matches_oracle(vren, w, h, L):
    (wkeys, wcounts, woffsets, windices, wtotal), (keys, counts, offsets, indices, status) = run_chain(vren, w, h, L, seed=w + L, intensity=(0.5, 3.0))
    assert np.array_equal(counts, wcounts)
    assert np.array_equal(offsets, woffsets)                          # exclusive scan over all 2^17 slots
    assert status[0] == wtotal and status[1] == 0
    assert np.array_equal(indices[:wtotal], windices[:wtotal])        # light order inside every list is part of the contract
    assert wtotal > 0 or L <= 40
    # keyed view (what a non-canonical list order would still have to satisfy)
    got = {int(k): indices[o:o + c].tolist() for k, o, c in zip(keys[: wkeys.size], offsets, counts)}
    want = {int(k): windices[o:o + c].tolist() for k, o, c in zip(wkeys, woffsets, wcounts)}
    assert got == want


def test_assign_lights_no_lights_and_overflow(vren):
    import torch

    w, h = 128, 64
    oc, vc = both_cameras(vren, w, h)
    depth = synthetic.depth_buffer(w, h, seed=5)
    keys, disp, ref = vren.find_unique_clusters(dev(depth), None, vc)
    # L == 0: counts zero-filled, nothing else touched (clustered_shading.cpp:494-496)
    dummy = torch.zeros(64, dtype=torch.uint8, device="cuda")
    counts, offsets, indices, status = vren.assign_lights(w, h, vc, keys, disp, dummy, 0, dummy, dummy.view(torch.float32), max_assigned=1024)
    assert int(counts.abs().sum()) == 0 and int(indices.abs().sum()) == 0
    # indices overflow is detected and writes are clamped (reference overruns VREN_MAX_ASSIGNED_LIGHT_COUNT silently)
    (wkeys, wcounts, woffsets, windices, wtotal), (keys, counts, offsets, indices, status) = run_chain(
        vren, 256, 144, 3000, seed=11, intensity=(20.0, 20.0), max_assigned=4096)
    assert wtotal > 4096 and status[0] == wtotal and status[1] == 1
    assert np.array_equal(counts, wcounts) and np.array_equal(indices, windices)


def test_cluster_chain_graph_replay_follows_new_inputs(vren):
    """ClusterAndShade.capture(): the chain recorded once in a CUDA graph gives the oracle's lists for whatever the
    buffers hold at replay time (a new frame: different depth buffer and light positions, same buffers and camera)"""
    import math

    import torch

    from vren_b200.pipeline import ClusterAndShade

    w, h, L = 640, 360, 5000
    oc, vc = both_cameras(vren, w, h)
    view = synthetic.view_matrix(0.2, 0.0, (0.0, 0.0, 0.0))
    depth = dev(synthetic.depth_buffer(w, h, seed=31))
    pos0, lights0 = synthetic.point_lights(L, seed=32, aspect=w / h, intensity=(0.5, 3.0))
    pos, lights = dev(pos0), dev(lights0)
    cs = ClusterAndShade(w, h, max_point_lights=L)
    graph = cs.capture(w, h, vc, view.tolist(), depth, None, pos, lights, L)
    for frame_seed in (41, 57):
        new_depth = synthetic.depth_buffer(w, h, seed=frame_seed)
        new_pos, new_lights = synthetic.point_lights(L, seed=frame_seed + 1, aspect=w / h, intensity=(0.5, 3.0))
        depth.copy_(dev(new_depth))
        pos.copy_(dev(new_pos))
        lights.copy_(dev(new_lights))
        graph.replay()
        torch.cuda.synchronize()
        wvp, wnodes, wpairs = oracle.construct_point_light_bvh(new_pos, new_lights, view)
        wkeys, _ = oracle.find_unique_clusters(new_depth, None, oc)
        wcounts, woffsets, windices, wtotal = oracle.assign_lights(w, h, oc, wkeys, cs.max_keys, wnodes, L, wpairs, wvp, cs.max_assigned)
        assert int(cs.dispatch_params[0]) == wkeys.size
        assert np.array_equal(host_u32(cs.cluster_keys)[: wkeys.size], wkeys)
        assert np.array_equal(host_u32(cs.counts), wcounts) and np.array_equal(host_u32(cs.offsets), woffsets)
        assert int(host_u32(cs.status)[0]) == wtotal
        assert np.array_equal(host_u32(cs.indices)[:wtotal], windices[:wtotal])


def test_cluster_chain_c5_full_size(vren):
    """BASELINE C5: 3840x2160 depth, 65 536 lights (intensity 1.0), one view"""
    (wkeys, wcounts, woffsets, windices, wtotal), (keys, counts, offsets, indices, status) = run_chain(vren, 3840, 2160, 65536, seed=2024)
    assert np.array_equal(keys[: wkeys.size], wkeys)
    assert np.array_equal(counts, wcounts) and np.array_equal(offsets, woffsets)
    assert status[0] == wtotal and status[1] == 0
    assert np.array_equal(indices[:wtotal], windices[:wtotal])


@pytest.mark.parametrize("w,h,L", [(96, 64, 50), (640, 360, 4000), (1920, 1080, 65536)])
def test_light_list_consumer_n1(vren, w, h, L):
    """per-pixel walk cluster_reference -> counts/offsets -> indices (shade.comp:101-105)"""
    depth = synthetic.depth_buffer(w, h, seed=w * 3 + L)
    pos, lights = synthetic.point_lights(L, seed=L + 1, aspect=w / h, intensity=(0.5, 3.0))
    view = synthetic.view_matrix(0.0, 0.0, (0.0, 0.0, 0.0))
    oc, vc = both_cameras(vren, w, h)
    vp, bvh, idx = vren.construct_point_light_bvh(dev(pos), dev(lights), view.tolist())
    keys, disp, ref = vren.find_unique_clusters(dev(depth), None, vc)
    counts, offsets, indices, status = vren.assign_lights(w, h, vc, keys, disp, bvh, L, idx, vp)
    got = host_u32(vren.light_list_hash(ref, disp, counts, offsets, indices)).reshape(h, w, 2)
    want = oracle.light_list_hash(host_u32(ref).reshape(h, w), host_u32(counts), host_u32(offsets), host_u32(indices))
    assert np.array_equal(got, want)
    assert int(got[..., 0].max()) > 0


@pytest.mark.parametrize("view_index", range(8))
def test_cluster_chain_c5_eight_views_with_normals(vren, view_index):
    """BASELINE C5 as SURVEY 8d spells it: 3840x2160, 65 536 lights, the 8 views of a batch (yaw += 45 degrees per view, a
    depth buffer per view; views 0 and 3 with RGBA16F normals, which multiply the clusters): keys, counts, offsets and
    ordered light lists bit-exact vs the oracle"""
    yaw = math.radians(45.0) * view_index
    (wkeys, wcounts, woffsets, windices, wtotal), (keys, counts, offsets, indices, status) = run_chain(
        vren, 3840, 2160, 65536, seed=2024 + view_index, yaw=yaw, with_normals=view_index in (0, 3), max_keys=1 << 20,
        max_assigned=(1 << 26) if view_index in (0, 3) else (1 << 23))     # normals: 35 M assignments in view 0
    assert np.array_equal(keys[: wkeys.size], wkeys)
    assert np.array_equal(counts, wcounts) and np.array_equal(offsets, woffsets)
    assert status[0] == wtotal and status[1] == 0
    assert np.array_equal(indices[:wtotal], windices[:wtotal])


def test_view_batch_single_gpu_matches_per_view_oracle(vren):
    """vren_b200.pipeline.ViewBatch with no process group: all views on this GPU, same results as view-by-view oracle runs"""
    import torch

    from vren_b200.pipeline import ViewBatch

    w, h, L, views = 640, 360, 4000, 4
    oc, vc = both_cameras(vren, w, h)
    pos0, lights0 = synthetic.point_lights(L, seed=71, aspect=w / h, intensity=(0.5, 3.0))
    vb = ViewBatch(w, h, L)
    vb.set_lights(dev(pos0), dev(lights0), L)
    frames, inputs = {}, {}
    for v in vb.my_views(views):
        view = synthetic.view_matrix(math.radians(45.0) * v, 0.0, (0.0, 0.0, 0.0))
        depth = synthetic.depth_buffer(w, h, seed=80 + v)
        frames[v] = (vc, view.tolist(), dev(depth), None)
        inputs[v] = (view, depth)
    got = {}

    def on_view(v, cs):
        got[v] = (host_u32(cs.cluster_keys).copy(), host_u32(cs.counts).copy(), host_u32(cs.offsets).copy(), host_u32(cs.indices).copy())

    res = vb(frames, on_view)
    torch.cuda.synchronize()
    assert sorted(res) == list(range(views))
    for v, (view, depth) in inputs.items():
        wvp, wnodes, wpairs = oracle.construct_point_light_bvh(pos0, lights0, view)
        wkeys, _ = oracle.find_unique_clusters(depth, None, oc)
        wcounts, woffsets, windices, wtotal = oracle.assign_lights(w, h, oc, wkeys, vb.cs.max_keys, wnodes, L, wpairs, wvp, vb.cs.max_assigned)
        keys, counts, offsets, indices = got[v]
        r = res[v].cpu().numpy().view(np.uint32)
        assert r[0] == wkeys.size and r[4] == wtotal
        assert np.array_equal(keys[: wkeys.size], wkeys) and np.array_equal(counts, wcounts) and np.array_equal(offsets, woffsets)
        assert np.array_equal(indices[:wtotal], windices[:wtotal])
