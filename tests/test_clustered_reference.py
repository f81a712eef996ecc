"""Parity of the clustered-shading pass (a6-a8) and of the rows next to it (n1-n3) PINNED TO THE REFERENCE'S OWN GLSL.

The functions that decide every integer output of the pass are pure: clustered_shading.glsl:7-43,58-110
(discretize_normal, decode_cluster_key, calc_cluster_aabb), assign_lights.comp:84-119 (test_aabb_aabb, test_sphere_aabb,
get_node_address), discretize_point_light_positions.comp:31-39 (Morton code), find_unique_clusters.comp:52-76 (depth ->
view z -> slice -> key), point_light_position_to_view_space.comp:30, init_light_array_bvh.comp:52-53, and for n1-n3
show_clusters.comp:111-116, depth_buffer_reduce.comp:12-34, bounce_point_lights.comp:37-71.  oracle/ref_extract.py compiles
them with g++ from the sources where they lie (through oracle/glsl_shim.hpp) into oracle/_ref/libvrenref_glsl.so;
tests/golden/make_clustered_golden.py ran them on the seeded inputs of tests/clustered_cases.py (>= 10^5 per function)
and stored the outputs in tests/golden/clustered_reference_outputs.npz.

  not gpu:  reference-compiled == oracle, bit for bit, live (where /root/reference exists); oracle == golden (everywhere)
  gpu:      CUDA == golden, through the product's entry points (a6, a7, n1-n3) and vrenb200_cluster_tests (a8's predicates)

Policy for GLSL builtins whose precision is implementation-defined (stated in oracle/glsl_shim.hpp): tan / log / pow = libm
tanf / logf / powf of the image; inverse(mat4) = the exact closed-form inverse of the perspective matrix.  The last test
quantifies what a generic cofactor inverse() changes: <= 1 ulp per matrix entry, <= 4 ulp per cluster corner, but up to
1.5 % in the view-space z of far pixels (w' = d iB + nAB is a cancellation), i.e. the slice of a far pixel is not fixed
by the reference's source alone.
"""
import ctypes as C
import math
from pathlib import Path

import numpy as np
import pytest

import clustered_cases as cc
import oracle

GOLDEN = Path(__file__).parent / "golden" / "clustered_reference_outputs.npz"


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def check_against_golden(golden, name, arr):
    """SHA-256 of the whole array must match; on a mismatch the verbatim head localises the first difference"""
    arr = np.ascontiguousarray(arr)
    assert tuple(golden[name + "__shape"]) == arr.shape, (name, arr.shape)
    if bytes.fromhex(cc.digest(arr)) == golden[name + "__sha256"].tobytes():
        return
    want, got = golden[name + "__head"], cc.head(arr)
    bad = np.nonzero(want.view(np.uint8).reshape(want.size, -1) != got.view(np.uint8).reshape(got.size, -1))[0]
    where = f"first difference at flat index {bad[0]}: golden {want[bad[0]]!r}, got {got[bad[0]]!r}" if bad.size else "difference beyond the stored head"
    raise AssertionError(f"{name}: differs from the reference-compiled output ({where})")


@pytest.fixture(scope="module")
def oracle_outputs(built):
    return cc.evaluate("oracle")


def test_golden_covers_at_least_1e5_inputs_per_function(golden):
    for name in ("normal_bins", "key_hi_4k_rows", "aabb_min", "aabb_aabb", "sphere_aabb", "view_pos", "morton", "leaf_min"):
        assert int(np.prod(golden[name + "__shape"])) >= 100_000, name


def test_oracle_equals_golden(golden, oracle_outputs):
    for name, arr in oracle_outputs.items():
        check_against_golden(golden, name, arr)


def test_reference_compiled_glsl_equals_oracle(built, oracle_outputs):
    if oracle.load_ref_glsl() is None:
        pytest.skip("oracle/_ref/libvrenref_glsl.so not built (no /root/reference here): covered by the golden file")
    ref = cc.evaluate("ref")
    assert set(ref) == set(oracle_outputs)
    for name, arr in ref.items():
        got = oracle_outputs[name]
        assert arr.shape == got.shape and arr.dtype == got.dtype, name
        assert np.array_equal(arr.view(np.uint8), got.view(np.uint8)), f"{name}: oracle differs from the reference-compiled GLSL"
    # the inputs exercise both outcomes of every predicate and the special cases
    assert 0.2 < ref["aabb_aabb"].mean() < 0.8 and 0.2 < ref["sphere_aabb"].mean() < 0.8
    assert (ref["normal_bins"] == 0xFFFFFFFF).sum() > 1000 and np.unique(ref["normal_bins"]).size >= 50
    assert np.unique(ref["key_hi_4k_rows"] & 0x3FF).size > 300


def test_generic_inverse_policy(built):
    """what the implementation-defined inverse() changes (see the module docstring)"""
    lib = oracle.load_ref_glsl()
    if lib is None:
        pytest.skip("needs oracle/_ref/libvrenref_glsl.so")
    _, proj, closed = cc.camera_matrices(16.0 / 9.0)
    generic = np.zeros(16, np.float32)
    lib.refglsl_set_inverse_override(None)
    lib.refglsl_inverse(proj, generic)
    nz = closed != 0
    assert np.array_equal(generic != 0, nz)
    ulp = np.abs(generic.view(np.int32).astype(np.int64) - closed.view(np.int32).astype(np.int64))[nz]
    assert ulp.max() <= 1
    pinned, loose = cc.evaluate("ref"), cc.evaluate("ref", generic_inverse=True)
    for name in ("aabb_min", "aabb_max"):
        d = np.abs(pinned[name].view(np.int32).astype(np.int64) - loose[name].view(np.int32).astype(np.int64))
        assert d.max() <= 4
        assert np.max(np.abs(pinned[name] - loose[name]) / np.maximum(np.abs(pinned[name]), np.float32(1e-30))) < 1e-6      # the north-star's AABB tolerance
    rel = np.abs(pinned["view_z_4k_rows"] - loose["view_z_4k_rows"]) / pinned["view_z_4k_rows"]
    assert 1e-3 < rel.max() < 0.05
    assert (pinned["key_hi_4k_rows"] != loose["key_hi_4k_rows"]).mean() > 0.01


# ---- CUDA == golden ---------------------------------------------------------------------------------------------------------
def _cam(vren, aspect):
    return vren.Camera(cc.FOV_Y, np.float32(aspect), cc.NEAR, cc.FAR)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(cc.IMAGES))
def test_cuda_cluster_keys_equal_reference(vren, golden, name):
    """a7 through vrenb200_find_unique_clusters: the key of every pixel (slice and normal bin) == reference-compiled"""
    import torch

    depth, nimg, _ = cc.image(name)
    h, w = depth.shape
    d = torch.from_numpy(depth).cuda()
    nr = torch.from_numpy(cc.rgba16f(nimg).view(np.int16)).cuda().view(torch.float16)
    keys, disp, ref = vren.find_unique_clusters(d, nr, _cam(vren, 16.0 / 9.0), max_keys=w * h)
    keys, ref = keys.cpu().numpy().view(np.uint32), ref.cpu().numpy().view(np.uint32)
    count = int(disp.cpu().numpy().view(np.uint32)[0])
    assert count <= w * h and int(disp.cpu().numpy().view(np.uint32)[3]) == 0
    px = keys[ref]                                             # key of every pixel through its cluster reference
    ys, xs = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    assert np.array_equal(px & 0xFFFF, ((xs >> 5) | ((ys >> 5) << 8)).astype(np.uint32))
    check_against_golden(golden, f"key_hi_{name}", (px >> 16).astype(np.uint16))


@pytest.mark.gpu
def test_cuda_light_chain_equals_reference(vren, golden):
    """a6 through vrenb200_construct_point_light_bvh: view positions, Morton codes and leaf boxes == reference-compiled"""
    import torch

    pos, intensity, view = cc.light_chain()
    lights = np.zeros((cc.N, 4), np.float32)
    lights[:, 3] = intensity
    vp, bvh, idx = vren.construct_point_light_bvh(torch.from_numpy(pos).cuda(), torch.from_numpy(lights).cuda(), view.tolist())
    check_against_golden(golden, "view_pos", vp.cpu().numpy())
    pairs = idx.cpu().numpy()[: cc.N * 8].view(np.uint32).reshape(cc.N, 2)
    morton = np.zeros(cc.N, np.uint32)
    morton[pairs[:, 1]] = pairs[:, 0]
    assert np.array_equal(np.sort(pairs[:, 1]), np.arange(cc.N, dtype=np.uint32))
    check_against_golden(golden, "morton", morton)
    nodes = bvh.cpu().numpy()[: cc.N * 32].view(oracle.BVH_NODE)
    lmin, lmax = np.zeros((cc.N, 3), np.float32), np.zeros((cc.N, 3), np.float32)
    lmin[pairs[:, 1]] = nodes["min"]                           # leaf i holds light pairs[i].y
    lmax[pairs[:, 1]] = nodes["max"]
    check_against_golden(golden, "leaf_min", lmin)
    check_against_golden(golden, "leaf_max", lmax)


@pytest.mark.gpu
def test_cuda_cluster_corners_and_overlap_tests_equal_reference(vren, golden):
    """a8's device functions through vrenb200_cluster_tests: calc_cluster_aabb, test_aabb_aabb, test_sphere_aabb"""
    import torch

    cells = cc.cluster_cells()
    keys = torch.from_numpy(cc.keys_of(cells).view(np.int32)).cuda()
    cam = _cam(vren, 3840.0 / 2160.0)
    cmin, cmax, _ = vren.cluster_tests(3840, 2160, cam, keys)
    cmin, cmax = cmin.cpu().numpy(), cmax.cpu().numpy()
    check_against_golden(golden, "aabb_min", cmin)
    check_against_golden(golden, "aabb_max", cmax)
    boxes, sph = cc.node_boxes(cmin, cmax, 71), cc.spheres(cmin, cmax, 81)
    _, _, flags = vren.cluster_tests(3840, 2160, cam, keys, torch.from_numpy(boxes).cuda(), torch.from_numpy(sph).cuda())
    flags = flags.cpu().numpy()
    check_against_golden(golden, "aabb_aabb", flags & 1)
    check_against_golden(golden, "sphere_aabb", (flags >> 1) & 1)


@pytest.mark.gpu
def test_cuda_next_rows_equal_reference(vren, golden):
    """n2 depth_buffer_reduce, n3 bounce_point_lights, n1 the list XOR of show_clusters.comp: CUDA == reference-compiled"""
    import torch

    dimg = cc.depth_image_n2()
    h, w = dimg.shape
    pyr = vren.depth_pyramid(torch.from_numpy(dimg).cuda()).cpu().numpy()
    check_against_golden(golden, "depth_reduce", pyr[w * h:w * h + (w >> 1) * (h >> 1)].reshape(h >> 1, w >> 1))
    bp, bd, lo, hi, speed, dt = cc.bounce_case()
    tp, td = torch.from_numpy(bp).cuda(), torch.from_numpy(bd).cuda()
    for _ in range(3):
        vren.bounce_point_lights(tp, td, lo, hi, float(speed), float(dt))
    check_against_golden(golden, "bounce_pos", tp.cpu().numpy())
    check_against_golden(golden, "bounce_dir", td.cpu().numpy())
    counts, offsets, indices = cc.light_lists()
    n = counts.size
    disp = torch.tensor([n, 1, 1, 0], dtype=torch.int32).cuda()
    cref = torch.arange(n, dtype=torch.int32).reshape(1, n).cuda()
    out = vren.light_list_hash(cref, disp, torch.from_numpy(counts.view(np.int32)).cuda(), torch.from_numpy(offsets.view(np.int32)).cuda(),
                               torch.from_numpy(indices.view(np.int32)).cuda()).cpu().numpy().view(np.uint32)
    assert np.array_equal(out[0, :, 0], counts)
    check_against_golden(golden, "list_xor", out[0, :, 1])
