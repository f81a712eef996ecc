"""TEST INFRASTRUCTURE: ctypes loader for the CPU oracle (oracle/liboracle.so, built from oracle/*.cpp).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package (vren_b200/) must never import this module.
"""
from __future__ import annotations

import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "liboracle.so"
REF_LIB_PATH = HERE / "_ref" / "libvrenref.so"

_lib = None
_ref = None

u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
voidp = C.c_void_p


def _build():
    sys.path.insert(0, str(HERE.parent))
    from vren_b200.build import build_oracle

    build_oracle()


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        _build()
    lib = C.CDLL(str(LIB_PATH))
    u32, u64, i32 = C.c_uint32, C.c_uint64, C.c_int

    def sig(name, res, *args):
        f = getattr(lib, name)
        f.restype = res
        f.argtypes = list(args)

    sig("oracle_is_power_of_2", i32, u32)
    sig("oracle_round_to_next_power_of_2", u32, u32)
    sig("oracle_round_to_next_multiple_of", u64, u64, u64)
    sig("oracle_divide_and_ceil", u32, u32, u32)
    sig("oracle_is_power_of", i32, u32, u32)
    sig("oracle_round_to_next_power_of", u32, u32, u32)
    sig("oracle_calc_bvh_padded_leaf_count", u32, u32)
    sig("oracle_calc_bvh_buffer_length", u32, u32)
    sig("oracle_calc_bvh_buffer_size", u64, u32)
    sig("oracle_calc_bvh_root_index", u32, u32)
    sig("oracle_calc_bvh_level_count", u32, u32)
    sig("oracle_reduce", None, i32, i32, voidp, u32, voidp, u32)
    sig("oracle_test_cpu_reduce_u32", None, i32, u32p, u32)
    sig("oracle_test_cpu_reduce_f32", None, i32, f32p, u32, u32)
    sig("oracle_exclusive_scan_u32", None, u32p, u32p, u32)
    sig("oracle_downsweep_u32", None, u32p, u32, u32, i32)
    sig("oracle_blelloch_scan_u32", None, u32p, u32, u32)
    sig("oracle_sort_keys", None, u32p, u32)
    sig("oracle_radix_sort_lsd4", None, u32p, u32)
    sig("oracle_sort_pairs", None, u32p, u32p, u32)
    sig("oracle_sort_pairs_interleaved", None, u64p, u32)
    sig("oracle_sort_pairs_interleaved_mt", None, u64p, u32, u32)
    sig("oracle_sort_keys_mt", None, u32p, u32, u32)
    sig("oracle_bucket_sort", None, u32p, u32, u32p, u32p)
    sig("oracle_build_bvh", None, voidp, u32)
    sig("oracle_bvh_traverse_point", u32, voidp, u32, f32p, u32p, u32)
    sig("oracle_bvh_linear_point", u32, voidp, u32, f32p, u32p, u32)
    for name, res, args in _LATE:
        if hasattr(lib, name):
            sig(name, res, *args)
    _lib = lib
    return lib


class Camera(C.Structure):
    _fields_ = [("fov_y", C.c_float), ("aspect_ratio", C.c_float), ("near_plane", C.c_float), ("far_plane", C.c_float)]


u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")

_LATE: list = [
    ("oracle_projection", None, (C.c_void_p, f32p, f32p)),
    ("oracle_discretize_normal", None, (f32p, C.c_uint32, u32p)),
    ("oracle_cluster_key", None, (f32p, f32p, C.c_uint32, C.c_uint32, C.c_void_p, u32p, f32p)),
    ("oracle_cluster_aabb", None, (u32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, f32p, f32p)),
    ("oracle_test_aabb_aabb", None, (f32p, C.c_uint32, u8p)),
    ("oracle_test_sphere_aabb", None, (f32p, C.c_uint32, u8p)),
    ("oracle_get_node_address", C.c_int32, (C.c_uint32, C.c_uint32, u32p)),
    ("oracle_morton_code", None, (f32p, C.c_uint32, f32p, f32p, u32p)),
    ("oracle_position_to_view_space", None, (f32p, f32p, C.c_uint32, f32p)),
    ("oracle_light_leaf_box", None, (f32p, f32p, C.c_uint32, f32p, f32p)),
    ("oracle_light_list_hash", None, (C.c_uint32, C.c_uint32, u32p, u32p, u32p, u32p, u32p)),
    ("oracle_depth_pyramid", C.c_uint32, (f32p, C.c_uint32, C.c_uint32, f32p)),
    ("oracle_bounce_point_lights", None, (f32p, f32p, C.c_uint32, f32p, f32p, C.c_float, C.c_float)),
    ("oracle_visualize_bvh", None, (u32p, C.c_uint32, u32p)),
    ("oracle_construct_point_light_bvh", None, (f32p, f32p, C.c_uint32, f32p, f32p, voidp, u32p)),
    ("oracle_find_unique_clusters", C.c_uint32, (f32p, voidp, C.c_uint32, C.c_uint32, C.POINTER(Camera), u32p, u32p)),
    ("oracle_assign_lights", C.c_uint64,
     (C.c_uint32, C.c_uint32, C.POINTER(Camera), u32p, C.c_uint32, C.c_uint32, voidp, C.c_uint32, C.c_uint32, u32p, f32p,
      u32p, C.c_uint64, u32p, u32p)),
]


def default_camera(width: int, height: int, fov_deg: float = 45.0, near: float = 0.01, far: float = 1000.0) -> Camera:
    """vren::camera defaults (camera.hpp:20-23) with the screen's aspect ratio"""
    import math

    return Camera(np.float32(math.radians(fov_deg)), np.float32(width / height), np.float32(near), np.float32(far))


def construct_point_light_bvh(positions: np.ndarray, lights: np.ndarray, view: np.ndarray):
    """positions [L,4] f32, lights [L,4] f32 (rgb, intensity), view: 16 floats column-major.
    Returns (view_pos [L,4], nodes BVH_NODE[len], sorted_pairs [L,2])"""
    lib = load()
    L = positions.shape[0]
    positions = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1)
    lights = np.ascontiguousarray(lights, dtype=np.float32).reshape(-1)
    view = np.ascontiguousarray(view, dtype=np.float32).reshape(-1)
    view_pos = np.zeros(L * 4, np.float32)
    nodes = np.zeros(int(lib.oracle_calc_bvh_buffer_length(L)), dtype=BVH_NODE)
    pairs = np.zeros(L * 2, np.uint32)
    lib.oracle_construct_point_light_bvh(positions, lights, L, view, view_pos, nodes.ctypes.data, pairs)
    return view_pos.reshape(L, 4), nodes, pairs.reshape(L, 2)


def find_unique_clusters(depth: np.ndarray, normals, cam: Camera):
    """depth [H,W] f32, normals [H,W,4] float16 or None -> (keys[count], cluster_ref [H,W])"""
    lib = load()
    H, W = depth.shape
    depth = np.ascontiguousarray(depth, dtype=np.float32)
    tiles = ((W + 31) // 32) * ((H + 31) // 32)
    keys = np.zeros(tiles * 1024, np.uint32)
    ref = np.zeros(H * W, np.uint32)
    nptr = None
    if normals is not None:
        normals = np.ascontiguousarray(normals, dtype=np.float16)
        nptr = normals.ctypes.data
    count = lib.oracle_find_unique_clusters(depth.reshape(-1), nptr, W, H, C.byref(cam), keys, ref)
    return keys[:count].copy(), ref.reshape(H, W)


def assign_lights(W, H, cam: Camera, keys: np.ndarray, max_keys: int, nodes: np.ndarray, light_count: int,
                  sorted_pairs: np.ndarray, view_pos: np.ndarray, max_assigned: int):
    """-> (counts[max_keys], offsets[max_keys], indices[max_assigned], total)"""
    lib = load()
    keys = np.ascontiguousarray(keys, dtype=np.uint32)
    counts = np.zeros(max_keys, np.uint32)
    offsets = np.zeros(max_keys, np.uint32)
    indices = np.zeros(max_assigned, np.uint32)
    nodes = np.ascontiguousarray(nodes)
    root = int(lib.oracle_calc_bvh_root_index(light_count))
    total = lib.oracle_assign_lights(W, H, C.byref(cam), keys, keys.size, max_keys, nodes.ctypes.data, root, light_count,
                                     np.ascontiguousarray(sorted_pairs, dtype=np.uint32).reshape(-1),
                                     np.ascontiguousarray(view_pos, dtype=np.float32).reshape(-1),
                                     indices, max_assigned, counts, offsets)
    return counts, offsets, indices, int(total)


def load_ref():
    """oracle/_ref/libvrenref.so — functions compiled from the reference sources where they lie (None if absent)."""
    global _ref
    if _ref is not None:
        return _ref
    if not REF_LIB_PATH.exists():
        return None
    _ref = C.CDLL(str(REF_LIB_PATH))
    return _ref


REF_GLSL_LIB_PATH = HERE / "_ref" / "libvrenref_glsl.so"
_ref_glsl = None


def load_ref_glsl():
    """oracle/_ref/libvrenref_glsl.so — pure functions of the reference's compute shaders compiled by g++ through
    oracle/glsl_shim.hpp from the GLSL where it lies (oracle/ref_extract.py).  None where /root/reference is absent."""
    global _ref_glsl
    if _ref_glsl is not None:
        return _ref_glsl
    if not REF_GLSL_LIB_PATH.exists():
        return None
    lib = C.CDLL(str(REF_GLSL_LIB_PATH))
    u32, f32, i32 = C.c_uint32, C.c_float, C.c_int
    u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")

    def sig(name, res, *args):
        f = getattr(lib, name)
        f.restype = res
        f.argtypes = list(args)

    sig("refglsl_set_inverse_override", None, voidp)
    sig("refglsl_inverse", None, f32p, f32p)
    sig("refglsl_discretize_normal", None, f32p, u32, u32p)
    sig("refglsl_decode_cluster_key", None, u32p, u32, u32p)
    sig("refglsl_calc_cluster_aabb", None, u32p, u32, u32, u32, f32, f32, f32p, f32p, f32p)
    sig("refglsl_test_aabb_aabb", None, f32p, u32, u8p)
    sig("refglsl_test_sphere_aabb", None, f32p, u32, u8p)
    sig("refglsl_get_node_address", C.c_int32, u32, u32, u32p)
    sig("refglsl_morton_code", None, f32p, u32, f32p, f32p, u32p)
    sig("refglsl_cluster_key", None, f32p, f32p, f32p, u32, u32, u32, u32, f32, f32, f32p, u32p, f32p)
    sig("refglsl_position_to_view_space", None, f32p, f32p, u32, f32p)
    sig("refglsl_light_leaf_box", None, f32p, f32p, u32, f32p, f32p)
    sig("refglsl_depth_reduce", None, f32p, i32, i32, f32p, i32, i32)
    sig("refglsl_bounce_point_lights", None, f32p, f32p, u32, f32p, f32p, f32, f32)
    sig("refglsl_light_list_xor", None, u32p, u32p, u32p, u32, u32, u32p)
    _ref_glsl = lib
    return lib


# ---- numpy convenience wrappers ---------------------------------------------------------------------------------
DT = {"u32": 0, "vec4": 1, "f32": 2}
OP = {"add": 0, "min": 1, "max": 2}


def next_pow2(n: int) -> int:
    return int(load().oracle_round_to_next_power_of_2(n))


def reduce(inp: np.ndarray, n: int, dtype: str, op: str, blocks: int = 1) -> np.ndarray:
    lib = load()
    P = next_pow2(n)
    comps = 4 if dtype == "vec4" else 1
    inp = np.ascontiguousarray(inp)
    out = np.zeros(blocks * P * comps, dtype=inp.dtype)
    lib.oracle_reduce(DT[dtype], OP[op], inp.ctypes.data, n, out.ctypes.data, blocks)
    return out


def exclusive_scan(inp: np.ndarray) -> np.ndarray:
    inp = np.ascontiguousarray(inp, dtype=np.uint32)
    out = np.empty_like(inp)
    load().oracle_exclusive_scan_u32(inp, out, inp.size)
    return out


def downsweep(buf: np.ndarray, n: int, blocks: int, clear_last: bool) -> np.ndarray:
    buf = np.ascontiguousarray(buf, dtype=np.uint32).copy()
    load().oracle_downsweep_u32(buf, n, blocks, int(clear_last))
    return buf


def blelloch_scan(buf: np.ndarray, n: int, blocks: int = 1) -> np.ndarray:
    buf = np.ascontiguousarray(buf, dtype=np.uint32).copy()
    load().oracle_blelloch_scan_u32(buf, n, blocks)
    return buf


def sort_keys(keys: np.ndarray) -> np.ndarray:
    keys = np.ascontiguousarray(keys, dtype=np.uint32).copy()
    load().oracle_sort_keys(keys, keys.size)
    return keys


def radix_sort_lsd4(keys: np.ndarray) -> np.ndarray:
    keys = np.ascontiguousarray(keys, dtype=np.uint32).copy()
    load().oracle_radix_sort_lsd4(keys, keys.size)
    return keys


def sort_pairs(keys: np.ndarray, vals: np.ndarray):
    keys = np.ascontiguousarray(keys, dtype=np.uint32).copy()
    vals = np.ascontiguousarray(vals, dtype=np.uint32).copy()
    load().oracle_sort_pairs(keys, vals, keys.size)
    return keys, vals


def bucket_sort(pairs: np.ndarray):
    """pairs: uint32 [n,2] -> (sorted [n,2], counters[65536])"""
    pairs = np.ascontiguousarray(pairs, dtype=np.uint32)
    n = pairs.shape[0]
    out = np.empty_like(pairs)
    counters = np.empty(65536, dtype=np.uint32)
    load().oracle_bucket_sort(pairs.reshape(-1), n, out.reshape(-1), counters)
    return out, counters


BVH_NODE = np.dtype([("min", np.float32, 3), ("next", np.uint32), ("max", np.float32, 3), ("pad", np.uint32)])


def build_bvh(nodes: np.ndarray, padded: int) -> np.ndarray:
    """nodes: structured BVH_NODE array of calc_bvh_buffer_length(padded) entries, leaves pre-filled"""
    nodes = np.ascontiguousarray(nodes).copy()
    load().oracle_build_bvh(nodes.ctypes.data, padded)
    return nodes


def depth_pyramid(depth: np.ndarray):
    """-> (flat pyramid float32, level_count)"""
    lib = load()
    H, W = depth.shape
    depth = np.ascontiguousarray(depth, dtype=np.float32)
    total, l = 0, 0
    while True:
        w, h = max(W >> l, 1), max(H >> l, 1)
        total += w * h
        l += 1
        if w == 1 and h == 1:
            break
    out = np.zeros(total, np.float32)
    levels = lib.oracle_depth_pyramid(depth.reshape(-1), W, H, out)
    return out, int(levels)


def bounce_point_lights(positions: np.ndarray, directions: np.ndarray, aabb_min, aabb_max, speed: float, dt: float):
    """bounce_point_lights.comp:33-73 -> (positions, directions), float32 [L,4] copies"""
    pos = np.ascontiguousarray(positions, dtype=np.float32).copy()
    dirs = np.ascontiguousarray(directions, dtype=np.float32).copy()
    lo, hi = np.asarray(aabb_min, np.float32).copy(), np.asarray(aabb_max, np.float32).copy()
    load().oracle_bounce_point_lights(pos.reshape(-1), dirs.reshape(-1), pos.shape[0], lo, hi, np.float32(speed), np.float32(dt))
    return pos, dirs


def visualize_bvh(nodes: np.ndarray, level_count: int) -> np.ndarray:
    """visualize_bvh.cpp:59-94 + show_bvh.comp:62-78.  nodes: the BVH buffer viewed as uint32 [N, 8] -> uint32 [N * 24, 4]"""
    nodes = np.ascontiguousarray(nodes).view(np.uint32).reshape(-1, 8)
    count = sum(32 ** l for l in range(level_count + 1))
    assert nodes.shape[0] >= count
    out = np.zeros((count * 24, 4), np.uint32)
    load().oracle_visualize_bvh(nodes.reshape(-1), level_count, out.reshape(-1))
    return out


def light_list_hash(cluster_ref: np.ndarray, counts: np.ndarray, offsets: np.ndarray, indices: np.ndarray) -> np.ndarray:
    H, W = cluster_ref.shape
    out = np.zeros(H * W * 2, np.uint32)
    load().oracle_light_list_hash(W, H, np.ascontiguousarray(cluster_ref, dtype=np.uint32).reshape(-1), np.ascontiguousarray(counts, dtype=np.uint32),
                                  np.ascontiguousarray(offsets, dtype=np.uint32), np.ascontiguousarray(indices, dtype=np.uint32), out)
    return out.reshape(H, W, 2)
