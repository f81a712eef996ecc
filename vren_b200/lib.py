"""ctypes binding of libvrenb200.so (the C ABI declared in include/vrenb200.h).

PyTorch is used only as the owner of device memory and streams; every compute call goes through the C ABI.
There is NO CPU fallback: if the shared library is missing the import of the symbols fails loudly.
"""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB_PATH = Path(__file__).resolve().parent / "libvrenb200.so"
HEADER = ROOT / "include" / "vrenb200.h"

OK, EINVAL_LENGTH, EALIGN, ESCRATCH, ECUDA, EINVAL_ARG, ELIMIT = range(7)
STATUS_NAMES = {0: "OK", 1: "EINVAL_LENGTH", 2: "EALIGN", 3: "ESCRATCH", 4: "ECUDA", 5: "EINVAL_ARG", 6: "ELIMIT"}
U32, VEC4, F32 = 0, 1, 2
ADD, MIN, MAX = 0, 1, 2
REDUCE_TREE, REDUCE_FINAL = 0, 1
BVH_LEAF_NODE = 0xFFFFFFFF
BVH_INVALID_NODE = 0xFFFFFFFE


class VrenError(RuntimeError):
    def __init__(self, status: int, what: str):
        self.status = status
        super().__init__(f"{what}: status {status} ({STATUS_NAMES.get(status, '?')})")


RANKING_AUTO, RANKING_MATCH, RANKING_ATOMIC_VERIFIED, RANKING_ATOMIC_SAMPLED, RANKING_ATOMIC_UNVERIFIED, RANKING_SELFTEST_REDO = range(6)
TILE_IDS_AUTO, TILE_IDS_BLOCK_INDEX, TILE_IDS_TICKET = range(3)
SCAN_TILE_IDS_TICKET = 1          # vrenb200_exclusive_scan_u32_ex flag: the scan's safe mode, opt-in
SORT_VARIANT_SINGLE_CTA = -1      # vrenb200_sort_config::variant: the whole sort in one CTA (n <= 8192), opt-in


class SortConfig(C.Structure):
    """vrenb200_sort_config: how one sort call runs (0 = library default)"""
    _fields_ = [("ranking", C.c_int), ("tile_ids", C.c_int), ("variant", C.c_int)]


class Camera(C.Structure):
    _fields_ = [("fov_y", C.c_float), ("aspect_ratio", C.c_float), ("near_plane", C.c_float), ("far_plane", C.c_float)]


def declared_symbols() -> list[str]:
    """every function name include/vrenb200.h declares (used by the CPU-side export test)"""
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"#ifdef VRENB200_TUNING.*?#endif", "", text, flags=re.S)     # hooks of tuning builds only
    return sorted(set(re.findall(r"\b(vrenb200_[a-z0-9_]+)\s*\(", text)))


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m vren_b200.build` (nvcc, sm_100a). "
            "There is no CPU fallback for the vren_b200 compute path."
        )
    lib = C.CDLL(str(LIB_PATH))
    vp, u32, u64, sz, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_size_t, C.c_int

    def sig(name, res, *args):
        f = getattr(lib, name)
        f.restype = res
        f.argtypes = list(args)

    sig("vrenb200_version", C.c_char_p)
    sig("vrenb200_status_string", C.c_char_p, i32)
    sig("vrenb200_last_cuda_error", i32)
    sig("vrenb200_is_power_of_2", i32, u32)
    sig("vrenb200_round_to_next_power_of_2", u32, u32)
    sig("vrenb200_round_to_next_multiple_of", u64, u64, u64)
    sig("vrenb200_divide_and_ceil", u32, u32, u32)
    sig("vrenb200_is_power_of", i32, u32, u32)
    sig("vrenb200_round_to_next_power_of", u32, u32, u32)
    sig("vrenb200_calc_reduce_output_buffer_length", u32, u32)
    sig("vrenb200_reduce_scratch_bytes", sz, i32, i32, u32, u32)
    sig("vrenb200_reduce", i32, vp, i32, i32, i32, vp, u32, vp, u32, vp, sz)
    sig("vrenb200_scan_scratch_bytes", sz, u32)
    sig("vrenb200_exclusive_scan_u32", i32, vp, vp, vp, u32, vp, sz)
    sig("vrenb200_blelloch_downsweep_u32", i32, vp, vp, u32, u32, i32)
    sig("vrenb200_radix_sort_scratch_bytes", sz, u32, i32)
    sig("vrenb200_radix_sort_keys", i32, vp, vp, u32, vp, sz)
    sig("vrenb200_radix_sort_pairs", i32, vp, vp, vp, u32, vp, sz)
    sig("vrenb200_radix_sort_scratch_buffer_1_bytes", sz, u32)
    sig("vrenb200_radix_sort_scratch_buffer_2_bytes", sz, u32)
    sig("vrenb200_radix_sort_compat", i32, vp, vp, u32, vp, sz, vp, sz)
    sig("vrenb200_radix_sort_host_work_bytes", sz, u32, i32)
    sig("vrenb200_radix_sort_pairs_host", i32, vp, vp, vp, u32, vp, sz)
    sig("vrenb200_radix_sort_selected_variant_name", C.c_char_p, u32, i32, vp)
    sig("vrenb200_radix_sort_num_variants", i32)
    sig("vrenb200_radix_sort_variant_name", C.c_char_p, i32)
    sig("vrenb200_sort_profile_create", vp)
    sig("vrenb200_sort_profile_destroy", None, vp)
    sig("vrenb200_sort_profile_read", i32, vp, C.POINTER(C.c_float))
    sig("vrenb200_radix_sort_ex", i32, vp, vp, vp, u32, vp, sz, vp, vp)
    sig("vrenb200_radix_sort_violation_word", vp, vp, u32, i32)
    for name, res, args in _LATE_SIGS:
        if hasattr(lib, name):
            sig(name, res, *args)
    _lib = lib
    return lib


_vp, _u32, _sz, _i32 = C.c_void_p, C.c_uint32, C.c_size_t, C.c_int
_LATE_SIGS = [
    ("vrenb200_sharded_sort_symmetric_bytes", _sz, (_u32,)),
    ("vrenb200_sharded_sort_local_bytes", _sz, (_u32, _u32, _vp)),
    ("vrenb200_sharded_sort_create", _i32, (_vp, _u32, _u32, _u32, _u32, _u32, _vp, _vp, _sz, _vp)),
    ("vrenb200_sharded_sort_destroy", None, (_vp,)),
    ("vrenb200_sharded_sort_pairs", _i32, (_vp, _vp, _vp, _vp, _u32, _i32)),
    ("vrenb200_sharded_sort_out_keys", _vp, (_vp,)),
    ("vrenb200_sharded_sort_out_values", _vp, (_vp,)),
    ("vrenb200_sharded_sort_status", _vp, (_vp,)),
    ("vrenb200_sharded_sort_phases", _i32, (_vp, C.POINTER(C.c_float))),
    ("vrenb200_cluster_tests", _i32, (_vp, _u32, _u32, _vp, _vp, _u32, _vp, _vp, _vp, _vp, _vp)),
    ("vrenb200_light_list_hash_scratch_bytes", _sz, (_u32,)),
    ("vrenb200_light_list_hash", _i32, (_vp, _u32, _u32, _vp, _vp, _u32, _vp, _vp, _vp, _u32, _vp, _vp, _sz)),
    ("vrenb200_depth_pyramid_level_count", _u32, (_u32, _u32)),
    ("vrenb200_depth_pyramid_level_width", _u32, (_u32, _u32)),
    ("vrenb200_depth_pyramid_level_height", _u32, (_u32, _u32)),
    ("vrenb200_depth_pyramid_level_offset", _sz, (_u32, _u32, _u32)),
    ("vrenb200_depth_pyramid_bytes", _sz, (_u32, _u32)),
    ("vrenb200_depth_pyramid_build", _i32, (_vp, _vp, _u32, _u32, _vp)),
    ("vrenb200_radix_top_digit_histogram", _i32, (_vp, _vp, _u32, _vp)),
    ("vrenb200_radix_partition_scatter", _i32, (_vp, _vp, _vp, _u32, _vp, _vp, _sz)),
    ("vrenb200_exclusive_scan_u32_base", _i32, (_vp, _vp, _vp, _u32, _u32, _vp, _sz)),
    ("vrenb200_exclusive_scan_u32_ex", _i32, (_vp, _vp, _vp, _u32, _u32, _vp, _sz, _u32)),
    ("vrenb200_radix_digit_histograms", _i32, (_vp, _vp, _u32, _vp)),
    ("vrenb200_bounce_point_lights", _i32, (_vp, _vp, _vp, _u32, C.POINTER(C.c_float * 3), C.POINTER(C.c_float * 3), C.c_float, C.c_float)),
    ("vrenb200_radix_sort_pairs_host_async", _i32, (_vp, _vp, _vp, _vp, _vp, _u32, _vp, _sz)),
    ("vrenb200_visualize_bvh_vertex_count", C.c_uint64, (_u32,)),
    ("vrenb200_visualize_bvh", _i32, (_vp, _vp, _u32, _vp)),
    ("vrenb200_radix_partition_set_shape", _i32, (_i32,)),
    ("vrenb200_kd_tree_build", _sz, (_vp, _sz, _vp, _sz, _vp, _sz)),
    ("vrenb200_kd_tree_search", None, (_vp, _sz, _vp, _vp, _vp, _vp, _vp, _vp)),
    ("vrenb200_kd_tree_search_batch", _i32, (_vp, _vp, _u32, _vp, _vp, _u32, _vp, _vp)),
    ("vrenb200_scan_set_variant", _i32, (_i32,)),          # tuning builds only
    ("vrenb200_scan_set_runahead", _i32, (_i32, _i32)),    # tuning builds only
    ("vrenb200_radix_sort_range_scratch_bytes", _sz, (_u32,)),
    ("vrenb200_radix_sort_pairs_range", _i32, (_vp, _vp, _vp, _vp, _vp, _u32, _i32, _i32, _vp, _sz, C.POINTER(C.c_int))),
    ("vrenb200_bucket_sort_output_bytes", _sz, (_u32,)),
    ("vrenb200_bucket_sort_scratch_bytes", _sz, (_u32,)),
    ("vrenb200_bucket_sort", _i32, (_vp, _vp, _u32, _vp, _vp, _sz)),
    ("vrenb200_bucket_sort_ex", _i32, (_vp, _vp, _u32, _vp, _vp, _sz, _vp, _i32)),
    ("vrenb200_calc_bvh_padded_leaf_count", _u32, (_u32,)),
    ("vrenb200_calc_bvh_buffer_length", _u32, (_u32,)),
    ("vrenb200_calc_bvh_buffer_size", _sz, (_u32,)),
    ("vrenb200_calc_bvh_root_index", _u32, (_u32,)),
    ("vrenb200_calc_bvh_level_count", _u32, (_u32,)),
    ("vrenb200_build_bvh", _i32, (_vp, _vp, _u32)),
    ("vrenb200_light_bvh_buffer_bytes", _sz, (_u32,)),
    ("vrenb200_light_index_buffer_bytes", _sz, (_u32,)),
    ("vrenb200_light_bvh_scratch_bytes", _sz, (_u32,)),
    ("vrenb200_construct_point_light_bvh", _i32, (_vp, _vp, _vp, _u32, _vp, _vp, _vp, _vp, _vp, _sz)),
    ("vrenb200_find_unique_clusters_scratch_bytes", _sz, (_u32, _u32)),
    ("vrenb200_find_unique_clusters", _i32,
     (_vp, _vp, _vp, _u32, _u32, C.POINTER(Camera), _vp, _u32, _vp, _vp, _vp, _sz)),
    ("vrenb200_assign_lights_scratch_bytes", _sz, (_u32, _u32)),
    ("vrenb200_assign_lights", _i32,
     (_vp, _u32, _u32, C.POINTER(Camera), _vp, _vp, _u32, _vp, _u32, _u32, _vp, _vp, _vp, _u32, _vp, _vp, _vp, _vp, _sz)),
]


def check(status: int, what: str) -> None:
    if status != OK:
        raise VrenError(status, what)


# ---- torch-tensor convenience layer (tests / bench); torch imported lazily -----------------------------------
def _ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def _stream() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream


def _scratch(nbytes: int):
    import torch

    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device="cuda")


_DT = {"u32": U32, "vec4": VEC4, "f32": F32}
_OP = {"add": ADD, "min": MIN, "max": MAX}


def reduce(inp, n: int, dtype: str, op: str, mode: str = "tree", blocks: int = 1, out=None):
    """vren::reduce<T,op>. inp: device tensor (int32 storage for u32, float32 for f32/vec4 [n,4])."""
    import torch

    lib = load()
    P = lib.vrenb200_round_to_next_power_of_2(n)
    comps = 4 if dtype == "vec4" else 1
    if out is None:
        out = torch.zeros(blocks * P * comps, dtype=inp.dtype, device=inp.device)
    m = REDUCE_TREE if mode == "tree" else REDUCE_FINAL
    sb = lib.vrenb200_reduce_scratch_bytes(_DT[dtype], m, n, blocks)
    scratch = _scratch(sb)
    check(lib.vrenb200_reduce(_stream(), _DT[dtype], _OP[op], m, _ptr(inp), n, _ptr(out), blocks, _ptr(scratch), sb),
          "vrenb200_reduce")
    return out


def exclusive_scan(inp, out=None, n: int | None = None):
    lib = load()
    n = inp.numel() if n is None else n
    out = inp if out is None else out
    sb = lib.vrenb200_scan_scratch_bytes(n)
    scratch = _scratch(sb)
    check(lib.vrenb200_exclusive_scan_u32(_stream(), _ptr(inp), _ptr(out), n, _ptr(scratch), sb), "vrenb200_exclusive_scan_u32")
    return out


def downsweep(buf, n: int, blocks: int = 1, clear_last: bool = True):
    lib = load()
    check(lib.vrenb200_blelloch_downsweep_u32(_stream(), _ptr(buf), n, blocks, int(clear_last)), "vrenb200_blelloch_downsweep_u32")
    return buf


def radix_sort_keys(keys, n: int | None = None):
    lib = load()
    n = keys.numel() if n is None else n
    sb = lib.vrenb200_radix_sort_scratch_bytes(n, 0)
    scratch = _scratch(sb)
    check(lib.vrenb200_radix_sort_keys(_stream(), _ptr(keys), n, _ptr(scratch), sb), "vrenb200_radix_sort_keys")
    return keys


def radix_sort_pairs(keys, vals, n: int | None = None, scratch=None, config: SortConfig | None = None):
    lib = load()
    n = keys.numel() if n is None else n
    sb = lib.vrenb200_radix_sort_scratch_bytes(n, 1)
    if scratch is None:
        scratch = _scratch(sb)
    if config is None:
        check(lib.vrenb200_radix_sort_pairs(_stream(), _ptr(keys), _ptr(vals), n, _ptr(scratch), sb), "vrenb200_radix_sort_pairs")
    else:
        check(lib.vrenb200_radix_sort_ex(_stream(), _ptr(keys), _ptr(vals), n, _ptr(scratch), sb, C.addressof(config), None), "vrenb200_radix_sort_ex")
    return keys, vals


def radix_sort_ex(keys, vals, config: SortConfig, n: int | None = None):
    """general form (vals may be None); returns the violation flag of the ranking check (0 = the check never failed)"""
    import torch

    lib = load()
    n = keys.numel() if n is None else n
    kv = 0 if vals is None else 1
    sb = lib.vrenb200_radix_sort_scratch_bytes(n, kv)
    scratch = _scratch(sb)
    check(lib.vrenb200_radix_sort_ex(_stream(), _ptr(keys), _ptr(vals), n, _ptr(scratch), sb, C.addressof(config), None), "vrenb200_radix_sort_ex")
    torch.cuda.synchronize()
    word = lib.vrenb200_radix_sort_violation_word(_ptr(scratch), n, kv)
    off = word - _ptr(scratch)
    return int(scratch[off: off + 4].view(torch.int32).item()) & 0xFFFFFFFF


def bucket_sort(pairs, n: int | None = None, config: SortConfig | None = None, end_offsets: int = -1):
    """vren::bucket_sort over uvec2 pairs [n,2] (int32 storage). Returns (out_buffer_u8, sorted_view [n,2], counters_view [65536])"""
    import torch

    lib = load()
    n = pairs.shape[0] if n is None else n
    ob = lib.vrenb200_bucket_sort_output_bytes(n)
    out = torch.zeros(ob, dtype=torch.uint8, device=pairs.device)
    sb = lib.vrenb200_bucket_sort_scratch_bytes(n)
    scratch = _scratch(sb)
    if config is None and end_offsets < 0:
        check(lib.vrenb200_bucket_sort(_stream(), _ptr(pairs), n, _ptr(out), _ptr(scratch), sb), "vrenb200_bucket_sort")
    else:
        check(lib.vrenb200_bucket_sort_ex(_stream(), _ptr(pairs), n, _ptr(out), _ptr(scratch), sb, C.addressof(config) if config is not None else None,
                                          end_offsets), "vrenb200_bucket_sort_ex")
    sorted_view = out[: n * 8].view(torch.int32).view(n, 2)
    coff = (n * 8 + 255) // 256 * 256
    counters = out[coff: coff + 65536 * 4].view(torch.int32)
    return out, sorted_view, counters


def build_bvh(nodes_u8, padded_leaf_count: int):
    """nodes_u8: uint8 device tensor holding calc_bvh_buffer_length(padded) 32-byte nodes, leaves pre-filled"""
    lib = load()
    check(lib.vrenb200_build_bvh(_stream(), _ptr(nodes_u8), padded_leaf_count), "vrenb200_build_bvh")
    return nodes_u8


def construct_point_light_bvh(positions, lights, view16, external_scratch: bool = True):
    """positions/lights: float32 [L,4] device tensors; view16: 16 python floats (column-major).
    Returns (view_pos [L,4], bvh_u8, index_u8)"""
    import torch

    lib = load()
    L = positions.shape[0]
    view_pos = torch.zeros(L, 4, dtype=torch.float32, device=positions.device)
    bvh = torch.zeros(lib.vrenb200_light_bvh_buffer_bytes(L), dtype=torch.uint8, device=positions.device)
    idx = torch.zeros(lib.vrenb200_light_index_buffer_bytes(L), dtype=torch.uint8, device=positions.device)
    view = (C.c_float * 16)(*[float(v) for v in view16])
    if external_scratch:
        sb = lib.vrenb200_light_bvh_scratch_bytes(L)
        scratch = _scratch(sb)
        sp = _ptr(scratch)
    else:
        sb, sp = 0, 0
    check(lib.vrenb200_construct_point_light_bvh(_stream(), _ptr(positions), _ptr(lights), L, C.cast(view, C.c_void_p),
                                                 _ptr(view_pos), _ptr(bvh), _ptr(idx), sp, sb),
          "vrenb200_construct_point_light_bvh")
    return view_pos, bvh, idx


DEFAULT_MAX_UNIQUE_CLUSTER_KEYS = 1 << 17     # VREN_MAX_UNIQUE_CLUSTER_KEY_COUNT, config.hpp:23
DEFAULT_MAX_ASSIGNED_LIGHTS = 1 << 23         # VREN_MAX_ASSIGNED_LIGHT_COUNT, config.hpp:24


def find_unique_clusters(depth, normals, camera: Camera, max_keys: int = DEFAULT_MAX_UNIQUE_CLUSTER_KEYS):
    """depth: float32 [H,W] device tensor; normals: float16 [H,W,4] or None.
    Returns (keys int32[max_keys], dispatch_params int32[4], cluster_ref int32 [H,W])"""
    import torch

    lib = load()
    H, W = depth.shape
    keys = torch.zeros(max_keys, dtype=torch.int32, device=depth.device)
    disp = torch.zeros(4, dtype=torch.int32, device=depth.device)
    ref = torch.zeros(H, W, dtype=torch.int32, device=depth.device)
    sb = lib.vrenb200_find_unique_clusters_scratch_bytes(W, H)
    scratch = _scratch(sb)
    check(lib.vrenb200_find_unique_clusters(_stream(), _ptr(depth), _ptr(normals), W, H, C.byref(camera), _ptr(keys), max_keys,
                                            _ptr(disp), _ptr(ref), _ptr(scratch), sb), "vrenb200_find_unique_clusters")
    return keys, disp, ref


def assign_lights(width: int, height: int, camera: Camera, keys, disp, bvh, light_count: int, index_buffer, view_pos,
                  max_keys: int = DEFAULT_MAX_UNIQUE_CLUSTER_KEYS, max_assigned: int = DEFAULT_MAX_ASSIGNED_LIGHTS):
    """Returns (counts[max_keys], offsets[max_keys], indices[max_assigned], status[4])"""
    import torch

    lib = load()
    dev = keys.device
    counts = torch.zeros(max_keys, dtype=torch.int32, device=dev)
    offsets = torch.zeros(max_keys, dtype=torch.int32, device=dev)
    indices = torch.zeros(max_assigned, dtype=torch.int32, device=dev)
    status = torch.zeros(4, dtype=torch.int32, device=dev)
    sb = lib.vrenb200_assign_lights_scratch_bytes(max_keys, max_assigned)
    scratch = _scratch(sb)
    root = lib.vrenb200_calc_bvh_root_index(light_count)
    check(lib.vrenb200_assign_lights(_stream(), width, height, C.byref(camera), _ptr(keys), _ptr(disp), max_keys, _ptr(bvh), root,
                                     light_count, _ptr(index_buffer), _ptr(view_pos), _ptr(indices), max_assigned,
                                     _ptr(counts), _ptr(offsets), _ptr(status), _ptr(scratch), sb), "vrenb200_assign_lights")
    return counts, offsets, indices, status


def cluster_tests(width: int, height: int, camera: Camera, keys, node_boxes=None, spheres=None):
    """diagnostics of a8 (vrenb200_cluster_tests): keys int32[n] -> (corner_min [n,3], corner_max [n,3], flags uint8[n]);
    node_boxes float32 [n,6] / spheres float32 [n,4] device tensors or None"""
    import torch

    lib = load()
    n = keys.numel()
    cmin = torch.zeros(n, 3, dtype=torch.float32, device=keys.device)
    cmax = torch.zeros(n, 3, dtype=torch.float32, device=keys.device)
    flags = torch.zeros(n, dtype=torch.uint8, device=keys.device)
    check(lib.vrenb200_cluster_tests(_stream(), width, height, C.byref(camera), _ptr(keys), n, _ptr(cmin), _ptr(cmax),
                                     _ptr(node_boxes), _ptr(spheres), _ptr(flags)), "vrenb200_cluster_tests")
    return cmin, cmax, flags


def depth_pyramid(depth):
    """vren::depth_buffer_reductor::copy_and_reduce. depth: float32 [H,W] device tensor -> flat float32 pyramid (levels back to back)"""
    import torch

    lib = load()
    H, W = depth.shape
    out = torch.zeros(lib.vrenb200_depth_pyramid_bytes(W, H) // 4, dtype=torch.float32, device=depth.device)
    check(lib.vrenb200_depth_pyramid_build(_stream(), _ptr(depth), W, H, _ptr(out)), "vrenb200_depth_pyramid_build")
    return out


def bounce_point_lights(positions, directions, aabb_min, aabb_max, speed: float, dt: float, count: int | None = None):
    """vren_demo::point_light_bouncer::bounce. positions / directions: float32 [L,4] device tensors, updated in place"""
    lib = load()
    count = positions.shape[0] if count is None else count
    lo, hi = (C.c_float * 3)(*[float(v) for v in aabb_min]), (C.c_float * 3)(*[float(v) for v in aabb_max])
    check(lib.vrenb200_bounce_point_lights(_stream(), _ptr(positions), _ptr(directions), count, C.byref(lo), C.byref(hi),
                                           C.c_float(speed), C.c_float(dt)), "vrenb200_bounce_point_lights")
    return positions, directions


def visualize_bvh(nodes_u8, level_count: int):
    """vren_demo::visualize_bvh::write -> float32 [vertices, 4] device tensor {x, y, z, color bits}"""
    import torch

    lib = load()
    count = lib.vrenb200_visualize_bvh_vertex_count(level_count)
    out = torch.empty(count, 4, dtype=torch.float32, device=nodes_u8.device)
    check(lib.vrenb200_visualize_bvh(_stream(), _ptr(nodes_u8), level_count, _ptr(out)), "vrenb200_visualize_bvh")
    return out


def light_list_hash(cluster_ref, disp, counts, offsets, indices):
    """per-pixel {light count, XOR of light indices} through the cluster reference image -> int32 [H,W,2]"""
    import torch

    lib = load()
    H, W = cluster_ref.shape
    max_keys, max_assigned = counts.numel(), indices.numel()
    out = torch.zeros(H, W, 2, dtype=torch.int32, device=cluster_ref.device)
    sb = lib.vrenb200_light_list_hash_scratch_bytes(max_keys)
    scratch = _scratch(sb)
    check(lib.vrenb200_light_list_hash(_stream(), W, H, _ptr(cluster_ref), _ptr(disp), max_keys, _ptr(counts), _ptr(offsets), _ptr(indices),
                                       max_assigned, _ptr(out), _ptr(scratch), sb), "vrenb200_light_list_hash")
    return out
