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


_LATE: list = []


def load_ref():
    """oracle/_ref/libvrenref.so — functions compiled from the reference sources where they lie (None if absent)."""
    global _ref
    if _ref is not None:
        return _ref
    if not REF_LIB_PATH.exists():
        return None
    _ref = C.CDLL(str(REF_LIB_PATH))
    return _ref


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
