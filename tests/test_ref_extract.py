"""Cross-checks the oracle (and the product's integer helpers) against the pieces of the REAL reference that compile
standalone (oracle/_ref, built by oracle/ref_extract.py from /root/reference where it lies).  Skipped where the
reference checkout is absent (the GPU box): nothing else depends on it."""
import ctypes as C

import numpy as np
import pytest

import oracle
from conftest import splitmix64


@pytest.fixture(scope="module")
def ref(built):
    lib = oracle.load_ref()
    if lib is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    for name in ("ref_round_to_next_power_of_2", "ref_round_to_next_power_of", "ref_divide_and_ceil", "ref_calc_bvh_padded_leaf_count",
                 "ref_calc_bvh_buffer_length", "ref_calc_bvh_root_index", "ref_calc_bvh_level_count"):
        getattr(lib, name).restype = C.c_uint32
    lib.ref_calc_bvh_buffer_size.restype = C.c_uint64
    lib.ref_round_to_next_multiple_of.restype = C.c_uint64
    lib.ref_round_to_next_multiple_of.argtypes = [C.c_uint64, C.c_uint64]
    return lib


VALUES = list(range(1, 200)) + [1000, 1023, 1024, 1025, 4095, 4096, 32767, 32768, 32769, 65536, 10**6, 1 << 20, (1 << 20) + 1, 2193819, 1 << 25]


def test_reference_helpers_equal_oracle_and_product(ref):
    from vren_b200 import lib as vlib

    orc, prod = oracle.load(), vlib.load()
    for v in VALUES:
        assert ref.ref_round_to_next_power_of_2(v) == orc.oracle_round_to_next_power_of_2(v) == prod.vrenb200_round_to_next_power_of_2(v)
        assert ref.ref_divide_and_ceil(v, 1024) == orc.oracle_divide_and_ceil(v, 1024) == prod.vrenb200_divide_and_ceil(v, 1024)
        assert bool(ref.ref_is_power_of(v, 32)) == bool(orc.oracle_is_power_of(v, 32)) == bool(prod.vrenb200_is_power_of(v, 32))
        assert ref.ref_round_to_next_power_of(v, 32) == orc.oracle_round_to_next_power_of(v, 32) == prod.vrenb200_round_to_next_power_of(v, 32)
        assert ref.ref_round_to_next_multiple_of(v, 256) == orc.oracle_round_to_next_multiple_of(v, 256) == prod.vrenb200_round_to_next_multiple_of(v, 256)
        for fn in ("calc_bvh_padded_leaf_count", "calc_bvh_buffer_length", "calc_bvh_buffer_size", "calc_bvh_root_index", "calc_bvh_level_count"):
            r = getattr(ref, "ref_" + fn)(v)
            assert r == getattr(orc, "oracle_" + fn)(v) == getattr(prod, "vrenb200_" + fn)(v), (fn, v)


@pytest.mark.parametrize("op", ["add", "min", "max"])
@pytest.mark.parametrize("n", [1, 10, 1000, 10000, 100000])
def test_reference_run_cpu_reduce_equals_oracle_tree(ref, op, n):
    """the reference test's own CPU reduce (vren_test/.../reduce.cpp:72-87) vs the oracle's restatement of reduce.comp"""
    x = np.ones(n, np.uint32) if op == "add" else (splitmix64(3, n) % np.uint64(100)).astype(np.uint32)
    P = oracle.next_pow2(n)
    padded = np.full(P, {"add": 0, "min": 0xFFFFFFFF, "max": 0}[op], np.uint32)
    padded[:n] = x
    ref.ref_run_cpu_reduce_u32(oracle.OP[op], padded.ctypes.data_as(C.c_void_p), P)
    assert np.array_equal(oracle.reduce(x, n, "u32", op), padded)
    f = (splitmix64(5, n) >> np.uint64(40)).astype(np.float32) / np.float32(1 << 24)
    fp = np.full(P, {"add": 0.0, "min": 1e35, "max": -1e35}[op], np.float32)
    fp[:n] = f
    ref.ref_run_cpu_reduce_f32(oracle.OP[op], fp.ctypes.data_as(C.c_void_p), P)
    assert np.array_equal(oracle.reduce(f, n, "f32", op).view(np.uint32), fp.view(np.uint32))   # fp32 add: same tree, same bits
