"""Golden vectors produced by the REAL reference's own code (tests/golden/reference_outputs.npz, generated in the build
container by tests/golden/make_golden.py from oracle/_ref): base.hpp integer helpers, build_bvh.cpp calc_bvh_* and the
reference test's run_cpu_reduce (vren_test/.../reduce.cpp:72-87).  Unlike tests/test_ref_extract.py these run where
/root/reference does not exist (the GPU box): the oracle, the product's host helpers and the CUDA reduce are all checked
against outputs of the reference itself."""
import sys
from pathlib import Path

import numpy as np
import pytest

import oracle

GOLDEN = Path(__file__).parent / "golden" / "reference_outputs.npz"
sys.path.insert(0, str(GOLDEN.parent))
from make_golden import HELPERS_U32, OPS, REDUCE_SIZES, reduce_inputs  # noqa: E402


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def test_helpers_oracle_and_product_equal_reference_outputs(golden, built):
    from vren_b200 import lib as vlib

    orc, prod = oracle.load(), vlib.load()
    values = [int(v) for v in golden["helper_values"]]
    for name in HELPERS_U32 + ["calc_bvh_buffer_size"]:
        want = [int(w) for w in golden["helper_" + name]]
        assert [getattr(orc, "oracle_" + name)(v) for v in values] == want, name
        assert [getattr(prod, "vrenb200_" + name)(v) for v in values] == want, name
    for key, args in (("divide_and_ceil_1024", ("divide_and_ceil", 1024)), ("round_to_next_power_of_32", ("round_to_next_power_of", 32)),
                      ("round_to_next_multiple_of_256", ("round_to_next_multiple_of", 256))):
        want = [int(w) for w in golden["helper_" + key]]
        assert [getattr(orc, "oracle_" + args[0])(v, args[1]) for v in values] == want, key
        assert [getattr(prod, "vrenb200_" + args[0])(v, args[1]) for v in values] == want, key
    want = [bool(w) for w in golden["helper_is_power_of_32"]]
    assert [bool(orc.oracle_is_power_of(v, 32)) for v in values] == want
    assert [bool(prod.vrenb200_is_power_of(v, 32)) for v in values] == want


@pytest.mark.parametrize("op", OPS)
@pytest.mark.parametrize("n", REDUCE_SIZES)
def test_oracle_reduce_equals_reference_run_cpu_reduce(golden, op, n):
    x, f = reduce_inputs(op, n)
    assert np.array_equal(oracle.reduce(x, n, "u32", op), golden[f"reduce_u32_{op}_{n}"])
    assert np.array_equal(oracle.reduce(f, n, "f32", op).view(np.uint32), golden[f"reduce_f32_{op}_{n}"])   # same tree, same bits


@pytest.mark.gpu
@pytest.mark.parametrize("op", OPS)
@pytest.mark.parametrize("n", REDUCE_SIZES)
def test_cuda_reduce_tree_equals_reference_run_cpu_reduce(golden, vren, op, n):
    """the CUDA path against the reference's own CPU reduce, full padded tree, fp32 add included (bit-exact)"""
    import torch

    x, f = reduce_inputs(op, n)
    got = vren.reduce(torch.from_numpy(x.view(np.int32)).cuda(), n, "u32", op, mode="tree").cpu().numpy().view(np.uint32)
    assert np.array_equal(got, golden[f"reduce_u32_{op}_{n}"])
    got = vren.reduce(torch.from_numpy(f).cuda(), n, "f32", op, mode="tree").cpu().numpy().view(np.uint32)
    assert np.array_equal(got, golden[f"reduce_f32_{op}_{n}"])
