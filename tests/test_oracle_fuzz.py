"""Property-based checks of the CHECKER: the oracle's primitives (oracle/oracle_primitives.cpp, plain C++ restatements of the
reference's shaders and CPU checks) against independent numpy formulations on generated inputs — sizes with ragged tails,
duplicate-heavy keys, wrap-around sums.  The GPU parity tests are only as good as the oracle they compare with."""
import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

import oracle

pytestmark = pytest.mark.usefixtures("built")
SETTINGS = dict(max_examples=60, deadline=None)


def keys_strategy(max_n=3000):
    # (n, seed, spread): spread selects how many distinct values the keys take (1 ... full range)
    return st.tuples(st.integers(0, max_n), st.integers(0, 2**31 - 1), st.sampled_from([1, 2, 7, 100, 1 << 16, 1 << 32]))


def make_keys(n, seed, spread):
    rng = np.random.Generator(np.random.PCG64(seed))
    k = rng.integers(0, spread, size=n, dtype=np.uint64)
    if spread <= 100:
        k = k * np.uint64(0x01010101)               # the same value in every digit: collisions in every radix pass
    return (k & np.uint64(0xFFFFFFFF)).astype(np.uint32)


@given(keys_strategy())
@settings(**SETTINGS)
def test_sort_keys_and_pairs(args):
    k = make_keys(*args)
    assert np.array_equal(oracle.sort_keys(k), np.sort(k))
    v = np.arange(k.size, dtype=np.uint32)[::-1].copy()
    sk, sv = oracle.sort_pairs(k, v)
    order = np.argsort(k, kind="stable")
    assert np.array_equal(sk, k[order]) and np.array_equal(sv, v[order])


@given(st.integers(10, 13), st.integers(0, 2**31 - 1), st.sampled_from([1, 7, 1 << 32]))
@settings(max_examples=12, deadline=None)
def test_reference_radix_sort_restatement(log2n, seed, spread):
    """the literal 4-bit x 8-pass algorithm of the reference (n a power of two >= 1024, radix_sort.cpp:158-161) sorts"""
    k = make_keys(1 << log2n, seed, spread)
    assert np.array_equal(oracle.radix_sort_lsd4(k), np.sort(k))


@given(keys_strategy(5000))
@settings(**SETTINGS)
def test_exclusive_scan_wraps_mod_2p32(args):
    x = make_keys(*args)
    want = np.concatenate([np.zeros(1, np.uint64), np.cumsum(x.astype(np.uint64))[:-1]]) & np.uint64(0xFFFFFFFF) if x.size else np.zeros(0, np.uint64)
    assert np.array_equal(oracle.exclusive_scan(x), want.astype(np.uint32))


@given(st.integers(0, 12), st.integers(1, 3), st.integers(0, 2**31 - 1))
@settings(max_examples=30, deadline=None)
def test_blelloch_scan_rows(log2n, blocks, seed):
    """vren::blelloch_scan on a power-of-two length (blelloch_scan.cpp:57-166): row 0 becomes its exclusive scan.  The reference
    reduces with blocks_num hard-coded to 1 (blelloch_scan.cpp:151), so further rows are only down-swept — the restatement keeps
    that (it is what the reference computes), and says so here"""
    n = 1 << log2n
    x = make_keys(n * blocks, seed, 1 << 32)
    got = oracle.blelloch_scan(x, n, blocks).reshape(blocks, n)
    rows = x.reshape(blocks, n)
    want = np.concatenate([np.zeros(1, np.uint64), np.cumsum(rows[0].astype(np.uint64))[:-1]]) & np.uint64(0xFFFFFFFF)
    assert np.array_equal(got[0], want.astype(np.uint32))
    for b in range(1, blocks):
        assert np.array_equal(got[b], oracle.downsweep(rows[b], n, 1, True))


@given(st.integers(1, 1500), st.sampled_from(["add", "min", "max"]), st.integers(0, 2**31 - 1))
@settings(**SETTINGS)
def test_reduce_tree_holds_every_aligned_block(n, op, seed):
    """after the call slot (j+1) 2^l - 1 (j even) holds the op over the aligned block of 2^l slots ending there — every
    slot holds the largest such block —, slots >= n reading as the identity (reduce.comp:50-87)"""
    x = make_keys(n, seed, 1 << 32)
    tree = oracle.reduce(x, n, "u32", op)
    P = oracle.next_pow2(n)
    ident = {"add": 0, "min": 0xFFFFFFFF, "max": 0}[op]
    padded = np.full(P, ident, np.uint64)
    padded[:n] = x
    f = {"add": lambda a: int(a.sum()) & 0xFFFFFFFF, "min": lambda a: int(a.min()), "max": lambda a: int(a.max())}[op]
    for i in range(P):                                   # slot i: the largest aligned block that ends at i
        size = (i + 1) & -(i + 1)
        assert int(tree[i]) == f(padded[i + 1 - size:i + 1]), i


@given(keys_strategy(4000))
@settings(**SETTINGS)
def test_bucket_sort_is_stable_and_counts_ends(args):
    """canonical (stable) outcome of vren::bucket_sort: ascending by x & 0xFFFF, input order inside a bucket, the whole 64-bit pair
    carried; counters[b] = number of pairs whose key is <= b (bucket_sort_write.comp:32)"""
    n, seed, spread = args
    x = make_keys(n, seed, spread)
    pairs = np.stack([x, np.arange(n, dtype=np.uint32)], axis=1)
    out, counters = oracle.bucket_sort(pairs)
    order = np.argsort(x & np.uint32(0xFFFF), kind="stable")
    assert np.array_equal(out, pairs[order])
    assert np.array_equal(counters, np.cumsum(np.bincount(x & np.uint32(0xFFFF), minlength=65536)).astype(np.uint32))
