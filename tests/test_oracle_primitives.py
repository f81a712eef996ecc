"""CPU suite: pins the oracle (oracle/oracle_primitives.cpp) against the reference's own test checks and KATs.

Reference checks restated here:
  - TEST(build_bvh, utils) exact KATs          vren_test/vren_test/primitives/build_bvh.cpp:225-253
  - reduce full-tree equality vs run_cpu_reduce vren_test/vren_test/primitives/reduce.cpp:245-291
  - scan == std::exclusive_scan, all-ones 2^0..2^19   .../blelloch_scan.cpp:122-176
  - radix == std::sort, reversed iota 2^10      .../radix_sort.cpp:82-88,122-132
  - bucket keys == std::sort on masked key, values a permutation  .../bucket_sort.cpp:129-179
  - BVH traversal hit set == linear scan        .../build_bvh.cpp:175-221
"""
import numpy as np
import pytest

import oracle
from conftest import splitmix64


@pytest.fixture(scope="module")
def orc(built):
    return oracle.load()


def test_bvh_sizing_kats(orc):
    # vren_test/vren_test/primitives/build_bvh.cpp:225-253, verbatim expectations
    pad = orc.oracle_calc_bvh_padded_leaf_count
    assert [pad(x) for x in (0, 1, 17, 32, 129, 582, 1024, 2193819)] == [32, 32, 32, 32, 1024, 1024, 1024, 33554432]
    ln = orc.oracle_calc_bvh_buffer_length
    assert [ln(x) for x in (0, 1, 17, 32)] == [33] * 4
    assert [ln(x) for x in (129, 582, 1024)] == [32 * 32 + 32 + 1] * 3
    assert ln(2193819) == 32**5 + 32**4 + 32**3 + 32**2 + 32 + 1
    root = orc.oracle_calc_bvh_root_index
    assert [root(x) for x in (0, 1)] == [32, 32]
    assert [root(x) for x in (129, 582)] == [32 * 32 + 32] * 2
    assert root(2193819) == 32**5 + 32**4 + 32**3 + 32**2 + 32
    lv = orc.oracle_calc_bvh_level_count
    assert [lv(x) for x in (0, 1, 129, 582, 2193819)] == [1, 1, 2, 2, 5]


def test_integer_helpers(orc):
    assert [orc.oracle_round_to_next_power_of_2(x) for x in (1, 2, 3, 1000, 1024, 1025, 10000)] == [1, 2, 4, 1024, 1024, 2048, 16384]
    assert orc.oracle_round_to_next_multiple_of(10, 256) == 256 and orc.oracle_round_to_next_multiple_of(512, 256) == 512
    assert orc.oracle_divide_and_ceil(1025, 1024) == 2 and orc.oracle_divide_and_ceil(1024, 1024) == 1
    assert orc.oracle_is_power_of(32768, 32) and not orc.oracle_is_power_of(32769, 32)


@pytest.mark.parametrize("op", ["add", "min", "max"])
@pytest.mark.parametrize("n", [1, 10, 100, 1000, 10000, 100000])
def test_reduce_tree_matches_reference_test_oracle_u32(orc, op, n):
    # inputs as fill_reduce_cpu_buffer: add -> ones, min/max -> rand()%100 (reduce.cpp:156-172)
    x = np.ones(n, np.uint32) if op == "add" else (splitmix64(3, n) % np.uint64(100)).astype(np.uint32)
    tree = oracle.reduce(x, n, "u32", op)
    P = oracle.next_pow2(n)
    ident = {"add": 0, "min": 0xFFFFFFFF, "max": 0}[op]
    padded = np.full(P, ident, np.uint32)
    padded[:n] = x
    orc.oracle_test_cpu_reduce_u32(oracle.OP[op], padded, P)
    assert np.array_equal(tree, padded)
    expect = {"add": x.sum(dtype=np.uint64) & 0xFFFFFFFF, "min": x.min(), "max": x.max()}[op]
    assert int(tree[P - 1]) == int(expect)


@pytest.mark.parametrize("op", ["add", "min", "max"])
def test_reduce_tree_matches_reference_test_oracle_vec4(orc, op):
    n = 10000  # TEST(reduce, type_vec4)
    x = (splitmix64(5, 4 * n) % np.uint64(100)).astype(np.float32).reshape(n, 4)
    tree = oracle.reduce(x.reshape(-1), n, "vec4", op).reshape(-1, 4)
    P = oracle.next_pow2(n)
    ident = {"add": 0.0, "min": 1e35, "max": -1e35}[op]
    padded = np.full((P, 4), ident, np.float32)
    padded[:n] = x
    flat = padded.reshape(-1).copy()
    orc.oracle_test_cpu_reduce_f32(oracle.OP[op], flat, P, 4)
    assert np.array_equal(tree.reshape(-1).view(np.uint32), flat.view(np.uint32))


def test_reduce_blocks_rows(orc):
    n, blocks = 300, 5
    x = (splitmix64(7, n * blocks) % np.uint64(1000)).astype(np.uint32)
    tree = oracle.reduce(x, n, "u32", "add", blocks=blocks)
    P = oracle.next_pow2(n)
    for y in range(blocks):
        assert int(tree[y * P + P - 1]) == int(x[y * n:(y + 1) * n].sum())


@pytest.mark.parametrize("log2n", range(0, 20))
def test_blelloch_scan_equals_exclusive_scan_all_ones(orc, log2n):
    n = 1 << log2n
    x = np.ones(n, np.uint32)
    assert np.array_equal(oracle.blelloch_scan(x, n), oracle.exclusive_scan(x))
    assert np.array_equal(oracle.exclusive_scan(x), np.arange(n, dtype=np.uint32))


@pytest.mark.parametrize("n", [1, 2, 512, 1024, 2048, 1 << 14, 1 << 21])
def test_blelloch_scan_random_wraps(orc, n):
    x = (splitmix64(11, n) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    got = oracle.blelloch_scan(x, n)
    want = np.concatenate([[0], np.cumsum(x.astype(np.uint64))[:-1]]).astype(np.uint64) & np.uint64(0xFFFFFFFF)
    assert np.array_equal(got, want.astype(np.uint32))
    assert np.array_equal(oracle.exclusive_scan(x), want.astype(np.uint32))


def test_downsweep_without_clear_adds_root(orc):
    n = 4096
    x = (splitmix64(13, n) % np.uint64(50)).astype(np.uint32)
    tree = oracle.reduce(x, n, "u32", "add")
    got = oracle.downsweep(tree, n, 1, clear_last=False)
    assert np.array_equal(got, oracle.exclusive_scan(x) + tree[n - 1])
    # n < 1024: the zero-filled workgroup tile swallows the root (blelloch_scan_downsweep.comp:68-97)
    n = 256
    tree = oracle.reduce(x[:n], n, "u32", "add")
    assert np.array_equal(oracle.downsweep(tree, n, 1, clear_last=False), oracle.exclusive_scan(x[:n]))


def test_radix_lsd4_restatement_equals_std_sort(orc):
    n = 1 << 10  # TEST(radix_sort, main): reversed iota
    x = np.arange(n, dtype=np.uint32)[::-1].copy()
    assert np.array_equal(oracle.radix_sort_lsd4(x), oracle.sort_keys(x))
    assert np.array_equal(oracle.sort_keys(x), np.arange(n, dtype=np.uint32))
    for n, seed in ((1 << 12, 1), (1 << 16, 2), (1 << 20, 3)):
        x = (splitmix64(seed, n) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        want = np.sort(x)
        assert np.array_equal(oracle.radix_sort_lsd4(x), want)
        assert np.array_equal(oracle.sort_keys(x), want)


def test_sort_pairs_is_stable(orc):
    n = 50000
    k = (splitmix64(17, n) % np.uint64(97)).astype(np.uint32)
    v = np.arange(n, dtype=np.uint32)
    sk, sv = oracle.sort_pairs(k, v)
    order = np.argsort(k, kind="stable")
    assert np.array_equal(sk, k[order]) and np.array_equal(sv, v[order])
    # the timed multi-threaded baseline must agree with it
    pairs = (k.astype(np.uint64) | (v.astype(np.uint64) << np.uint64(32))).copy()
    orc.oracle_sort_pairs_interleaved_mt(pairs, n, 4)
    assert np.array_equal((pairs & np.uint64(0xFFFFFFFF)).astype(np.uint32), sk)
    assert np.array_equal((pairs >> np.uint64(32)).astype(np.uint32), sv)
    big = (splitmix64(19, 1 << 18) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    mt = big.copy()
    orc.oracle_sort_keys_mt(mt, mt.size, 3)
    assert np.array_equal(mt, np.sort(big))


@pytest.mark.parametrize("n", [1, 5, 4 * 5 * 6, 4 * 5 * 6 * 4 * 5, 100003])
def test_bucket_sort_contract(orc, n):
    # key = rand()%65536, value = i (vren_test/.../bucket_sort.cpp:97-101)
    keys = (splitmix64(23, n) & np.uint64(0xFFFFFFFF)).astype(np.uint32)  # full 32-bit x: only low 16 bits sort
    pairs = np.stack([keys, np.arange(n, dtype=np.uint32)], axis=1)
    out, counters = oracle.bucket_sort(pairs)
    masked = out[:, 0] & 0xFFFF
    assert np.all(masked[:-1] <= masked[1:])
    assert np.array_equal(np.sort(out[:, 1]), np.arange(n, dtype=np.uint32))       # values form a permutation
    order = np.argsort(keys & 0xFFFF, kind="stable")                               # canonical tie-break
    assert np.array_equal(out, pairs[order])
    hist = np.bincount(keys & 0xFFFF, minlength=65536)
    assert np.array_equal(counters, np.cumsum(hist).astype(np.uint32))             # bucket END offsets


def make_leaves(leaf_count, seed):
    padded = int(oracle.load().oracle_calc_bvh_padded_leaf_count(leaf_count))
    length = int(oracle.load().oracle_calc_bvh_buffer_length(leaf_count))
    nodes = np.zeros(length, dtype=oracle.BVH_NODE)
    r = (splitmix64(seed, 6 * max(leaf_count, 1)) >> np.uint64(40)).astype(np.float64) / float(1 << 24) * 100.0
    r = r.astype(np.float32).reshape(-1, 6)
    nodes["next"][:padded] = 0xFFFFFFFE
    if leaf_count:
        pos, ext = r[:leaf_count, :3], r[:leaf_count, 3:]
        nodes["min"][:leaf_count] = pos - ext
        nodes["max"][:leaf_count] = pos + ext
        nodes["next"][:leaf_count] = 0xFFFFFFFF
    return nodes, padded, length


@pytest.mark.parametrize("leaf_count,queries", [(0, 16), (1, 16), (10, 16), (100, 8), (1000, 4), (10000, 2), (40000, 2)])
def test_bvh_traversal_equals_linear_scan(orc, leaf_count, queries):
    nodes, padded, length = make_leaves(leaf_count, 29)
    built = oracle.build_bvh(nodes, padded)
    root = length - 1
    pts = ((splitmix64(31, 3 * queries) >> np.uint64(40)).astype(np.float64) / float(1 << 24) * 100.0).astype(np.float32)
    a = np.zeros(padded, np.uint32)
    b = np.zeros(padded, np.uint32)
    for q in range(queries):
        p = np.ascontiguousarray(pts[3 * q:3 * q + 3])
        na = orc.oracle_bvh_traverse_point(built.ctypes.data, root, p, a, padded) if leaf_count else 0
        nb = orc.oracle_bvh_linear_point(built.ctypes.data, padded, p, b, padded)
        assert na == nb
        assert np.array_equal(a[:na], b[:nb])
    if leaf_count:
        assert built["next"][root] == root - 32 if padded > 32 else built["next"][root] == 0
        assert np.array_equal(built["min"][root], nodes["min"][:leaf_count].min(axis=0))
        assert np.array_equal(built["max"][root], nodes["max"][:leaf_count].max(axis=0))
    else:
        assert built["next"][root] == 0xFFFFFFFE
