"""GPU parity: CUDA kernels (through the C ABI) vs the CPU oracle on the same seeded inputs.

Sizes and input patterns mirror the reference's own tests (vren_test/vren_test/primitives/*.cpp); the
BASELINE.json full sizes are covered through size-independent properties (sortedness + permutation checksum,
closed-form scans).  Integer results are compared bit-exactly; fp32 results bit-exactly as well (same tree).
"""
import numpy as np
import pytest

import oracle
from conftest import splitmix64

pytestmark = pytest.mark.gpu


def dev_u32(arr):
    import torch

    return torch.from_numpy(np.ascontiguousarray(arr, dtype=np.uint32).view(np.int32)).cuda()


def host_u32(t):
    return t.cpu().numpy().view(np.uint32)


def rand_u32(seed, n):
    return (splitmix64(seed, n) & np.uint64(0xFFFFFFFF)).astype(np.uint32)


# ---- a1 reduce --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("op", ["add", "min", "max"])
@pytest.mark.parametrize("n", [1, 10, 100, 1000, 4096, 10000, 100000, (1 << 20) + 17])
def test_reduce_u32_tree_and_final(vren, op, n):
    x = np.ones(n, np.uint32) if op == "add" and n <= 100000 else (rand_u32(3, n) % np.uint32(100)).astype(np.uint32)
    want = oracle.reduce(x, n, "u32", op)
    got = host_u32(vren.reduce(dev_u32(x), n, "u32", op, mode="tree"))
    assert np.array_equal(got, want)                 # the whole padded tree, like the reference test
    P = oracle.next_pow2(n)
    fin = host_u32(vren.reduce(dev_u32(x), n, "u32", op, mode="final"))
    assert fin[P - 1] == want[P - 1]


@pytest.mark.parametrize("op", ["add", "min", "max"])
@pytest.mark.parametrize("n", [1, 7, 1000, 1024, 10000, 70001])
def test_reduce_vec4_bit_exact(vren, op, n):
    import torch

    # TEST(reduce, type_vec4): rand()%100 as float; plus non-integers so the add order matters
    x = (rand_u32(5, 4 * n) % np.uint32(100)).astype(np.float32) + (rand_u32(6, 4 * n) >> np.uint32(8)).astype(np.float32) / np.float32(1 << 24)
    want = oracle.reduce(x, n, "vec4", op)
    got = vren.reduce(torch.from_numpy(x).cuda(), n, "vec4", op, mode="tree").cpu().numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    P = oracle.next_pow2(n)
    fin = vren.reduce(torch.from_numpy(x).cuda(), n, "vec4", op, mode="final").cpu().numpy()
    assert np.array_equal(fin.reshape(-1, 4)[P - 1].view(np.uint32), want.reshape(-1, 4)[P - 1].view(np.uint32))


@pytest.mark.parametrize("n", [5, 4096, 5000, 1 << 16, 300001])
def test_reduce_f32_add_tree_order(vren, n):
    import torch

    x = (rand_u32(9, n) >> np.uint32(8)).astype(np.float32) / np.float32(1 << 24)   # U[0,1)
    want = oracle.reduce(x, n, "f32", "add")
    got = vren.reduce(torch.from_numpy(x).cuda(), n, "f32", "add", mode="tree").cpu().numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    P = oracle.next_pow2(n)
    fin = vren.reduce(torch.from_numpy(x).cuda(), n, "f32", "add", mode="final").cpu().numpy()
    assert fin.view(np.uint32)[P - 1] == want.view(np.uint32)[P - 1]
    # north-star tolerance for fp32 reduce: 1e-6 relative vs an fp64 sum
    assert abs(float(fin[P - 1]) - float(x.astype(np.float64).sum())) <= 1e-6 * float(x.astype(np.float64).sum()) + 1e-6


def test_reduce_rows_and_in_place(vren):
    n, blocks = 2048, 16   # the shape radix_sort.cpp:230 uses (blocks = 16 digit rows)
    x = (rand_u32(10, n * blocks) % np.uint32(1000)).astype(np.uint32)
    want = oracle.reduce(x, n, "u32", "add", blocks=blocks)
    d = dev_u32(x)
    got = host_u32(vren.reduce(d, n, "u32", "add", mode="tree", blocks=blocks, out=d))   # in place (reduce.cpp:117-128)
    assert np.array_equal(got, want)
    n = 300                   # non-pow2 rows: in stride n, out stride P
    x = (rand_u32(11, n * 3) % np.uint32(1000)).astype(np.uint32)
    want = oracle.reduce(x, n, "u32", "max", blocks=3)
    assert np.array_equal(host_u32(vren.reduce(dev_u32(x), n, "u32", "max", mode="tree", blocks=3)), want)
    fin = host_u32(vren.reduce(dev_u32(x), n, "u32", "max", mode="final", blocks=3))
    assert np.array_equal(fin[511::512], want[511::512])


def test_reduce_full_size_c2(vren):
    import torch

    n = 1 << 28   # BASELINE C2
    x = torch.ones(n, dtype=torch.int32, device="cuda")
    fin = vren.reduce(x, n, "u32", "add", mode="final")
    assert int(fin[n - 1].item()) == n
    x = torch.arange(n, dtype=torch.int32, device="cuda")
    assert int(vren.reduce(x, n, "u32", "max", mode="final")[n - 1].item()) == n - 1
    tree = vren.reduce(x, n, "u32", "add", mode="tree")
    # closed form for slots: sum over the aligned block ending at i
    for i in (0, 1, 3, 4095, 4096 * 7 - 1, (1 << 22) - 1, n - 1, n // 2 - 1):
        size = ((i + 1) & -(i + 1))
        lo = i - size + 1
        want = (size * (lo + i) // 2) & 0xFFFFFFFF
        assert (int(tree[i].item()) & 0xFFFFFFFF) == want
    f = torch.full((n,), 0.25, dtype=torch.float32, device="cuda")
    assert float(vren.reduce(f, n, "f32", "add", mode="final")[n - 1].item()) == n * 0.25


# ---- a2 scan ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("log2n", range(0, 20))
def test_scan_all_ones_pow2(vren, log2n):
    n = 1 << log2n   # TEST(blelloch_scan, main)
    x = np.ones(n, np.uint32)
    got = host_u32(vren.exclusive_scan(dev_u32(x)))
    assert np.array_equal(got, oracle.exclusive_scan(x))


@pytest.mark.parametrize("n", [3, 1000, 4096, 4097, 123457, (1 << 21) + 5, (1 << 22) + 5, (1 << 24) + 3 * 16384 + 1])
def test_scan_random_wraparound_any_length(vren, n):
    x = rand_u32(13, n)
    want = oracle.exclusive_scan(x)
    assert np.array_equal(host_u32(vren.exclusive_scan(dev_u32(x))), want)          # in place
    import torch

    src = dev_u32(x)
    dst = torch.zeros_like(src)
    vren.exclusive_scan(src, out=dst)                                                # out of place
    assert np.array_equal(host_u32(dst), want) and np.array_equal(host_u32(src), x)


@pytest.mark.parametrize("n,offset", [(16384 * 700 + 77, 0), (16384 * 700 + 77, 3), ((1 << 22) + 16384, 1)])
def test_scan_kernels_agree(vren, n, offset):
    """every scan kernel behind the tuning hook (register tile, staged tile, run-ahead with and without L2 hints) gives
    the oracle's result, also on views that are not 16-byte aligned (no bulk copy / vector path)"""
    import torch

    lib = vren.load()
    x = rand_u32(29, n + 8)
    want = oracle.exclusive_scan(x[offset:offset + n])
    tuning = hasattr(lib, "vrenb200_scan_set_variant")     # VRENB200_TUNING builds expose the experimental kernels
    try:
        for variant in ((0, 1, 2, 3, 4, 8, 11, 12, 14, 15) if tuning else (0,)):
            assert not tuning or lib.vrenb200_scan_set_variant(variant) == 0
            src = dev_u32(x)
            dst = torch.zeros_like(src)
            vren.exclusive_scan(src[offset:offset + n], out=dst[offset:offset + n])
            got = host_u32(dst)
            assert np.array_equal(got[offset:offset + n], want), variant
            assert not got[:offset].any() and not got[offset + n:].any(), variant
            for _ in range(4):                                               # in place (out == in), a few times: the
                src = dev_u32(x)                                             # run-ahead kernel must order its roles
                vren.exclusive_scan(src[offset:offset + n])
                assert np.array_equal(host_u32(src)[offset:offset + n], want), variant
    finally:
        if tuning:
            lib.vrenb200_scan_set_variant(0)


def test_scan_full_size_c2(vren):
    import torch

    n = 1 << 28
    x = torch.ones(n, dtype=torch.int32, device="cuda")
    vren.exclusive_scan(x)
    assert torch.equal(x, torch.arange(n, dtype=torch.int32, device="cuda"))
    y = torch.full((n,), 0x01000193, dtype=torch.int32, device="cuda")   # wraps many times
    vren.exclusive_scan(y)
    idx = torch.tensor([0, 1, 4095, 4096, n // 3, n - 1], device="cuda")
    want = [(i * 0x01000193) & 0xFFFFFFFF for i in idx.tolist()]
    assert [v & 0xFFFFFFFF for v in y[idx].tolist()] == want


@pytest.mark.parametrize("n,blocks,clear", [(1, 1, True), (256, 1, False), (1024, 1, True), (2048, 3, True),
                                             (1 << 15, 16, True), (1 << 15, 2, False), (1 << 21, 1, True)])
def test_downsweep_matches_reference_levels(vren, n, blocks, clear):
    x = (rand_u32(15, n * blocks) % np.uint32(1000)).astype(np.uint32)
    tree = oracle.reduce(x, n, "u32", "add", blocks=blocks)
    want = oracle.downsweep(tree, n, blocks, clear)
    got = host_u32(vren.downsweep(dev_u32(tree), n, blocks, clear))
    assert np.array_equal(got, want)
    if clear:
        for y in range(blocks):
            assert np.array_equal(got[y * n:(y + 1) * n], oracle.exclusive_scan(x[y * n:(y + 1) * n]))


# ---- a3 radix sort ----------------------------------------------------------------------------------------------
def test_radix_reference_case_reversed_iota_1024(vren):
    n = 1 << 10   # TEST(radix_sort, main)
    x = np.arange(n, dtype=np.uint32)[::-1].copy()
    assert np.array_equal(host_u32(vren.radix_sort_keys(dev_u32(x))), oracle.sort_keys(x))


@pytest.mark.parametrize("pattern", ["uniform", "reversed", "mod100", "equal", "top_digit_only"])
@pytest.mark.parametrize("n", [1, 2, 31, 1000, 6144, 6145, 8191, 8192, 8193, 1 << 16, (1 << 20), (1 << 20) + 4097])
def test_radix_keys_vs_std_sort(vren, pattern, n):
    if pattern == "uniform":
        x = rand_u32(1, n)
    elif pattern == "reversed":
        x = np.arange(n, dtype=np.uint32)[::-1].copy()
    elif pattern == "mod100":
        x = (rand_u32(2, n) % np.uint32(100)).astype(np.uint32)
    elif pattern == "equal":
        x = np.full(n, 0xFFFFFFFF, np.uint32)
    else:
        x = (rand_u32(4, n) & np.uint32(0xFF000000)).astype(np.uint32)
    assert np.array_equal(host_u32(vren.radix_sort_keys(dev_u32(x))), oracle.sort_keys(x))


@pytest.mark.parametrize("n", [1, 77, 6144, 8191, 8193, 1 << 14, 3 * 8192 + 1, (1 << 20), (1 << 20) + 1234])
def test_radix_pairs_stable(vren, n):
    # few distinct keys -> stability is observable through the values
    k = (rand_u32(21, n) % np.uint32(1000) * np.uint32(0x00410041)).astype(np.uint32)
    v = np.arange(n, dtype=np.uint32)
    wk, wv = oracle.sort_pairs(k, v)
    gk, gv = vren.radix_sort_pairs(dev_u32(k), dev_u32(v))
    assert np.array_equal(host_u32(gk), wk) and np.array_equal(host_u32(gv), wv)
    k = rand_u32(22, n)
    wk, wv = oracle.sort_pairs(k, v)
    gk, gv = vren.radix_sort_pairs(dev_u32(k), dev_u32(v))
    assert np.array_equal(host_u32(gk), wk) and np.array_equal(host_u32(gv), wv)


RANKINGS = ["match", "verified", "sampled", "unverified", "selftest_redo"]


def _cfg(vren, ranking, tile_ids="auto", variant=0):
    r = {"match": vren.RANKING_MATCH, "verified": vren.RANKING_ATOMIC_VERIFIED, "sampled": vren.RANKING_ATOMIC_SAMPLED,
         "unverified": vren.RANKING_ATOMIC_UNVERIFIED, "selftest_redo": vren.RANKING_SELFTEST_REDO, "auto": vren.RANKING_AUTO}[ranking]
    t = {"auto": vren.TILE_IDS_AUTO, "block": vren.TILE_IDS_BLOCK_INDEX, "ticket": vren.TILE_IDS_TICKET}[tile_ids]
    return vren.SortConfig(r, t, variant)


@pytest.mark.parametrize("ranking", RANKINGS)
@pytest.mark.parametrize("pattern", ["equal", "two_values", "mod100", "low_byte_only", "uniform"])
@pytest.mark.parametrize("n", [33, 4097, 100003, (1 << 20) + 77, (1 << 22) + 12345])
def test_radix_pairs_every_ranking_collision_heavy(vren, ranking, pattern, n):
    """every lane-collision pattern of the ranking step (32 lanes on one counter ... all different) under every ranking
    mode, keys and values bit-exact against the stable oracle.  The verified modes must never see their check fail on this
    device (the hardware serves same-address lanes in lane order); "selftest_redo" makes the main passes write nothing and
    report a failed check, so there the result is what the by-construction repeat passes produce."""
    if pattern == "equal":
        k = np.full(n, 0x12345678, np.uint32)
    elif pattern == "two_values":
        k = np.where(rand_u32(51, n) & np.uint32(1), np.uint32(0x01010101), np.uint32(0x02020202)).astype(np.uint32)
    elif pattern == "mod100":
        k = (rand_u32(52, n) % np.uint32(100)).astype(np.uint32)
    elif pattern == "low_byte_only":
        k = (rand_u32(53, n) & np.uint32(0x7)).astype(np.uint32) * np.uint32(0x01010101)
    else:
        k = rand_u32(54, n)
    v = np.arange(n, dtype=np.uint32)
    wk, wv = oracle.sort_pairs(k, v)
    cfg = _cfg(vren, ranking)
    gk, gv = dev_u32(k), dev_u32(v)
    violation = vren.radix_sort_ex(gk, gv, cfg)
    assert np.array_equal(host_u32(gk), wk) and np.array_equal(host_u32(gv), wv)
    assert violation == (0xF if ranking == "selftest_redo" else 0)
    gk = dev_u32(k)
    violation = vren.radix_sort_ex(gk, None, cfg)
    assert np.array_equal(host_u32(gk), wk)
    assert violation == (0xF if ranking == "selftest_redo" else 0)


@pytest.mark.parametrize("ranking", ["match", "verified"])
@pytest.mark.parametrize("n", [1, 4096, 4097, 3 * 12288 + 5, (1 << 21) + 999, (1 << 23) + 1])
def test_radix_ticket_tile_ids(vren, ranking, n):
    """tile ids taken from an atomic ticket (start order by construction) give the same result as the block index"""
    k = rand_u32(55, n)
    v = np.arange(n, dtype=np.uint32)
    wk, wv = oracle.sort_pairs(k, v)
    gk, gv = dev_u32(k), dev_u32(v)
    assert vren.radix_sort_ex(gk, gv, _cfg(vren, ranking, "ticket")) == 0
    assert np.array_equal(host_u32(gk), wk) and np.array_equal(host_u32(gv), wv)
    gk = dev_u32(k)
    vren.radix_sort_ex(gk, None, _cfg(vren, ranking, "ticket"))
    assert np.array_equal(host_u32(gk), wk)


def test_bucket_sort_redo_path(vren):
    """the interleaved (uvec2) layout through the repeat passes"""
    import torch

    n = 300007
    pairs = np.stack([rand_u32(56, n), np.arange(n, dtype=np.uint32)], axis=1)
    want, wc = oracle.bucket_sort(pairs)
    t = torch.from_numpy(pairs.view(np.int32)).cuda()
    for ranking in ("selftest_redo", "match", "sampled"):
        _, got, counters = vren.bucket_sort(t, config=_cfg(vren, ranking))
        assert np.array_equal(got.cpu().numpy().view(np.uint32), want), ranking
        assert np.array_equal(counters.cpu().numpy().view(np.uint32), wc), ranking


@pytest.mark.parametrize("delta", [-1, 0, 1])
@pytest.mark.parametrize("tile", [12288, 11776, 16384])
def test_radix_sizes_around_the_default_tiles(vren, tile, delta):
    """whole number of default tiles (12288 pairs atomic order, 11776 ballot match, 16384 keys only) and one element either
    side, above the 2^21 switch to the large tiles"""
    n = tile * 180 + delta
    k = rand_u32(61, n)
    v = np.arange(n, dtype=np.uint32)
    wk, wv = oracle.sort_pairs(k, v)
    gk, gv = vren.radix_sort_pairs(dev_u32(k), dev_u32(v))
    assert np.array_equal(host_u32(gk), wk) and np.array_equal(host_u32(gv), wv)
    assert np.array_equal(host_u32(vren.radix_sort_keys(dev_u32(k))), wk)


def test_radix_all_variants_agree(vren):
    """every entry of the kernel table (vrenb200_sort_config::variant), pairs and keys"""
    lib = vren.load()
    n = (1 << 18) + 333
    k = rand_u32(23, n)
    v = np.arange(n, dtype=np.uint32)
    wk, wv = oracle.sort_pairs(k, v)
    for var in range(1, lib.vrenb200_radix_sort_num_variants() + 1):
        name = lib.vrenb200_radix_sort_variant_name(var)
        gk, gv = dev_u32(k), dev_u32(v)
        if b"x64/" not in name:            # 64 rows per thread: keys-only tiles
            vren.radix_sort_ex(gk, gv, _cfg(vren, "auto", variant=var))
            assert np.array_equal(host_u32(gk), wk) and np.array_equal(host_u32(gv), wv), name
        gk = dev_u32(k)
        vren.radix_sort_ex(gk, None, _cfg(vren, "auto", variant=var))
        assert np.array_equal(host_u32(gk), wk), name


@pytest.mark.parametrize("n", [1, 1000, 8192 * 3 + 5, (1 << 20) + 3])
def test_radix_host_buffers_sync_and_async(vren, n):
    """vrenb200_radix_sort_pairs_host (in place, synchronising) and _host_async (separate destination, two streams in
    flight at once, each with its own device work buffer) against the oracle's stable sort by key"""
    import torch

    lib = vren.load()
    k = rand_u32(41, n) % np.uint32(max(n // 3, 1))                       # plenty of duplicate keys: stability is visible
    v = np.arange(n, dtype=np.uint32)
    wk, wv = oracle.sort_pairs(k, v)
    wb = lib.vrenb200_radix_sort_host_work_bytes(n, 1)
    pin = lambda a: torch.from_numpy(a.view(np.int32).copy()).pin_memory()
    # synchronous, in place
    hk, hv = pin(k), pin(v)
    work = torch.empty(wb, dtype=torch.uint8, device="cuda")
    assert lib.vrenb200_radix_sort_pairs_host(None, hk.data_ptr(), hv.data_ptr(), n, work.data_ptr(), wb) == 0
    assert np.array_equal(hk.numpy().view(np.uint32), wk) and np.array_equal(hv.numpy().view(np.uint32), wv)
    # two asynchronous calls on two streams from the same read-only source
    src_k, src_v = pin(k), pin(v)
    lanes = []
    for _ in range(2):
        lanes.append((torch.cuda.Stream(), torch.empty(n, dtype=torch.int32).pin_memory(), torch.empty(n, dtype=torch.int32).pin_memory(),
                      torch.empty(wb, dtype=torch.uint8, device="cuda")))
    for rep in range(2):
        for st, ok, ov, wk_buf in lanes:
            assert lib.vrenb200_radix_sort_pairs_host_async(st.cuda_stream, src_k.data_ptr(), src_v.data_ptr(), ok.data_ptr(), ov.data_ptr(),
                                                            n, wk_buf.data_ptr(), wb) == 0
    for st, ok, ov, _ in lanes:
        st.synchronize()
        assert np.array_equal(ok.numpy().view(np.uint32), wk) and np.array_equal(ov.numpy().view(np.uint32), wv)
    assert np.array_equal(src_k.numpy().view(np.uint32), k)               # the source is not touched
    # argument checks
    assert lib.vrenb200_radix_sort_pairs_host_async(None, src_k.data_ptr(), None, hk.data_ptr(), hv.data_ptr(), n, work.data_ptr(), wb) == vren.EINVAL_ARG
    assert lib.vrenb200_radix_sort_pairs_host_async(None, src_k.data_ptr(), src_v.data_ptr(), hk.data_ptr(), hv.data_ptr(), n, work.data_ptr(), 16) == vren.ESCRATCH


def test_radix_compat_preconditions(vren):
    import torch

    lib = vren.load()
    buf = torch.zeros(4096, dtype=torch.int32, device="cuda")
    s1 = torch.zeros(lib.vrenb200_radix_sort_scratch_buffer_1_bytes(4096), dtype=torch.uint8, device="cuda")
    s2 = torch.zeros(lib.vrenb200_radix_sort_scratch_buffer_2_bytes(4096), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    # radix_sort.cpp:158-161 -> std::invalid_argument
    for bad in (0, 512, 1000, 3000):
        assert lib.vrenb200_radix_sort_compat(st, buf.data_ptr(), bad, s1.data_ptr(), s1.numel(), s2.data_ptr(), s2.numel()) == 1
    x = np.arange(4096, dtype=np.uint32)[::-1].copy()
    d = dev_u32(x)
    assert lib.vrenb200_radix_sort_compat(st, d.data_ptr(), 4096, s1.data_ptr(), s1.numel(), s2.data_ptr(), s2.numel()) == 0
    assert np.array_equal(host_u32(d), np.arange(4096, dtype=np.uint32))
    assert lib.vrenb200_radix_sort_compat(st, d.data_ptr(), 4096, s1.data_ptr(), 16, s2.data_ptr(), s2.numel()) == 3


def test_radix_pairs_full_size_headline(vren):
    """2^28 pairs (BASELINE metric size): sortedness, permutation, key/value consistency, stability."""
    import torch

    n = 1 << 28
    g = torch.Generator(device="cuda")
    g.manual_seed(1234)
    keys = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
    vals = torch.arange(n, dtype=torch.int32, device="cuda")
    orig = keys.clone()
    vren.radix_sort_pairs(keys, vals)
    torch.cuda.synchronize()
    # unsigned order == signed order after flipping the sign bit
    flipped = keys ^ torch.tensor(-(1 << 31), dtype=torch.int32, device="cuda")
    assert bool((flipped[1:] >= flipped[:-1]).all())
    assert torch.equal(orig[vals.long()], keys)                      # every pair kept together
    marks = torch.zeros(n, dtype=torch.bool, device="cuda")
    marks[vals.long()] = True
    assert bool(marks.all())                                          # values are a permutation
    ties = flipped[1:] == flipped[:-1]
    assert bool((vals[1:][ties] > vals[:-1][ties]).all())             # stable
