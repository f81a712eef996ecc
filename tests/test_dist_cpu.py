"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in vren_b200/dist.py: digit-range planning, all-to-all-v
plumbing, scan bases.  The local compute steps are replaced by a numpy stand-in (LocalOps protocol) because the CUDA
kernels cannot run here; the GPU versions of the same steps are covered by tests/test_gpu_*.py and, on >= 2 GPUs, by
tests/test_dist_gpu.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vren_b200 import dist as vdist


class NumpyOps:
    """test double for CudaOps: same contracts, numpy on CPU tensors"""

    @staticmethod
    def _u32(t):
        return t.numpy().view(np.uint32)

    def digit_histograms(self, keys):
        k = self._u32(keys)
        h = np.stack([np.bincount((k >> (8 * d)) & 0xFF, minlength=256) for d in range(4)])
        return torch.from_numpy(h.astype(np.int32))

    def partition_by_top_digit(self, keys, vals):
        order = np.argsort(self._u32(keys) >> 24, kind="stable")
        return keys[torch.from_numpy(order)], vals[torch.from_numpy(order)]

    def sort_pairs(self, keys, vals):
        order = torch.from_numpy(np.argsort(self._u32(keys), kind="stable"))
        keys.copy_(keys[order])
        vals.copy_(vals[order])
        return keys, vals

    def reduce_add(self, x):
        return int(self._u32(x).sum(dtype=np.uint64)) & 0xFFFFFFFF

    def exclusive_scan(self, x, base):
        a = self._u32(x).astype(np.uint64)
        out = (np.concatenate([np.zeros(1, np.uint64), np.cumsum(a)[:-1]]) + np.uint64(base)) & np.uint64(0xFFFFFFFF)
        x.copy_(torch.from_numpy(out.astype(np.uint32).view(np.int32)))
        return x


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _shard(rank, world, n, skew):
    rng = np.random.Generator(np.random.PCG64(100 + rank))
    if skew:
        keys = (rng.integers(0, 5, size=n, dtype=np.uint64) << np.uint64(28)) | rng.integers(0, 1 << 20, size=n, dtype=np.uint64)
    else:
        keys = rng.integers(0, 1 << 32, size=n, dtype=np.uint64)
    vals = (np.arange(n, dtype=np.uint64) + rank * (1 << 24)).astype(np.uint32)
    return keys.astype(np.uint32), vals


def _worker(rank, world, port, sizes, skew, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        keys, vals = _shard(rank, world, sizes[rank], skew)
        tk = torch.from_numpy(keys.view(np.int32).copy())
        tv = torch.from_numpy(vals.view(np.int32).copy())
        rk, rv, plan = vdist.sharded_sort_pairs(tk, tv, ops=NumpyOps())
        np.save(os.path.join(out_dir, f"k{rank}.npy"), rk.numpy().view(np.uint32))
        np.save(os.path.join(out_dir, f"v{rank}.npy"), rv.numpy().view(np.uint32))
        np.save(os.path.join(out_dir, f"b{rank}.npy"), np.array(plan.digit_lo))
        x = torch.from_numpy((keys % 1000).astype(np.uint32).view(np.int32).copy())
        total = vdist.sharded_reduce_add(x.clone(), ops=NumpyOps())
        sx = vdist.sharded_exclusive_scan(x, ops=NumpyOps())
        np.save(os.path.join(out_dir, f"s{rank}.npy"), sx.numpy().view(np.uint32))
        np.save(os.path.join(out_dir, f"t{rank}.npy"), np.array([total], dtype=np.uint64))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("sizes,skew", [((5000, 5000), False), ((7001, 123), False), ((4000, 6000), True), ((0, 3000), False)])
def test_sharded_sort_and_scan_world2_gloo(tmp_path, sizes, skew):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, sizes, skew, str(tmp_path)), nprocs=world, join=True)
    shards = [_shard(r, world, sizes[r], skew) for r in range(world)]
    all_k = np.concatenate([s[0] for s in shards])
    all_v = np.concatenate([s[1] for s in shards])
    order = np.argsort(all_k, kind="stable")
    got_k = np.concatenate([np.load(tmp_path / f"k{r}.npy") for r in range(world)])
    got_v = np.concatenate([np.load(tmp_path / f"v{r}.npy") for r in range(world)])
    assert np.array_equal(got_k, all_k[order])                   # rank-order concatenation is globally sorted ...
    assert np.array_equal(got_v, all_v[order])                   # ... and stable
    bounds = np.load(tmp_path / "b0.npy")
    assert np.array_equal(bounds, np.load(tmp_path / "b1.npy")) and bounds[0] == 0 and bounds[-1] == 256
    for r in range(world):                                       # every rank only holds its digit range
        k = np.load(tmp_path / f"k{r}.npy")
        if k.size:
            assert (k >> 24).min() >= bounds[r] and (k >> 24).max() < bounds[r + 1]
    if not skew and min(sizes) > 1000:
        assert abs(np.load(tmp_path / "k0.npy").size - sum(sizes) / 2) < 0.05 * sum(sizes)   # balanced for uniform keys
    x = (all_k % 1000).astype(np.uint64)
    want = (np.concatenate([np.zeros(1, np.uint64), np.cumsum(x)[:-1]]) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    got = np.concatenate([np.load(tmp_path / f"s{r}.npy") for r in range(world)])
    assert np.array_equal(got, want)
    assert int(np.load(tmp_path / "t0.npy")[0]) == int(x.sum()) & 0xFFFFFFFF


def test_plan_digit_ranges_properties():
    rng = np.random.Generator(np.random.PCG64(1))
    for world in (1, 2, 4, 8):
        for _ in range(20):
            h = torch.from_numpy(rng.integers(0, 1000, size=256).astype(np.int64))
            h[rng.integers(0, 256, size=rng.integers(0, 200))] = 0
            b = vdist.plan_digit_ranges(h, world)
            assert len(b) == world + 1 and b[0] == 0 and b[-1] == 256 and all(x <= y for x, y in zip(b, b[1:]))
    # everything in one digit: one rank gets it all, nothing is lost
    h = torch.zeros(256, dtype=torch.int64)
    h[200] = 12345
    b = vdist.plan_digit_ranges(h, 8)
    assert sum(int(h[b[r]:b[r + 1]].sum()) for r in range(8)) == 12345


def _digit_hists(keys):
    return np.stack([np.bincount((keys >> np.uint32(8 * p)) & np.uint32(0xFF), minlength=256) for p in range(4)]).astype(np.int64)


@pytest.mark.parametrize("case", ["uniform", "skewed_half", "keys_below_2p24", "keys_below_2p13", "all_equal", "empty_ranks"])
@pytest.mark.parametrize("world,rounds", [(1, 1), (2, 1), (2, 3), (4, 4), (8, 2)])
def test_exchange_plan_then_segment_sorts_give_the_global_stable_order(case, world, rounds):
    """numpy simulation of the multi-GPU sort as csrc/sharded_sort.cu runs it, driven by the host mirror of its device plan
    (vdist.exchange_plan): the partition digit is the highest byte in which the keys differ; every source writes its pairs of
    digit value d, in input order, at dst_off[source][d] of the owner's tile-aligned receive buffer; every owner sorts each
    segment stably by the bytes below the partition digit, round by round, and emits the segments back to back.  The
    concatenation over the ranks must be the stable sort of the concatenated input."""
    rng = np.random.Generator(np.random.PCG64(12 + world))
    tile = 64
    sizes = [int(rng.integers(500, 3000)) for _ in range(world)]
    if case == "empty_ranks" and world > 1:
        sizes[0] = 0
        sizes[-1] = 0 if world > 2 else sizes[-1]
    keys = [rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32) for n in sizes]
    if case == "skewed_half":
        keys[-1][: sizes[-1] // 2] &= np.uint32(0x03FF00FF)
    elif case == "keys_below_2p24":
        keys = [k & np.uint32(0x00FFFFFF) for k in keys]
    elif case == "keys_below_2p13":
        keys = [k & np.uint32(0x1FFF) for k in keys]
    elif case == "all_equal":
        keys = [np.full_like(k, 0xABCD1234) for k in keys]
    vals, base = [], 0
    for n in sizes:
        vals.append(np.arange(base, base + n, dtype=np.uint32))
        base += n
    hists = np.stack([_digit_hists(k) for k in keys])
    total_tiles = sum(sizes) // tile + 257
    plan = vdist.exchange_plan(hists, 4, tile, rounds, total_tiles, total_tiles)
    assert plan["error"] == 0
    p, owner, bounds = plan["pstar"], plan["owner"], plan["bounds"]
    expect_p = {"keys_below_2p24": 2, "keys_below_2p13": 1, "all_equal": 0}.get(case, 3)
    assert p == (expect_p if sum(sizes) > 1 else 0)
    cap = total_tiles * tile
    bufs_k = [np.zeros(cap, np.uint32) for _ in range(world)]
    bufs_v = [np.zeros(cap, np.uint32) for _ in range(world)]
    filled = [np.zeros(cap, bool) for _ in range(world)]
    for src in range(world):
        digit = (keys[src] >> np.uint32(8 * p)) & np.uint32(0xFF)
        seen = np.zeros(256, np.int64)
        for k, v, d in zip(keys[src], vals[src], digit):                       # local partition + transfer: stable in (source, digit)
            dst = int(owner[d])
            pos = int(plan["dst_off"][src][d]) + seen[d]
            seen[d] += 1
            assert not filled[dst][pos]
            filled[dst][pos] = True
            bufs_k[dst][pos], bufs_v[dst][pos] = k, v
    out_k, out_v = [], []
    low_mask = np.uint32((1 << (8 * p)) - 1)
    for dst in range(world):
        rd = plan["round_digit"][dst]
        assert rd[0] == bounds[dst] and rd[-1] == bounds[dst + 1] and all(a <= b for a, b in zip(rd, rd[1:]))
        got = 0
        for k_round in range(rounds):
            for d in range(int(rd[k_round]), int(rd[k_round + 1])):            # the segmented passes of one round
                lo = int(plan["first_tile"][d]) * tile
                n = int(plan["seg_len"][d])
                assert filled[dst][lo:lo + n].all() and not filled[dst][lo + n:(lo + n + tile - 1) // tile * tile].any()
                k, v = bufs_k[dst][lo:lo + n], bufs_v[dst][lo:lo + n]
                assert n == 0 or (((k >> np.uint32(8 * p)) & np.uint32(0xFF)) == d).all()
                order = np.argsort(k & low_mask, kind="stable")
                out_k.append(k[order]); out_v.append(v[order])
                got += n
        assert got == plan["out_count"][dst] == int(filled[dst].sum())
    all_k, all_v = np.concatenate(keys), np.concatenate(vals)
    want = np.argsort(all_k, kind="stable")
    assert np.array_equal(np.concatenate(out_k), all_k[want]) and np.array_equal(np.concatenate(out_v), all_v[want])


def test_exchange_plan_reports_what_does_not_fit():
    """everything in one value of the partition digit: the owner needs room for all of it (error bit 0); a round that
    outgrows its launch bound is error bit 1"""
    k = np.concatenate([np.full(1000, 0x05000000, np.uint32), np.full(10, 0x06000000, np.uint32)])
    hists = np.stack([_digit_hists(k), _digit_hists(k)])
    assert vdist.exchange_plan(hists, 4, 64, 1, 40, 40)["error"] == 0
    assert vdist.exchange_plan(hists, 4, 64, 1, 30, 30)["error"] & 1
    assert vdist.exchange_plan(hists, 4, 64, 2, 40, 20)["error"] & 2


def test_required_capacity_is_the_smallest_that_fits():
    """the documented overflow policy: a skewed input (two values of the top digit, 90 % in one) does not fit the default capacity — the plan
    says so — and fits exactly from vdist.required_capacity on"""
    rng = np.random.Generator(np.random.PCG64(5))
    world, n, tile = 4, 2_000_000, 4096
    keys = [(((rng.random(n) < 0.1).astype(np.uint64) << np.uint64(30)) | rng.integers(0, 1 << 12, size=n, dtype=np.uint64)).astype(np.uint32)
            for _ in range(world)]                           # 90 % of all pairs share one value of the top digit
    hists = np.stack([_digit_hists(k) for k in keys])
    default = vdist.default_capacity(n, world)
    assert vdist.exchange_plan(hists, 4, tile, 1, default // tile, default // tile)["error"] & 1
    need = vdist.required_capacity(hists, 4, tile)
    assert need % tile == 0 and need > default
    assert vdist.exchange_plan(hists, 4, tile, 1, need // tile, need // tile)["error"] == 0
    assert vdist.exchange_plan(hists, 4, tile, 1, need // tile - 1, need // tile)["error"] & 1
    # balanced keys: the requirement stays below the default (25 % slack + a partial tile per segment)
    uni = np.stack([_digit_hists(rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)) for _ in range(world)])
    assert vdist.required_capacity(uni, 4, 12288) <= vdist.default_capacity(n, world)
