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


def test_plan_p2p_offsets_tile_the_receive_buffers():
    """the fused exchange writes every source's block at plan_p2p_offsets: blocks must tile each receive buffer exactly, in
    source-rank order, so that the final local stable sort yields the global stable order"""
    rng = np.random.Generator(np.random.PCG64(7))
    for world in (1, 2, 4, 8):
        hists = torch.from_numpy(rng.integers(0, 50, size=(world, 256)).astype(np.int64))
        bounds = vdist.plan_digit_ranges(hists.sum(0), world)
        plans = [vdist.plan_p2p_offsets(hists, bounds, src) for src in range(world)]
        rank_of = plans[0][0]
        for d in range(256):
            r = int(rank_of[d])
            assert bounds[r] <= d < bounds[r + 1]
        for dst in range(world):
            pos = 0
            for src in range(world):
                _, my_offset, recv_counts, per_dest = plans[src]
                assert int(my_offset[dst]) == pos
                pos += int(hists[src, bounds[dst]:bounds[dst + 1]].sum())
                assert int(per_dest[src, dst]) == int(hists[src, bounds[dst]:bounds[dst + 1]].sum())
            assert pos == plans[0][2][dst]


def test_plan_digit_exchange_then_segment_sorts_give_the_global_stable_order():
    """numpy simulation of the planned full-top-digit exchange: every source writes its pairs of top digit d at
    my_digit_offset[d] of the owner's buffer (stable inside (source, digit)), every destination then sorts each
    top-digit segment by the low 24 bits (three LSD passes); the concatenation must be the stable sort of the input"""
    rng = np.random.Generator(np.random.PCG64(12))
    for world, sizes in ((1, [1000]), (2, [5000, 3000]), (4, [4096, 0, 777, 9000]), (8, [2000] * 8)):
        keys = [rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32) for n in sizes]
        keys[0][: sizes[0] // 2] &= np.uint32(0x03FF00FF)                      # skew + many equal keys
        vals, base = [], 0
        for n in sizes:
            vals.append(np.arange(base, base + n, dtype=np.uint32))
            base += n
        hists = torch.from_numpy(np.stack([np.bincount(k >> 24, minlength=256) for k in keys]).astype(np.int64))
        bounds = vdist.plan_digit_ranges(hists.sum(0), world)
        plans = [vdist.plan_digit_exchange(hists, bounds, r) for r in range(world)]
        recv_counts = plans[0][3]
        bufs_k = [np.zeros(c, np.uint32) for c in recv_counts]
        bufs_v = [np.zeros(c, np.uint32) for c in recv_counts]
        filled = [np.zeros(c, bool) for c in recv_counts]
        for src in range(world):
            rank_of, my_off, seg_start, _ = plans[src]
            seen = np.zeros(256, np.int64)
            for k, v in zip(keys[src], vals[src]):                               # the exchange pass: stable in (source, digit)
                d = int(k >> 24)
                dst = int(rank_of[d])
                pos = int(my_off[d]) + seen[d]
                seen[d] += 1
                assert not filled[dst][pos]
                filled[dst][pos] = True
                bufs_k[dst][pos], bufs_v[dst][pos] = k, v
        out_k, out_v = [], []
        for dst in range(world):
            assert filled[dst].all()
            seg_start = plans[dst][2][dst]
            for d in range(256):                                                 # the three low passes, per segment
                lo, hi = int(seg_start[d]), int(seg_start[d + 1])
                k, v = bufs_k[dst][lo:hi], bufs_v[dst][lo:hi]
                assert k.size == 0 or ((k >> 24) == d).all()
                order = np.argsort(k & np.uint32(0x00FFFFFF), kind="stable")
                out_k.append(k[order]); out_v.append(v[order])
        all_k, all_v = np.concatenate(keys), np.concatenate(vals)
        want = np.argsort(all_k, kind="stable")
        assert np.array_equal(np.concatenate(out_k), all_k[want]) and np.array_equal(np.concatenate(out_v), all_v[want])
