"""multi-GPU parity of the sharded sort / bucket sort / scan against the oracle: one process per GPU, NVLink peer stores
through torch symmetric memory (the product path, vren_b200.dist.ShardedSort) and the NCCL baseline.  World sizes 2, 4, 8:
a case is skipped when the box has fewer GPUs (gpurun --gpus N; logs under profiles/r2*_pytest_dist_gpu_n*.log)."""
import os
import socket

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _shard(rank, n, case):
    rng = np.random.Generator(np.random.PCG64(500 + rank))
    if case == "three_values":
        keys = (rng.integers(0, 3, size=n, dtype=np.uint64) << np.uint64(30)) | rng.integers(0, 1 << 12, size=n, dtype=np.uint64)
    elif case == "below_2p24":
        keys = rng.integers(0, 1 << 24, size=n, dtype=np.uint64)
    else:
        keys = rng.integers(0, 1 << 32, size=n, dtype=np.uint64)
    return keys.astype(np.uint32), (np.arange(n, dtype=np.uint64) + rank * (1 << 26)).astype(np.uint32)


def _worker(rank, world, port, sizes, case, out_dir, path, rounds):
    import torch
    import torch.distributed as dist

    from vren_b200 import dist as vdist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        keys, vals = _shard(rank, sizes[rank], case)
        tk = torch.from_numpy(keys.view(np.int32).copy()).cuda()
        tv = torch.from_numpy(vals.view(np.int32).copy()).cuda()
        if path == "p2p":
            # skewed cases: room for everything on one rank (the documented overflow policy: a plan that does not fit is
            # reported, the caller retries with a larger capacity)
            cap = sum(sizes) + 257 * 12288 if case == "three_values" else None
            ctx = vdist.ShardedSort.for_process_group(max(max(sizes), 1), cap, rounds)
            for _ in range(3):      # three sorts: the receive buffers, flags and epochs are reused
                ctx.sort(tk, tv)
            torch.cuda.synchronize()
            rk, rv = ctx.result()
            rk, rv = rk.clone(), rv.clone()
            # sharded bucket sort on the same context: stable by the low 16 bits, END offsets of the global sequence
            pairs = torch.stack([tk, tv], dim=1).contiguous()
            bp, ends = vdist.sharded_bucket_sort(ctx, pairs)
            np.save(os.path.join(out_dir, f"b{rank}.npy"), bp.cpu().numpy().view(np.uint32))
            if rank == 0:
                np.save(os.path.join(out_dir, "ends.npy"), ends.cpu().numpy())
            dist.barrier()
            ctx.close()
        else:
            rk, rv, _ = vdist.sharded_sort_pairs(tk, tv)
        np.save(os.path.join(out_dir, f"k{rank}.npy"), rk.cpu().numpy().view(np.uint32))
        np.save(os.path.join(out_dir, f"v{rank}.npy"), rv.cpu().numpy().view(np.uint32))
        x = torch.from_numpy((keys % 1000).astype(np.uint32).view(np.int32).copy()).cuda()
        total = vdist.sharded_reduce_add(x)
        sx = vdist.sharded_exclusive_scan(x)
        np.save(os.path.join(out_dir, f"s{rank}.npy"), sx.cpu().numpy().view(np.uint32))
        if rank == 0:
            np.save(os.path.join(out_dir, "total.npy"), np.array([total], np.uint64))
    finally:
        dist.destroy_process_group()


CASES = [
    ("uniform", 1 << 20, 1),
    ("uniform", (1 << 21) + 12345, 4),     # large tiles, four overlapped rounds
    ("ragged", None, 2),
    ("three_values", 1 << 18, 1),
    ("below_2p24", 300000, 2),
]


# the NCCL baseline has no rounds: it runs the cases once
RUNS = [(c, n, r, "p2p") for c, n, r in CASES] + [(c, n, 1, "nccl") for c, n, r in CASES if c in ("uniform", "ragged", "three_values") and n != (1 << 21) + 12345]


@pytest.mark.parametrize("case,n,rounds,path", RUNS)
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_sort_bucket_sort_and_scan(vren, tmp_path, world, case, n, rounds, path):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    if case == "ragged":
        sizes = tuple([300001, 77, 0, 4096, 1, 150000, 12288, 99999][:world])
        case = "uniform"
    else:
        sizes = tuple([n + 3 * r for r in range(world)])
    mp.spawn(_worker, args=(world, _free_port(), sizes, case, str(tmp_path), path, rounds), nprocs=world, join=True)
    shards = [_shard(r, sizes[r], case) for r in range(world)]
    all_k = np.concatenate([s[0] for s in shards])
    all_v = np.concatenate([s[1] for s in shards])
    wk, wv = oracle.sort_pairs(all_k, all_v)
    assert np.array_equal(np.concatenate([np.load(tmp_path / f"k{r}.npy") for r in range(world)]), wk)
    assert np.array_equal(np.concatenate([np.load(tmp_path / f"v{r}.npy") for r in range(world)]), wv)
    x = (all_k % 1000).astype(np.uint32)
    assert np.array_equal(np.concatenate([np.load(tmp_path / f"s{r}.npy") for r in range(world)]), oracle.exclusive_scan(x))
    assert int(np.load(tmp_path / "total.npy")[0]) == int(x.astype(np.uint64).sum() & 0xFFFFFFFF)
    if path == "p2p":
        want, counters = oracle.bucket_sort(np.stack([all_k, all_v], axis=1))
        got = np.concatenate([np.load(tmp_path / f"b{r}.npy").reshape(-1, 2) for r in range(world)])
        assert np.array_equal(got, want)
        assert np.array_equal(np.load(tmp_path / "ends.npy").astype(np.uint32), counters)
