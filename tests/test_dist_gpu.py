"""multi-GPU (NCCL) parity of the sharded sort / scan against the oracle; needs >= 2 visible GPUs (gpurun --gpus 2)."""
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


def _shard(rank, n, skew):
    rng = np.random.Generator(np.random.PCG64(500 + rank))
    if skew:
        keys = (rng.integers(0, 3, size=n, dtype=np.uint64) << np.uint64(30)) | rng.integers(0, 1 << 12, size=n, dtype=np.uint64)
    else:
        keys = rng.integers(0, 1 << 32, size=n, dtype=np.uint64)
    return keys.astype(np.uint32), (np.arange(n, dtype=np.uint64) + rank * (1 << 26)).astype(np.uint32)


def _worker(rank, world, port, sizes, skew, out_dir, p2p=False):
    import torch
    import torch.distributed as dist

    from vren_b200 import dist as vdist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        keys, vals = _shard(rank, sizes[rank], skew)
        tk = torch.from_numpy(keys.view(np.int32).copy()).cuda()
        tv = torch.from_numpy(vals.view(np.int32).copy()).cuda()
        if p2p:
            ex = vdist.P2PExchange(int(sum(sizes) * 1.5) + 4096, torch.device("cuda", rank))
            sentinel = 0x5EED5EED
            ex.keys.fill_(sentinel)
            ex.vals.fill_(sentinel)
            for _ in range(2):      # twice: the receive buffers are reused
                rk, rv, plan = vdist.sharded_sort_pairs_p2p(tk, tv, ex)
            # nothing may be stored beyond the pairs this rank receives (the padding keys of a ragged last tile of a
            # sender would land in the block of the next sender, or here)
            assert bool((ex.keys[rk.numel():] == sentinel).all()) and bool((ex.vals[rk.numel():] == sentinel).all())
            rk, rv = rk.clone(), rv.clone()
        else:
            rk, rv, plan = vdist.sharded_sort_pairs(tk, tv)
        np.save(os.path.join(out_dir, f"k{rank}.npy"), rk.cpu().numpy().view(np.uint32))
        np.save(os.path.join(out_dir, f"v{rank}.npy"), rv.cpu().numpy().view(np.uint32))
        x = torch.from_numpy((keys % 1000).astype(np.uint32).view(np.int32).copy()).cuda()
        sx = vdist.sharded_exclusive_scan(x)
        np.save(os.path.join(out_dir, f"s{rank}.npy"), sx.cpu().numpy().view(np.uint32))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("p2p", [False, True])
@pytest.mark.parametrize("sizes,skew", [((1 << 20, 1 << 20), False), ((300001, 77), False), ((1 << 18, 1 << 19), True)])
def test_sharded_sort_and_scan_nccl(vren, tmp_path, sizes, skew, p2p):
    import torch
    import torch.multiprocessing as mp

    world = 2
    if torch.cuda.device_count() < world:
        pytest.skip("needs 2 GPUs")
    mp.spawn(_worker, args=(world, _free_port(), sizes, skew, str(tmp_path), p2p), nprocs=world, join=True)
    shards = [_shard(r, sizes[r], skew) for r in range(world)]
    all_k = np.concatenate([s[0] for s in shards])
    all_v = np.concatenate([s[1] for s in shards])
    wk, wv = oracle.sort_pairs(all_k, all_v)
    assert np.array_equal(np.concatenate([np.load(tmp_path / f"k{r}.npy") for r in range(world)]), wk)
    assert np.array_equal(np.concatenate([np.load(tmp_path / f"v{r}.npy") for r in range(world)]), wv)
    want = oracle.exclusive_scan((all_k % 1000).astype(np.uint32))
    assert np.array_equal(np.concatenate([np.load(tmp_path / f"s{r}.npy") for r in range(world)]), want)
