"""multi-GPU parity of the sharded sort / bucket sort / scan against the oracle: one process per GPU, NVLink peer stores
through torch symmetric memory (the product path, vren_b200.dist.ShardedSort) and the NCCL baseline.  World sizes 2, 4, 8:
a case is skipped when the box has fewer GPUs (gpurun --gpus N; logs under profiles/r2*_pytest_dist_gpu_n*.log)."""
import datetime
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
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank), timeout=datetime.timedelta(seconds=90))
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


def _view_worker(rank, world, port, out_dir, w, h, L, views):
    import math

    import torch
    import torch.distributed as dist

    from vren_b200 import lib as vlib
    from vren_b200 import synthetic
    from vren_b200.pipeline import ViewBatch

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank), timeout=datetime.timedelta(seconds=90))
    try:
        vlib.load()
        oc = oracle.default_camera(w, h)
        vc = vlib.Camera(oc.fov_y, oc.aspect_ratio, oc.near_plane, oc.far_plane)
        vb = ViewBatch(w, h, L)
        pos = torch.zeros(L, 4, dtype=torch.float32, device="cuda")
        lights = torch.zeros(L, 4, dtype=torch.float32, device="cuda")
        if rank == 0:       # only the source rank has the frame's lights; the others get them by broadcast
            p0, l0 = synthetic.point_lights(L, seed=71, aspect=w / h, intensity=(0.5, 3.0))
            pos, lights = torch.from_numpy(p0).cuda(), torch.from_numpy(l0).cuda()
        vb.set_lights(pos, lights, L, src=0)
        frames = {v: (vc, synthetic.view_matrix(math.radians(45.0) * v, 0.0, (0.0, 0.0, 0.0)).tolist(),
                      torch.from_numpy(synthetic.depth_buffer(w, h, seed=80 + v)).cuda(), None) for v in vb.my_views(views)}

        def on_view(v, cs):
            torch.cuda.synchronize()
            n = int(cs.dispatch_params[0])
            total = int(cs.status.cpu().numpy().view(np.uint32)[0])
            np.savez(os.path.join(out_dir, f"view{v}.npz"), keys=cs.cluster_keys[:n].cpu().numpy().view(np.uint32),
                     counts=cs.counts.cpu().numpy().view(np.uint32), offsets=cs.offsets.cpu().numpy().view(np.uint32),
                     indices=cs.indices[:total].cpu().numpy().view(np.uint32), rank=np.array([rank]))

        vb(frames, on_view)
        torch.cuda.synchronize()
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_view_batch_one_view_per_gpu(vren, tmp_path, world):
    """SURVEY 8e row 4 (C5 batched): 8 views (yaw += 45 degrees) over the ranks, lights broadcast once from rank 0; every view's
    cluster keys, counts, offsets and ordered light lists bit-exact vs the oracle run view by view"""
    import math

    import torch
    import torch.multiprocessing as mp

    from vren_b200 import synthetic

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    w, h, L, views = 640, 360, 4000, 8
    mp.spawn(_view_worker, args=(world, _free_port(), str(tmp_path), w, h, L, views), nprocs=world, join=True)
    oc = oracle.default_camera(w, h)
    pos, lights = synthetic.point_lights(L, seed=71, aspect=w / h, intensity=(0.5, 3.0))
    for v in range(views):
        got = np.load(tmp_path / f"view{v}.npz")
        assert int(got["rank"][0]) == v % world
        view = synthetic.view_matrix(math.radians(45.0) * v, 0.0, (0.0, 0.0, 0.0))
        wvp, wnodes, wpairs = oracle.construct_point_light_bvh(pos, lights, view)
        wkeys, _ = oracle.find_unique_clusters(synthetic.depth_buffer(w, h, seed=80 + v), None, oc)
        wcounts, woffsets, windices, wtotal = oracle.assign_lights(w, h, oc, wkeys, 1 << 17, wnodes, L, wpairs, wvp, 1 << 23)
        assert np.array_equal(got["keys"], wkeys) and np.array_equal(got["counts"], wcounts) and np.array_equal(got["offsets"], woffsets)
        assert np.array_equal(got["indices"], windices[:wtotal])
