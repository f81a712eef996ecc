"""torchrun script: phase timing of the fused P2P sharded sort (rank 0 prints)"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from vren_b200 import dist as vdist  # noqa: E402

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
n = 1 << 28
g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
keys0 = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
vals0 = torch.arange(n, dtype=torch.int32, device=dev)
ops = vdist.CudaOps()
ex = vdist.P2PExchange(int(n * 1.25) + 4096, dev)


def sync():
    torch.cuda.synchronize()
    return time.perf_counter()


for it in range(4):
    keys, vals = keys0.clone(), vals0.clone()
    dist.barrier(); t0 = sync()
    hist = ops.top_digit_histogram(keys).to(torch.int64)
    t1 = sync()
    gathered = [torch.empty_like(hist) for _ in range(world)]
    dist.all_gather(gathered, hist)
    hists = torch.stack(gathered).cpu()
    t2 = sync()
    bounds = vdist.plan_digit_ranges(hists.sum(0), world)
    dest_rank, dest_off, recv_counts = vdist.plan_p2p_offsets(hists, bounds, rank)
    kp = torch.tensor(ex.key_ptrs, dtype=torch.int64)[dest_rank] + 4 * dest_off
    vp = torch.tensor(ex.val_ptrs, dtype=torch.int64)[dest_rank] + 4 * dest_off
    table = torch.stack([kp, vp]).contiguous().to(dev)
    t3 = sync()
    ex.barrier()
    t4 = sync()
    ops.partition_scatter(keys, vals, table)
    t5 = sync()
    ex.barrier()
    t6 = sync()
    nr = recv_counts[rank]
    ops.sort_pairs(ex.keys[:nr], ex.vals[:nr])
    t7 = sync()
    if rank == 0:
        print("it %d ms: hist %.2f gather %.2f plan %.2f barrier %.2f scatter %.2f barrier %.2f localsort %.2f total %.2f" % (
            it, *[(b - a) * 1e3 for a, b in ((t0, t1), (t1, t2), (t2, t3), (t3, t4), (t4, t5), (t5, t6), (t6, t7), (t0, t7))]), flush=True)
    # same local sort on ordinary memory for comparison
    k2, v2 = ex.keys[:nr].clone(), ex.vals[:nr].clone()
    k2.copy_(keys0[:nr]) if nr <= n else None
    ta = sync(); ops.sort_pairs(k2, v2); tb = sync()
    if rank == 0:
        print("   same-size sort in cudaMalloc memory: %.2f ms" % ((tb - ta) * 1e3), flush=True)
dist.destroy_process_group()
