"""Peer-store bandwidth ceiling between GPUs of one node: every rank copies a 1 GiB buffer into the next rank's
symmetric-memory buffer with a plain vectorised copy kernel (torch .copy_ on the peer-mapped tensor), all ranks at once.
Run: torchrun --nproc-per-node N tools/p2p_bw.py"""
import json
import os

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
    dev = torch.device("cuda", torch.cuda.current_device())
    n = 1 << 28
    buf = symm.empty(n, dtype=torch.int32, device=dev)
    h = symm.rendezvous(buf, dist.group.WORLD)
    src = torch.arange(n, dtype=torch.int32, device=dev)
    peer = h.get_buffer((rank + 1) % world, (n,), torch.int32)
    for mode in ("peer_store", "local_copy"):
        dst = peer if mode == "peer_store" else torch.empty_like(src)
        for _ in range(2):
            dst.copy_(src)
        h.barrier()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            h.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dst.copy_(src)
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"mode": mode, "world": world, "ms": float(t.item()), "GB/s_per_gpu_egress": 4 * n / float(t.item()) / 1e6}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
