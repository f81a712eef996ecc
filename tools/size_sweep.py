"""Throughput of sort (pairs / keys), scan and reduce versus problem size on one GPU (CUDA events, median of 10)."""
import json

import numpy as np
import torch

from vren_b200 import lib as vlib

lib = vlib.load()
dev = torch.device("cuda")
stream = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device=dev)
g.manual_seed(9)


def timed(fn, restore=None, iters=10):
    ts = []
    for it in range(iters + 2):
        if restore:
            restore()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        if it >= 2:
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


for log2n in (10, 12, 14, 16, 18, 20, 22, 24, 26, 28):
    n = 1 << log2n
    k0 = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
    v0 = torch.arange(n, dtype=torch.int32, device=dev)
    k, v = k0.clone(), v0.clone()
    sb = lib.vrenb200_radix_sort_scratch_bytes(n, 1)
    scr = torch.empty(sb, dtype=torch.uint8, device=dev)
    ms_pairs = timed(lambda: vlib.check(lib.vrenb200_radix_sort_pairs(stream, k.data_ptr(), v.data_ptr(), n, scr.data_ptr(), sb), "pairs"),
                     restore=lambda: (k.copy_(k0), v.copy_(v0)))
    ms_keys = timed(lambda: vlib.check(lib.vrenb200_radix_sort_keys(stream, k.data_ptr(), n, scr.data_ptr(), sb), "keys"), restore=lambda: k.copy_(k0))
    ssb = lib.vrenb200_scan_scratch_bytes(n)
    sscr = torch.empty(ssb, dtype=torch.uint8, device=dev)
    ms_scan = timed(lambda: vlib.check(lib.vrenb200_exclusive_scan_u32(stream, v0.data_ptr(), v.data_ptr(), n, sscr.data_ptr(), ssb), "scan"))
    print(json.dumps({"log2n": log2n, "sort_pairs_ms": round(ms_pairs, 4), "Gpairs/s": round(n / ms_pairs / 1e6, 2),
                      "sort_keys_ms": round(ms_keys, 4), "Gkeys/s": round(n / ms_keys / 1e6, 2),
                      "scan_ms": round(ms_scan, 4), "scan_GB/s": round(8 * n / ms_scan / 1e6, 1)}), flush=True)
    del k, v, k0, v0, scr, sscr
