"""Sweeps the two run-ahead distances of the scan kernel at 2^28 uint32.  Run on the GPU box."""
import json
import sys

import numpy as np
import torch

from vren_b200 import lib as vlib

lib = vlib.load()
dev = torch.device("cuda:0")
stream = torch.cuda.current_stream().cuda_stream
n = 1 << 28
x = torch.randint(0, 16, (n,), dtype=torch.int32, device=dev)
y = torch.empty_like(x)
sb = lib.vrenb200_scan_scratch_bytes(n)
scr = torch.empty(sb, dtype=torch.uint8, device=dev)
want = ((torch.cumsum(x, 0, dtype=torch.int64) - x) & 0xFFFFFFFF)
grid = [(a, b) for a in (64, 128, 192, 256, 384, 512) for b in (32, 64, 128, 192, 256, 384) if a + b <= 768]
best = None
for a, b in grid:
    vlib.check(lib.vrenb200_scan_set_runahead(a, b), "lags")
    call = lambda: vlib.check(lib.vrenb200_exclusive_scan_u32(stream, x.data_ptr(), y.data_ptr(), n, scr.data_ptr(), sb), "scan")
    call(); torch.cuda.synchronize()
    ok = bool(torch.equal(y.to(torch.int64) & 0xFFFFFFFF, want))
    ts = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    print(json.dumps({"finalize_lag": a, "scan_lag": b, "ok": ok, "ms": round(ms, 4), "GB/s": round(8 * n / ms / 1e6, 1)}), flush=True)
    if ok and (best is None or ms < best[0]):
        best = (ms, a, b)
print("best", best)
vlib.check(lib.vrenb200_scan_set_variant(0), "variant")
