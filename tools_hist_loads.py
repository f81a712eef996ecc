"""Histogram kernel of the sort: 2 vs 4 128-bit loads in flight per thread (per-kernel time from the profiled sort)."""
import ctypes as C
import json

import numpy as np
import torch

from vren_b200 import lib as vlib

lib = vlib.load()
dev = torch.device("cuda")
stream = torch.cuda.current_stream().cuda_stream
n = 1 << 28
g = torch.Generator(device=dev)
g.manual_seed(3)
k0 = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
v0 = torch.arange(n, dtype=torch.int32, device=dev)
k, v = k0.clone(), v0.clone()
sb = lib.vrenb200_radix_sort_scratch_bytes(n, 1)
scr = torch.empty(sb, dtype=torch.uint8, device=dev)
prof = lib.vrenb200_sort_profile_create()
ms = (C.c_float * 6)()
for loads in (2, 4, 2, 4):
    vlib.check(lib.vrenb200_radix_sort_set_hist_loads(loads), "hist loads")
    hist, total = [], []
    for it in range(8):
        k.copy_(k0); v.copy_(v0)
        vlib.check(lib.vrenb200_radix_sort_pairs_profiled(stream, k.data_ptr(), v.data_ptr(), n, scr.data_ptr(), sb, prof), "sort")
        torch.cuda.synchronize()
        vlib.check(lib.vrenb200_sort_profile_read(prof, ms), "read")
        if it >= 2:
            hist.append(ms[0]); total.append(sum(ms))
    print(json.dumps({"hist_loads_in_flight": loads, "histogram_ms": round(float(np.median(hist)), 4), "sort_ms": round(float(np.median(total)), 4)}), flush=True)
vlib.check(lib.vrenb200_radix_sort_set_hist_loads(4), "hist loads")
