"""DEPHASE sweep of the sort pass: start delay of the second CTA of every SM (first wave only) vs whole-sort time at 2^28 pairs.

usage: python tools_sort_dephase_sweep.py [variant [delays_ns ...]]
"""
import json
import sys

import numpy as np
import torch

from vren_b200 import lib as vlib

lib = vlib.load()
dev = torch.device("cuda")
stream = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device=dev)
g.manual_seed(11)
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 34
delays = [int(a) for a in sys.argv[2:]] or [0, 1000, 2000, 3000, 4000, 5000, 6000, 8000]
n = 1 << 28
k0 = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
v0 = torch.arange(n, dtype=torch.int32, device=dev)
k, v = k0.clone(), v0.clone()
sb = lib.vrenb200_radix_sort_scratch_bytes(n, 1)
scr = torch.empty(sb, dtype=torch.uint8, device=dev)


def timed(var):
    vlib.check(lib.vrenb200_radix_sort_set_variant(var), "variant")
    k.copy_(k0); v.copy_(v0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    vlib.check(lib.vrenb200_radix_sort_pairs(stream, k.data_ptr(), v.data_ptr(), n, scr.data_ptr(), sb), "pairs")
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1)


def ok():
    ku = k.to(torch.int64) & 0xFFFFFFFF
    d = ku[1:] - ku[:-1]
    vv = v.to(torch.int64)
    return bool((d >= 0).all()) and bool((k0[vv] == k).all()) and bool(((d > 0) | (vv[1:] > vv[:-1])).all())


configs = [(0, 0, 0)] + [(variant, ns, rule) for rule in (0, 1) for ns in delays]
times = {c: [] for c in configs}
good = {}
for rnd in range(8):
    for c in configs:
        vlib.check(lib.vrenb200_radix_sort_set_dephase(c[1], c[2]), "dephase")
        t = timed(c[0])
        if rnd == 0:
            good[c] = ok()
        else:
            times[c].append(t)
for c in configs:
    med = float(np.median(times[c]))
    print(json.dumps({"variant": c[0], "dephase_ns": c[1], "rule": c[2], "ok": good[c], "sort_ms_median": round(med, 4),
                      "sort_ms_min": round(float(np.min(times[c])), 4), "Gpairs/s": round(n / med / 1e6, 2)}), flush=True)
vlib.check(lib.vrenb200_radix_sort_set_variant(0), "variant")
