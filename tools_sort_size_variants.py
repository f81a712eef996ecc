"""Whole-sort time (pairs) versus problem size for several pass variants: where a smaller tile pays.

usage: python tools_sort_size_variants.py variant [variant ...]
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
g.manual_seed(9)
variants = [int(a) for a in sys.argv[1:]] or [0]
for log2n in (14, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26):
    n = 1 << log2n
    k0 = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
    v0 = torch.arange(n, dtype=torch.int32, device=dev)
    k, v = k0.clone(), v0.clone()
    sb = lib.vrenb200_radix_sort_scratch_bytes(n, 1)
    scr = torch.empty(sb, dtype=torch.uint8, device=dev)
    row = {"log2n": log2n}
    for var in variants:
        vlib.check(lib.vrenb200_radix_sort_set_variant(var), "variant")
        ts = []
        for it in range(14):
            k.copy_(k0); v.copy_(v0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            vlib.check(lib.vrenb200_radix_sort_pairs(stream, k.data_ptr(), v.data_ptr(), n, scr.data_ptr(), sb), "pairs")
            e1.record(); e1.synchronize()
            if it >= 4:
                ts.append(e0.elapsed_time(e1))
        row[f"v{var}_us"] = round(float(np.median(ts)) * 1000, 1)
    print(json.dumps(row), flush=True)
vlib.check(lib.vrenb200_radix_sort_set_variant(0), "variant")
