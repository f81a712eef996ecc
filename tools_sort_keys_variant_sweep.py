"""Keys-only sort (a3 as the reference ships it) at 2^28 keys: pass variants interleaved round-robin, sortedness checked.

usage: python tools_sort_keys_variant_sweep.py variant [variant ...]
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
g.manual_seed(13)
variants = [int(a) for a in sys.argv[1:]] or [0]
n = 1 << 28
k0 = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
want = None
k = k0.clone()
sb = lib.vrenb200_radix_sort_scratch_bytes(n, 0)
scr = torch.empty(sb, dtype=torch.uint8, device=dev)
times = {v: [] for v in variants}
ok = {}
for rnd in range(10):
    for var in variants:
        vlib.check(lib.vrenb200_radix_sort_set_variant(var), "variant")
        k.copy_(k0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        vlib.check(lib.vrenb200_radix_sort_keys(stream, k.data_ptr(), n, scr.data_ptr(), sb), "keys")
        e1.record(); e1.synchronize()
        if rnd == 0:
            u = k.to(torch.int64) & 0xFFFFFFFF
            if want is None:
                want = torch.sort(k0.to(torch.int64) & 0xFFFFFFFF).values
            ok[var] = bool(torch.equal(u, want))
            del u
        else:
            times[var].append(e0.elapsed_time(e1))
for var in variants:
    med = float(np.median(times[var]))
    print(json.dumps({"variant": var, "name": lib.vrenb200_radix_sort_variant_name(var).decode(), "ok": ok[var],
                      "sort_ms_median": round(med, 4), "Gkeys/s": round(n / med / 1e6, 2)}), flush=True)
vlib.check(lib.vrenb200_radix_sort_set_variant(0), "variant")
