"""Sort-pass variants on one GPU: correctness (sorted, stable, pairs intact) and time of the whole 2^28-pair sort.

usage: python tools_sort_variant_sweep.py [variant ...]     (default: 0 and every variant from 17 up)
"""
import json
import os
import sys

import numpy as np
import torch

from vren_b200 import lib as vlib

lib = vlib.load()
dev = torch.device("cuda")
stream = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device=dev)
g.manual_seed(11)

nvar = lib.vrenb200_radix_sort_num_variants()
variants = [int(a) for a in sys.argv[1:]] or [v for v in [0] + list(range(17, nvar))
                                             if b"[retired]" not in lib.vrenb200_radix_sort_variant_name(v)]   # 0 = ballot-match default


def check_sorted(k0, k, v):
    ku = k.to(torch.int64) & 0xFFFFFFFF
    d = ku[1:] - ku[:-1]
    ok_sorted = bool((d >= 0).all())
    ok_pairs = bool((k0[v.to(torch.int64)] == k).all())
    vv = v.to(torch.int64)
    ok_stable = bool(((d > 0) | (vv[1:] > vv[:-1])).all())
    return ok_sorted and ok_pairs and ok_stable


def run(n, variant, iters):
    k0 = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
    v0 = torch.arange(n, dtype=torch.int32, device=dev)
    k, v = k0.clone(), v0.clone()
    sb = lib.vrenb200_radix_sort_scratch_bytes(n, 1)
    scr = torch.empty(sb, dtype=torch.uint8, device=dev)
    vlib.check(lib.vrenb200_radix_sort_set_variant(variant), "variant")
    ts = []
    ok = True
    for it in range(iters + 2):
        k.copy_(k0); v.copy_(v0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        vlib.check(lib.vrenb200_radix_sort_pairs(stream, k.data_ptr(), v.data_ptr(), n, scr.data_ptr(), sb), "pairs")
        e1.record(); e1.synchronize()
        if it == 0:
            ok = check_sorted(k0, k, v)
        if it >= 2:
            ts.append(e0.elapsed_time(e1))
    return ok, float(np.median(ts)), float(np.min(ts))


if os.environ.get("VREN_SWEEP_DISTANCES"):
    for dist in (int(d) for d in os.environ["VREN_SWEEP_DISTANCES"].split(",")):     # PREFETCH_L2 distance
        vlib.check(lib.vrenb200_radix_sort_set_prefetch_tiles(dist), "prefetch distance")
        ok, med, mn = run(1 << 28, 0, 7)
        print(json.dumps({"variant": 0, "prefetch_tiles": dist, "ok_2p28": ok, "sort_ms_median": round(med, 4), "sort_ms_min": round(mn, 4)}), flush=True)
vlib.check(lib.vrenb200_radix_sort_set_prefetch_tiles(int(os.environ.get("VREN_PREFETCH_TILES", "148"))), "prefetch distance")

# variants interleaved round-robin on the same buffers, so that clock / thermal drift hits all of them alike
n = 1 << 28
ok_small = {var: run((1 << 22) + 12345, var, 1)[0] for var in variants}       # ragged last tile
k0 = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
v0 = torch.arange(n, dtype=torch.int32, device=dev)
k, v = k0.clone(), v0.clone()
sb = lib.vrenb200_radix_sort_scratch_bytes(n, 1)
scr = torch.empty(sb, dtype=torch.uint8, device=dev)
times = {var: [] for var in variants}
ok_big = {}
ROUNDS = 11
for rnd in range(ROUNDS + 1):
    for var in variants:
        vlib.check(lib.vrenb200_radix_sort_set_variant(var), "variant")
        k.copy_(k0); v.copy_(v0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        vlib.check(lib.vrenb200_radix_sort_pairs(stream, k.data_ptr(), v.data_ptr(), n, scr.data_ptr(), sb), "pairs")
        e1.record(); e1.synchronize()
        if rnd == 0:
            ok_big[var] = check_sorted(k0, k, v)
        else:
            times[var].append(e0.elapsed_time(e1))
for var in variants:
    name = lib.vrenb200_radix_sort_variant_name(var).decode()
    med, mn = float(np.median(times[var])), float(np.min(times[var]))
    print(json.dumps({"variant": var, "name": name, "ok_ragged_2p22": ok_small[var], "ok_2p28": ok_big[var], "sort_ms_median": round(med, 4),
                      "sort_ms_min": round(mn, 4), "Gpairs/s": round(n / med / 1e6, 2)}), flush=True)
vlib.check(lib.vrenb200_radix_sort_set_variant(0), "variant")
