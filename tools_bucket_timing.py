"""Bucket sort (a4) at 2^26 uvec2 pairs: END offsets by per-key global atomics vs by search in the sorted output."""
import json

import numpy as np
import torch

from vren_b200 import lib as vlib

lib = vlib.load()
dev = torch.device("cuda")
stream = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device=dev)
g.manual_seed(5)
for logn in (20, 23, 26):
    nb = 1 << logn
    pairs = torch.randint(0, 1 << 16, (nb, 2), dtype=torch.int32, device=dev, generator=g)
    ob = lib.vrenb200_bucket_sort_output_bytes(nb)
    bb = lib.vrenb200_bucket_sort_scratch_bytes(nb)
    bscr = torch.empty(bb, dtype=torch.uint8, device=dev)
    outs = {}
    for name, smin in (("atomics", 0xFFFFFFFF), ("search", 0)):
        vlib.check(lib.vrenb200_bucket_sort_set_search_min(smin), "search_min")
        bout = torch.zeros(ob, dtype=torch.uint8, device=dev)
        ts = []
        for it in range(12):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            vlib.check(lib.vrenb200_bucket_sort(stream, pairs.data_ptr(), nb, bout.data_ptr(), bscr.data_ptr(), bb), "bucket_sort")
            e1.record(); e1.synchronize()
            if it >= 2:
                ts.append(e0.elapsed_time(e1))
        outs[name] = bout
        ms = float(np.median(ts))
        print(json.dumps({"n": nb, "end_offsets": name, "ms": round(ms, 4), "Gpairs/s": round(nb / ms / 1e6, 2),
                          "frac_hbm_40B": round(40 * nb / ms / 1e6 / 6553.3, 3)}), flush=True)
    print(json.dumps({"n": nb, "outputs_identical": bool(torch.equal(outs["atomics"], outs["search"]))}), flush=True)
vlib.check(lib.vrenb200_bucket_sort_set_search_min(1 << 20), "search_min")
