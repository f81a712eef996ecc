"""scan kernels at mid sizes (where to switch from the register-tile kernel to the run-ahead kernel)"""
import json
import numpy as np
import torch
from vren_b200 import lib as vlib
lib = vlib.load()
dev = torch.device("cuda")
stream = torch.cuda.current_stream().cuda_stream
for log2n in (20, 21, 22, 23, 24, 25):
    n = 1 << log2n
    x = torch.randint(0, 16, (n,), dtype=torch.int32, device=dev)
    y = torch.empty_like(x)
    sb = lib.vrenb200_scan_scratch_bytes(n)
    scr = torch.empty(sb, dtype=torch.uint8, device=dev)
    row = {"log2n": log2n}
    for name, v in (("reg256", 1), ("reg1024", 3), ("staged", 4), ("ra64h", 12), ("ra128h", 13), ("ra256h", 14)):
        vlib.check(lib.vrenb200_scan_set_variant(v), "variant")
        call = lambda: vlib.check(lib.vrenb200_exclusive_scan_u32(stream, x.data_ptr(), y.data_ptr(), n, scr.data_ptr(), sb), "scan")
        ts = []
        for it in range(25):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); call(); e1.record(); e1.synchronize()
            if it >= 5:
                ts.append(e0.elapsed_time(e1))
        row[name] = round(float(np.median(ts)) * 1000, 1)
    print(json.dumps(row), flush=True)
vlib.check(lib.vrenb200_scan_set_variant(0), "variant")
