"""Times the exclusive-scan variants (CTA size) at 2^28 and 2^24 uint32 and checks them against torch.cumsum.

Run on the GPU box: python tools_scan_sweep.py
"""
import json

import numpy as np
import torch

from vren_b200 import lib as vlib


def main():
    lib = vlib.load()
    dev = torch.device("cuda:0")
    stream = torch.cuda.current_stream().cuda_stream
    peak = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", 6458.1) if __import__("os").path.exists("MEASURED_PEAKS.json") else 6458.1
    for log2n in (24, 28):
        n = 1 << log2n
        x = torch.randint(0, 16, (n,), dtype=torch.int32, device=dev)
        y = torch.empty_like(x)
        sb = lib.vrenb200_scan_scratch_bytes(n)
        scr = torch.empty(sb, dtype=torch.uint8, device=dev)
        want = torch.cumsum(x, 0, dtype=torch.int64) - x
        want = (want & 0xFFFFFFFF).to(torch.int64)
        for v in range(3):
            vlib.check(lib.vrenb200_scan_set_variant(v), "variant")
            call = lambda: vlib.check(lib.vrenb200_exclusive_scan_u32(stream, x.data_ptr(), y.data_ptr(), n, scr.data_ptr(), sb), "scan")
            call()
            torch.cuda.synchronize()
            got = y.to(torch.int64) & 0xFFFFFFFF
            ok = bool(torch.equal(got, want))
            ts = []
            for _ in range(20):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); call(); e1.record(); e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = float(np.median(ts))
            print(json.dumps({"log2n": log2n, "variant": v, "ok": ok, "ms": ms, "GB/s": 8 * n / ms / 1e6, "frac_hbm": 8 * n / ms / 1e6 / peak}))
    vlib.check(lib.vrenb200_scan_set_variant(0), "variant")


if __name__ == "__main__":
    main()
