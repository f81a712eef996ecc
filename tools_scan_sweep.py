"""Times the exclusive-scan variants (CTA size) at 2^28 and 2^24 uint32 and checks them against torch.cumsum.

Run on the GPU box: python tools_scan_sweep.py
"""
import json

import numpy as np
import torch

from vren_b200 import lib as vlib


def main():
    import os
    fast = os.environ.get("SCAN_SWEEP_FAST") == "1"    # under ncu: one launch per variant at 2^28, no ragged cases
    lib = vlib.load()
    dev = torch.device("cuda:0")
    stream = torch.cuda.current_stream().cuda_stream
    peak = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", 6458.1) if __import__("os").path.exists("MEASURED_PEAKS.json") else 6458.1
    for log2n in ((28,) if fast else (24, 28)):
        n = 1 << log2n
        x = torch.randint(0, 16, (n,), dtype=torch.int32, device=dev)
        y = torch.empty_like(x)
        sb = lib.vrenb200_scan_scratch_bytes(n)
        scr = torch.empty(sb, dtype=torch.uint8, device=dev)
        want = torch.cumsum(x, 0, dtype=torch.int64) - x
        want = (want & 0xFFFFFFFF).to(torch.int64)
        for v in range(16):
            vlib.check(lib.vrenb200_scan_set_variant(v), "variant")
            call = lambda: vlib.check(lib.vrenb200_exclusive_scan_u32(stream, x.data_ptr(), y.data_ptr(), n, scr.data_ptr(), sb), "scan")
            call()
            torch.cuda.synchronize()
            got = y.to(torch.int64) & 0xFFFFFFFF
            ok = bool(torch.equal(got, want))
            ts = []
            for _ in range(0 if fast else 20):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); call(); e1.record(); e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = float(np.median(ts)) if ts else 0.0
            print(json.dumps({"log2n": log2n, "variant": v, "ok": ok, "ms": ms, "GB/s": 8 * n / max(ms, 1e-9) / 1e6,
                              "frac_hbm": 8 * n / max(ms, 1e-9) / 1e6 / peak}))
    # ragged sizes and unaligned views
    for v in (() if fast else (0, 1, 2, 3, 4, 8, 11, 12, 14, 15)):
        vlib.check(lib.vrenb200_scan_set_variant(v), "variant")
        bad = 0
        for n, off in ((1, 0), (5, 1), (16383, 0), (16384, 0), (16385, 3), (100003, 1), ((1 << 22) + 12345, 0), ((1 << 22) + 7, 2)):
            buf = torch.randint(0, 1 << 30, (n + 8,), dtype=torch.int32, device=dev)
            outb = torch.zeros(n + 8, dtype=torch.int32, device=dev)
            x, y = buf[off:off + n], outb[off:off + n]
            sb = lib.vrenb200_scan_scratch_bytes(n)
            scr = torch.empty(sb, dtype=torch.uint8, device=dev)
            vlib.check(lib.vrenb200_exclusive_scan_u32(stream, x.data_ptr(), y.data_ptr(), n, scr.data_ptr(), sb), "scan")
            torch.cuda.synchronize()
            want = ((torch.cumsum(x.to(torch.int64), 0) - x.to(torch.int64)) & 0xFFFFFFFF)
            got = y.to(torch.int64) & 0xFFFFFFFF
            ok = bool(torch.equal(got, want)) and int(outb[off + n:].abs().sum()) == 0 and int(outb[:off].abs().sum()) == 0
            bad += 0 if ok else 1
        print(json.dumps({"variant": v, "ragged_failures": bad}))
    vlib.check(lib.vrenb200_scan_set_variant(0), "variant")


if __name__ == "__main__":
    main()
