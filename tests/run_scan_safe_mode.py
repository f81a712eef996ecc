"""The scan's safe mode (vrenb200_exclusive_scan_u32_ex with VRENB200_SCAN_TILE_IDS_TICKET: ticket tile ids, csrc/scan.cu) on the
GPU against the oracle.  Run by tests/test_scan_safe_mode.py in a SUBPROCESS: the two kernel instances it launches had never run
on hardware when they were committed.

    python tests/run_scan_safe_mode.py
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import oracle  # noqa: E402
from vren_b200 import lib as vlib  # noqa: E402


def main():
    torch.cuda.set_device(0)
    lib = vlib.load()
    stream = torch.cuda.current_stream().cuda_stream
    rng = np.random.Generator(np.random.PCG64(99))
    cases = 0
    for n in (1, 1000, 8192, 8193, 100003, (1 << 20) + 5, (1 << 22) + 16384 + 1, (1 << 24) + 3):
        x = rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
        base = 0x9E3779B9
        want = (oracle.exclusive_scan(x).astype(np.uint64) + np.uint64(base)).astype(np.uint32) if n else x
        sb = lib.vrenb200_scan_scratch_bytes(n)
        scratch = torch.empty(max(sb, 256), dtype=torch.uint8, device="cuda")
        src = torch.from_numpy(x.view(np.int32)).cuda()
        dst = torch.zeros_like(src)
        vlib.check(lib.vrenb200_exclusive_scan_u32_ex(stream, src.data_ptr(), dst.data_ptr(), n, base, scratch.data_ptr(), sb, vlib.SCAN_TILE_IDS_TICKET),
                   "exclusive_scan_u32_ex")
        assert np.array_equal(dst.cpu().numpy().view(np.uint32), want), (n, "out of place")
        for _ in range(3):                                   # in place, repeatedly: the ticket is re-zeroed by every call
            buf = src.clone()
            vlib.check(lib.vrenb200_exclusive_scan_u32_ex(stream, buf.data_ptr(), buf.data_ptr(), n, base, scratch.data_ptr(), sb, vlib.SCAN_TILE_IDS_TICKET),
                       "exclusive_scan_u32_ex")
            assert np.array_equal(buf.cpu().numpy().view(np.uint32), want), (n, "in place")
        cases += 1
    # an unknown flag is refused
    assert lib.vrenb200_exclusive_scan_u32_ex(stream, src.data_ptr(), dst.data_ptr(), n, 0, scratch.data_ptr(), sb, 2) != 0
    torch.cuda.synchronize()
    print(f"ok {cases} cases")


if __name__ == "__main__":
    main()
