"""Single-CTA radix sort (VRENB200_SORT_VARIANT_SINGLE_CTA, csrc/small_sort.cu) on the GPU against the oracle.  Run by
tests/test_small_sort.py in a SUBPROCESS: the kernel had never run on hardware when it was committed, and a device fault in
it must not poison the CUDA context of the test session.

    python tests/run_small_sort.py
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import oracle  # noqa: E402
from vren_b200 import lib as vlib  # noqa: E402


def keys_for(pattern, n, rng):
    r = rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
    if pattern == "uniform":
        return r
    if pattern == "reversed":                     # TEST(radix_sort, main): reversed iota
        return np.arange(n, dtype=np.uint32)[::-1].copy()
    if pattern == "few_values":
        return ((r % np.uint32(7)) * np.uint32(0x01010101)).astype(np.uint32)
    if pattern == "all_ones":                     # equal to the padding key
        return np.full(n, 0xFFFFFFFF, np.uint32)
    return np.where(r & np.uint32(1), np.uint32(0xFFFFFFFF), r & np.uint32(0xFF00FF00)).astype(np.uint32)


def main():
    torch.cuda.set_device(0)
    vlib.load()
    cfg = vlib.SortConfig(vlib.RANKING_AUTO, vlib.TILE_IDS_AUTO, vlib.SORT_VARIANT_SINGLE_CTA)
    rng = np.random.Generator(np.random.PCG64(4242))
    cases = 0
    for n in (1, 2, 31, 32, 33, 1000, 1023, 1024, 1025, 2048, 2049, 4096, 5000, 8191, 8192):
        for pattern in ("uniform", "reversed", "few_values", "all_ones", "half_ones"):
            k = keys_for(pattern, n, rng)
            v = np.arange(n, dtype=np.uint32)
            wk, wv = oracle.sort_pairs(k, v)
            # guard elements either side: the in-place kernel must not write outside [0, n)
            gk = torch.full((n + 8,), 0x5A5A5A5A, dtype=torch.int32, device="cuda")
            gv = torch.full((n + 8,), 0x5A5A5A5A, dtype=torch.int32, device="cuda")
            gk[4:4 + n] = torch.from_numpy(k.view(np.int32)).cuda()
            gv[4:4 + n] = torch.from_numpy(v.view(np.int32)).cuda()
            vlib.radix_sort_pairs(gk[4:4 + n], gv[4:4 + n], n, config=cfg)
            hk, hv = gk.cpu().numpy().view(np.uint32), gv.cpu().numpy().view(np.uint32)
            assert np.array_equal(hk[4:4 + n], wk) and np.array_equal(hv[4:4 + n], wv), (n, pattern, "pairs")
            assert (hk[:4] == 0x5A5A5A5A).all() and (hk[4 + n:] == 0x5A5A5A5A).all() and (hv[:4] == 0x5A5A5A5A).all() and (hv[4 + n:] == 0x5A5A5A5A).all()
            kk = torch.from_numpy(k.view(np.int32)).cuda()
            vlib.radix_sort_ex(kk, None, cfg)
            assert np.array_equal(kk.cpu().numpy().view(np.uint32), wk), (n, pattern, "keys")
            cases += 1
    # larger than one CTA's capacity: the same configuration silently takes the tiled path
    n = 8193
    k = keys_for("uniform", n, rng)
    kk = torch.from_numpy(k.view(np.int32)).cuda()
    vlib.radix_sort_ex(kk, None, cfg)
    assert np.array_equal(kk.cpu().numpy().view(np.uint32), oracle.sort_keys(k))
    torch.cuda.synchronize()
    print(f"ok {cases} cases")


if __name__ == "__main__":
    main()
