"""Several ranks of the multi-GPU sort inside ONE process on ONE device (vren_b200.dist.ShardedSort.emulated): the peers'
symmetric regions are ordinary device tensors, every rank has its own stream.  Run by tests/test_sharded_sort.py in a
subprocess, because the ranks' spin-wait kernels need every stream on its own hardware queue
(CUDA_DEVICE_MAX_CONNECTIONS, read when CUDA initialises).

    python tests/run_sharded_emulated.py <world> <rounds> <case> <key_bits> <ranking> <n0,n1,...>
"""
import os
import sys
from pathlib import Path

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import oracle  # noqa: E402
from vren_b200 import dist as vdist  # noqa: E402
from vren_b200 import lib as vlib  # noqa: E402


def make_keys(case, rank, n):
    rng = np.random.Generator(np.random.PCG64(900 + rank))
    k = rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
    if case == "uniform":
        return k
    if case == "below_2p24":
        return k & np.uint32(0x00FFFFFF)
    if case == "below_2p13":
        return k & np.uint32(0x1FFF)
    if case == "all_equal":
        return np.full(n, 0xABCD1234, np.uint32)
    if case == "three_values":       # three values of the top digit, many equal keys
        return ((rng.integers(0, 3, size=n, dtype=np.uint64) << np.uint64(30)) | rng.integers(0, 1 << 12, size=n, dtype=np.uint64)).astype(np.uint32)
    if case == "high_noise_low16":   # bucket-sort keys: only the low 16 bits count, the high half is carried
        return k
    raise SystemExit(f"unknown case {case}")


def main():
    world, rounds, case, key_bits, ranking = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4]), sys.argv[5]
    sizes = [int(x) for x in sys.argv[6].split(",")]
    assert len(sizes) == world
    torch.cuda.set_device(0)
    vlib.load()
    r = {"auto": vlib.RANKING_AUTO, "match": vlib.RANKING_MATCH, "selftest_redo": vlib.RANKING_SELFTEST_REDO}[ranking.split("+")[0]]
    t = vlib.TILE_IDS_TICKET if ranking.endswith("+ticket") else vlib.TILE_IDS_AUTO
    cfg = vlib.SortConfig(r, t, 0)
    max_n = max(max(sizes), 1)
    capacity = sum(sizes) + 257 * 12288 if case in ("all_equal", "three_values") else None     # skew: room for everything
    ctxs = vdist.ShardedSort.emulated(world, max_n, capacity, rounds, cfg)
    streams = [torch.cuda.Stream() for _ in range(world)]
    keys = [make_keys(case, rank, n) for rank, n in enumerate(sizes)]
    vals, base = [], 0
    for n in sizes:
        vals.append(np.arange(base, base + n, dtype=np.uint32))
        base += n
    dk = [torch.from_numpy(k.view(np.int32).copy()).cuda() for k in keys]
    dv = [torch.from_numpy(v.view(np.int32).copy()).cuda() for v in vals]
    all_k, all_v = np.concatenate(keys), np.concatenate(vals)
    mask = np.uint32(0xFFFFFFFF if key_bits == 32 else (1 << key_bits) - 1)
    order = np.argsort(all_k & mask, kind="stable")
    want_k, want_v = all_k[order], all_v[order]
    if key_bits == 32:
        ok, ov = oracle.sort_pairs(all_k, all_v)       # the oracle agrees with the numpy statement above
        assert np.array_equal(ok, want_k) and np.array_equal(ov, want_v)
    torch.cuda.synchronize()
    for rep in range(3):                                # the buffers, flags and epochs are reused
        for rank in range(world):
            with torch.cuda.stream(streams[rank]):
                ctxs[rank].sort(dk[rank], dv[rank], key_bits=key_bits)
        torch.cuda.synchronize()
        got_k = np.concatenate([c.result()[0].cpu().numpy().view(np.uint32) for c in ctxs])
        got_v = np.concatenate([c.result()[1].cpu().numpy().view(np.uint32) for c in ctxs])
        assert got_k.size == want_k.size, (got_k.size, want_k.size)
        assert np.array_equal(got_k, want_k), f"keys differ (rep {rep}), first at {np.nonzero(got_k != want_k)[0][:5]}"
        assert np.array_equal(got_v, want_v), f"values differ (rep {rep}), first at {np.nonzero(got_v != want_v)[0][:5]}"
        lo_hi = [c.owned_digits() for c in ctxs]
        assert lo_hi[0][0] == 0 and lo_hi[-1][1] == 256 and all(a[1] == b[0] for a, b in zip(lo_hi, lo_hi[1:]))
    print(f"ok world={world} rounds={rounds} case={case} key_bits={key_bits} ranking={ranking} sizes={sizes} pstar={lo_hi[0][2]}")


if __name__ == "__main__":
    main()
