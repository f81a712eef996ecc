"""The DEVICE plan of the multi-GPU sort (the body of csrc/sharded_sort.cu::plan_kernel, taken from that file between its
[[plan-*]] markers) executed on the host, thread for thread (tests/cpp/cta_emulator.hpp, 256 OS threads), against its numpy mirror
vren_b200.dist.exchange_plan — the mirror that tests/test_dist_cpu.py simulates the whole sort with.  So the chain
"device plan == mirror == globally stable sort" holds on CPU too, for 1-8 ranks, 1-8 rounds, skewed and empty shards, narrow key
ranges, 16-bit keys and plans that do not fit."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from vren_b200 import dist as vdist

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "build" / "emulation"


def extract_plan_sources():
    """plan_defs.inc / plan_body.inc from sharded_sort.cu (comment markers: the compiled kernel is untouched)"""
    text = (ROOT / "vren_b200" / "csrc" / "sharded_sort.cu").read_text()
    inc = OUT / "plan_inc"
    inc.mkdir(parents=True, exist_ok=True)
    first = re.search(r"^(constexpr int kMaxRanks = \d+;)[^\n]*\[\[plan-defs-a\]\]", text, flags=re.M)
    defs = re.search(r"\[\[plan-defs-b-begin\]\][^\n]*\n(.*?)\n[^\n]*\[\[plan-defs-b-end\]\]", text, flags=re.S)
    body = re.search(r"\[\[plan-body-begin\]\][^\n]*\n(.*?)\n[^\n]*\[\[plan-body-end\]\]", text, flags=re.S)
    assert first and defs and body
    defs_text = re.sub(r"^size_t sym_", "inline size_t sym_", defs.group(1), flags=re.M)
    (inc / "plan_defs.inc").write_text(first.group(1) + "\n" + defs_text + "\n")
    assert "wait_epoch(&mine->hist_ready[d], sp.epoch)" in body.group(1) and "__syncthreads_count" in body.group(1)
    (inc / "plan_body.inc").write_text(body.group(1) + "\n")
    return inc


@pytest.fixture(scope="module")
def emu():
    OUT.mkdir(parents=True, exist_ok=True)
    so = OUT / "libplan_emulation.so"
    cuda_inc = "/usr/local/cuda/include"
    if not Path(cuda_inc, "cuda_runtime.h").exists():
        pytest.skip("CUDA headers not found (host types of radix_internal.cuh)")
    inc = extract_plan_sources()
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-Wall", "-Wno-unknown-pragmas", "-Wno-attributes", "-Wno-unused-variable", "-Wno-unused-but-set-variable",
           "-Wno-unused-function", "-pthread", "-fPIC", "-shared", f"-I{cuda_inc}", f"-I{inc}", str(ROOT / "tests" / "cpp" / "plan_emulation.cpp"), "-o", str(so)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lib = C.CDLL(str(so))
    u32p, u16p = np.ctypeslib.ndpointer(np.uint32, flags="C"), np.ctypeslib.ndpointer(np.uint16, flags="C")
    lib.emu_plan.argtypes = [u32p] + [C.c_uint32] * 7 + [u32p] * 6 + [u16p, u32p]
    lib.emu_plan.restype = C.c_int
    return lib


def device_plan(lib, hists, rank, key_digits, tile, rounds, cap_tiles, round_bound):
    world = hists.shape[0]
    out = {"scalars": np.zeros(8, np.uint32), "seg": np.zeros((256, 3), np.uint32), "xfer": np.zeros((5, 256), np.uint32),
           "cum_pairs": np.zeros((rounds, 256), np.uint32), "round_digit": np.zeros(rounds + 1, np.uint32),
           "round_tile": np.zeros(rounds + 1, np.uint32), "tile_seg": np.full(cap_tiles + 1, 0xFFFF, np.uint16), "part": np.zeros(256, np.uint32)}
    rc = lib.emu_plan(np.ascontiguousarray(hists, np.uint32), world, rank, key_digits, tile, rounds, cap_tiles, round_bound,
                      out["scalars"], out["seg"], out["xfer"], out["cum_pairs"], out["round_digit"], out["round_tile"], out["tile_seg"], out["part"])
    if rc == 2:
        pytest.skip("this host cannot run the 256 threads of the emulated CTA")
    assert rc == 0
    return out


def check_against_mirror(lib, hists, key_digits, tile, rounds, cap_tiles, round_bound):
    world = hists.shape[0]
    want = vdist.exchange_plan(hists, key_digits, tile, rounds, cap_tiles, round_bound)
    p = want["pstar"]
    bounds = want["bounds"]
    for rank in range(world):
        got = device_plan(lib, hists, rank, key_digits, tile, rounds, cap_tiles, round_bound)
        pstar, error, num_tiles, out_count, lo, hi, st0, st1 = [int(x) for x in got["scalars"]]
        assert pstar == p and error == want["error"] == st0
        assert (lo, hi) == (bounds[rank], bounds[rank + 1])
        if error:
            assert st1 == 0
            continue
        assert out_count == want["out_count"][rank] == st1
        assert np.array_equal(got["xfer"][3], want["owner"])                                       # owner of every digit value
        assert np.array_equal(got["seg"][:, 0], want["first_tile"]) and np.array_equal(got["seg"][:, 1], want["seg_len"])
        assert np.array_equal(got["xfer"][2], want["dst_off"][rank])                               # where this rank's blocks land
        assert np.array_equal(got["xfer"][1], hists[rank, p])                                      # ... and how long they are
        assert np.array_equal(got["round_digit"], want["round_digit"][rank])
        # what the mirror does not model: the sender-side layout, the rounds' prefix tables and the tile -> segment map
        src_off, length, dst_off, round_of = got["xfer"][0].astype(np.int64), got["xfer"][1].astype(np.int64), got["xfer"][2].astype(np.int64), got["xfer"][4]
        assert np.array_equal(got["part"], got["xfer"][0])                                         # the partition pass scatters to src_off
        assert ((src_off - dst_off) % 4 == 0).all()                                                # 16-byte co-alignment of every block
        ends = src_off + length
        assert src_off[0] < 4 and (src_off[1:] >= ends[:-1]).all() and (src_off[1:] - ends[:-1] < 4).all()
        assert ends[-1] <= hists[rank, p].sum() + 1024                                             # kPartSlack
        rd = want["round_digit"]
        for d in range(256):
            o = int(want["owner"][d])
            assert rd[o][round_of[d]] <= d and (d < rd[o][round_of[d] + 1] or round_of[d] == rounds - 1)
        for k in range(rounds):
            assert np.array_equal(got["cum_pairs"][k], np.cumsum(np.where(round_of == k, length, 0)))
        tiles = (want["seg_len"] + tile - 1) // tile
        assert num_tiles == int(tiles[lo:hi].sum())
        ex = np.concatenate([[0], np.cumsum(tiles[lo:hi])])
        assert np.array_equal(got["round_tile"], [ex[int(x) - lo] for x in got["round_digit"]])
        want_seg = np.concatenate([np.full(int(tiles[d]), d, np.uint16) for d in range(lo, hi)] + [np.zeros(0, np.uint16)])
        assert np.array_equal(got["tile_seg"][:num_tiles], want_seg)
        outs = np.concatenate([[0], np.cumsum(want["seg_len"][lo:hi])])
        assert np.array_equal(got["seg"][lo:hi, 2], outs[:-1])


def _digit_hists(keys):
    return np.stack([np.bincount((keys >> np.uint32(8 * p)) & np.uint32(0xFF), minlength=256) for p in range(4)]).astype(np.int64)


@pytest.mark.parametrize("case", ["uniform", "skewed_half", "keys_below_2p24", "keys_below_2p13", "all_equal", "empty_ranks"])
@pytest.mark.parametrize("world,rounds", [(1, 1), (2, 3), (4, 4), (8, 2)])
def test_device_plan_equals_host_mirror_on_key_sets(emu, case, world, rounds):
    rng = np.random.Generator(np.random.PCG64(12 + world))
    tile = 64
    sizes = [int(rng.integers(500, 3000)) for _ in range(world)]
    if case == "empty_ranks" and world > 1:
        sizes[0] = 0
        sizes[-1] = 0 if world > 2 else sizes[-1]
    keys = [rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32) for n in sizes]
    if case == "skewed_half":
        keys[-1][: sizes[-1] // 2] &= np.uint32(0x03FF00FF)
    elif case == "keys_below_2p24":
        keys = [k & np.uint32(0x00FFFFFF) for k in keys]
    elif case == "keys_below_2p13":
        keys = [k & np.uint32(0x1FFF) for k in keys]
    elif case == "all_equal":
        keys = [np.full_like(k, 0xABCD1234) for k in keys]
    hists = np.stack([_digit_hists(k) for k in keys])
    total_tiles = sum(sizes) // tile + 257
    check_against_mirror(emu, hists, 4, tile, rounds, total_tiles, total_tiles)
    check_against_mirror(emu, hists, 2, tile, rounds, total_tiles, total_tiles)          # the 16-bit bucket-sort key


def test_device_plan_equals_host_mirror_on_random_histograms(emu):
    """the plan is a function of the histograms alone: random counts with holes, spikes and the production tile sizes"""
    rng = np.random.Generator(np.random.PCG64(99))
    for trial in range(12):
        world = int(rng.choice([2, 3, 5, 8]))
        rounds = int(rng.integers(1, 9))
        tile = int(rng.choice([4096, 12288]))
        h = rng.integers(0, 40000, size=(world, 4, 256)).astype(np.int64)
        h[:, :, rng.integers(0, 256, size=60)] = 0                                       # digit values nobody has
        if trial % 3 == 0:
            h[rng.integers(0, world), 3, rng.integers(0, 256)] += 3_000_000              # a hot top digit
        if trial % 4 == 1:
            h[:, 3, :] = 0
            h[:, 3, 7] = h[:, 2, :].sum(1)                                               # keys below 2^24
        total = int(h[:, 3, :].sum())
        cap_tiles = total // tile // world * 2 + 300
        check_against_mirror(emu, h, 4, tile, rounds, cap_tiles, max(cap_tiles // rounds + cap_tiles // 16 + 1, 1))


def test_device_plan_reports_what_does_not_fit(emu):
    k = np.concatenate([np.full(1000, 0x05000000, np.uint32), np.full(10, 0x06000000, np.uint32)])
    hists = np.stack([_digit_hists(k), _digit_hists(k)])
    for cap, bound, rounds, bit in ((40, 40, 1, 0), (30, 30, 1, 1), (40, 20, 2, 2)):
        got = int(device_plan(emu, hists, 0, 4, 64, rounds, cap, bound)["scalars"][1])
        assert got == vdist.exchange_plan(hists, 4, 64, rounds, cap, bound)["error"]
        assert (got & bit) == bit and (bit != 0 or got == 0)


def test_device_plan_is_race_free_under_thread_sanitizer():
    """barriers of plan_body: a ThreadSanitizer build of the same emulation (stand-alone driver) reports no data race"""
    OUT.mkdir(parents=True, exist_ok=True)
    exe = OUT / "plan_emulation_tsan"
    if not Path("/usr/local/cuda/include/cuda_runtime.h").exists():
        pytest.skip("CUDA headers not found")
    inc = extract_plan_sources()
    cmd = ["g++", "-fsanitize=thread", "-DPLAN_EMULATION_MAIN", "-std=c++17", "-O1", "-g", "-Wno-unknown-pragmas", "-Wno-attributes", "-pthread",
           "-I/usr/local/cuda/include", f"-I{inc}", str(ROOT / "tests" / "cpp" / "plan_emulation.cpp"), "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and ("tsan" in r.stderr or "sanitize" in r.stderr or "cuda_runtime.h" in r.stderr):
        pytest.skip("ThreadSanitizer runtime or CUDA headers not available")
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DONE" in r.stdout and "ThreadSanitizer" not in (r.stdout + r.stderr), (r.stdout + r.stderr)[-3000:]
