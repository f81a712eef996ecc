#!/usr/bin/env python
"""bench.py — headline benchmark: radix sort of uint32 key-value pairs (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--log2n 28]

A "step" is one complete sort of one batch of synthetic pairs (uniform keys, value = global index).
  value        whole-job Gpairs/s, inputs resident in HBM, device-timed with CUDA events (max over ranks)
  e2e          same metric through the host-buffer path (H2D + sort + D2H inside the timed region); at N > 1 through the
               SHARDED sort (host shards in, sorted shards out)
  roofline     dominant kernel = onesweep pass: algorithmic 16 B/pair/launch over its CUDA-event duration
  cpu_baseline the oracle's restatement of the reference's CPU check (std::stable_sort by key), all host threads, same size
N>1 (torchrun): ONE global sort of N x 2^log2n pairs, 2^log2n per rank (weak scaling) through the C ABI's multi-GPU sort
(vren_b200/csrc/sharded_sort.cu): device plan, local partition, per-round NVLink peer-store transfers overlapped with
segmented onesweep passes.  Before the timed steps a smaller global sort (2^22 per rank) is compared with the oracle on rank 0;
the timed output is checked by an all-reduced (key, value) multiset checksum, per-shard sortedness with value order inside
equal keys, and last-key(r) <= first-key(r+1) across ranks.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "radix_sort_pairs_throughput"
UNIT = "Gpairs/s"
BYTES_PER_PAIR_SORT = 68          # 4 (histogram read) + 4 passes x 16
BYTES_PER_PAIR_PASS = 16          # one onesweep launch: read k+v, write k+v


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """samples SM clock + throttle reasons during the timed region (NVML; nvidia-smi as a fallback)"""

    def __init__(self, index: int):
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    _REASONS = {
        0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
        0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting",
    }

    def _run(self):
        n = self._nvml
        while not self._stop.is_set():
            try:
                self.samples.append(int(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
                try:
                    r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                except Exception:
                    r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                for bit, name in self._REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def __enter__(self):
        if self._nvml is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def cpu_sort_sample(log2_sample: int, threads: int, steps: int = 1, warmup: int = 0):
    """times the oracle's stable sort by key on 2^log2_sample interleaved pairs; returns (Gpairs/s, seconds/step)"""
    import numpy as np

    import oracle

    lib = oracle.load()
    n = 1 << log2_sample
    rng = np.random.Generator(np.random.PCG64(1234))
    keys = rng.integers(0, 1 << 32, size=n, dtype=np.uint64)
    pairs0 = keys | (np.arange(n, dtype=np.uint64) << np.uint64(32))
    times = []
    for it in range(warmup + steps):
        pairs = pairs0.copy()
        t0 = time.perf_counter()
        if threads > 1:
            lib.oracle_sort_pairs_interleaved_mt(pairs, n, threads)
        else:
            lib.oracle_sort_pairs_interleaved(pairs, n)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    k = (pairs & np.uint64(0xFFFFFFFF))
    assert bool((k[1:] >= k[:-1]).all()), "oracle sort produced unsorted output"
    sec = sum(times) / len(times)
    return n / sec / 1e9, sec


def workload_string(log2n):
    return f"radix sort of 2^{log2n} uint32 key-value pairs per GPU (uniform keys, value=index)"


def config_object(log2n):
    """the same object in both arms (byte-identical): what is sorted, and how L2 is kept cold between steps"""
    return {"workload": workload_string(log2n), "l2": f"inputs larger than L2: {(8 << log2n) >> 20} MiB of pairs per GPU and step, 126 MB of L2"}


def run_reference(args):
    """--impl reference: the reference's CPU check (std::stable_sort by key; oracle port) on all host threads, on the SAME
    size as the headline configuration (2^log2n pairs per step)"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    value, sec = cpu_sort_sample(args.log2n, threads, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": config_object(args.log2n),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"std::stable_sort by key (vren_test radix_sort.cpp:88 check, pairs extension) of 2^{args.log2n} "
                                   f"pairs per step on {threads} threads (chunk sort + parallel merges); one shard's size at every N"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_secondary_baselines(small: bool = False):
    """SURVEY 8d "CPU baseline beside it" for the secondary rows: the oracle's restatements of the reference's CPU checks (the
    code vren_test compares the GPU results with), ONE host thread each as in vren_test, on bounded samples of the configurations
    the GPU rows use.  Needs no GPU.  small=True: tiny sizes (the CPU test of this function)."""
    import math

    import numpy as np

    import oracle
    from vren_b200 import synthetic

    def once(fn):
        t0 = time.perf_counter()
        r = fn()
        return time.perf_counter() - t0, r

    out = {"threads": 1, "kind": "port (oracle/), one thread as in vren_test"}
    rng = np.random.Generator(np.random.PCG64(77))
    n = 1 << (14 if small else 26)
    x = rng.integers(0, 100, size=n, dtype=np.uint64).astype(np.uint32)
    sec, _ = once(lambda: oracle.exclusive_scan(x))
    out["scan_u32"] = {"sample": f"2^{int(math.log2(n))} u32", "ms": sec * 1e3, "GB/s": 8 * n / sec / 1e9, "what": "std::exclusive_scan (blelloch_scan.cpp:122-176 check)"}
    sec, _ = once(lambda: oracle.reduce(x, n, "u32", "add"))
    out["reduce_u32_add_tree"] = {"sample": f"2^{int(math.log2(n))} u32", "ms": sec * 1e3, "GB/s": 8 * n / sec / 1e9, "what": "run_cpu_reduce, whole padded tree (reduce.cpp:72-98)"}
    n = 1 << (14 if small else 24)
    k = rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
    sec, _ = once(lambda: oracle.sort_keys(k))
    out["radix_sort_keys"] = {"sample": f"2^{int(math.log2(n))} keys", "ms": sec * 1e3, "Gkeys/s": n / sec / 1e9, "what": "std::sort (radix_sort.cpp:88 check)"}
    pairs = np.stack([rng.integers(0, 1 << 16, size=n, dtype=np.uint64).astype(np.uint32), np.arange(n, dtype=np.uint32)], axis=1)
    sec, _ = once(lambda: oracle.bucket_sort(pairs))
    out["bucket_sort_uvec2"] = {"sample": f"2^{int(math.log2(n))} pairs", "ms": sec * 1e3, "Gpairs/s": n / sec / 1e9, "what": "stable counting sort by the 16-bit key + END offsets"}
    leaves = 1 << (10 if small else 20)
    length = int(oracle.load().oracle_calc_bvh_buffer_length(leaves))
    nodes = np.zeros(length, dtype=oracle.BVH_NODE)
    nodes["min"][:leaves] = rng.random((leaves, 3), dtype=np.float32) * 100
    nodes["max"][:leaves] = nodes["min"][:leaves] + rng.random((leaves, 3), dtype=np.float32) * 10
    nodes["next"][:leaves] = 0xFFFFFFFF
    sec, _ = once(lambda: oracle.build_bvh(nodes, leaves))
    out["build_bvh"] = {"sample": f"{leaves} leaves (the GPU row's size)" if not small else f"{leaves} leaves", "ms": sec * 1e3, "what": "build_bvh.comp:32-55 level by level"}
    w, h, L = (3840, 2160, 65536) if not small else (160, 96, 300)
    depth = synthetic.depth_buffer(w, h, seed=2024)
    pos, lights = synthetic.point_lights(L, seed=2025, aspect=w / h, intensity=(1.0, 1.0))
    view = synthetic.view_matrix(0.0, 0.0, (0, 0, 0))
    cam = oracle.default_camera(w, h)
    s6, (vp, nd, pr) = once(lambda: oracle.construct_point_light_bvh(pos, lights, view))
    s7, (keys, _) = once(lambda: oracle.find_unique_clusters(depth, None, cam))
    s8, (_, _, _, total) = once(lambda: oracle.assign_lights(w, h, cam, keys, 1 << 17, nd, L, pr, vp, 1 << 23))
    out["light_assign"] = {"sample": f"one {w}x{h} view, {L} lights (the GPU row's size)" if not small else f"one {w}x{h} view, {L} lights",
                           "ms_per_view": (s6 + s7 + s8) * 1e3, "ms_light_bvh": s6 * 1e3, "ms_cluster_keys": s7 * 1e3, "ms_assign_lights": s8 * 1e3,
                           "clusters": int(keys.size), "assigned_lights": int(total)}
    return out


def secondary_metrics(lib, vlib, dev, out=None, big_log2=28):
    """the other BASELINE.json configs, each timed with CUDA events after 3 warm-ups (inputs larger than L2):
    C2 scan + reduce over 2^28 u32, C4 BuildBVH over 2^20 leaves, C5 clustered light assignment at 4K / 65 536 lights.
    big_log2: the size of the 2^28 rows (28 always, except in tests/bench_dry_run.py, which walks this code without a GPU)"""
    import math

    import numpy as np
    import torch

    from vren_b200 import synthetic
    from vren_b200.pipeline import ClusterAndShade

    peak, _ = measured_peaks()
    out = {} if out is None else out       # filled row by row: the caller's watchdog prints what is there if a later row hangs
    stream = torch.cuda.current_stream().cuda_stream

    def timed(fn, iters=20):       # >= 20 iterations after 3 warm-ups, median
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    n = 1 << big_log2
    x = torch.ones(n, dtype=torch.int32, device=dev)
    y = torch.empty_like(x)
    sb = lib.vrenb200_scan_scratch_bytes(n)
    scr = torch.empty(sb, dtype=torch.uint8, device=dev)
    ms = timed(lambda: vlib.check(lib.vrenb200_exclusive_scan_u32(stream, x.data_ptr(), y.data_ptr(), n, scr.data_ptr(), sb), "scan"))
    out["scan_u32_2p28"] = {"ms": ms, "GB/s": 8 * n / ms / 1e6, "frac_hbm": 8 * n / ms / 1e6 / peak, "bytes_per_elt": 8}
    rb = lib.vrenb200_reduce_scratch_bytes(vlib.U32, vlib.REDUCE_FINAL, n, 1)
    rscr = torch.empty(max(rb, 256), dtype=torch.uint8, device=dev)
    ms = timed(lambda: vlib.check(lib.vrenb200_reduce(stream, vlib.U32, vlib.ADD, vlib.REDUCE_FINAL, x.data_ptr(), n, y.data_ptr(), 1, rscr.data_ptr(), rb), "reduce"))
    out["reduce_u32_add_final_2p28"] = {"ms": ms, "GB/s": 4 * n / ms / 1e6, "frac_hbm": 4 * n / ms / 1e6 / peak, "bytes_per_elt": 4}
    xf = x.view(torch.float32)
    ms = timed(lambda: vlib.check(lib.vrenb200_reduce(stream, vlib.F32, vlib.ADD, vlib.REDUCE_FINAL, xf.data_ptr(), n, y.data_ptr(), 1, rscr.data_ptr(), rb), "reduce"))
    out["reduce_f32_add_final_2p28"] = {"ms": ms, "GB/s": 4 * n / ms / 1e6, "frac_hbm": 4 * n / ms / 1e6 / peak, "bytes_per_elt": 4}
    ms = timed(lambda: vlib.check(lib.vrenb200_reduce(stream, vlib.U32, vlib.ADD, vlib.REDUCE_TREE, x.data_ptr(), n, y.data_ptr(), 1, 0, 0), "reduce"))
    out["reduce_u32_add_tree_2p28"] = {"ms": ms, "GB/s": 8 * n / ms / 1e6, "frac_hbm": 8 * n / ms / 1e6 / peak, "bytes_per_elt": 8}
    del x, y, xf, scr

    # a3 as the reference ships it (keys only, 4 + 4x8 = 36 B/key) and a4 (bucket sort of uvec2, 8 + 2x16 = 40 B/pair).
    # The sorts run in place, so every timed call after the first sees sorted input; onesweep's work does not depend on
    # the key order (same histogram, same number of tiles and passes), the scatter just becomes more regular.
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    keys = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
    kb = lib.vrenb200_radix_sort_scratch_bytes(n, 0)
    kscr = torch.empty(kb, dtype=torch.uint8, device=dev)
    shuffled = keys.clone()

    def sort_keys():
        keys.copy_(shuffled)
        vlib.check(lib.vrenb200_radix_sort_keys(stream, keys.data_ptr(), n, kscr.data_ptr(), kb), "radix_sort_keys")

    def restore_only():
        keys.copy_(shuffled)

    ms = timed(sort_keys) - timed(restore_only)
    out["radix_sort_keys_2p28"] = {"ms": ms, "Gkeys/s": n / ms / 1e6, "GB/s": 36 * n / ms / 1e6, "frac_hbm": 36 * n / ms / 1e6 / peak,
                                   "bytes_per_key": 36, "note": "time of (restore + sort) minus time of restore"}
    del keys, shuffled, kscr
    # the headline sort under every ranking / tile-id configuration of vrenb200_sort_config (all of them kernels the GPU suite
    # verifies), on this box, so that the cost of the checked default is in the same line as the headline
    try:
        pk = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
        pv = torch.arange(n, dtype=torch.int32, device=dev)
        wk, wv = torch.empty_like(pk), torch.empty_like(pv)
        pb = lib.vrenb200_radix_sort_scratch_bytes(n, 1)
        pscr = torch.empty(pb, dtype=torch.uint8, device=dev)

        def restore_pairs():
            wk.copy_(pk)
            wv.copy_(pv)

        t_restore = timed(restore_pairs, iters=10)
        row = {}
        for label, ranking, tile_ids in (("sampled_ticket (default)", vlib.RANKING_ATOMIC_SAMPLED, vlib.TILE_IDS_TICKET),
                                         ("sampled_block", vlib.RANKING_ATOMIC_SAMPLED, vlib.TILE_IDS_BLOCK_INDEX),
                                         ("unverified_block", vlib.RANKING_ATOMIC_UNVERIFIED, vlib.TILE_IDS_BLOCK_INDEX),
                                         ("verified_ticket", vlib.RANKING_ATOMIC_VERIFIED, vlib.TILE_IDS_TICKET),
                                         ("match_ticket", vlib.RANKING_MATCH, vlib.TILE_IDS_TICKET)):
            cfg_r = vlib.SortConfig(ranking, tile_ids, 0)

            def sort_cfg():
                restore_pairs()
                vlib.check(lib.vrenb200_radix_sort_ex(stream, wk.data_ptr(), wv.data_ptr(), n, pscr.data_ptr(), pb, C.addressof(cfg_r), None), "radix_sort_ex")

            ms_cfg = timed(sort_cfg, iters=10) - t_restore
            row[label] = {"ms": ms_cfg, "Gpairs/s": n / ms_cfg / 1e6, "frac_hbm_68B": 68 * n / ms_cfg / 1e6 / peak}
        row["note"] = "key-value pairs, time of (restore + sort) minus time of restore; CUDA events around the whole call"
        out["radix_sort_pairs_2p28_by_config"] = row
        del pk, pv, wk, wv, pscr
    except Exception as exc:  # noqa: BLE001
        out["radix_sort_pairs_2p28_by_config"] = {"error": f"{type(exc).__name__}: {exc}"}
    # C1 (BASELINE configs[0]: the reference's own radix-sort sizes, vren_test radix_sort.cpp:82-143): launch-bound small sorts.
    # "auto" is what a caller gets (ballot-match passes below 2^21 elements: no repeat kernel behind a pass); "sampled" is the
    # large-input default forced onto the same size, for comparison
    try:
        for log2s in (10, 20):
            ns = 1 << log2s
            src = torch.randint(-(1 << 31), (1 << 31) - 1, (ns,), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
            work = src.clone()
            sb_s = lib.vrenb200_radix_sort_scratch_bytes(ns, 0)
            scr_s = torch.empty(sb_s, dtype=torch.uint8, device=dev)
            row = {}
            t_restore = timed(lambda: work.copy_(src), iters=30)
            for label, cfg_s in (("auto", vlib.SortConfig(vlib.RANKING_AUTO, vlib.TILE_IDS_AUTO, 0)),
                                 ("sampled", vlib.SortConfig(vlib.RANKING_ATOMIC_SAMPLED, vlib.TILE_IDS_AUTO, 0))):
                def sort_small():
                    work.copy_(src)
                    vlib.check(lib.vrenb200_radix_sort_ex(stream, work.data_ptr(), None, ns, scr_s.data_ptr(), sb_s, C.addressof(cfg_s), None), "radix_sort_ex")

                row[f"us_{label}"] = (timed(sort_small, iters=30) - t_restore) * 1e3
            wk = (work ^ torch.tensor(-(1 << 31), dtype=torch.int32, device=dev)).to(torch.int64)
            assert bool((wk[1:] >= wk[:-1]).all()), "bench: small sort not sorted"
            row["note"] = "keys only, one call on an idle stream; time of (restore + sort) minus time of restore"
            out[f"radix_sort_keys_2p{log2s}"] = row
            del src, work, scr_s
    except Exception as exc:  # noqa: BLE001 - rows added without a GPU at hand: they must not take the other rows with them
        out["radix_sort_keys_small"] = {"error": f"{type(exc).__name__}: {exc}"}
    nb = 1 << max(big_log2 - 2, 10)
    pairs = torch.randint(0, 1 << 16, (nb, 2), dtype=torch.int32, device=dev, generator=g)
    ob = lib.vrenb200_bucket_sort_output_bytes(nb)
    bout = torch.empty(ob, dtype=torch.uint8, device=dev)
    bb = lib.vrenb200_bucket_sort_scratch_bytes(nb)
    bscr = torch.empty(bb, dtype=torch.uint8, device=dev)
    ms = timed(lambda: vlib.check(lib.vrenb200_bucket_sort(stream, pairs.data_ptr(), nb, bout.data_ptr(), bscr.data_ptr(), bb), "bucket_sort"))
    out["bucket_sort_2p26_uvec2"] = {"ms": ms, "Gpairs/s": nb / ms / 1e6, "GB/s": 40 * nb / ms / 1e6, "frac_hbm": 40 * nb / ms / 1e6 / peak,
                                     "bytes_per_pair": 40, "frac_hbm_of_24_B_per_pair": 24 * nb / ms / 1e6 / peak,
                                     "note": "40 B/pair = what this stable two-pass sort moves (8 histogram read + 2 x 16); SURVEY 8d's 24 B/pair is the "
                                             "reference's unstable direct scatter (count + write), which a deterministic result cannot use"}
    del pairs, bout, bscr

    # C4: BuildBVH over 2^20 pre-filled leaves (33 B/leaf)
    leaves = 1 << 20
    length = lib.vrenb200_calc_bvh_buffer_length(leaves)
    nodes = torch.zeros(length * 8, dtype=torch.float32, device=dev)
    nv = nodes.view(length, 8)
    nv[:leaves, 0:3] = torch.rand(leaves, 3, device=dev) * 100
    nv[:leaves, 4:7] = nv[:leaves, 0:3] + torch.rand(leaves, 3, device=dev) * 10
    nodes.view(torch.int32).view(length, 8)[:leaves, 3] = -1
    ms = timed(lambda: vlib.check(lib.vrenb200_build_bvh(stream, nodes.data_ptr(), leaves), "build_bvh"))
    out["build_bvh_2p20_leaves"] = {"ms": ms, "GB/s": 33 * leaves / ms / 1e6, "frac_hbm": 33 * leaves / ms / 1e6 / peak, "bytes_per_leaf": 33,
                                    "note": "34.6 MB problem: L2-resident and launch-bound, not HBM-bound"}
    del nodes, nv

    # C5: one view of the clustered light assignment
    w, h, L = 3840, 2160, 65536
    depth = torch.from_numpy(synthetic.depth_buffer(w, h, seed=2024)).to(dev)
    pos, lights = synthetic.point_lights(L, seed=2025, aspect=w / h, intensity=(1.0, 1.0))
    pos, lights = torch.from_numpy(pos).to(dev), torch.from_numpy(lights).to(dev)
    view = synthetic.view_matrix(0.0, 0.0, (0, 0, 0)).tolist()
    cam = vlib.Camera(np.float32(math.radians(45.0)), np.float32(w / h), np.float32(0.01), np.float32(1000.0))
    cs = ClusterAndShade(w, h, max_point_lights=L)
    ms = timed(lambda: cs(w, h, cam, view, depth, None, pos, lights, L), iters=20)
    st = cs.status.cpu().numpy()
    graph_ms = None
    try:
        graph = cs.capture(w, h, cam, view, depth, None, pos, lights, L)
        graph_ms = timed(graph.replay, iters=20)
    except Exception as exc:  # noqa: BLE001 - the stream-launched number above stands on its own
        graph_ms = f"capture failed: {exc}"
    out["light_assign_4k_65536_lights"] = {"ms_per_view": ms, "ms_per_view_graph_replay": graph_ms, "target_ms": 0.5,
                                           "clusters": int(cs.dispatch_params[0]),
                                           "assigned_lights": int(st[0]), "node_tests": int(st[2]), "leaf_tests": int(st[3]),
                                           "stages": "construct_point_light_bvh + find_unique_cluster_list + assign_lights"}
    # SURVEY 8f rows on the same inputs: n1 per-pixel light-list walk, n2 depth-buffer pyramid
    hb = lib.vrenb200_light_list_hash_scratch_bytes(cs.max_keys)
    hscr = torch.empty(hb, dtype=torch.uint8, device=dev)
    px = torch.empty(h, w, 2, dtype=torch.int32, device=dev)
    ms = timed(lambda: vlib.check(lib.vrenb200_light_list_hash(stream, w, h, cs.cluster_ref.data_ptr(), cs.dispatch_params.data_ptr(), cs.max_keys,
                                                               cs.counts.data_ptr(), cs.offsets.data_ptr(), cs.indices.data_ptr(), cs.max_assigned,
                                                               px.data_ptr(), hscr.data_ptr(), hb), "light_list_hash"))
    out["light_list_consumer_4k"] = {"ms": ms, "GB/s": 12 * w * h / ms / 1e6, "bytes_per_px": 12}
    pyr = torch.empty(lib.vrenb200_depth_pyramid_bytes(w, h) // 4, dtype=torch.float32, device=dev)
    ms = timed(lambda: vlib.check(lib.vrenb200_depth_pyramid_build(stream, depth.data_ptr(), w, h, pyr.data_ptr()), "depth_pyramid"))
    out["depth_pyramid_4k"] = {"ms": ms, "GB/s": (4 + 16 / 3) * w * h / ms / 1e6, "bytes_per_px": 9.33,
                               "note": "77 MB problem: launch/latency-bound"}
    # n3: the producer of the light positions (65536 lights, 64 B/light: launch-bound)
    bpos = torch.cat([pos[:, :3].clone(), torch.ones(L, 1, device=dev)], 1).contiguous()
    bdir = torch.nn.functional.normalize(torch.randn(L, 3, device=dev), dim=1)
    bdir = torch.cat([bdir, torch.zeros(L, 1, device=dev)], 1).contiguous()
    ms = timed(lambda: vlib.bounce_point_lights(bpos, bdir, (-10.0, -10.0, -10.0), (10.0, 10.0, 10.0), 3.0, 0.016))
    out["bounce_point_lights_65536"] = {"ms": ms, "GB/s": 64 * L / ms / 1e6, "bytes_per_light": 64, "note": "4 MB problem: launch-bound"}
    # n4: debug lines of the 2^20-leaf BVH built above the C4 row's size (32 B read + 384 B written per node)
    leaves = 1 << 20
    levels = lib.vrenb200_calc_bvh_level_count(leaves)
    length = lib.vrenb200_calc_bvh_buffer_length(leaves)
    nodes = torch.zeros(length * 8, dtype=torch.float32, device=dev)
    nv = nodes.view(length, 8)
    nv[:leaves, 0:3] = torch.rand(leaves, 3, device=dev) * 100
    nv[:leaves, 4:7] = nv[:leaves, 0:3] + torch.rand(leaves, 3, device=dev) * 10
    nodes.view(torch.int32).view(length, 8)[:leaves, 3] = -1
    vlib.check(lib.vrenb200_build_bvh(stream, nodes.data_ptr(), leaves), "build_bvh")
    verts = torch.empty(lib.vrenb200_visualize_bvh_vertex_count(levels), 4, dtype=torch.float32, device=dev)
    ms = timed(lambda: vlib.check(lib.vrenb200_visualize_bvh(stream, nodes.data_ptr(), levels, verts.data_ptr()), "visualize_bvh"))
    out["visualize_bvh_2p20_leaves"] = {"ms": ms, "GB/s": 416 * length / ms / 1e6, "frac_hbm": 416 * length / ms / 1e6 / peak, "bytes_per_node": 416}
    # LAST (a fault here must not cost any other row): the opt-in single-CTA sort (csrc/small_sort.cu, one launch for n <= 8192), never
    # run on hardware before the round's final driver run — tests/test_small_sort.py is its parity check
    try:
        cfg_1 = vlib.SortConfig(vlib.RANKING_AUTO, vlib.TILE_IDS_AUTO, vlib.SORT_VARIANT_SINGLE_CTA)
        row = {}
        for log2s in (10, 13):
            ns = 1 << log2s
            src = torch.randint(-(1 << 31), (1 << 31) - 1, (ns,), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
            work = src.clone()
            sb_s = lib.vrenb200_radix_sort_scratch_bytes(ns, 0)
            scr_s = torch.empty(sb_s, dtype=torch.uint8, device=dev)
            t_restore = timed(lambda: work.copy_(src), iters=30)

            def sort_one_cta():
                work.copy_(src)
                vlib.check(lib.vrenb200_radix_sort_ex(stream, work.data_ptr(), None, ns, scr_s.data_ptr(), sb_s, C.addressof(cfg_1), None), "radix_sort_ex")

            us = (timed(sort_one_cta, iters=30) - t_restore) * 1e3
            want = torch.sort(src.to(torch.int64) & 0xFFFFFFFF).values
            row[f"us_2p{log2s}_keys"] = us if torch.equal(work.to(torch.int64) & 0xFFFFFFFF, want) else "WRONG RESULT"
        row["note"] = "opt-in (vrenb200_sort_config::variant = -1): all four passes in one CTA, one launch; time of (restore + sort) minus restore"
        out["radix_sort_single_cta"] = row
    except Exception as exc:  # noqa: BLE001
        out["radix_sort_single_cta"] = {"error": f"{type(exc).__name__}: {exc}"}
    # ... and the scan's opt-in safe mode (ticket tile ids, vrenb200_exclusive_scan_u32_ex): the chained register-tile kernel at
    # every size; tests/test_scan_safe_mode.py is its parity check
    try:
        n = 1 << big_log2
        x = torch.ones(n, dtype=torch.int32, device=dev)
        y = torch.empty_like(x)
        sb = lib.vrenb200_scan_scratch_bytes(n)
        scr = torch.empty(sb, dtype=torch.uint8, device=dev)
        ms = timed(lambda: vlib.check(lib.vrenb200_exclusive_scan_u32_ex(stream, x.data_ptr(), y.data_ptr(), n, 0, scr.data_ptr(), sb,
                                                                         vlib.SCAN_TILE_IDS_TICKET), "scan_ex"))
        ok = bool(torch.equal(y[-4:].cpu(), torch.arange(n - 4, n, dtype=torch.int32))) and int(y[0].item()) == 0
        out["scan_u32_2p28_safe_mode"] = {"ms": ms if ok else "WRONG RESULT", "GB/s": 8 * n / ms / 1e6, "frac_hbm": 8 * n / ms / 1e6 / peak, "bytes_per_elt": 8,
                                          "note": "opt-in: ticket tile ids, no assumption about the CTA dispatch order"}
        del x, y, scr
    except Exception as exc:  # noqa: BLE001
        out["scan_u32_2p28_safe_mode"] = {"error": f"{type(exc).__name__}: {exc}"}
    return out


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from vren_b200 import lib as vlib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the vren_b200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = vlib.load()
    ranking = {"auto": vlib.RANKING_AUTO, "match": vlib.RANKING_MATCH, "verified": vlib.RANKING_ATOMIC_VERIFIED,
               "sampled": vlib.RANKING_ATOMIC_SAMPLED, "unverified": vlib.RANKING_ATOMIC_UNVERIFIED}[args.ranking]
    tile_ids = {"auto": vlib.TILE_IDS_AUTO, "block": vlib.TILE_IDS_BLOCK_INDEX, "ticket": vlib.TILE_IDS_TICKET}[args.tile_ids]
    cfg = vlib.SortConfig(ranking, tile_ids, args.variant or 0)
    cfg_ptr = C.addressof(cfg)

    n = 1 << args.log2n
    dev = torch.device("cuda", local_rank)
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    keys0 = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
    vals0 = torch.arange(rank * n, (rank + 1) * n, dtype=torch.int64, device=dev).to(torch.int32)      # global index: unique, rank-major
    stream = torch.cuda.current_stream().cuda_stream
    prof = lib.vrenb200_sort_profile_create()
    kern_ms = (C.c_float * 6)()
    sign = torch.tensor(-(1 << 31), dtype=torch.int32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def checksum(k, v):
        """order-independent 64-bit fingerprint of a multiset of pairs (sum of a mix of every pair, mod 2^63)"""
        x = (k.to(torch.int64) & 0xFFFFFFFF) | ((v.to(torch.int64) & 0xFFFFFFFF) << 32)
        x = (x ^ (x >> 31)) * 0x2545F4914F6CDD1D
        x = x ^ (x >> 29)
        return torch.stack([(x & 0x7FFFFFFF).sum(), ((x >> 31) & 0x7FFFFFFF).sum(), torch.tensor(k.numel(), dtype=torch.int64, device=k.device)])

    def verify_sorted_shard(k, v, what):
        """keys ascending (unsigned); inside equal keys the values (global indices) ascending = stable"""
        if k.numel() < 2:
            return
        kk = (k ^ sign).to(torch.int64)
        vv = v.to(torch.int64) & 0xFFFFFFFF
        dk = kk[1:] - kk[:-1]
        assert bool((dk >= 0).all()), f"bench: {what}: keys not sorted"
        assert bool(((dk > 0) | (vv[1:] > vv[:-1])).all()), f"bench: {what}: equal keys out of input order (not stable)"

    sorter = None
    single = None
    if world == 1:
        keys, vals = torch.empty_like(keys0), torch.empty_like(vals0)
        sbytes = lib.vrenb200_radix_sort_scratch_bytes(n, 1)
        scratch = torch.empty(sbytes, dtype=torch.uint8, device=dev)
    else:
        from vren_b200 import dist as vdist

        # ---- parity of the multi-GPU path against the oracle before anything is timed (2^22 pairs per rank, rank 0 checks)
        import oracle

        nv = 1 << 22
        small = vdist.ShardedSort.for_process_group(nv, None, 2, config=cfg)
        for rep in range(2):
            small.sort(keys0[:nv], vals0[:nv])
        torch.cuda.synchronize()
        sk, sv = small.result()
        sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([sk.numel()], dtype=torch.int64, device=dev))
        sizes = [int(t.item()) for t in sizes]
        pad = max(sizes)
        gk = [torch.zeros(pad, dtype=torch.int32, device=dev) for _ in range(world)]
        gv = [torch.zeros(pad, dtype=torch.int32, device=dev) for _ in range(world)]
        ik = [torch.zeros(nv, dtype=torch.int32, device=dev) for _ in range(world)]
        iv = [torch.zeros(nv, dtype=torch.int32, device=dev) for _ in range(world)]
        dist.all_gather(gk, torch.nn.functional.pad(sk, (0, pad - sk.numel())))
        dist.all_gather(gv, torch.nn.functional.pad(sv, (0, pad - sv.numel())))
        dist.all_gather(ik, keys0[:nv].contiguous())
        dist.all_gather(iv, vals0[:nv].contiguous())
        if rank == 0:
            got_k = np.concatenate([gk[r][:sizes[r]].cpu().numpy().view(np.uint32) for r in range(world)])
            got_v = np.concatenate([gv[r][:sizes[r]].cpu().numpy().view(np.uint32) for r in range(world)])
            wk, wv = oracle.sort_pairs(np.concatenate([t.cpu().numpy().view(np.uint32) for t in ik]),
                                       np.concatenate([t.cpu().numpy().view(np.uint32) for t in iv]))
            assert np.array_equal(got_k, wk) and np.array_equal(got_v, wv), "bench: sharded sort differs from the oracle"
        del gk, gv, ik, iv
        small.close()
        del small
        torch.cuda.empty_cache()
        sorter = vdist.ShardedSort.for_process_group(n, None, args.rounds, config=cfg)
    in_sum = checksum(keys0, vals0)
    if world > 1:
        dist.all_reduce(in_sum)

    def one_step(profile):
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world == 1:
            keys.copy_(keys0)      # restore the unsorted batch (untimed; 2 GiB of traffic also evicts L2)
            vals.copy_(vals0)
            e0.record()
            vlib.check(lib.vrenb200_radix_sort_ex(stream, keys.data_ptr(), vals.data_ptr(), n, scratch.data_ptr(), sbytes, cfg_ptr,
                                                  prof if profile else None), "radix_sort_ex")
        else:
            e0.record()
            sorter.sort(keys0, vals0)          # out of place: the input shard stays as it is; 2 GiB >> L2 between steps
        e1.record()
        return e0, e1

    for _ in range(max(args.warmup, 3)):
        one_step(False)
    barrier()
    step_ms, pass_ms, hist_ms = [], [], []
    with ClockSampler(local_rank) as clocks:
        for _ in range(args.steps):
            e0, e1 = one_step(True)
            e1.synchronize()
            step_ms.append(e0.elapsed_time(e1))
            if world == 1:
                vlib.check(lib.vrenb200_sort_profile_read(prof, kern_ms), "profile_read")
                hist_ms.append(kern_ms[0])
                pass_ms.extend(kern_ms[2:6])
        barrier()
    # ---- correctness of the last timed step
    violation = 0
    phases = None
    if world == 1:
        verify_sorted_shard(keys, vals, "output")
        assert torch.equal(keys0[(vals.to(torch.int64) & 0xFFFFFFFF)], keys), "bench: pairs broken"
        word = lib.vrenb200_radix_sort_violation_word(scratch.data_ptr(), n, 1) - scratch.data_ptr()
        violation = int(scratch[word: word + 4].view(torch.int32).item())
        out_sum = checksum(keys, vals)
    else:
        phases = sorter.phases()
        ok, ov = sorter.result()
        verify_sorted_shard(ok, ov, f"shard of rank {rank}")
        out_sum = checksum(ok, ov)
        dist.all_reduce(out_sum)
        # last key of every rank <= first key of the next non-empty rank (unsigned), ties in rank-major value order
        edge = torch.full((4,), -1, dtype=torch.int64, device=dev)
        if ok.numel():
            edge = torch.stack([(ok[0] ^ sign).to(torch.int64), ov[0].to(torch.int64) & 0xFFFFFFFF,
                                (ok[-1] ^ sign).to(torch.int64), ov[-1].to(torch.int64) & 0xFFFFFFFF])
        edges = [torch.zeros_like(edge) for _ in range(world)]
        dist.all_gather(edges, edge)
        nonempty = [e.tolist() for e in edges if int(e[1]) >= 0]
        for a, b in zip(nonempty, nonempty[1:]):
            assert (a[2], a[3]) < (b[0], b[1]), "bench: shards overlap / out of order across ranks"
    assert torch.equal(in_sum, out_sum), "bench: the output is not a permutation of the input pairs"

    if world == 1:
        # one profiled run per kernel family for the roofline object happened above; nothing more to do
        pass
    else:
        # the dominant kernel of the multi-GPU sort is the same pass kernel: profile one local sort for the roofline object
        keys, vals = keys0.clone(), vals0.clone()
        sbytes = lib.vrenb200_radix_sort_scratch_bytes(n, 1)
        scratch = torch.empty(sbytes, dtype=torch.uint8, device=dev)
        for _ in range(3):
            keys.copy_(keys0); vals.copy_(vals0)
            vlib.check(lib.vrenb200_radix_sort_ex(stream, keys.data_ptr(), vals.data_ptr(), n, scratch.data_ptr(), sbytes, cfg_ptr, prof), "radix_sort_ex")
            torch.cuda.synchronize()
            vlib.check(lib.vrenb200_sort_profile_read(prof, kern_ms), "profile_read")
            hist_ms.append(kern_ms[0])
            pass_ms.extend(kern_ms[2:6])
        del keys, vals, scratch

    ms = sum(step_ms) / len(step_ms)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * n / (ms_max * 1e-3) / 1e9

    # ---- e2e: HOST buffers in, HOST buffers out, H2D + sort + D2H timed (wall clock around a drained pipeline), max over ranks
    e2e_steps = max(2, min(args.steps, 12))
    hk_in = torch.empty(n, dtype=torch.int32, pin_memory=True)
    hv_in = torch.empty(n, dtype=torch.int32, pin_memory=True)
    hk_in.copy_(keys0)
    hv_in.copy_(vals0)
    if world == 1:
        # consecutive steps are issued on two streams (two device work buffers, vrenb200_radix_sort_pairs_host_async), so step
        # i+1 uploads while step i downloads; a single call is upload -> sort -> download, one PCIe direction at a time
        del keys, vals, scratch
        wbytes = lib.vrenb200_radix_sort_host_work_bytes(n, 1)
        lanes = [{"stream": torch.cuda.Stream(device=dev), "hk": torch.empty(n, dtype=torch.int32, pin_memory=True),
                  "hv": torch.empty(n, dtype=torch.int32, pin_memory=True), "work": torch.empty(wbytes, dtype=torch.uint8, device=dev)}
                 for _ in range(2)]
        torch.cuda.synchronize()

        def issue(i, k=2):
            lane = lanes[i % k]
            vlib.check(lib.vrenb200_radix_sort_pairs_host_async(lane["stream"].cuda_stream, hk_in.data_ptr(), hv_in.data_ptr(), lane["hk"].data_ptr(),
                                                                lane["hv"].data_ptr(), n, lane["work"].data_ptr(), wbytes), "radix_sort_pairs_host_async")

        def drain():
            for lane in lanes:
                lane["stream"].synchronize()

        issue(0); issue(1); drain()                      # warm-up
        t0 = time.perf_counter()
        issue(0); drain()
        single_ms = (time.perf_counter() - t0) * 1e3     # one call alone: upload, sort, download back to back
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            issue(i)
        drain()
        e2e_step_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        for lane in lanes:
            hkn = lane["hk"].numpy().view("uint32")
            assert bool((hkn[1:] >= hkn[:-1]).all()), "bench e2e: output not sorted"
        assert bool(torch.equal(hk_in[(lanes[0]["hv"].to(torch.int64) & 0xFFFFFFFF) - rank * n], lanes[0]["hk"])), "bench e2e: pairs broken"
        d2h_bytes = 8 * n
        e2e_note = "consecutive steps alternate between two streams / device work buffers (upload of step i+1 overlaps download of step i)"
        # the same with a third lane (the upload of step i+2 can start while step i still downloads): taken if it is faster.  Added
        # after the round's GPU time was spent — same call, same checks; the two-lane number stands if anything goes wrong
        try:
            lanes.append({"stream": torch.cuda.Stream(device=dev), "hk": torch.empty(n, dtype=torch.int32, pin_memory=True),
                          "hv": torch.empty(n, dtype=torch.int32, pin_memory=True), "work": torch.empty(wbytes, dtype=torch.uint8, device=dev)})
            torch.cuda.synchronize()
            for i in range(3):
                issue(i, 3)
            drain()
            t0 = time.perf_counter()
            for i in range(e2e_steps):
                issue(i, 3)
            drain()
            three_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
            hkn = lanes[2]["hk"].numpy().view("uint32")
            assert bool((hkn[1:] >= hkn[:-1]).all()) and bool(torch.equal(lanes[2]["hk"], lanes[0]["hk"])), "third lane: wrong output"
            if three_ms < e2e_step_ms:
                e2e_note = (f"consecutive steps rotate over three streams / device work buffers (uploads overlap downloads); "
                            f"with two: {e2e_step_ms:.1f} ms per step")
                e2e_step_ms = three_ms
            else:
                e2e_note += f"; a third lane is not faster ({three_ms:.1f} ms per step)"
        except Exception as exc:  # noqa: BLE001
            e2e_note += f" (three lanes not measured: {type(exc).__name__}: {exc})"
        del lanes
    else:
        # the sharded sort with host shards: every step uploads the rank's 2 x 4n input bytes, runs the collective sort and
        # downloads the rank's sorted shard (its share of the global result)
        dk, dv = torch.empty_like(keys0), torch.empty_like(vals0)
        hk_out = torch.empty(sorter.capacity, dtype=torch.int32, pin_memory=True)
        hv_out = torch.empty(sorter.capacity, dtype=torch.int32, pin_memory=True)
        ok, _ = sorter.result()
        m = ok.numel()                                    # same keys every step: the shard size does not change

        def e2e_step():
            dk.copy_(hk_in, non_blocking=True)
            dv.copy_(hv_in, non_blocking=True)
            sorter.sort(dk, dv)
            hk_out[:m].copy_(sorter.out_keys[:m], non_blocking=True)
            hv_out[:m].copy_(sorter.out_vals[:m], non_blocking=True)

        e2e_step(); barrier()
        t0 = time.perf_counter()
        e2e_step(); torch.cuda.synchronize()
        single_ms = (time.perf_counter() - t0) * 1e3
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        e2e_step_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        hkn = hk_out[:m].numpy().view("uint32")
        assert bool((hkn[1:] >= hkn[:-1]).all()), "bench e2e: shard not sorted"
        d2h_bytes = 8 * m
        e2e_note = "sharded sort with host shards: H2D of the rank's shard, collective sort, D2H of the rank's sorted shard"
        e2e_serial_ms = e2e_step_ms
        # the same steps with the upload of step i+1 on a second stream and a second input buffer, so that it overlaps the
        # download of step i (PCIe is full duplex; the N = 1 leg does the same with two lanes).  Added after the round's GPU time
        # was spent: if anything goes wrong here the serial number above stands
        try:
            up = torch.cuda.Stream(device=dev)
            main_stream = torch.cuda.current_stream()
            bufs = [(dk, dv), (torch.empty_like(keys0), torch.empty_like(vals0))]
            uploaded = [torch.cuda.Event(), torch.cuda.Event()]
            consumed = [torch.cuda.Event(), torch.cuda.Event()]

            def e2e_pipelined(count):
                for i in range(count):
                    bk, bv = bufs[i % 2]
                    with torch.cuda.stream(up):
                        if i >= 2:
                            up.wait_event(consumed[i % 2])      # the sort that read this buffer two steps ago has finished
                        bk.copy_(hk_in, non_blocking=True)
                        bv.copy_(hv_in, non_blocking=True)
                        uploaded[i % 2].record(up)
                    main_stream.wait_event(uploaded[i % 2])
                    sorter.sort(bk, bv)
                    consumed[i % 2].record(main_stream)
                    hk_out[:m].copy_(sorter.out_keys[:m], non_blocking=True)    # on the sort's stream: the next sort starts after it
                    hv_out[:m].copy_(sorter.out_vals[:m], non_blocking=True)

            up.wait_stream(main_stream)
            e2e_pipelined(2); barrier()
            hk_out.zero_()
            t0 = time.perf_counter()
            e2e_pipelined(e2e_steps)
            torch.cuda.synchronize()
            piped_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
            hkn = hk_out[:m].numpy().view("uint32")
            piped_ok = bool((hkn[1:] >= hkn[:-1]).all()) and bool(hkn.any())
        except Exception as exc:  # noqa: BLE001
            piped_ms, piped_ok = None, False
            e2e_note += f" (pipelined form failed: {type(exc).__name__}: {exc})"
        # every rank must take the same branch: agree on the outcome
        agree = torch.tensor([1.0 if piped_ok else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN)
        worst = torch.tensor([piped_ms if piped_ok else 0.0, e2e_serial_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        if float(agree[0].item()) == 1.0 and float(worst[0].item()) < float(worst[1].item()):
            e2e_step_ms = piped_ms
            e2e_note = ("sharded sort with host shards, steps pipelined: the H2D of step i+1 (second stream, second input buffer) overlaps "
                        f"the D2H of step i; one step after the other: {e2e_serial_ms:.1f} ms on this rank")
    te = torch.tensor([e2e_step_ms, single_ms, float(d2h_bytes)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    te = [float(x) for x in te.tolist()]          # host numbers from here on: nothing below may need the device to print the line
    e2e_value = world * n / (te[0] * 1e-3) / 1e9

    # ---- the line: everything it needs is a host value from here on, so it can be printed whatever the secondary rows do
    make_line = None
    if rank == 0:
        peak, peak_src = measured_peaks()
        pass_avg_ms = sum(pass_ms) / len(pass_ms)
        variant_name = lib.vrenb200_radix_sort_selected_variant_name(n, 1, cfg_ptr).decode()
        pass_kernel = "onesweep_pass_kernel"
        achieved = BYTES_PER_PAIR_PASS * n / (pass_avg_ms * 1e-3) / 1e9
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            try:
                v, sec = cpu_sort_sample(args.log2n, threads)
                cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                       "sample": f"std::stable_sort by key of 2^{args.log2n} pairs (the whole step), {threads} threads, {sec:.1f} s"}
            except MemoryError:     # a host with less than ~8 GB free: a quarter of the step
                v, sec = cpu_sort_sample(args.log2n - 2, threads)
                cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                       "sample": f"std::stable_sort by key of 2^{args.log2n - 2} pairs (host memory too small for the whole step), {threads} threads, {sec:.1f} s"}
        if cpu is not None:
            # SURVEY 8d asks for the per-core figure beside the all-cores one: std::stable_sort on ONE thread (as vren_test runs its
            # check), on a 2^24-pair sample so that it takes a couple of seconds (n log n: the full step is ~15 % slower per pair)
            try:
                v1, sec1 = cpu_sort_sample(min(args.log2n, 24), 1)
                cpu["single_thread"] = {"value": v1, "unit": UNIT, "cores": 1, "sample": f"2^{min(args.log2n, 24)} pairs, {sec1:.1f} s"}
            except Exception as exc:  # noqa: BLE001
                cpu["single_thread"] = {"error": f"{type(exc).__name__}: {exc}"}
        secondary_cpu = None
        if world == 1 and not args.no_cpu_baseline and not args.no_secondary:
            try:
                secondary_cpu = cpu_secondary_baselines()
            except Exception as exc:  # noqa: BLE001
                secondary_cpu = {"error": f"{type(exc).__name__}: {exc}"}
        traffic = ncu_traffic(pass_kernel, args.log2n)
        passes = 4
        launches_single = 2 + 2 * passes                                   # histogram, offsets, 4 passes + their (idle) redo kernels
        launches_multi = 4 + 2 + 1 + args.rounds * (1 + 3 + 6) + 2         # hist, publish, offsets, plan, partition + redo, wait, per round: transfer + wait + segment histograms + scan + 3 passes + 3 redos, compact, done
        clock_summary = clocks.summary()

        def make_line(secondary):
            return {
                "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_max, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u32", "data": "synthetic",
                "config": config_object(args.log2n),
                "details": {"variant": variant_name,
                            "ranking_check_failures": violation,
                            "parallelism": "1 GPU" if world == 1 else
                            f"one global sort of {world}x2^{args.log2n} pairs: device plan, local partition by the top digit, {args.rounds} rounds of "
                            "NVLink peer-store transfers overlapped with segmented 3-pass onesweep of what has arrived (no NCCL on the data path)"},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic.get("dram_bytes_per_launch") if traffic else None,
                             "traffic_source": traffic, "kernel": pass_kernel, "peak_source": peak_src,
                             "kernel_ms": pass_avg_ms, "histogram_ms": sum(hist_ms) / len(hist_ms),
                             "whole_sort_frac": BYTES_PER_PAIR_SORT * n / (ms_max * 1e-3) / 1e9 / peak},
                "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": int(te[2]),
                        "ms_per_step": te[0], "steps": e2e_steps, "issue": e2e_note,
                        "single_call_ms": te[1], "single_call_value": world * n / (te[1] * 1e-3) / 1e9},
                "gpu_launches": (launches_single if world == 1 else launches_multi) * args.steps,
                "verification": "sortedness + stability + pair integrity" if world == 1 else
                                "2^22/rank global sort == oracle.sort_pairs; timed output: all-reduced pair-multiset checksum, per-shard sortedness and stability, cross-rank boundary order",
                "phases_rank0_last_step": phases,
                "secondary": secondary,
                "secondary_cpu_baseline": secondary_cpu,
                "clocks": clock_summary,
            }

    # ---- secondary rows (the other BASELINE configs).  They come after the headline has been measured and must not be able to
    # lose it: an exception becomes an "error" entry, and a watchdog on every rank prints the line without them and ends the
    # process if they have not finished in time (a rank that failed alone would leave the others waiting in a collective)
    secondary = None
    if not args.no_secondary:
        finished = threading.Event()
        rows = {}            # filled row by row by secondary_metrics*

        def bail():
            if finished.is_set():
                return
            if rank == 0:
                partial = dict(rows)
                partial["error"] = (f"the secondary rows did not finish within {args.secondary_timeout} s and were abandoned after the rows above; "
                                    "every other entry of this line was measured before they started")
                print(json.dumps(make_line(partial)), flush=True)
            os._exit(0)

        watchdog = threading.Timer(args.secondary_timeout, bail)
        watchdog.daemon = True
        watchdog.start()
        torch.cuda.empty_cache()
        try:
            secondary = secondary_metrics(lib, vlib, dev, rows) if world == 1 else secondary_metrics_multi(lib, vlib, dev, sorter, rank, world, rows)
        except Exception as exc:  # noqa: BLE001
            secondary = dict(rows)
            secondary["error"] = f"{type(exc).__name__}: {exc} (rows above were measured before the failure)"
        finished.set()
        watchdog.cancel()
    if rank == 0:
        print(json.dumps(make_line(secondary)), flush=True)
    lib.vrenb200_sort_profile_destroy(prof)
    if world > 1:
        sorter.close()
        dist.destroy_process_group()


def secondary_metrics_multi(lib, vlib, dev, sorter, rank, world, out=None, big_log2=28):
    """N > 1: the other sharded paths of SURVEY 8e, device-timed (max over ranks): sharded exclusive scan and reduce over
    2^28 u32 per GPU, sharded bucket sort (16-bit key) of 2^26 pairs per GPU, clustered shading of 8 views at 4K over the ranks"""
    import math

    import numpy as np
    import torch
    import torch.distributed as dist

    from vren_b200 import dist as vdist
    from vren_b200 import synthetic
    from vren_b200.pipeline import ViewBatch

    peak, _ = measured_peaks()
    out = {} if out is None else out

    def timed(fn, iters=5):
        ts = []
        for i in range(iters + 2):
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); e1.synchronize()
            if i >= 2:
                ts.append(e0.elapsed_time(e1))
        t = torch.tensor([float(np.median(ts))], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = 1 << big_log2                      # 28 always, except in tests/bench_dry_run.py
    ops = vdist.CudaOps()
    x = torch.ones(n, dtype=torch.int32, device=dev)
    ms = timed(lambda: vdist.sharded_exclusive_scan(x, ops=ops))
    out["sharded_scan_u32_2p28_per_gpu"] = {"ms": ms, "GB/s_per_gpu": 12 * n / ms / 1e6, "frac_hbm": 12 * n / ms / 1e6 / peak, "bytes_per_elt": 12,
                                            "note": "local reduce (4 B) + all_gather of the partial sums + local scan with base (8 B); host reads the partials"}
    ms = timed(lambda: vdist.sharded_reduce_add(x, ops=ops))
    out["sharded_reduce_u32_2p28_per_gpu"] = {"ms": ms, "GB/s_per_gpu": 4 * n / ms / 1e6, "frac_hbm": 4 * n / ms / 1e6 / peak, "bytes_per_elt": 4}
    del x
    nb = 1 << max(big_log2 - 2, 10)
    g = torch.Generator(device=dev)
    g.manual_seed(77 + rank)
    bk = torch.randint(0, 1 << 16, (nb,), dtype=torch.int32, device=dev, generator=g)
    bv = torch.arange(nb, dtype=torch.int32, device=dev)
    ms = timed(lambda: sorter.sort(bk, bv, key_bits=16))
    out["sharded_bucket_sort_2p26_per_gpu"] = {"ms": ms, "Gpairs/s": world * nb / ms / 1e6,
                                               "note": "16-bit key: partition digit = byte 1, one segmented pass; END offsets (all_reduce of 65536 counters) not timed"}
    del bk, bv
    # C5 batched: 8 views (yaw += 45 degrees) of a 3840x2160 frame, 65 536 lights broadcast once, one view per rank
    w, h, L, views = 3840, 2160, 65536, 8
    vb = ViewBatch(w, h, L)
    pos = torch.zeros(L, 4, dtype=torch.float32, device=dev)
    lights = torch.zeros(L, 4, dtype=torch.float32, device=dev)
    if rank == 0:
        p0, l0 = synthetic.point_lights(L, seed=2025, aspect=w / h, intensity=(1.0, 1.0))
        pos, lights = torch.from_numpy(p0).to(dev), torch.from_numpy(l0).to(dev)
    cam = vlib.Camera(np.float32(math.radians(45.0)), np.float32(w / h), np.float32(0.01), np.float32(1000.0))
    frames = {v: (cam, synthetic.view_matrix(math.radians(45.0) * v, 0.0, (0, 0, 0)).tolist(),
                  torch.from_numpy(synthetic.depth_buffer(w, h, seed=2024 + v)).to(dev), None) for v in vb.my_views(views)}
    res = {}

    def frame():
        vb.set_lights(pos, lights, L)
        res.update(vb(frames))

    ms = timed(frame)
    mine = {v: [int(t) for t in r.tolist()] for v, r in res.items()}
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    per_view = {}
    for d in gathered:
        per_view.update(d)
    out["light_assign_4k_65536_lights_8_views"] = {"ms_per_frame": ms, "ms_per_view": ms / views, "views": views,
                                                   "views_per_rank": len(frames), "target_ms_per_view": 0.5,
                                                   "clusters_per_view": [per_view[v][0] for v in sorted(per_view)],
                                                   "assigned_lights_per_view": [per_view[v][4] for v in sorted(per_view)],
                                                   "note": "lights broadcast once per frame (NCCL, inside the timed region), view v on rank v % world"}
    # LAST: the headline multi-GPU sort cut into other numbers of rounds (each of them ran on hardware in round 2: profiles/r2e_, r2g_,
    # r2j_), on this box: how much the overlap of transfers and passes is worth at this N
    try:
        ns = 1 << big_log2
        g2 = torch.Generator(device=dev)
        g2.manual_seed(4321 + rank)
        sk = torch.randint(-(1 << 31), (1 << 31) - 1, (ns,), dtype=torch.int64, device=dev, generator=g2).to(torch.int32)
        sv = torch.arange(ns, dtype=torch.int32, device=dev)
        row = {}
        for rounds in (1, 2, 4, 8):
            ctx = vdist.ShardedSort.for_process_group(ns, None, rounds)
            ms = timed(lambda: ctx.sort(sk, sv), iters=3)
            torch.cuda.synchronize()
            ctx.result()                         # raises (on every rank alike) if the plan did not fit
            row[f"rounds_{rounds}"] = {"ms": ms, "Gpairs/s": world * ns / ms / 1e6, "phases_this_rank_last_sort": ctx.phases()}
            ctx.close()
            del ctx
            torch.cuda.empty_cache()
        out["sharded_sort_by_rounds"] = row
        del sk, sv
    except Exception as exc:  # noqa: BLE001
        out["sharded_sort_by_rounds"] = {"error": f"{type(exc).__name__}: {exc}"}
    return out


def ncu_traffic(kernel, log2n):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu --set full capture
    (profiles/ncu_traffic.json records the commit the capture was taken at)."""
    try:
        rec = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic.json")))[kernel]
        return rec if rec["log2n"] == log2n else None
    except (OSError, KeyError, ValueError):
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2n", type=int, default=28)
    ap.add_argument("--variant", type=int, default=None, help="1-based entry of the kernel table (vrenb200_sort_config::variant)")
    ap.add_argument("--ranking", default="auto", choices=["auto", "match", "verified", "sampled", "unverified"])
    ap.add_argument("--tile-ids", default="auto", choices=["auto", "block", "ticket"])
    ap.add_argument("--rounds", type=int, default=None,
                    help="N>1: pieces the exchange is cut into (transfer of one overlaps the sorting of the previous); default 2 at N = 2, else 4 "
                         "(measured: profiles/r2g_bench_n2_tma.log, r2j_bench_n4_exclusive.log, r2h_bench_n8_tma.log)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--secondary-timeout", type=float, default=300.0, help="seconds the secondary rows may take before the line is printed without them")
    args = ap.parse_args()
    if args.rounds is None:
        args.rounds = 2 if int(os.environ.get("WORLD_SIZE", "1")) == 2 else 4
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
