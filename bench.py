#!/usr/bin/env python
"""bench.py — headline benchmark: radix sort of uint32 key-value pairs (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--log2n 28]

A "step" is one complete sort (histogram + 4 onesweep passes) of one batch of synthetic pairs.
  value        whole-job Gpairs/s, inputs resident in HBM, device-timed with CUDA events (max over ranks)
  e2e          same metric through the host-buffer C-ABI call (H2D + sort + D2H inside the timed region)
  roofline     dominant kernel = onesweep pass: algorithmic 16 B/pair/launch over its CUDA-event duration
  cpu_baseline the oracle's restatement of the reference's CPU check (std::stable_sort by key) on a bounded sample
N>1 (torchrun): every rank sorts its own shard of n pairs (weak scaling, no data-path collective).
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


def run_reference(args):
    """--impl reference: the reference's CPU check (std::stable_sort by key; oracle port) on all host threads"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    log2_sample = min(args.log2n, 25)
    value, sec = cpu_sort_sample(log2_sample, threads, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": f"radix sort of 2^{args.log2n} uint32 key-value pairs per GPU (uniform keys, value=index)",
                   "l2": "inputs larger than L2"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"std::stable_sort by key (vren_test radix_sort.cpp:88 check, pairs extension) of 2^{log2_sample} "
                                   f"pairs per step on {threads} threads (chunk sort + parallel merges)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
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
    if args.variant is not None:
        vlib.check(lib.vrenb200_radix_sort_set_variant(args.variant), "set_variant")

    n = 1 << args.log2n
    dev = torch.device("cuda", local_rank)
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    keys0 = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
    vals0 = torch.arange(n, dtype=torch.int32, device=dev)
    keys, vals = torch.empty_like(keys0), torch.empty_like(vals0)
    sbytes = lib.vrenb200_radix_sort_scratch_bytes(n, 1)
    scratch = torch.empty(sbytes, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    prof = lib.vrenb200_sort_profile_create()
    kern_ms = (C.c_float * 6)()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(profile):
        keys.copy_(keys0)      # restore the unsorted batch (untimed; 2 GiB of traffic also evicts L2)
        vals.copy_(vals0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        vlib.check(lib.vrenb200_radix_sort_pairs_profiled(stream, keys.data_ptr(), vals.data_ptr(), n, scratch.data_ptr(),
                                                          sbytes, prof if profile else None), "radix_sort_pairs")
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
            vlib.check(lib.vrenb200_sort_profile_read(prof, kern_ms), "profile_read")
            hist_ms.append(kern_ms[0])
            pass_ms.extend(kern_ms[2:6])
        barrier()
    # correctness of the last step (cheap device-side property check; full parity lives in tests/)
    flipped = keys ^ torch.tensor(-(1 << 31), dtype=torch.int32, device=dev)
    assert bool((flipped[1:] >= flipped[:-1]).all()), "bench: output not sorted"
    assert torch.equal(keys0[vals.long()], keys), "bench: pairs broken"

    ms = sum(step_ms) / len(step_ms)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * n / (ms_max * 1e-3) / 1e9

    # ---- e2e: host buffers through the C ABI (pinned), H2D + sort + D2H timed, max over ranks -----------------
    e2e_steps = max(1, min(args.steps, 3))
    hk0 = keys0.cpu()
    hv0 = vals0.cpu()
    hk = torch.empty(n, dtype=torch.int32, pin_memory=True)
    hv = torch.empty(n, dtype=torch.int32, pin_memory=True)
    wbytes = lib.vrenb200_radix_sort_host_work_bytes(n, 1)
    del keys, vals, scratch
    work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
    e2e_ms = []
    for it in range(1 + e2e_steps):
        hk.copy_(hk0)
        hv.copy_(hv0)
        barrier()
        t0 = time.perf_counter()
        vlib.check(lib.vrenb200_radix_sort_pairs_host(stream, hk.data_ptr(), hv.data_ptr(), n, work.data_ptr(), wbytes),
                   "radix_sort_pairs_host")
        dt = (time.perf_counter() - t0) * 1e3   # the call ends with a stream sync: wall == device + copies
        if it > 0:
            e2e_ms.append(dt)
    hkn = hk.numpy().view("uint32")
    assert bool((hkn[1:] >= hkn[:-1]).all()), "bench e2e: output not sorted"
    te = torch.tensor([sum(e2e_ms) / len(e2e_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * n / (float(te.item()) * 1e-3) / 1e9

    if rank == 0:
        peak, peak_src = measured_peaks()
        pass_avg_ms = sum(pass_ms) / len(pass_ms)
        achieved = BYTES_PER_PAIR_PASS * n / (pass_avg_ms * 1e-3) / 1e9
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, sec = cpu_sort_sample(min(args.log2n, 26), 1)
            cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"std::stable_sort by key of 2^{min(args.log2n, 26)} pairs, 1 thread, {sec:.1f} s"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_max, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": f"radix sort of 2^{args.log2n} uint32 key-value pairs per GPU (uniform keys, value=index)",
                       "l2": "inputs larger than L2 (2 GiB restored between steps)",
                       "variant": lib.vrenb200_radix_sort_variant_name(args.variant or 0).decode(),
                       "parallelism": "1 GPU" if world == 1 else f"{world} independent shards, one per GPU"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "kernel": "onesweep_pass_kernel", "peak_source": peak_src,
                         "kernel_ms": pass_avg_ms, "histogram_ms": sum(hist_ms) / len(hist_ms),
                         "whole_sort_frac": BYTES_PER_PAIR_SORT * n / (ms_max * 1e-3) / 1e9 / peak},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 8 * n,
                    "ms_per_step": float(te.item())},
            "gpu_launches": 6 * args.steps,
            "clocks": clocks.summary(),
        }
        print(json.dumps(line), flush=True)
    lib.vrenb200_sort_profile_destroy(prof)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2n", type=int, default=28)
    ap.add_argument("--variant", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
