"""Walks bench.py's GPU arm WITHOUT a GPU: every tensor is placed on the CPU (a TorchFunctionMode rewrites cuda devices and drops
pin_memory), torch.cuda's streams / events / synchronisation are stand-ins, and the device entry points of the C ABI are replaced
by numpy fakes working on the same pointers (sorts sort, scans scan, everything else is a no-op); the host-only entry points
(sizes, names, the violation-word address) are the real library's.  Nothing here measures anything — it exists so that the
Python control flow of the bench (the timed loop, the e2e legs, the secondary rows added after the round's GPU time was spent, the
assembly of the JSON line, the watchdog) is executed before the driver executes it on the GPU box.

    python tests/bench_dry_run.py [bench.py arguments]      prints bench.py's line; run by tests/test_bench_contract.py
    python tests/bench_dry_run.py --script tests/run_small_sort.py      walks another GPU script the same way (exact fakes at any size)
"""
import contextlib
import ctypes as C
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402
from torch.overrides import TorchFunctionMode  # noqa: E402

CPU = torch.device("cpu")
EXACT_BELOW = 1 << 21        # the fakes really sort / scan below this many elements (--script: always)


def is_cuda(d):
    return (isinstance(d, str) and d.startswith("cuda")) or (isinstance(d, torch.device) and d.type == "cuda")


class OnCpu(TorchFunctionMode):
    def __torch_function__(self, func, types, args=(), kwargs=None):
        kwargs = dict(kwargs or {})
        if is_cuda(kwargs.get("device")):
            kwargs["device"] = CPU
        kwargs.pop("pin_memory", None) if kwargs.get("pin_memory") else None
        name = getattr(func, "__name__", "")
        if name == "cuda" and args and isinstance(args[0], torch.Tensor):
            return args[0]
        if name == "to" and args and isinstance(args[0], torch.Tensor):
            args = tuple(CPU if is_cuda(a) else a for a in args)
        if name == "copy_":
            kwargs.pop("non_blocking", None)
        return func(*args, **kwargs)


class FakeStream:
    cuda_stream = 0

    def synchronize(self): pass
    def wait_event(self, e): pass
    def wait_stream(self, s): pass


class FakeEvent:
    def __init__(self, enable_timing=False): self.t = 0.0
    def record(self, stream=None): self.t = time.perf_counter()
    def synchronize(self): pass
    def elapsed_time(self, other): return max((other.t - self.t) * 1e3, 1e-3)


def patch_torch_cuda():
    tc = torch.cuda
    tc.is_available = lambda: True
    tc.set_device = lambda *a, **k: None
    tc.synchronize = lambda *a, **k: None
    tc.empty_cache = lambda: None
    tc.current_device = lambda: 0
    tc.device_count = lambda: 1
    stream = FakeStream()
    tc.current_stream = lambda *a, **k: stream
    tc.Stream = lambda *a, **k: FakeStream()
    tc.Event = FakeEvent
    tc.stream = lambda s: contextlib.nullcontext()
    real_generator = torch.Generator
    torch.Generator = lambda device=None: real_generator(device=CPU if device is None or is_cuda(device) else device)


def u32(ptr, n):
    return np.ctypeslib.as_array((C.c_uint32 * n).from_address(ptr)) if n else np.zeros(0, np.uint32)


class FakeLib:
    """device entry points as numpy fakes on the callers' pointers; everything else falls through to the real library"""

    def __init__(self, real):
        self._real = real

    def __getattr__(self, name):
        return getattr(self._real, name)

    # ---- sorts (small ones really sort; the 2^28-sized secondary rows are not worth minutes of numpy)
    def _sort(self, keys, vals, n):
        if n > EXACT_BELOW:
            return
        k = u32(keys, n)
        if vals:
            v = u32(vals, n)
            order = np.argsort(k, kind="stable")
            k[:], v[:] = k[order], v[order]
        else:
            k.sort()

    def vrenb200_radix_sort_ex(self, stream, keys, vals, n, scratch, nbytes, cfg, prof):
        self._sort(keys, vals, n)
        word = self._real.vrenb200_radix_sort_violation_word(scratch, n, 1 if vals else 0)
        u32(word, 1)[0] = 0
        return 0

    def vrenb200_radix_sort_keys(self, stream, keys, n, scratch, nbytes):
        self._sort(keys, None, n)
        return 0

    def vrenb200_radix_sort_pairs_host_async(self, stream, kin, vin, kout, vout, n, work, nbytes):
        u32(kout, n)[:] = u32(kin, n)
        u32(vout, n)[:] = u32(vin, n)
        self._sort(kout, vout, n)
        return 0

    def vrenb200_sort_profile_create(self): return 1
    def vrenb200_sort_profile_destroy(self, p): return None

    def vrenb200_sort_profile_read(self, p, ms):
        for i in range(6):
            ms[i] = 0.5
        return 0

    # ---- scans
    def _scan(self, src, dst, n, base):
        if n > EXACT_BELOW:                     # input is all ones in the bench: the closed form at the ends is all that is looked at
            u32(dst, 1)[0] = base
            u32(dst + 4 * (n - 4), 4)[:] = np.arange(n - 4, n, dtype=np.uint32) + np.uint32(base)
            return
        x = u32(src, n).astype(np.uint64)
        ex = np.concatenate([np.zeros(1, np.uint64), np.cumsum(x)[:-1]]) + np.uint64(base)
        u32(dst, n)[:] = (ex & np.uint64(0xFFFFFFFF)).astype(np.uint32)

    def vrenb200_exclusive_scan_u32(self, stream, src, dst, n, scratch, nbytes):
        self._scan(src, dst, n, 0)
        return 0

    def vrenb200_exclusive_scan_u32_ex(self, stream, src, dst, n, base, scratch, nbytes, flags):
        if flags & ~1:
            return 5                            # VRENB200_EINVAL_ARG, as the library answers an unknown flag
        self._scan(src, dst, n, base)
        return 0


    def vrenb200_exclusive_scan_u32_base(self, stream, src, dst, n, base, scratch, nbytes):
        self._scan(src, dst, n, base)
        return 0


class FakeShardedSort:
    """stand-in of vren_b200.dist.ShardedSort for the N > 1 walk (gloo): all-gather, stable sort of the concatenation, every rank
    keeps a contiguous slice — what the bench checks (oracle comparison, checksums, boundaries) holds for any such split"""

    def __init__(self, max_n, capacity):
        self.max_n, self.capacity = max_n, capacity if capacity is not None else int(max_n * 1.25) + 257 * 12288
        self.out_keys = torch.zeros(self.capacity, dtype=torch.int32)
        self.out_vals = torch.zeros(self.capacity, dtype=torch.int32)
        self.count = 0

    @classmethod
    def for_process_group(cls, max_n, capacity=None, rounds=1, group=None, config=None):
        return cls(max_n, capacity)

    def sort(self, keys, vals, key_bits=32, stream=None):
        import torch.distributed as dist

        world, rank = dist.get_world_size(), dist.get_rank()
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([keys.numel()], dtype=torch.int64))
        pad = max(int(t.item()) for t in sizes)
        gk = [torch.zeros(pad, dtype=torch.int32) for _ in range(world)]
        gv = [torch.zeros(pad, dtype=torch.int32) for _ in range(world)]
        dist.all_gather(gk, torch.nn.functional.pad(keys.contiguous(), (0, pad - keys.numel())))
        dist.all_gather(gv, torch.nn.functional.pad(vals.contiguous(), (0, pad - vals.numel())))
        k = np.concatenate([g[:int(s.item())].numpy().view(np.uint32) for g, s in zip(gk, sizes)])
        v = np.concatenate([g[:int(s.item())].numpy().view(np.uint32) for g, s in zip(gv, sizes)])
        mask = np.uint32(0xFFFFFFFF if key_bits == 32 else (1 << key_bits) - 1)
        order = np.argsort(k & mask, kind="stable")
        lo, hi = k.size * rank // world, k.size * (rank + 1) // world
        self.count = hi - lo
        self.out_keys[:self.count] = torch.from_numpy(k[order][lo:hi].view(np.int32).copy())
        self.out_vals[:self.count] = torch.from_numpy(v[order][lo:hi].view(np.int32).copy())

    def result(self):
        return self.out_keys[:self.count], self.out_vals[:self.count]

    def phases(self):
        return {"plan_and_partition_ms": 0.0, "transfers_ms": 0.0, "segmented_passes_ms": 0.0}

    def close(self): pass


def noop(*a, **k):
    return 0


for _name in ("vrenb200_reduce", "vrenb200_bucket_sort", "vrenb200_build_bvh", "vrenb200_visualize_bvh", "vrenb200_light_list_hash",
              "vrenb200_depth_pyramid_build", "vrenb200_bounce_point_lights", "vrenb200_construct_point_light_bvh", "vrenb200_find_unique_clusters",
              "vrenb200_assign_lights"):
    setattr(FakeLib, _name, staticmethod(noop))


def main():
    patch_torch_cuda()
    from vren_b200 import lib as vlib

    real = vlib.load()
    fake = FakeLib(real)
    vlib.load = lambda: fake
    if len(sys.argv) > 2 and sys.argv[1] == "--script":
        import runpy

        global EXACT_BELOW
        EXACT_BELOW = 1 << 40
        script = sys.argv[2]
        sys.argv = [script] + sys.argv[3:]
        with OnCpu():
            runpy.run_path(script, run_name="__main__")
        return
    import functools

    import bench

    bench.secondary_metrics = functools.partial(bench.secondary_metrics, big_log2=22)
    bench.secondary_metrics_multi = functools.partial(bench.secondary_metrics_multi, big_log2=20)
    if int(__import__("os").environ.get("WORLD_SIZE", "1")) > 1:
        # one process per "GPU" over gloo; the multi-GPU sort is the stand-in above
        import torch.distributed as dist

        from vren_b200 import dist as vdist

        real_init = dist.init_process_group
        dist.init_process_group = lambda backend=None, **k: real_init("gloo", **{kk: vv for kk, vv in k.items() if kk != "device_id"})
        vdist.ShardedSort = FakeShardedSort
        vdist.CudaOps.__init__ = lambda self: (setattr(self, "vlib", vlib), setattr(self, "lib", fake), setattr(self, "device", CPU))[0]
    bench.ClockSampler.__enter__ = lambda self: self              # no NVML here
    sys.argv = ["bench.py"] + sys.argv[1:]
    with OnCpu():
        bench.main()


if __name__ == "__main__":
    main()
