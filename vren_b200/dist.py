"""Multi-GPU (one process per GPU, torch.distributed) forms of the primitives that shard — SURVEY 8e.

The reference is single-device; this is new work specified by the north-star:
  * ShardedSort            : the product path.  Thin host wrapper of the C ABI's vrenb200_sharded_sort_* (csrc/sharded_sort.cu):
                             histograms published into every peer, the plan computed on the device, local partition by the
                             highest differing digit, per-round peer-store transfers over NVLink overlapped with segmented
                             onesweep passes of what has arrived.  No NCCL on the data path, no host synchronisation inside a
                             sort; torch only provides the symmetric (peer-mapped) memory and the process group for set-up.
  * sharded_sort_pairs     : the NCCL form kept as the baseline — local top-digit partition, all_to_all_single of keys and
                             values, local 4-pass sort of what was received.
  * sharded_bucket_sort    : vren::bucket_sort sharded — ShardedSort on the 16-bit key, END offsets by all_reduce of the ranks'
                             local END tables.
  * sharded_exclusive_scan / sharded_reduce_add : local reduce -> all_gather / all_reduce of the partial sums -> local scan with base.
  * views are independent: batched clustered shading needs no exchange (one view per rank, vren_b200/pipeline.py::ViewBatch).

The collective plumbing of the NCCL forms is backend-agnostic (NCCL on GPUs, gloo in the CPU tests); the local compute steps
come from a `LocalOps` object.  The product `CudaOps` calls the C ABI; tests may inject a numpy stand-in to exercise the exchange
logic on CPU.  There is no CPU fallback in this module: without CUDA, `CudaOps()` raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch
import torch.distributed as dist


class CudaOps:
    """local steps on the current CUDA device through libvrenb200.so"""

    def __init__(self):
        if not torch.cuda.is_available():
            raise RuntimeError("vren_b200.dist.CudaOps needs a CUDA device (no CPU fallback)")
        from . import lib as vlib

        self.vlib = vlib
        self.lib = vlib.load()
        self.device = torch.device("cuda", torch.cuda.current_device())

    def _stream(self):
        return torch.cuda.current_stream().cuda_stream

    def digit_histograms(self, keys: torch.Tensor) -> torch.Tensor:
        hist = torch.empty(4, 256, dtype=torch.int32, device=keys.device)
        self.vlib.check(self.lib.vrenb200_radix_digit_histograms(self._stream(), keys.data_ptr(), keys.numel(), hist.data_ptr()),
                        "radix_digit_histograms")
        return hist

    def partition_by_top_digit(self, keys: torch.Tensor, vals: torch.Tensor):
        """stable partition by bits 24..31 (one onesweep pass); returns new tensors"""
        n = keys.numel()
        ok, ov = torch.empty_like(keys), torch.empty_like(vals)
        sb = self.lib.vrenb200_radix_sort_range_scratch_bytes(n)
        scratch = torch.empty(max(sb, 256), dtype=torch.uint8, device=keys.device)
        self.vlib.check(self.lib.vrenb200_radix_sort_pairs_range(self._stream(), keys.data_ptr(), vals.data_ptr(), ok.data_ptr(), ov.data_ptr(),
                                                                 n, 3, 1, scratch.data_ptr(), sb, None), "radix_sort_pairs_range")
        return ok, ov

    def sort_pairs(self, keys: torch.Tensor, vals: torch.Tensor):
        self.vlib.radix_sort_pairs(keys, vals)
        return keys, vals

    def reduce_add(self, x: torch.Tensor) -> int:
        n = x.numel()
        if n == 0:              # an empty shard contributes the identity
            return 0
        out = self.vlib.reduce(x, n, "u32", "add", mode="final")
        P = self.lib.vrenb200_round_to_next_power_of_2(n)
        return int(out[P - 1].item()) & 0xFFFFFFFF

    def exclusive_scan(self, x: torch.Tensor, base: int) -> torch.Tensor:
        n = x.numel()
        if n == 0:
            return x
        sb = self.lib.vrenb200_scan_scratch_bytes(n)
        scratch = torch.empty(max(sb, 256), dtype=torch.uint8, device=x.device)
        self.vlib.check(self.lib.vrenb200_exclusive_scan_u32_base(self._stream(), x.data_ptr(), x.data_ptr(), n, base & 0xFFFFFFFF,
                                                                  scratch.data_ptr(), sb), "exclusive_scan_u32_base")
        return x


@dataclass
class SortPlan:
    """what every rank derives from the gathered top-digit histograms (identical on all ranks)"""
    digit_lo: list            # rank r owns top digits [digit_lo[r], digit_lo[r+1])
    send_counts: list         # pairs this rank sends to each destination
    recv_counts: list         # pairs this rank receives from each source


def plan_digit_ranges(total_hist: torch.Tensor, world: int) -> list:
    """split the 256 values of a digit into `world` contiguous ranges with (nearly) equal pair counts.
    total_hist: int64[256] on CPU.  Returns world+1 boundaries, boundaries[0] = 0, boundaries[world] = 256.
    (The device plan of csrc/sharded_sort.cu::plan_kernel applies the same rule.)"""
    cum = torch.cumsum(total_hist, 0)
    total = int(cum[-1])
    bounds = [0]
    for r in range(1, world):
        target = (total * r + world - 1) // world
        # first digit whose inclusive cumulative count reaches the target, rounded to the nearer boundary
        d = int(torch.searchsorted(cum, torch.tensor(target, dtype=cum.dtype)).item())
        d = min(max(d, bounds[-1]), 255)
        before = int(cum[d - 1]) if d > 0 else 0
        after = int(cum[d])
        cut = d if (target - before) <= (after - target) else d + 1
        bounds.append(min(max(cut, bounds[-1]), 256))
    bounds.append(256)
    return bounds


def plan_exchange_counts(hists: torch.Tensor, bounds: list):
    """hists: int64 [G, 256] (CPU) -> per_dest [src, dst]: pairs every source sends to every destination"""
    world = hists.shape[0]
    b = torch.tensor(bounds, dtype=torch.int64)
    csum = torch.cat([torch.zeros(world, 1, dtype=torch.int64), torch.cumsum(hists, 1)], dim=1)        # [G, 257]
    return csum[:, b[1:]] - csum[:, b[:-1]]


def make_sort_plan(local_top_hist: torch.Tensor, group=None) -> SortPlan:
    """local_top_hist: int64[256] on the rank's device (top-digit counts of the local keys)"""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    gathered = [torch.empty_like(local_top_hist) for _ in range(world)]
    dist.all_gather(gathered, local_top_hist, group=group)
    hists = torch.stack(gathered).cpu()                      # [world, 256]
    bounds = plan_digit_ranges(hists.sum(0), world)
    per_dest = plan_exchange_counts(hists, bounds)                                                     # [src, dst]
    return SortPlan(bounds, [int(v) for v in per_dest[rank]], [int(v) for v in per_dest[:, rank]])


def sharded_sort_pairs(keys: torch.Tensor, vals: torch.Tensor, ops=None, group=None):
    """NCCL baseline: globally stable sort by key of the pairs held by all ranks.  keys/vals: int32 storage of uint32
    values.  Returns (keys_out, vals_out, plan): rank r holds the pairs whose top digit falls in its range, sorted."""
    ops = ops or CudaOps()
    hist = ops.digit_histograms(keys)[3]
    plan = make_sort_plan(hist.to(torch.int64), group)
    pk, pv = ops.partition_by_top_digit(keys, vals)
    n_recv = sum(plan.recv_counts)
    rk = torch.empty(n_recv, dtype=keys.dtype, device=keys.device)
    rv = torch.empty(n_recv, dtype=vals.dtype, device=vals.device)
    dist.all_to_all_single(rk, pk, output_split_sizes=plan.recv_counts, input_split_sizes=plan.send_counts, group=group)
    dist.all_to_all_single(rv, pv, output_split_sizes=plan.recv_counts, input_split_sizes=plan.send_counts, group=group)
    ops.sort_pairs(rk, rv)
    return rk, rv, plan


def exchange_plan(hists, key_digits: int, tile: int, rounds: int, cap_tiles: int, round_bound: int):
    """Host mirror of csrc/sharded_sort.cu::plan_kernel (same arithmetic, numpy) — used by the CPU tests of the exchange
    logic and for sizing the receive capacity ahead of a sort; the product computes its plan on the device.
    hists: integer array [G, 4, 256], digit counts of every rank's shard.  Returns a dict:
      pstar, error, bounds[G+1], owner[256], first_tile[256] (in the owner's buffer), seg_len[256], dst_off[G][256] (where
      source s writes its block of digit d), round_digit[G][rounds+1], out_count[G]"""
    import numpy as np

    h = np.asarray(hists, dtype=np.int64)
    world = h.shape[0]
    pstar = 0
    for p in range(key_digits - 1, -1, -1):
        if int((h[:, p, :].sum(0) != 0).sum()) >= 2:
            pstar = p
            break
    total = h[:, pstar, :].sum(0)
    bounds = plan_digit_ranges(torch.from_numpy(total), world)
    owner = np.zeros(256, np.int64)
    for r in range(1, world):
        owner += (np.arange(256) >= bounds[r])
    tiles = (total + tile - 1) // tile
    tiles_ex = np.concatenate([[0], np.cumsum(tiles)])
    first_tile = tiles_ex[:-1] - tiles_ex[np.array(bounds)[owner]]
    error = 0
    round_digit = np.zeros((world, rounds + 1), np.int64)
    for r in range(world):
        lo, hi = bounds[r], bounds[r + 1]
        if tiles_ex[hi] - tiles_ex[lo] > cap_tiles:
            error |= 1
        base, all_tiles = tiles_ex[lo], tiles_ex[hi] - tiles_ex[lo]
        dd = lo
        round_digit[r, 0] = lo
        for k in range(1, rounds):
            target = (all_tiles * k + rounds - 1) // rounds
            while dd < hi and tiles_ex[dd] - base < target:
                dd += 1
            round_digit[r, k] = dd
        round_digit[r, rounds] = hi
        for k in range(rounds):
            if tiles_ex[round_digit[r, k + 1]] - tiles_ex[round_digit[r, k]] > round_bound:
                error |= 2
    before = np.concatenate([np.zeros((1, 256), np.int64), np.cumsum(h[:, pstar, :], axis=0)[:-1]])     # [G, 256]: lower sources
    dst_off = first_tile[None, :] * tile + before
    out_count = [int(total[bounds[r]:bounds[r + 1]].sum()) for r in range(world)]
    return {"pstar": pstar, "error": error, "bounds": bounds, "owner": owner, "first_tile": first_tile, "seg_len": total,
            "dst_off": dst_off, "round_digit": round_digit, "out_count": out_count}


def required_capacity(hists, key_digits: int = 4, tile: int = 12288) -> int:
    """The overflow policy of the multi-GPU sort, host side: a plan that does not fit is REPORTED (status word, ShardedSort.result()
    raises), never truncated; the caller then sizes a new context with this function.  hists: [G, 4, 256] digit counts of every
    rank's shard (vrenb200_radix_digit_histograms, all-gathered — or what the ranks already know about their keys).  Returns the
    smallest receive capacity, in pairs, with which the device plan fits on every rank: a value of the partition digit is the
    unit of ownership (it cannot be split), so a hot value needs room for all of its pairs on its owner."""
    import numpy as np

    h = np.asarray(hists, dtype=np.int64)
    world = h.shape[0]
    big = 1 << 40
    plan = exchange_plan(h, key_digits, tile, 1, big, big)
    tiles = (plan["seg_len"] + tile - 1) // tile
    b = plan["bounds"]
    return int(max(int(tiles[b[r]:b[r + 1]].sum()) for r in range(world)) * tile)


def default_capacity(max_n: int, world: int, rounds: int = 1) -> int:
    """receive capacity for balanced keys: the rank's share with 25 % slack plus a partial tile per segment (256 segments
    of up to 12288 pairs in all) — see include/vrenb200.h"""
    return int(max_n * 1.25) + 257 * 12288


class ShardedSort:
    """One rank's context of the multi-GPU sort (vrenb200_sharded_sort_*).

    `regions`: the symmetric regions of ALL ranks as uint8 tensors addressable from this rank, in rank order.  Use
    ShardedSort.for_process_group() (torch symmetric memory, one process per GPU) or ShardedSort.emulated() (several ranks on
    one device inside one process: the peers are ordinary device tensors).  The output tensors are owned by the context and
    overwritten by the next sort."""

    def __init__(self, rank: int, world: int, max_n: int, capacity: int, regions, rounds: int = 1, config=None):
        from . import lib as vlib

        self.vlib, self.lib = vlib, vlib.load()
        self.rank, self.world, self.max_n, self.capacity, self.rounds = rank, world, max_n, capacity, rounds
        self.regions = regions
        dev = regions[rank].device
        cfg_ptr = C.addressof(config) if config is not None else None
        self.local = torch.empty(self.lib.vrenb200_sharded_sort_local_bytes(max_n, capacity, cfg_ptr), dtype=torch.uint8, device=dev)
        ptrs = (C.c_void_p * world)(*[int(r.data_ptr()) for r in regions])
        handle = C.c_void_p()
        self.config = config
        vlib.check(self.lib.vrenb200_sharded_sort_create(C.byref(handle), rank, world, max_n, capacity, rounds, ptrs, self.local.data_ptr(),
                                                         self.local.numel(), C.addressof(config) if config is not None else None),
                   "vrenb200_sharded_sort_create")
        self.handle = handle

        def view(ptr, count):
            off = ptr - self.local.data_ptr()
            return self.local[off: off + 4 * count].view(torch.int32)

        self.out_keys = view(self.lib.vrenb200_sharded_sort_out_keys(handle), capacity)
        self.out_vals = view(self.lib.vrenb200_sharded_sort_out_values(handle), capacity)
        self.status = view(self.lib.vrenb200_sharded_sort_status(handle), 8)

    @staticmethod
    def symmetric_bytes(capacity: int) -> int:
        from . import lib as vlib

        return int(vlib.load().vrenb200_sharded_sort_symmetric_bytes(capacity))

    @classmethod
    def for_process_group(cls, max_n: int, capacity: int | None = None, rounds: int = 1, group=None, config=None):
        """one process per GPU: the symmetric regions come from torch symmetric memory (peer-mapped over NVLink)"""
        import torch.distributed._symmetric_memory as symm

        group = group if group is not None else dist.group.WORLD
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        capacity = capacity if capacity is not None else default_capacity(max_n, world, rounds)
        dev = torch.device("cuda", torch.cuda.current_device())
        nbytes = cls.symmetric_bytes(capacity)
        region = symm.empty(nbytes, dtype=torch.uint8, device=dev)
        hdl = symm.rendezvous(region, group)
        regions = [hdl.get_buffer(r, (nbytes,), torch.uint8) for r in range(world)]
        self = cls(rank, world, max_n, capacity, regions, rounds, config)
        self._symm = (region, hdl)
        torch.cuda.synchronize()
        dist.barrier(group)              # every rank has zeroed its flag words before anyone sorts
        return self

    @classmethod
    def emulated(cls, world: int, max_n: int, capacity: int | None = None, rounds: int = 1, config=None, device="cuda"):
        """all ranks in THIS process on one device (tests): returns the list of contexts"""
        capacity = capacity if capacity is not None else default_capacity(max_n, world, rounds)
        nbytes = cls.symmetric_bytes(capacity)
        regions = [torch.zeros(nbytes, dtype=torch.uint8, device=device) for _ in range(world)]
        ctxs = [cls(r, world, max_n, capacity, regions, rounds, config) for r in range(world)]
        torch.cuda.synchronize()
        return ctxs

    def sort(self, keys: torch.Tensor, vals: torch.Tensor, key_bits: int = 32, stream=None):
        """enqueue this rank's part of one collective sort (no synchronisation).  keys/vals: int32 storage, 16-byte aligned"""
        s = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        self.vlib.check(self.lib.vrenb200_sharded_sort_pairs(self.handle, s, keys.data_ptr(), vals.data_ptr(), keys.numel(), key_bits),
                        "vrenb200_sharded_sort_pairs")

    def result(self):
        """after the stream has finished the call: (keys, vals) views of this rank's sorted output; raises if the plan did
        not fit (status word)"""
        st = self.status.cpu().numpy().view("uint32")
        if st[0] != 0:
            raise RuntimeError(f"sharded sort: the exchange plan does not fit (status {int(st[0])}: "
                               f"{'receive capacity exceeded' if st[0] & 1 else 'a round exceeds its launch bound'}); "
                               f"capacity {self.capacity}, rounds {self.rounds} — retry with rounds=1 and/or a larger capacity "
                               "(vren_b200.dist.required_capacity sizes it from the ranks' digit histograms)")
        n = int(st[1])
        return self.out_keys[:n], self.out_vals[:n]

    def phases(self):
        """CUDA-event milliseconds of the last sort on this rank (after a synchronise): {plan + partition, transfers, passes}"""
        ms = (C.c_float * 3)()
        self.vlib.check(self.lib.vrenb200_sharded_sort_phases(self.handle, ms), "vrenb200_sharded_sort_phases")
        return {"plan_and_partition_ms": float(ms[0]), "transfers_ms": float(ms[1]), "segmented_passes_ms": float(ms[2])}

    def owned_digits(self):
        st = self.status.cpu().numpy().view("uint32")
        return int(st[2]), int(st[3]), int(st[4])

    def close(self):
        if getattr(self, "handle", None):
            self.lib.vrenb200_sharded_sort_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sharded_bucket_sort(ctx: ShardedSort, pairs: torch.Tensor, group=None):
    """vren::bucket_sort (bucket_sort.cpp:62-161) over the pairs of all ranks: stable by x & 0xFFFF.  pairs: int32 [n, 2]
    (uvec2).  Returns (sorted pairs [m, 2] of this rank, END offsets int64[65536] of the GLOBAL sorted sequence — what the
    reference's counter region holds after the call, bucket_sort_write.comp:32)."""
    keys = pairs[:, 0].contiguous()
    vals = pairs[:, 1].contiguous()
    ctx.sort(keys, vals, key_bits=16)
    torch.cuda.current_stream().synchronize()
    k, v = ctx.result()
    # local END table: number of local keys <= b; the global one is the sum over the ranks (shards are disjoint key ranges)
    k16 = (k.to(torch.int64) & 0xFFFF)
    ends = torch.searchsorted(k16, torch.arange(65536, dtype=torch.int64, device=k.device), right=True)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(ends, group=group)
    return torch.stack([k, v], dim=1), ends


def sharded_exclusive_scan(x: torch.Tensor, ops=None, group=None) -> torch.Tensor:
    """in-place exclusive add-scan (mod 2^32) of the concatenation of every rank's `x` (rank order)"""
    ops = ops or CudaOps()
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    total = torch.tensor([ops.reduce_add(x)], dtype=torch.int64, device=x.device)
    gathered = [torch.empty_like(total) for _ in range(world)]
    dist.all_gather(gathered, total, group=group)
    base = sum(int(t.item()) for t in gathered[:rank]) & 0xFFFFFFFF
    return ops.exclusive_scan(x, base)


def sharded_reduce_add(x: torch.Tensor, ops=None, group=None) -> int:
    ops = ops or CudaOps()
    total = torch.tensor([ops.reduce_add(x)], dtype=torch.int64, device=x.device)
    dist.all_reduce(total, group=group)
    return int(total.item()) & 0xFFFFFFFF
