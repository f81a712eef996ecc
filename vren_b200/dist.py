"""Multi-GPU (one process per GPU, torch.distributed) versions of the primitives that shard — SURVEY 8e.

The reference is single-device; this is new work specified by the north-star:
  * sharded_sort_pairs  : every rank holds n_r pairs.  Local top-digit histogram -> all_gather -> contiguous digit
                          ranges per rank (balanced by count) -> stable local partition by the top digit (one
                          onesweep pass) -> all-to-all-v of keys and values over NVLink (NCCL) -> local 4-pass
                          onesweep of the received pairs.  Concatenating the ranks' outputs in rank order gives the
                          stable sort of the concatenated input.
  * sharded_exclusive_scan : local reduce -> all_gather of the partial sums -> local scan with base.
  * views are independent: batched clustered shading needs no exchange (one view per rank, see bench.py).

The collective plumbing is backend-agnostic (NCCL on GPUs, gloo in the CPU tests); the local compute steps come from
a `LocalOps` object.  The product `CudaOps` calls the C ABI; tests may inject a numpy stand-in to exercise the
exchange logic on CPU.  There is no CPU fallback in this module: without CUDA, `CudaOps()` raises.
"""
from __future__ import annotations

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

    def top_digit_histogram(self, keys: torch.Tensor) -> torch.Tensor:
        hist = torch.empty(256, dtype=torch.int32, device=keys.device)
        self.vlib.check(self.lib.vrenb200_radix_top_digit_histogram(self._stream(), keys.data_ptr(), keys.numel(), hist.data_ptr()),
                        "radix_top_digit_histogram")
        return hist

    def partition_scatter(self, keys: torch.Tensor, vals: torch.Tensor, dest_table: torch.Tensor):
        """fused partition + exchange: pairs are stored straight at dest_table[0/1][top digit] (peer-mapped or local)"""
        n = keys.numel()
        sb = self.lib.vrenb200_radix_sort_range_scratch_bytes(n)
        scratch = torch.empty(max(sb, 256), dtype=torch.uint8, device=keys.device)
        self.vlib.check(self.lib.vrenb200_radix_partition_scatter(self._stream(), keys.data_ptr(), vals.data_ptr(), n, dest_table.data_ptr(),
                                                                  scratch.data_ptr(), sb), "radix_partition_scatter")

    def sort_pairs(self, keys: torch.Tensor, vals: torch.Tensor):
        self.vlib.radix_sort_pairs(keys, vals)
        return keys, vals

    def reduce_add(self, x: torch.Tensor) -> int:
        n = x.numel()
        out = self.vlib.reduce(x, n, "u32", "add", mode="final")
        P = self.lib.vrenb200_round_to_next_power_of_2(n)
        return int(out[P - 1].item()) & 0xFFFFFFFF

    def exclusive_scan(self, x: torch.Tensor, base: int) -> torch.Tensor:
        n = x.numel()
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
    """split the 256 top digits into `world` contiguous ranges with (nearly) equal pair counts.
    total_hist: int64[256] on CPU.  Returns world+1 boundaries, boundaries[0] = 0, boundaries[world] = 256."""
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


def make_sort_plan(local_top_hist: torch.Tensor, group=None) -> SortPlan:
    """local_top_hist: int64[256] on the rank's device (top-digit counts of the local keys)"""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    gathered = [torch.empty_like(local_top_hist) for _ in range(world)]
    dist.all_gather(gathered, local_top_hist, group=group)
    hists = torch.stack(gathered).cpu()                      # [world, 256]
    bounds = plan_digit_ranges(hists.sum(0), world)
    per_dest = plan_p2p_offsets(hists, bounds, rank)[3]                                               # [src, dst]
    return SortPlan(bounds, [int(v) for v in per_dest[rank]], [int(v) for v in per_dest[:, rank]])


def sharded_sort_pairs(keys: torch.Tensor, vals: torch.Tensor, ops=None, group=None):
    """globally stable sort by key of the pairs held by all ranks.  keys/vals: int32 storage of uint32 values.
    Returns (keys_out, vals_out, plan): rank r holds the pairs whose top digit falls in its range, sorted."""
    ops = ops or CudaOps()
    hist = ops.top_digit_histogram(keys) if hasattr(ops, "top_digit_histogram") else ops.digit_histograms(keys)[3]
    plan = make_sort_plan(hist.to(torch.int64), group)
    pk, pv = ops.partition_by_top_digit(keys, vals)
    n_recv = sum(plan.recv_counts)
    rk = torch.empty(n_recv, dtype=keys.dtype, device=keys.device)
    rv = torch.empty(n_recv, dtype=vals.dtype, device=vals.device)
    dist.all_to_all_single(rk, pk, output_split_sizes=plan.recv_counts, input_split_sizes=plan.send_counts, group=group)
    dist.all_to_all_single(rv, pv, output_split_sizes=plan.recv_counts, input_split_sizes=plan.send_counts, group=group)
    ops.sort_pairs(rk, rv)
    return rk, rv, plan


def plan_p2p_offsets(hists: torch.Tensor, bounds: list, rank: int):
    """receive-buffer layout of the fused exchange: rank r's buffer holds one block per source rank, in source-rank order,
    each block = that source's pairs whose top digit falls in r's range, in the source's original order.
    hists: int64 [G, 256] (CPU).  Returns (rank_of[256] uint8, my_offset[G] = element offset of THIS rank's block inside
    every destination's buffer, recv_counts[G])."""
    world = hists.shape[0]
    b = torch.tensor(bounds, dtype=torch.int64)
    rank_of = (torch.bucketize(torch.arange(256), b[1:-1], right=True)).to(torch.uint8) if world > 1 else torch.zeros(256, dtype=torch.uint8)
    csum = torch.cat([torch.zeros(world, 1, dtype=torch.int64), torch.cumsum(hists, 1)], dim=1)        # [G, 257]
    per_dest = csum[:, b[1:]] - csum[:, b[:-1]]                                                        # [src, dst]
    my_offset = per_dest[:rank].sum(0)
    return rank_of, my_offset, [int(v) for v in per_dest.sum(0)], per_dest


def plan_digit_exchange(hists: torch.Tensor, bounds: list, rank: int):
    """Planning for an exchange pass that partitions by the FULL top digit (host side only, the device side
    is not wired yet).  Destination r's receive buffer is laid out by top digit, and inside a digit by source rank, so it
    arrives grouped into top-digit segments and the local sort needs only the three low passes per segment.
    hists: int64 [G, 256] top-digit counts of every rank (CPU); bounds: plan_digit_ranges().
    Returns (rank_of uint8[256], my_digit_offset int64[256] = where THIS rank's pairs of top digit d start inside the
    owner's buffer, segment_start int64[G][257] = start of every digit segment in every destination's buffer
    (segment_start[r][d] == segment_start[r][d+1] outside r's range), recv_counts[G])."""
    world = hists.shape[0]
    b = torch.tensor(bounds, dtype=torch.int64)
    rank_of = (torch.bucketize(torch.arange(256), b[1:-1], right=True)).to(torch.uint8) if world > 1 else torch.zeros(256, dtype=torch.uint8)
    total = hists.sum(0)                                                   # [256]
    owner = rank_of.to(torch.int64)
    in_range = torch.nn.functional.one_hot(owner, world).T.to(torch.int64)  # [G, 256]: 1 where digit d belongs to rank r
    seg_len = in_range * total                                             # [G, 256]
    segment_start = torch.cat([torch.zeros(world, 1, dtype=torch.int64), torch.cumsum(seg_len, 1)], dim=1)   # [G, 257]
    my_digit_offset = segment_start[owner, torch.arange(256)] + hists[:rank].sum(0)
    return rank_of, my_digit_offset, segment_start, [int(v) for v in seg_len.sum(1)]


class P2PExchange:
    """symmetric-memory receive buffers (keys, values) of one process group, created once and reused"""

    def __init__(self, capacity: int, device, group=None):
        import torch.distributed._symmetric_memory as symm

        self.group = group if group is not None else dist.group.WORLD
        self.capacity = capacity
        self.keys = symm.empty(capacity, dtype=torch.int32, device=device)
        self.vals = symm.empty(capacity, dtype=torch.int32, device=device)
        self.hk = symm.rendezvous(self.keys, self.group)
        self.hv = symm.rendezvous(self.vals, self.group)
        self.key_ptrs = [int(p) for p in self.hk.buffer_ptrs]
        self.val_ptrs = [int(p) for p in self.hv.buffer_ptrs]

    def barrier(self):
        self.hk.barrier()


def sharded_sort_pairs_p2p(keys: torch.Tensor, vals: torch.Tensor, exchange: P2PExchange, ops=None, group=None, phases: dict | None = None):
    """same contract as sharded_sort_pairs, but the all-to-all is fused into the partition kernel: every rank's
    onesweep pass on the top digit stores its pairs directly into the destination ranks' receive buffers through
    NVLink peer pointers (no NCCL on the data path, no send-side staging copy).
    `phases` (optional dict): filled with CUDA-event milliseconds of {plan, exchange, local_sort} for this call."""
    ops = ops or CudaOps()
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if phases is not None else None
    if ev:
        ev[0].record()
    # every rank has entered this call, i.e. is done with the previous contents of the receive buffers; issued first so
    # that it completes under the planning below instead of in front of the exchange pass
    exchange.barrier()
    hist = ops.top_digit_histogram(keys).to(torch.int64)
    gathered = [torch.empty_like(hist) for _ in range(world)]
    dist.all_gather(gathered, hist, group=group)
    hists = torch.stack(gathered).cpu()
    bounds = plan_digit_ranges(hists.sum(0), world)
    rank_of, my_offset, recv_counts, per_dest = plan_p2p_offsets(hists, bounds, rank)
    if max(recv_counts) > exchange.capacity:
        raise RuntimeError(f"P2P receive buffer too small: need {max(recv_counts)} pairs, capacity {exchange.capacity}")
    if world > 32:
        raise RuntimeError("fused exchange supports up to 32 ranks")
    # device table {u64 kptr[32]; u64 vptr[32]; u8 rank_of[256]} (vrenb200_radix_partition_scatter)
    kp = torch.zeros(32, dtype=torch.int64)
    vp = torch.zeros(32, dtype=torch.int64)
    kp[:world] = torch.tensor(exchange.key_ptrs, dtype=torch.int64) + 4 * my_offset
    vp[:world] = torch.tensor(exchange.val_ptrs, dtype=torch.int64) + 4 * my_offset
    table = torch.cat([kp.view(torch.uint8), vp.view(torch.uint8), rank_of]).to(keys.device, non_blocking=True)
    if ev:
        ev[1].record()
    ops.partition_scatter(keys, vals, table)
    exchange.barrier()                       # all remote stores into my buffers have completed
    if ev:
        ev[2].record()
    n_recv = recv_counts[rank]
    rk, rv = exchange.keys[:n_recv], exchange.vals[:n_recv]
    ops.sort_pairs(rk, rv)
    if ev:
        ev[3].record()
        ev[3].synchronize()
        phases["plan_ms"] = ev[0].elapsed_time(ev[1])
        phases["exchange_ms"] = ev[1].elapsed_time(ev[2])
        phases["local_sort_ms"] = ev[2].elapsed_time(ev[3])
        phases["received_pairs"] = int(n_recv)
    plan = SortPlan(bounds, [int(v) for v in per_dest[rank]], [int(v) for v in per_dest[:, rank]])
    return rk, rv, plan


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
