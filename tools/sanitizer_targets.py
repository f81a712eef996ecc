"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
compute-sanitizer --tool memcheck python tools/sanitizer_targets.py   (from the repo root)"""
import math
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from vren_b200 import lib as vlib, synthetic
from vren_b200.pipeline import ClusterAndShade

lib = vlib.load()
dev = torch.device("cuda")
g = torch.Generator(device=dev)
g.manual_seed(5)
# scan: register-tile kernel, run-ahead kernel (n >= 2^22), ragged and unaligned
for n, off in ((1000, 0), (70001, 1), ((1 << 22) + 16384 * 3 + 5, 0), ((1 << 22) + 77, 3)):
    buf = torch.randint(0, 100, (n + 8,), dtype=torch.int32, device=dev, generator=g)
    x = buf[off:off + n]
    want = torch.cumsum(x, 0, dtype=torch.int64) - x
    vlib.exclusive_scan(x)
    assert torch.equal(x.to(torch.int64) & 0xFFFFFFFF, want & 0xFFFFFFFF), ("scan", n, off)
# sort: pairs, keys, every non-experimental variant at a ragged size
n = 3 * 8192 + 1234
k = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
import os  # noqa: E402
only = [int(a) for a in os.environ.get("VREN_SANITIZER_VARIANTS", "").split(",") if a]
for var in only or range(1, lib.vrenb200_radix_sort_num_variants() + 1):
    keys_only_tile = b"x64/" in lib.vrenb200_radix_sort_variant_name(var)
    for tile_ids in (vlib.TILE_IDS_BLOCK_INDEX, vlib.TILE_IDS_TICKET):
        cfg = vlib.SortConfig(vlib.RANKING_AUTO, tile_ids, var)
        if not keys_only_tile:
            kk, vv = k.clone(), torch.arange(n, dtype=torch.int32, device=dev)
            vlib.radix_sort_ex(kk, vv, cfg)
            u = kk.to(torch.int64) & 0xFFFFFFFF
            assert bool((u[1:] >= u[:-1]).all()) and torch.equal(k[vv.long()], kk), ("pairs", var)
        kk = k.clone()
        vlib.radix_sort_ex(kk, None, cfg)
        u = kk.to(torch.int64) & 0xFFFFFFFF
        assert bool((u[1:] >= u[:-1]).all()), ("keys", var)
# the repeat passes of the verified ranking
kk, vv = k.clone(), torch.arange(n, dtype=torch.int32, device=dev)
assert vlib.radix_sort_ex(kk, vv, vlib.SortConfig(vlib.RANKING_SELFTEST_REDO, 0, 0)) == 0xF
assert torch.equal(k[vv.long()], kk)
# variant 0 takes the small tile below 2^21 pairs: exercise the default large tiles too (ragged last tile)
n = (1 << 21) + 12345
k = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
kk, vv = k.clone(), torch.arange(n, dtype=torch.int32, device=dev)
vlib.radix_sort_pairs(kk, vv)
u = kk.to(torch.int64) & 0xFFFFFFFF
assert bool((u[1:] >= u[:-1]).all()) and torch.equal(k[vv.long()], kk), "pairs, default tile"
kk = k.clone()
vlib.radix_sort_keys(kk)
u = kk.to(torch.int64) & 0xFFFFFFFF
assert bool((u[1:] >= u[:-1]).all()), "keys, default tile"
# bucket sort: END offsets by search in the sorted output (the path of inputs >= 2^20 pairs), then by counting
pairs = torch.randint(0, 1 << 16, (50001, 2), dtype=torch.int32, device=dev, generator=g)
_, sorted_pairs, counters = vlib.bucket_sort(pairs, end_offsets=1)
assert int(counters[-1]) == 50001
pairs = torch.randint(0, 1 << 16, (50001, 2), dtype=torch.int32, device=dev, generator=g)
_, sorted_pairs, counters = vlib.bucket_sort(pairs, end_offsets=0)
keys16 = sorted_pairs[:, 0] & 0xFFFF
assert bool((keys16[1:] >= keys16[:-1]).all()) and int(counters[-1]) == 50001
# reduce
x = torch.randint(0, 100, (100003,), dtype=torch.int32, device=dev, generator=g)
for mode in ("tree", "final"):
    vlib.reduce(x, x.numel(), "u32", "add", mode=mode)
# clustered chain + consumers
w, h, L = 320, 200, 3000
depth = torch.from_numpy(synthetic.depth_buffer(w, h, seed=3)).to(dev)
nrm = torch.from_numpy(synthetic.normal_buffer(w, h, seed=4)).to(dev)
pos, lights = synthetic.point_lights(L, seed=6, aspect=w / h, intensity=(0.5, 3.0))
pos, lights = torch.from_numpy(pos).to(dev), torch.from_numpy(lights).to(dev)
view = synthetic.view_matrix(0.1, 0.0, (0, 0, 0)).tolist()
cam = vlib.Camera(np.float32(math.radians(45.0)), np.float32(w / h), np.float32(0.01), np.float32(1000.0))
cs = ClusterAndShade(w, h, max_point_lights=L)
for normals in (None, nrm):
    cs(w, h, cam, view, depth, normals, pos, lights, L)
    torch.cuda.synchronize()
    assert int(cs.dispatch_params[0]) > 0 and int(cs.status[0]) > 0
vlib.light_list_hash(cs.cluster_ref[:h, :w].contiguous(), cs.dispatch_params, cs.counts, cs.offsets, cs.indices)
vlib.depth_pyramid(depth)
d = torch.nn.functional.normalize(torch.randn(L, 3, device=dev), dim=1)
d = torch.cat([d, torch.zeros(L, 1, device=dev)], 1).contiguous()
p4 = torch.cat([pos[:, :3].clone(), torch.ones(L, 1, device=dev)], 1).contiguous()
vlib.bounce_point_lights(p4, d, (-5.0, -5.0, -5.0), (5.0, 5.0, 5.0), 30.0, 0.5)
levels = lib.vrenb200_calc_bvh_level_count(L)
vlib.visualize_bvh(cs.bvh, levels)
torch.cuda.synchronize()
print("sanitizer targets done")
