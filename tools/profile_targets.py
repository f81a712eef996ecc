"""launches each secondary kernel a few times so that ncu can capture them (scan, reduce, clustered chain, sort)"""
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from vren_b200 import lib as vlib, synthetic  # noqa: E402
from vren_b200.pipeline import ClusterAndShade  # noqa: E402

lib = vlib.load()
dev = torch.device("cuda")
stream = torch.cuda.current_stream().cuda_stream
n = 1 << 28
ONLY = os.environ.get("VREN_PROFILE_ONLY", "")   # "sort": skip the scan and the clustered chain
x = torch.ones(n, dtype=torch.int32, device=dev)
y = torch.empty_like(x)
sb = lib.vrenb200_scan_scratch_bytes(n)
scr = torch.empty(sb, dtype=torch.uint8, device=dev)
for _ in range(0 if ONLY == "sort" else 3):
    vlib.check(lib.vrenb200_exclusive_scan_u32(stream, x.data_ptr(), y.data_ptr(), n, scr.data_ptr(), sb), "scan")
torch.cuda.synchronize()
del x, y
w, h, L = 3840, 2160, 65536
depth = torch.from_numpy(synthetic.depth_buffer(w, h, seed=2024)).to(dev)
pos, lights = synthetic.point_lights(L, seed=2025, aspect=w / h, intensity=(1.0, 1.0))
pos, lights = torch.from_numpy(pos).to(dev), torch.from_numpy(lights).to(dev)
view = synthetic.view_matrix(0.0, 0.0, (0, 0, 0)).tolist()
cam = vlib.Camera(np.float32(math.radians(45.0)), np.float32(w / h), np.float32(0.01), np.float32(1000.0))
cs = ClusterAndShade(w, h, max_point_lights=L)
for _ in range(0 if ONLY == "sort" else 3):
    cs(w, h, cam, view, depth, None, pos, lights, L)
torch.cuda.synchronize()
g = torch.Generator(device=dev)
g.manual_seed(1)
keys = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
vals = torch.arange(n, dtype=torch.int32, device=dev)
cfg = vlib.SortConfig(vlib.RANKING_AUTO, vlib.TILE_IDS_AUTO, int(os.environ.get("VREN_SORT_VARIANT", "0")))
for _ in range(2):
    vlib.radix_sort_pairs(keys, vals, config=cfg)
torch.cuda.synchronize()
print("done")
