"""times the a6 -> a7 -> a8 chain at BASELINE C5 (3840x2160, 65536 lights) with CUDA events; prints one JSON line"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from vren_b200 import lib as vlib, synthetic  # noqa: E402


def main(w=3840, h=2160, L=65536, iters=20, intensity=1.0):
    lib = vlib.load()
    depth = torch.from_numpy(synthetic.depth_buffer(w, h, seed=2024)).cuda()
    pos, lights = synthetic.point_lights(L, seed=2025, aspect=w / h, intensity=(intensity, intensity))
    pos, lights = torch.from_numpy(pos).cuda(), torch.from_numpy(lights).cuda()
    view = synthetic.view_matrix(0.0, 0.0, (0, 0, 0)).tolist()
    import math
    cam = vlib.Camera(np.float32(math.radians(45.0)), np.float32(w / h), np.float32(0.01), np.float32(1000.0))
    ev = lambda: torch.cuda.Event(enable_timing=True)
    t6, t7, t8 = [], [], []
    for it in range(iters + 3):
        e = [ev() for _ in range(4)]
        e[0].record()
        vp, bvh, idx = vlib.construct_point_light_bvh(pos, lights, view)
        e[1].record()
        keys, disp, ref = vlib.find_unique_clusters(depth, None, cam)
        e[2].record()
        counts, offsets, indices, status = vlib.assign_lights(w, h, cam, keys, disp, bvh, L, idx, vp)
        e[3].record()
        torch.cuda.synchronize()
        if it >= 3:
            t6.append(e[0].elapsed_time(e[1])); t7.append(e[1].elapsed_time(e[2])); t8.append(e[2].elapsed_time(e[3]))
    st = status.cpu().numpy()
    out = {"w": w, "h": h, "lights": L, "intensity": intensity, "clusters": int(disp[0]), "assigned": int(st[0]),
           "node_tests": int(st[2]), "leaf_tests": int(st[3]),
           "a6_light_bvh_ms": float(np.median(t6)), "a7_cluster_keys_ms": float(np.median(t7)), "a8_assign_ms": float(np.median(t8)),
           "view_ms": float(np.median(t6) + np.median(t7) + np.median(t8)),
           "note": "includes per-call torch allocations of the python harness (zeros of the output buffers)"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
    main(intensity=0.01)
    main(w=1920, h=1080)
