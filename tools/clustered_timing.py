"""times the a6 -> a7 -> a8 chain at BASELINE C5 (3840x2160, 65536 lights) with CUDA events, no per-frame allocation"""
import json
import math
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from vren_b200 import lib as vlib, synthetic  # noqa: E402
from vren_b200.pipeline import ClusterAndShade  # noqa: E402


def main(w=3840, h=2160, L=65536, iters=20, intensity=1.0, normals=False):
    lib = vlib.load()
    depth = torch.from_numpy(synthetic.depth_buffer(w, h, seed=2024)).cuda()
    nrm = torch.from_numpy(synthetic.normal_buffer(w, h, seed=7)).cuda() if normals else None
    pos, lights = synthetic.point_lights(L, seed=2025, aspect=w / h, intensity=(intensity, intensity))
    pos, lights = torch.from_numpy(pos).cuda(), torch.from_numpy(lights).cuda()
    view = synthetic.view_matrix(0.0, 0.0, (0, 0, 0)).tolist()
    cam = vlib.Camera(np.float32(math.radians(45.0)), np.float32(w / h), np.float32(0.01), np.float32(1000.0))
    cs = ClusterAndShade(w, h, max_point_lights=L)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    times = []
    for it in range(iters + 3):
        e0, e1 = ev(), ev()
        e0.record()
        cs(w, h, cam, view, depth, nrm, pos, lights, L)
        e1.record()
        torch.cuda.synchronize()
        if it >= 3:
            times.append(e0.elapsed_time(e1))
    # the same chain replayed from a CUDA graph (buffers and camera fixed, contents free to change between replays)
    graph_ms = None
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            cs(w, h, cam, view, depth, nrm, pos, lights, L)      # warm-up on the capture stream
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                cs(w, h, cam, view, depth, nrm, pos, lights, L)
        torch.cuda.current_stream().wait_stream(side)
        gt = []
        for it in range(iters + 3):
            e0, e1 = ev(), ev()
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            if it >= 3:
                gt.append(e0.elapsed_time(e1))
        graph_ms = float(np.median(gt))
    except Exception as exc:  # noqa: BLE001
        graph_ms = f"capture failed: {exc}"
    st = cs.status.cpu().numpy()
    print(json.dumps({"w": w, "h": h, "lights": L, "intensity": intensity, "normals": normals, "clusters": int(cs.dispatch_params[0]),
                      "assigned": int(st[0]), "node_tests": int(st[2]), "leaf_tests": int(st[3]),
                      "view_ms_median": float(np.median(times)), "view_ms_min": float(np.min(times)), "graph_replay_ms_median": graph_ms}))


if __name__ == "__main__":
    main()
    main(intensity=0.01)
    main(normals=True)
    main(w=1920, h=1080)
