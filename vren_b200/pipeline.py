"""Python mirror of vren::cluster_and_shade (steps 1-3; clustered_shading.cpp:817-1147) for the test/bench harness:
owns every intermediate buffer once (like the reference's constructor) and sequences a6 -> a7 -> a8 through the C ABI
without any per-frame allocation."""
from __future__ import annotations

import ctypes as C

import torch

from . import lib as vlib


class ClusterAndShade:
    def __init__(self, max_width=1920, max_height=1080, max_point_lights=1 << 20, max_unique_cluster_keys=1 << 17,
                 max_assigned_lights=1 << 23, device=None):
        self.lib = vlib.load()
        dev = device or torch.device("cuda", torch.cuda.current_device())
        L = max_point_lights
        u8 = lambda n: torch.empty(max(int(n), 256), dtype=torch.uint8, device=dev)
        self.max_keys, self.max_assigned = max_unique_cluster_keys, max_assigned_lights
        self.view_pos = torch.empty(L, 4, dtype=torch.float32, device=dev)
        self.bvh = u8(self.lib.vrenb200_light_bvh_buffer_bytes(L))
        self.index = u8(self.lib.vrenb200_light_index_buffer_bytes(L))
        self.cluster_keys = torch.empty(max_unique_cluster_keys, dtype=torch.int32, device=dev)
        self.dispatch_params = torch.zeros(4, dtype=torch.int32, device=dev)
        self.cluster_ref = torch.empty(max_height, max_width, dtype=torch.int32, device=dev)
        self.indices = torch.empty(max_assigned_lights, dtype=torch.int32, device=dev)
        self.counts = torch.empty(max_unique_cluster_keys, dtype=torch.int32, device=dev)
        self.offsets = torch.empty(max_unique_cluster_keys, dtype=torch.int32, device=dev)
        self.status = torch.zeros(4, dtype=torch.int32, device=dev)
        self.scratch_keys = u8(self.lib.vrenb200_find_unique_clusters_scratch_bytes(max_width, max_height))
        self.scratch_assign = u8(self.lib.vrenb200_assign_lights_scratch_bytes(max_unique_cluster_keys, max_assigned_lights))

    def __call__(self, width, height, camera: vlib.Camera, view16, depth, normals, positions, lights, light_count, stream=None):
        """enqueues steps 1-3 on the current stream; every argument is a device tensor or a scalar"""
        lib = self.lib
        s = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        if light_count > 0:   # clustered_shading.cpp:979
            view = (C.c_float * 16)(*[float(v) for v in view16])
            vlib.check(lib.vrenb200_construct_point_light_bvh(s, positions.data_ptr(), lights.data_ptr(), light_count, C.cast(view, C.c_void_p),
                                                             self.view_pos.data_ptr(), self.bvh.data_ptr(), self.index.data_ptr(), 0, 0),
                       "construct_point_light_bvh")
        vlib.check(lib.vrenb200_find_unique_clusters(s, depth.data_ptr(), 0 if normals is None else normals.data_ptr(), width, height,
                                                     C.byref(camera), self.cluster_keys.data_ptr(), self.max_keys,
                                                     self.dispatch_params.data_ptr(), self.cluster_ref.data_ptr(),
                                                     self.scratch_keys.data_ptr(), self.scratch_keys.numel()), "find_unique_clusters")
        root = lib.vrenb200_calc_bvh_root_index(light_count)
        vlib.check(lib.vrenb200_assign_lights(s, width, height, C.byref(camera), self.cluster_keys.data_ptr(), self.dispatch_params.data_ptr(),
                                              self.max_keys, self.bvh.data_ptr(), root, light_count, self.index.data_ptr(),
                                              self.view_pos.data_ptr(), self.indices.data_ptr(), self.max_assigned,
                                              self.counts.data_ptr(), self.offsets.data_ptr(), self.status.data_ptr(),
                                              self.scratch_assign.data_ptr(), self.scratch_assign.numel()), "assign_lights")

    def capture(self, width, height, camera: vlib.Camera, view16, depth, normals, positions, lights, light_count):
        """Records steps 1-3 once into a CUDA graph and returns it: `graph.replay()` re-runs the chain on the contents the
        buffers hold at that moment (new depth / light positions every frame, same buffers, same camera and view matrix).
        A dozen small launches and memsets per view become one graph launch."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            # first call outside the capture: one-time allocations (device copies of the camera tables, kernel attributes)
            self(width, height, camera, view16, depth, normals, positions, lights, light_count)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                self(width, height, camera, view16, depth, normals, positions, lights, light_count)
        torch.cuda.current_stream().wait_stream(side)
        return graph


class ViewBatch:
    """Batched multi-view clustered shading, one view per GPU (the reference runs one view per frame:
    cluster_and_shade::operator(), clustered_shading.cpp:924-1172).

    Views are independent — each has its own depth buffer, normals, camera and view matrix — so view v runs on rank
    v % world with no exchange during the pass; the point lights of the frame are broadcast once (NCCL) from the rank that
    holds them.  Every rank owns one ClusterAndShade (buffers allocated once) and runs its views back to back on its
    stream.  With world == 1 (or no process group) all views run on this GPU."""

    def __init__(self, width, height, max_point_lights, group=None, **limits):
        import torch.distributed as dist

        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.width, self.height = width, height
        self.cs = ClusterAndShade(width, height, max_point_lights=max_point_lights, **limits)
        self.positions = self.lights = None
        self.light_count = 0

    def my_views(self, num_views: int) -> list[int]:
        return list(range(self.rank, num_views, self.world))

    def set_lights(self, positions, lights, light_count: int, src: int = 0):
        """positions / lights: float32 [L,4] device tensors (contents matter on rank `src` only): broadcast once per frame"""
        import torch.distributed as dist

        if self.world > 1:
            dist.broadcast(positions, src, group=self.group)
            dist.broadcast(lights, src, group=self.group)
        self.positions, self.lights, self.light_count = positions, lights, light_count

    def __call__(self, frames: dict, on_view=None):
        """frames: {view index: (camera, view16, depth, normals or None)} for THIS rank's views.  Runs them in view order on
        the current stream; on_view(view index, cluster_and_shade) is called after each view has been enqueued (copy out what
        you need: the buffers are reused by the next view).  Returns {view index: int32[8] device tensor} =
        {cluster count, 1, 1, key overflow, assigned lights, index overflow, node tests, leaf tests}."""
        out = {}
        for v in sorted(frames):
            camera, view16, depth, normals = frames[v]
            self.cs(self.width, self.height, camera, view16, depth, normals, self.positions, self.lights, self.light_count)
            out[v] = torch.cat([self.cs.dispatch_params, self.cs.status])
            if on_view is not None:
                on_view(v, self.cs)
        return out
