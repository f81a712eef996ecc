"""Synthetic inputs of the clustered-shading path: depth buffer, normals, point lights, camera.

Pure numpy, deterministic (counter-based splitmix64), shared by tests/ and bench so that every run sees the same data.
"""
from __future__ import annotations

import math

import numpy as np


def splitmix64(seed: int, n: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) + np.uint64(seed) * np.uint64(0xD1B54A32D192ED03)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def uniform(seed: int, n: int, lo: float, hi: float) -> np.ndarray:
    u = (splitmix64(seed, n) >> np.uint64(40)).astype(np.float64) / float(1 << 24)
    return (lo + u * (hi - lo)).astype(np.float32)


def depth_buffer(width: int, height: int, near: float = 0.01, far: float = 1000.0, fov_y: float = math.radians(45.0),
                 seed: int = 1, background: float = 0.10) -> np.ndarray:
    """z_view from a few tilted planes + low-frequency value noise, ~10 % pixels at the far plane (d = 1.0).
    d = f/(f-n) - f*n/((f-n)*z)  (camera.cpp:46-48)"""
    ys, xs = np.mgrid[0:height, 0:width].astype(np.float32)
    u = xs / np.float32(width) * 2 - 1
    v = ys / np.float32(height) * 2 - 1
    z = np.full((height, width), np.float32(far), np.float32)
    planes = [(6.0, 2.5, 0.8), (14.0, -4.0, 3.0), (30.0, 9.0, -6.0), (3.0, 0.2, 1.2), (55.0, -20.0, 10.0)]
    for i, (z0, ax, ay) in enumerate(planes):
        zp = np.float32(z0) + np.float32(ax) * u + np.float32(ay) * v
        mask = ((xs // 97 + ys // 61) % len(planes)) == i
        z = np.where(mask & (zp > near * 2), zp, z).astype(np.float32)
    # value noise on a coarse grid, bilinear up-sampled
    gh, gw = height // 48 + 2, width // 48 + 2
    g = uniform(seed, gh * gw, -0.5, 0.5).reshape(gh, gw)
    gy, gx = ys / 48.0, xs / 48.0
    y0, x0 = gy.astype(np.int32), gx.astype(np.int32)
    fy, fx = gy - y0, gx - x0
    noise = (g[y0, x0] * (1 - fx) * (1 - fy) + g[y0, x0 + 1] * fx * (1 - fy) + g[y0 + 1, x0] * (1 - fx) * fy + g[y0 + 1, x0 + 1] * fx * fy)
    z = (z * (1 + 0.2 * noise)).astype(np.float32)
    bg = uniform(seed + 7, (height // 16 + 1) * (width // 16 + 1), 0, 1).reshape(height // 16 + 1, width // 16 + 1)
    z = np.where(bg[(ys // 16).astype(np.int32), (xs // 16).astype(np.int32)] < background, np.float32(far), z)
    z = np.clip(z, np.float32(near), np.float32(far)).astype(np.float32)
    f, n = np.float32(far), np.float32(near)
    d = f / (f - n) - (f * n) / ((f - n) * z)
    d = np.clip(d, 0, 1).astype(np.float32)
    d[z >= far] = 1.0
    return np.ascontiguousarray(d)


def normal_buffer(width: int, height: int, seed: int = 3, zero_fraction: float = 0.1) -> np.ndarray:
    """RGBA16F world-space normals (not re-normalised, like deferred.frag:19); `zero_fraction` of pixels are (0,0,0)"""
    n = uniform(seed, width * height * 3, -1, 1).reshape(height, width, 3)
    blocks = uniform(seed + 1, (height // 8 + 1) * (width // 8 + 1) * 3, -1, 1).reshape(height // 8 + 1, width // 8 + 1, 3)
    ys, xs = np.mgrid[0:height, 0:width]
    n = 0.15 * n + blocks[ys // 8, xs // 8]
    out = np.zeros((height, width, 4), np.float16)
    out[..., :3] = n.astype(np.float16)
    zero = uniform(seed + 2, (height // 32 + 1) * (width // 32 + 1), 0, 1).reshape(height // 32 + 1, width // 32 + 1)
    out[zero[ys // 32, xs // 32] < zero_fraction] = 0
    return np.ascontiguousarray(out)


def point_lights(count: int, seed: int = 5, depth_range=(1.0, 120.0), fov_y: float = math.radians(45.0), aspect: float = 16 / 9,
                 intensity=(1.0, 1.0)):
    """lights uniform in the visible frustum slab (view space == world space for the identity view);
    returns (positions [L,4] f32, lights [L,4] f32 = rgb + intensity)"""
    z = uniform(seed, count, *depth_range)
    t = math.tan(fov_y / 2)
    x = uniform(seed + 1, count, -1, 1) * z * np.float32(t * aspect)
    y = uniform(seed + 2, count, -1, 1) * z * np.float32(t)
    pos = np.stack([x, y, z, np.ones(count, np.float32)], axis=1).astype(np.float32)
    lights = np.zeros((count, 4), np.float32)
    lights[:, :3] = uniform(seed + 3, count * 3, 0, 1).reshape(count, 3)
    lights[:, 3] = uniform(seed + 4, count, *intensity) if intensity[0] != intensity[1] else np.float32(intensity[0])
    return np.ascontiguousarray(pos), np.ascontiguousarray(lights)


def view_matrix(yaw: float, pitch: float, position) -> np.ndarray:
    """vren::camera::get_view (camera.cpp:11-38), column-major float32[16]; evaluated in float64: it is an INPUT"""
    cy, sy, cp, sp = math.cos(yaw), math.sin(yaw), math.cos(-pitch), math.sin(-pitch)
    ry = np.array([[cy, 0, sy, 0], [0, 1, 0, 0], [-sy, 0, cy, 0], [0, 0, 0, 1]])
    rx = np.array([[1, 0, 0, 0], [0, cp, -sp, 0], [0, sp, cp, 0], [0, 0, 0, 1]])
    t = np.eye(4)
    t[:3, 3] = -np.asarray(position, dtype=np.float64)
    v = np.linalg.inv(ry @ rx) @ t
    return np.ascontiguousarray(v.T.astype(np.float32).reshape(-1))
