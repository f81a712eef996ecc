"""n3 — vren_demo::point_light_bouncer (bounce_point_lights.comp:33-73): oracle properties on the CPU, bit-exact parity
of the CUDA kernel against the oracle on the GPU."""
import numpy as np
import pytest

import oracle

LO = np.array([-1.0, -2.0, -3.0], np.float32)
HI = np.array([4.0, 2.0, 1.0], np.float32)


def make_lights(L, seed, inside=True):
    rng = np.random.default_rng(seed)
    span = (HI - LO)
    p = rng.uniform(LO - (0 if inside else 1) * span, HI + (0 if inside else 1) * span, (L, 3))
    d = rng.normal(size=(L, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    pos = np.concatenate([p, rng.uniform(0, 5, (L, 1))], 1).astype(np.float32)      # w is overwritten by the shader
    dirs = np.concatenate([d, rng.uniform(0, 5, (L, 1))], 1).astype(np.float32)
    return pos, dirs


def test_oracle_zero_time_only_clamps():
    pos, dirs = make_lights(257, 1, inside=False)
    p, d = oracle.bounce_point_lights(pos, dirs, LO, HI, 3.0, 0.0)
    assert np.array_equal(p[:, :3], np.minimum(np.maximum(pos[:, :3], LO), HI))       # :40
    assert np.array_equal(d[:, :3], dirs[:, :3])
    assert (p[:, 3] == 1).all() and (d[:, 3] == 0).all()                                # :70-71


def test_oracle_straight_segment_without_hit():
    pos = np.array([[0, 0, 0, 1]], np.float32)
    dirs = np.array([[1, 0, 0, 0]], np.float32)
    p, d = oracle.bounce_point_lights(pos, dirs, LO, HI, 2.0, 0.5)
    assert np.array_equal(p[0], np.array([1, 0, 0, 1], np.float32)) and np.array_equal(d[0], dirs[0])


def test_oracle_reflects_off_a_face():
    pos = np.array([[3.5, 0, 0, 1]], np.float32)
    dirs = np.array([[1, 0, 0, 0]], np.float32)
    p, d = oracle.bounce_point_lights(pos, dirs, LO, HI, 1.0, 1.0)       # 0.5 to the face (minus EPS), then back
    assert d[0, 0] == -1.0
    assert abs(p[0, 0] - 3.5) < 1e-3 and p[0, 0] < 4.0


def test_oracle_stays_inside_and_keeps_speed_over_many_frames():
    pos, dirs = make_lights(2000, 2)
    p, d = pos, dirs
    for _ in range(40):
        p, d = oracle.bounce_point_lights(p, d, LO, HI, 3.0, 0.1)
    assert (p[:, :3] >= LO).all() and (p[:, :3] <= HI).all()
    assert np.array_equal(np.abs(d[:, :3]), np.abs(dirs[:, :3]))          # reflections only flip signs


def test_oracle_iteration_cap():
    # a huge step in a thin box: at most 32 reflections are followed (MAX_BOUNCING_ITER), the rest of the time is dropped
    lo, hi = np.array([0, 0, 0], np.float32), np.array([1e-3, 1, 1], np.float32)
    pos = np.array([[5e-4, 0.5, 0.5, 1]], np.float32)
    dirs = np.array([[1, 0, 0, 0]], np.float32)
    p, d = oracle.bounce_point_lights(pos, dirs, lo, hi, 1000.0, 1.0)
    assert 0 <= p[0, 0] <= 1e-3 and abs(d[0, 0]) == 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("L", [1, 255, 1024, 65536])
def test_gpu_matches_oracle_bit_exact_over_frames(L):
    import torch

    from vren_b200 import lib as vren

    pos, dirs = make_lights(L, 7 + L, inside=False)
    p_gpu, d_gpu = torch.from_numpy(pos).cuda(), torch.from_numpy(dirs).cuda()
    p_ref, d_ref = pos, dirs
    for frame in range(6):
        speed, dt = (3.0, 0.1) if frame % 2 == 0 else (40.0, 0.25)        # the second setting reflects several times per call
        vren.bounce_point_lights(p_gpu, d_gpu, LO, HI, speed, dt)
        p_ref, d_ref = oracle.bounce_point_lights(p_ref, d_ref, LO, HI, speed, dt)
        assert np.array_equal(p_gpu.cpu().numpy().view(np.uint32), p_ref.view(np.uint32)), frame
        assert np.array_equal(d_gpu.cpu().numpy().view(np.uint32), d_ref.view(np.uint32)), frame


@pytest.mark.gpu
def test_gpu_degenerate_directions_match_oracle():
    """axis-aligned directions divide by zero (+-inf / NaN face times): same canonical fmin semantics on both sides"""
    import torch

    from vren_b200 import lib as vren

    pos = np.array([[0, 0, 0, 1], [4, 2, 1, 1], [-1, -2, -3, 1], [1, 1, 0, 1]], np.float32)
    dirs = np.array([[0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 0], [0.6, 0, -0.8, 0]], np.float32)
    p_gpu, d_gpu = torch.from_numpy(pos).cuda(), torch.from_numpy(dirs).cuda()
    vren.bounce_point_lights(p_gpu, d_gpu, LO, HI, 7.0, 1.0)
    p_ref, d_ref = oracle.bounce_point_lights(pos, dirs, LO, HI, 7.0, 1.0)
    assert np.array_equal(p_gpu.cpu().numpy().view(np.uint32), p_ref.view(np.uint32))
    assert np.array_equal(d_gpu.cpu().numpy().view(np.uint32), d_ref.view(np.uint32))


@pytest.mark.gpu
def test_gpu_argument_errors():
    import ctypes as C

    import torch

    from vren_b200 import lib as vren

    lib = vren.load()
    lo, hi = (C.c_float * 3)(0, 0, 0), (C.c_float * 3)(1, 1, 1)
    buf = torch.zeros(9, dtype=torch.float32, device="cuda")
    assert lib.vrenb200_bounce_point_lights(None, None, buf.data_ptr(), 1, C.byref(lo), C.byref(hi), 1.0, 1.0) == vren.EINVAL_ARG
    assert lib.vrenb200_bounce_point_lights(None, buf.data_ptr() + 4, buf.data_ptr(), 1, C.byref(lo), C.byref(hi), 1.0, 1.0) == vren.EALIGN
    assert lib.vrenb200_bounce_point_lights(None, buf.data_ptr(), buf.data_ptr(), 0, C.byref(lo), C.byref(hi), 1.0, 1.0) == vren.OK
