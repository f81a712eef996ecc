"""bench.py's contract, as far as it can be checked without a GPU: the reference arm (the oracle's restatement of the
reference's CPU check, all host threads) prints ONE JSON line with the keys the driver reads, on the byte-identical `config`
object of the GPU arm (same_config), and only on rank 0; the GPU arm refuses to run without a device."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def run(args, **env):
    e = dict(os.environ)
    e.pop("RANK", None)
    e.pop("WORLD_SIZE", None)
    e.update(env)
    return subprocess.run([sys.executable, str(ROOT / "bench.py")] + args, capture_output=True, text=True, timeout=300, env=e, cwd=str(ROOT))


def test_reference_arm_line():
    sys.path.insert(0, str(ROOT))
    import bench

    r = run(["--impl", "reference", "--log2n", "16", "--steps", "2", "--warmup", "1", "--gpus", "1"])
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"] == bench.config_object(16)                       # what the GPU arm prints for the same size
    assert json.dumps(d["config"]) == json.dumps(bench.config_object(16))
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "2^16" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None


def test_reference_arm_runs_on_rank_0_only():
    r = run(["--impl", "reference", "--log2n", "12", "--steps", "1", "--warmup", "0", "--gpus", "2"], RANK="1", WORLD_SIZE="2")
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        return
    r = run(["--steps", "1", "--warmup", "0", "--log2n", "12"])
    assert r.returncode != 0 and "no CUDA device" in (r.stdout + r.stderr)


def test_cpu_baselines_of_the_secondary_rows():
    """the reference's CPU checks timed beside the GPU rows (scan, reduce tree, std::sort, bucket sort, BuildBVH, one view
    of the light assignment) — the function runs without a GPU, here on tiny sizes"""
    sys.path.insert(0, str(ROOT))
    import bench

    r = bench.cpu_secondary_baselines(small=True)
    for row in ("scan_u32", "reduce_u32_add_tree", "radix_sort_keys", "bucket_sort_uvec2", "build_bvh"):
        assert r[row]["ms"] > 0 and "sample" in r[row]
    la = r["light_assign"]
    assert la["ms_per_view"] > 0 and la["clusters"] > 0 and la["assigned_lights"] > 0
    json.dumps(r)


def _line(stdout):
    lines = [l for l in stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, stdout[-2000:]
    return json.loads(lines[0])


EXPECTED_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                 "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks", "secondary"}


def test_gpu_arm_control_flow_without_a_gpu_single():
    """tests/bench_dry_run.py: the whole N = 1 flow of bench.py — timed loop, checks, both e2e forms, CPU baselines, every secondary
    row, assembly of the line — executed on CPU tensors with numpy stand-ins for the device entry points.  No row may fail for a
    reason of its own (the CUDA-graph capture cannot work without a driver and says so)."""
    sys.path.insert(0, str(ROOT))
    import bench

    r = subprocess.run([sys.executable, str(ROOT / "tests" / "bench_dry_run.py"), "--log2n", "14", "--steps", "3", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-3000:]
    d = _line(r.stdout)
    assert EXPECTED_KEYS <= set(d) and d["n_gpus"] == 1 and d["steps"] == 3 and d["config"] == bench.config_object(14)
    assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and d["roofline"]["bound"] == "hbm"
    assert d["cpu_baseline"]["kind"] == "port" and "single_thread" in d["cpu_baseline"]
    assert d["e2e"]["h2d_bytes_per_step"] == 8 << 14 and d["e2e"]["d2h_bytes_per_step"] == 8 << 14 and d["e2e"]["value"] > 0
    sec = d["secondary"]
    assert "error" not in sec, sec.get("error")
    for name in ("scan_u32_2p28", "reduce_u32_add_final_2p28", "radix_sort_keys_2p28", "radix_sort_pairs_2p28_by_config", "radix_sort_keys_2p10", "radix_sort_keys_2p20", "bucket_sort_2p26_uvec2",
                 "build_bvh_2p20_leaves", "light_assign_4k_65536_lights", "light_list_consumer_4k", "depth_pyramid_4k", "bounce_point_lights_65536",
                 "visualize_bvh_2p20_leaves", "radix_sort_single_cta", "scan_u32_2p28_safe_mode"):
        assert name in sec and "error" not in sec[name], (name, sec.get(name))
    assert isinstance(sec["radix_sort_single_cta"]["us_2p10_keys"], float) and isinstance(sec["scan_u32_2p28_safe_mode"]["ms"], float)
    assert set(d["secondary_cpu_baseline"]) >= {"scan_u32", "radix_sort_keys", "light_assign"}


def test_gpu_arm_control_flow_without_a_gpu_two_ranks():
    """the N = 2 flow over gloo with a stand-in for the multi-GPU sort: oracle comparison before the timed steps, checksums and
    boundary checks after them, both forms of the host-shard e2e leg, the four multi-GPU secondary rows, one line from rank 0"""
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(ROOT / "tests" / "bench_dry_run.py"), "--gpus", "2", "--log2n", "22", "--steps", "2", "--warmup", "3"],
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, cwd=str(ROOT)))
    outs = [p.communicate(timeout=900) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs[0][1][-3000:] + outs[1][1][-3000:]
    assert outs[1][0].strip() == ""                                   # rank 0 alone prints
    d = _line(outs[0][0])
    assert EXPECTED_KEYS <= set(d) and d["n_gpus"] == 2 and d["scaling"] == "weak" and d["cpu_baseline"] is None
    assert d["e2e"]["h2d_bytes_per_step"] == 8 << 22 and d["e2e"]["value"] > 0 and "failed" not in d["e2e"]["issue"]
    assert d["phases_rank0_last_step"] is not None
    sec = d["secondary"]
    assert "error" not in sec, sec.get("error")
    assert set(sec) == {"sharded_scan_u32_2p28_per_gpu", "sharded_reduce_u32_2p28_per_gpu", "sharded_bucket_sort_2p26_per_gpu", "light_assign_4k_65536_lights_8_views",
                        "sharded_sort_by_rounds"}
    assert set(sec["sharded_sort_by_rounds"]) == {"rounds_1", "rounds_2", "rounds_4", "rounds_8"}
    assert sec["light_assign_4k_65536_lights_8_views"]["views"] == 8 and len(sec["light_assign_4k_65536_lights_8_views"]["clusters_per_view"]) == 8


def test_watchdog_prints_the_line_without_the_secondary_rows():
    """a secondary row that does not finish must not cost the headline: with a tiny limit the watchdog prints the line (with
    whatever rows were done) and ends the process with exit code 0"""
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "bench_dry_run.py"), "--log2n", "14", "--steps", "2", "--warmup", "3", "--no-cpu-baseline",
                        "--secondary-timeout", "0.05"], capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-3000:]
    d = _line(r.stdout)
    assert d["value"] > 0 and d["e2e"]["value"] > 0 and "abandoned" in d["secondary"]["error"]


@pytest.mark.parametrize("script,expect", [("run_small_sort.py", "ok 75 cases"), ("run_scan_safe_mode.py", "ok 8 cases")])
def test_first_hardware_run_scripts_walk_without_a_gpu(script, expect):
    """the scripts behind the non-gating first-hardware-run tests (tests/test_small_sort.py, tests/test_scan_safe_mode.py) executed
    with the stand-ins: if one of them fails on the GPU box, it is the kernel, not the script"""
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "bench_dry_run.py"), "--script", str(ROOT / "tests" / script)],
                       capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert r.returncode == 0 and expect in r.stdout, r.stdout[-1000:] + r.stderr[-3000:]
