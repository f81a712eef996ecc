"""bench.py's contract, as far as it can be checked without a GPU: the reference arm (the oracle's restatement of the
reference's CPU check, all host threads) prints ONE JSON line with the keys the driver reads, on the byte-identical `config`
object of the GPU arm (same_config), and only on rank 0; the GPU arm refuses to run without a device."""
import json
import os
import subprocess
import sys
from pathlib import Path

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
