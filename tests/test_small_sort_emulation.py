"""The single-CTA radix sort (vren_b200/csrc/small_sort_body.cuh, the kernel behind VRENB200_SORT_VARIANT_SINGLE_CTA) executed
on the HOST: tests/cpp/cta_emulator.hpp runs the unchanged kernel body with one OS thread per CUDA thread (barriers and warp
intrinsics on pthread barriers) and compares with std::stable_sort by key.  A ThreadSanitizer build plays the role of
compute-sanitizer's racecheck: a missing __syncthreads() is a data race — shown by the negative control, which removes one."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "cpp" / "small_sort_emulation.cpp"
BODY = ROOT / "vren_b200" / "csrc" / "small_sort_body.cuh"
OUT = ROOT / "build" / "emulation"


def build(name, src=SRC, tsan=False):
    OUT.mkdir(parents=True, exist_ok=True)
    exe = OUT / name
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-Wall", "-Wno-unknown-pragmas", "-pthread", str(src), "-o", str(exe)]
    if tsan:
        cmd.insert(1, "-fsanitize=thread")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and tsan and ("tsan" in r.stderr or "sanitize" in r.stderr):
        pytest.skip("ThreadSanitizer runtime not available")
    assert r.returncode == 0, r.stderr
    return exe


def run(exe, with_values, sizes):
    r = subprocess.run([str(exe), str(int(with_values))] + [str(n) for n in sizes], capture_output=True, text=True, timeout=600)
    if r.returncode == 3 and "cta_emulator: this host allows only" in r.stderr:
        pytest.skip("this host cannot run the 1024 threads of the emulated CTA")
    return r.returncode, r.stdout + r.stderr


def test_kernel_body_sorts_on_the_host():
    """1 ... 8192 elements (1-8 rows per warp, ragged last rows), five key patterns each (uniform, the reference test's reversed
    iota, few values, all equal to the padding key, half padding key), key/value and keys-only instances"""
    exe = build("small_sort_emulation")
    rc, out = run(exe, True, [1, 2, 33, 1024, 1025, 2049, 5000, 8191, 8192])
    assert rc == 0 and "ALL PASS 45 cases" in out, out
    rc, out = run(exe, False, [1, 1023, 1024, 3000, 8192])
    assert rc == 0 and "ALL PASS 25 cases" in out, out


def test_kernel_body_is_race_free_under_thread_sanitizer():
    exe = build("small_sort_emulation_tsan", tsan=True)
    rc, out = run(exe, True, [1000, 2049])
    assert rc == 0 and "ALL PASS 10 cases" in out and "ThreadSanitizer" not in out, out[-3000:]


def test_thread_sanitizer_sees_a_missing_barrier(tmp_path):
    """negative control of the race check: the same body without the barrier between the ranking and the offset step"""
    text = BODY.read_text()
    marker = "        __syncthreads();\n        // 3. counts -> offsets"
    assert text.count(marker) == 1
    tree = tmp_path / "copy"
    (tree / "tests" / "cpp").mkdir(parents=True)
    (tree / "vren_b200" / "csrc").mkdir(parents=True)
    (tree / "vren_b200" / "csrc" / BODY.name).write_text(text.replace(marker, "        // 3. counts -> offsets"))
    for f in ("cta_emulator.hpp", "small_sort_emulation.cpp"):
        shutil.copy(ROOT / "tests" / "cpp" / f, tree / "tests" / "cpp" / f)
    exe = build("small_sort_emulation_tsan_broken", src=tree / "tests" / "cpp" / "small_sort_emulation.cpp", tsan=True)
    rc, out = run(exe, True, [2049])
    assert "ThreadSanitizer: data race" in out, out[-3000:]
