"""The chained scan kernel (csrc/scan.cu::exclusive_scan_u32_kernel) executed on the HOST, one CTA after the other in an order the
test chooses (tests/cpp/cta_emulator.hpp).  The kernel text is taken from scan.cu itself (between the [[chained-scan-*]]
markers), not from a copy.  What it shows: with ticket tile ids — the scan's safe mode, vrenb200_exclusive_scan_u32_ex — ANY
start order of the CTAs gives the exclusive scan (reverse and shuffled orders here), which is exactly the property the
block-index form has to assume of the hardware; the block-index form itself is run in order, as a check of the harness against
the kernel the GPU suite verifies."""
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "build" / "emulation"


def extract(text, name):
    m = re.search(r"\[\[chained-scan-" + name + r"-begin\]\][^\n]*\n(.*?)\n[^\n]*\[\[chained-scan-" + name + r"-end\]\]", text, flags=re.S)
    assert m, name
    return m.group(1) + "\n"


def build(name, tsan=False):
    if not Path("/usr/local/cuda/include/cuda_runtime.h").exists():
        pytest.skip("CUDA headers not found (uint4, host types of common.cuh)")
    inc = OUT / "scan_inc"
    inc.mkdir(parents=True, exist_ok=True)
    text = (ROOT / "vren_b200" / "csrc" / "scan.cu").read_text()
    (inc / "scan_chained_defs.inc").write_text(extract(text, "defs"))
    body = extract(text, "body")
    assert "blockIdx.x" in body and "atomicAdd(&state->ticket, 1u)" in body and "st_relaxed_u64" in body
    (inc / "scan_chained_body.inc").write_text(body)
    out = OUT / name
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-Wall", "-Wno-unknown-pragmas", "-Wno-attributes", "-Wno-unused-variable", "-pthread",
           "-I/usr/local/cuda/include", f"-I{inc}", str(ROOT / "tests" / "cpp" / "scan_emulation.cpp"), "-o", str(out)]
    if tsan:
        cmd.insert(1, "-fsanitize=thread")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and tsan and ("tsan" in r.stderr or "sanitize" in r.stderr):
        pytest.skip("ThreadSanitizer runtime not available")
    assert r.returncode == 0, r.stderr
    return out


@pytest.fixture(scope="module")
def exe():
    return build("scan_emulation")


def run(exe, ticket, order, misalign, sizes):
    r = subprocess.run([str(exe), str(int(ticket)), order, str(misalign)] + [str(n) for n in sizes], capture_output=True, text=True, timeout=600)
    if r.returncode == 3:
        pytest.skip("this host cannot run the 256 threads of the emulated CTA")
    return r.returncode, r.stdout + r.stderr


SIZES = [1, 100, 8192, 8193, 3 * 8192 + 5, 100000]          # 1 ... 13 tiles of 8192 elements, ragged tails


@pytest.mark.parametrize("order", ["forward", "reverse", "shuffle"])
def test_ticket_tile_ids_give_the_scan_in_any_cta_order(exe, order):
    rc, out = run(exe, True, order, 0, SIZES)
    assert rc == 0 and f"ALL PASS {len(SIZES)} cases" in out, out
    rc, out = run(exe, True, order, 3, [8193, 50001])       # pointers that are not 16-byte aligned: the guarded scalar path
    assert rc == 0 and "ALL PASS 2 cases" in out, out


def test_block_index_tile_ids_in_order(exe):
    rc, out = run(exe, False, "forward", 0, SIZES)
    assert rc == 0 and f"ALL PASS {len(SIZES)} cases" in out, out


def test_ticket_form_is_race_free_under_thread_sanitizer():
    """the ticket prologue (thread 0 takes the ticket, barrier, everybody reads it) and the rest of the body under ThreadSanitizer"""
    rc, out = run(build("scan_emulation_tsan", tsan=True), True, "shuffle", 0, [3 * 8192 + 5])
    assert rc == 0 and "ALL PASS 1 cases" in out and "ThreadSanitizer" not in out, out[-3000:]
