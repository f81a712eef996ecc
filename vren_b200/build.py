"""In-tree build of the CUDA library (libvrenb200.so) and of the CPU oracle (oracle/liboracle.so).

nvcc cross-compiles sm_100a without a GPU; the resulting .so files travel to the GPU box with the repo
snapshot (they are git-ignored, not gpurun-ignored).  Nothing here imports torch.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "vren_b200" / "csrc"
LIB = ROOT / "vren_b200" / "libvrenb200.so"
ORACLE_DIR = ROOT / "oracle"
ORACLE_LIB = ORACLE_DIR / "liboracle.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",  # fp32 parity: no FMA contraction anywhere
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
    "-shared",
]
if os.environ.get("VRENB200_TUNING") == "1":     # experimental kernel table + process-global tuning hooks (never the shipped build)
    NVCC_FLAGS.insert(0, "-DVRENB200_TUNING")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def _newer(target: Path, sources) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(Path(s).stat().st_mtime <= t for s in sources)


def build_cuda(force: bool = False, verbose: bool = False) -> Path:
    sources = sorted(CSRC.glob("*.cu"))
    deps = sources + sorted(CSRC.glob("*.cuh")) + sorted((ROOT / "include").glob("*.h")) + [CSRC / "exports.map"]
    if not force and _newer(LIB, deps):
        return LIB
    objs = []
    build_dir = ROOT / "build" / "obj"
    build_dir.mkdir(parents=True, exist_ok=True)
    procs = []
    for src in sources:
        obj = build_dir / (src.stem + ".o")
        objs.append(obj)
        if not force and _newer(obj, [src] + sorted(CSRC.glob("*.cuh")) + sorted((ROOT / "include").glob("*.h"))):
            continue
        cmd = [_nvcc()] + [f for f in NVCC_FLAGS if f != "-shared"] + ["-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed for {src}:\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("CUDA build failed")
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xlinker", f"--version-script={CSRC / 'exports.map'}",
           "-o", str(LIB)] + [str(o) for o in objs] + ["-lcudart"]
    subprocess.run(cmd, check=True)
    return LIB


def build_oracle(force: bool = False) -> Path:
    sources = sorted(ORACLE_DIR.glob("*.cpp"))
    deps = sources + sorted(ORACLE_DIR.glob("*.h"))
    if not force and _newer(ORACLE_LIB, deps):
        return ORACLE_LIB
    # -ffp-contract=off + no -march: the oracle fixes an IEEE fp32 op sequence
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-Wall",
           "-o", str(ORACLE_LIB)] + [str(s) for s in sources] + ["-lpthread"]
    subprocess.run(cmd, check=True)
    return ORACLE_LIB


FACADE_TEST = ROOT / "build" / "facade_test"


def build_facade_test(force: bool = False) -> Path:
    """tests/cpp/facade_test.cpp: the C++ facade (include/vren/) driven like the reference's own tests"""
    src = ROOT / "tests" / "cpp" / "facade_test.cpp"
    deps = [src, LIB] + sorted((ROOT / "include").rglob("*.h*"))
    if not force and _newer(FACADE_TEST, deps):
        return FACADE_TEST
    FACADE_TEST.parent.mkdir(parents=True, exist_ok=True)
    cuda_home = Path(_nvcc()).resolve().parent.parent
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", f"-I{ROOT / 'include'}", f"-I{cuda_home / 'include'}", str(src), "-o", str(FACADE_TEST),
           f"-L{LIB.parent}", "-lvrenb200", f"-L{cuda_home / 'lib64'}", "-lcudart", f"-Wl,-rpath,{LIB.parent}",
           f"-Wl,-rpath,{cuda_home / 'lib64'}"]
    subprocess.run(cmd, check=True)
    return FACADE_TEST


SHARDED_HOST = ROOT / "build" / "sharded_sort_host"


def build_sharded_sort_host(force: bool = False) -> Path:
    """tests/cpp/sharded_sort_host.cpp: a C++ host driving the multi-GPU sort through the C ABI alone (INTEGRATION.md's sketch)"""
    src = ROOT / "tests" / "cpp" / "sharded_sort_host.cpp"
    deps = [src, LIB] + sorted((ROOT / "include").glob("*.h"))
    if not force and _newer(SHARDED_HOST, deps):
        return SHARDED_HOST
    SHARDED_HOST.parent.mkdir(parents=True, exist_ok=True)
    cuda_home = Path(_nvcc()).resolve().parent.parent
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", f"-I{ROOT / 'include'}", f"-I{cuda_home / 'include'}", str(src), "-o", str(SHARDED_HOST),
           f"-L{LIB.parent}", "-lvrenb200", f"-L{cuda_home / 'lib64'}", "-lcudart", f"-Wl,-rpath,{LIB.parent}",
           f"-Wl,-rpath,{cuda_home / 'lib64'}"]
    subprocess.run(cmd, check=True)
    return SHARDED_HOST


def build_reference_extract(force: bool = False):
    """oracle/_ref: the few reference functions that compile standalone (see oracle/ref_extract.py)."""
    script = ORACLE_DIR / "ref_extract.py"
    if not script.exists() or not Path("/root/reference").exists():
        return None
    r = subprocess.run([sys.executable, str(script)] + (["--force"] if force else []), capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        return None
    return ORACLE_DIR / "_ref" / "libvrenref.so"


if __name__ == "__main__":
    force = "--force" in sys.argv
    verbose = "-v" in sys.argv
    print(build_cuda(force=force, verbose=verbose))
    if list(ORACLE_DIR.glob("*.cpp")):
        print(build_oracle(force=force))
    print(build_reference_extract(force=force))
