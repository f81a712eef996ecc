"""Generates tests/golden/clustered_reference_outputs.npz: outputs of the REFERENCE'S OWN GLSL — the pure functions of
the clustered-shading shaders (and of the rows next to it), compiled by g++ through oracle/glsl_shim.hpp from the
sources where they lie under /root/reference (oracle/ref_extract.py -> oracle/_ref/libvrenref_glsl.so) — on the seeded
inputs of tests/clustered_cases.py.  Per output: a SHA-256 of the whole array and its first 16384 scalars verbatim.
GLSL's inverse() is implementation-defined; the fixture is generated with the exact closed-form inverse of the
perspective matrix injected (the policy is stated in oracle/glsl_shim.hpp and DESIGN.md section 3).

The fixture travels to the GPU box (where /root/reference does not exist): tests/test_clustered_reference.py checks the
oracle and the CUDA path against it there.

usage (in the build container, repo root):  python tests/golden/make_clustered_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import clustered_cases as cc  # noqa: E402


def main():
    from vren_b200 import build as b

    b.build_oracle()
    b.build_reference_extract()
    ref = cc.evaluate("ref")
    out = {}
    for name, arr in ref.items():
        out[name + "__sha256"] = np.frombuffer(bytes.fromhex(cc.digest(arr)), np.uint8)
        out[name + "__head"] = cc.head(arr)
        out[name + "__shape"] = np.array(arr.shape, np.int64)
    path = Path(__file__).with_name("clustered_reference_outputs.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({path.stat().st_size} bytes, {len(ref)} outputs)")


if __name__ == "__main__":
    main()
