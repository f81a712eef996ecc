#!/usr/bin/env python
"""TEST INFRASTRUCTURE: oracle/_ref — the few reference functions of the path that compile standalone.

The reference as a whole cannot be built here (Vulkan SDK, glslc, glm, volk, VMA, gtest ... are absent and its CMake
fetches from the network), but three pure-CPU pieces can be compiled FROM THE SOURCES WHERE THEY LIE:
  * vren/vren/base/base.hpp:32-79                                 integer helpers (double log/pow formulation)
  * vren/vren/primitives/build_bvh.cpp:101-136                    calc_bvh_* sizing functions
  * vren_test/vren_test/primitives/reduce.cpp:72-87               run_cpu_reduce, the test's CPU tree reduce
This script reads those line ranges (checked by anchor strings, never committed), wraps them with a 20-line glm shim of
our own (glm is not installed) and compiles oracle/_ref/libvrenref.so.  Only tests/ uses it, to cross-check the oracle.
Nothing is written outside oracle/_ref/, which is git-ignored.
"""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent / "_ref"

SHIM = r"""
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <limits>
#include <type_traits>
namespace glm {   // minimal stand-in for the glm calls the extracted lines make
    inline double log(double v) { return std::log(v); }
    inline double floor(double v) { return std::floor(v); }
    inline double ceil(double v) { return std::ceil(v); }
    template <typename B> inline double pow(B b, double e) { return std::pow((double) b, e); }
    template <typename T> inline T log2(T v)            // glm/gtc/integer.hpp: floor(log2) for integers
    {
        if constexpr (std::is_integral_v<T>) { T r = 0; while (v >>= 1) r++; return r; }
        else return std::log2(v);
    }
    template <typename T> inline T max(T a, T b) { return a < b ? b : a; }
    template <typename T> inline T min(T a, T b) { return b < a ? b : a; }
}
namespace vren {
    struct bvh_node { float m_min[3]; uint32_t m_next; float m_max[3]; uint32_t _pad; };
    uint32_t calc_bvh_padded_leaf_count(uint32_t leaf_count);
    uint32_t calc_bvh_buffer_length(uint32_t leaf_count);
    size_t calc_bvh_buffer_size(uint32_t leaf_count);
    uint32_t calc_bvh_root_index(uint32_t leaf_count);
    uint32_t calc_bvh_level_count(uint32_t leaf_count);
"""

WRAPPERS = r"""
extern "C" {
uint32_t ref_round_to_next_power_of_2(uint32_t v) { return vren::round_to_next_power_of_2(v); }
int ref_is_power_of_2(uint32_t v) { return vren::is_power_of_2(v); }
uint64_t ref_round_to_next_multiple_of(uint64_t v, uint64_t m) { return vren::round_to_next_multiple_of<uint64_t>(v, m); }
int ref_is_power_of(uint32_t n, uint32_t b) { return vren::is_power_of<uint32_t>(n, b); }
uint32_t ref_round_to_next_power_of(uint32_t n, uint32_t b) { return vren::round_to_next_power_of<uint32_t>(n, b); }
uint32_t ref_divide_and_ceil(uint32_t v, uint32_t d) { return vren::divide_and_ceil(v, d); }
uint32_t ref_calc_bvh_padded_leaf_count(uint32_t n) { return vren::calc_bvh_padded_leaf_count(n); }
uint32_t ref_calc_bvh_buffer_length(uint32_t n) { return vren::calc_bvh_buffer_length(n); }
uint64_t ref_calc_bvh_buffer_size(uint32_t n) { return vren::calc_bvh_buffer_size(n); }
uint32_t ref_calc_bvh_root_index(uint32_t n) { return vren::calc_bvh_root_index(n); }
uint32_t ref_calc_bvh_level_count(uint32_t n) { return vren::calc_bvh_level_count(n); }
// op: 0 add, 1 min, 2 max (glm::min / glm::max argument order of the test, reduce.cpp:92-97)
void ref_run_cpu_reduce_u32(int op, uint32_t* data, uint32_t length)
{
    run_cpu_reduce<uint32_t>(data, length, [op](uint32_t const& a, uint32_t const& b) -> uint32_t {
        return op == 0 ? a + b : op == 1 ? glm::min(a, b) : glm::max(a, b); });
}
void ref_run_cpu_reduce_f32(int op, float* data, uint32_t length)
{
    run_cpu_reduce<float>(data, length, [op](float const& a, float const& b) -> float {
        return op == 0 ? a + b : op == 1 ? glm::min(a, b) : glm::max(a, b); });
}
}
"""


def lines(path: Path, first: int, last: int, anchors: list[str]) -> str:
    text = path.read_text().splitlines()
    chunk = "\n".join(text[first - 1:last])
    for a in anchors:
        if a not in chunk:
            raise SystemExit(f"ref_extract: anchor {a!r} not found in {path}:{first}-{last}; the reference changed")
    return chunk


def main() -> int:
    if not REF.exists():
        print("ref_extract: /root/reference not present, nothing to do")
        return 0
    lib = OUT / "libvrenref.so"
    force = "--force" in sys.argv
    if lib.exists() and not force:
        return 0
    OUT.mkdir(parents=True, exist_ok=True)
    base = lines(REF / "vren/vren/base/base.hpp", 32, 79, ["is_power_of_2", "round_to_next_power_of_2", "divide_and_ceil", "round_to_next_power_of"])
    bvh = lines(REF / "vren/vren/primitives/build_bvh.cpp", 101, 136, ["calc_bvh_padded_leaf_count", "calc_bvh_level_count"])
    red = lines(REF / "vren_test/vren_test/primitives/reduce.cpp", 72, 87, ["run_cpu_reduce", "operation(data[a], data[b])"])
    src = OUT / "extracted.cpp"
    src.write_text(SHIM + base + "\n}\n" + bvh + "\n" + red + "\n" + WRAPPERS)
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-w", "-o", str(lib), str(src)], check=True)
    print(lib)
    return 0


if __name__ == "__main__":
    sys.exit(main())
