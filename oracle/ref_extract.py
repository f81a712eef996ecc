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

Second library, oracle/_ref/libvrenref_glsl.so — the pure functions of the reference's COMPUTE SHADERS that decide every
integer output of the clustered-shading pass (and of the rows next to it), compiled by g++ through oracle/glsl_shim.hpp
(our own vec/mat types and builtins; its header states which builtin precisions are ours) from the GLSL where it lies:
  * resources/shaders/clustered_shading.glsl:7-43, 58-110          discretize_normal, decode_cluster_key, calc_cluster_aabb
  * clustered_shading/assign_lights.comp:84-119                    test_aabb_aabb, test_sphere_aabb, get_node_address
  * clustered_shading/discretize_point_light_positions.comp:31-39  Morton code of a view-space position
  * clustered_shading/find_unique_clusters.comp:52-76              depth -> view z -> slice -> cluster key
  * clustered_shading/point_light_position_to_view_space.comp:30   K9 statement
  * clustered_shading/init_light_array_bvh.comp:52-53              leaf box of a light
  * depth_buffer_reduce.comp:12-34 (n2), vren_demo bounce_point_lights.comp:37-71 (n3), show_clusters.comp:111-116 (n1)
The GLSL text is made C++ by three mechanical rewrites, applied at extraction time and nowhere stored: floating literals
get an `f` suffix (a GLSL `1.0` is a float), `out T name` parameters become `T& name`, and statement ranges taken from
inside a main() are wrapped in a function whose parameters are the shader inputs those statements read.
"""
from __future__ import annotations

import re
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


SHADERS = REF / "vren/resources/shaders"
DEMO_SHADERS = REF / "vren_demo/resources/shaders"


def glsl(text: str) -> str:
    """the mechanical GLSL -> C++ rewrites (see the module docstring)"""
    text = re.sub(r"(?<![\w.])(\d+\.\d*(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])", r"\1f", text)   # 1.0 -> 1.0f
    text = re.sub(r"\b(?:in)?out\s+(\w+)\s+(\w+)", r"\1& \2", text)                                  # out uvec3 v -> uvec3& v
    return text


GLSL_WRAPPERS = r"""
}  // namespace glsl
using namespace glsl;
extern "C" {
void refglsl_set_inverse_override(const float* m16)   // NULL: generic cofactor inverse of the shim
{
    static mat4 m;
    if (m16) { for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) m.c[c][r] = m16[c * 4 + r]; inverse_override() = &m; }
    else inverse_override() = nullptr;
}
void refglsl_inverse(const float* m16, float* out16)
{
    mat4 m; for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) m.c[c][r] = m16[c * 4 + r];
    const mat4 i = inverse(m);
    for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) out16[c * 4 + r] = i.c[c][r];
}
void refglsl_discretize_normal(const float* n3, uint32_t count, uint32_t* out)
{
    for (uint32_t i = 0; i < count; i++) out[i] = clustered_shading_discretize_normal(vec3(n3[3 * i], n3[3 * i + 1], n3[3 * i + 2]));
}
void refglsl_decode_cluster_key(const uint32_t* keys, uint32_t count, uint32_t* out4)
{
    for (uint32_t i = 0; i < count; i++)
    {
        uvec3 ijk; uint nidx;
        clustered_shading_decode_cluster_key(keys[i], ijk, nidx);
        out4[4 * i] = ijk.x; out4[4 * i + 1] = ijk.y; out4[4 * i + 2] = ijk.z; out4[4 * i + 3] = nidx;
    }
}
void refglsl_calc_cluster_aabb(const uint32_t* ijk3, uint32_t count, uint32_t tiles_x, uint32_t tiles_y, float camera_near, float camera_half_fov,
                               const float* proj16, float* out_min4, float* out_max4)
{
    mat4 p; for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) p.c[c][r] = proj16[c * 4 + r];
    for (uint32_t i = 0; i < count; i++)
    {
        vec4 mn, mx;
        clustered_shading_calc_cluster_aabb(uvec3(ijk3[3 * i], ijk3[3 * i + 1], ijk3[3 * i + 2]), uvec2(tiles_x, tiles_y), camera_near, camera_half_fov, p, mn, mx);
        for (int k = 0; k < 4; k++) { out_min4[4 * i + k] = mn[k]; out_max4[4 * i + k] = mx[k]; }
    }
}
void refglsl_test_aabb_aabb(const float* b12, uint32_t count, uint8_t* out)
{
    for (uint32_t i = 0; i < count; i++)
    {
        const float* b = b12 + 12 * i;
        out[i] = test_aabb_aabb(vec3(b[0], b[1], b[2]), vec3(b[3], b[4], b[5]), vec3(b[6], b[7], b[8]), vec3(b[9], b[10], b[11])) ? 1 : 0;
    }
}
void refglsl_test_sphere_aabb(const float* s10, uint32_t count, uint8_t* out)
{
    for (uint32_t i = 0; i < count; i++)
    {
        const float* s = s10 + 10 * i;
        out[i] = test_sphere_aabb(vec3(s[0], s[1], s[2]), s[3], vec3(s[4], s[5], s[6]), vec3(s[7], s[8], s[9])) ? 1 : 0;
    }
}
int32_t refglsl_get_node_address(uint32_t bvh_root_idx, uint32_t level, const uint32_t* level_overlaps4)
{
    for (int i = 0; i < 4; i++) g_level_overlaps[i] = level_overlaps4[i];
    return get_node_address(bvh_root_idx, level);
}
void refglsl_morton_code(const float* pos3, uint32_t count, const float* min4, const float* max4, uint32_t* out)
{
    const vec4 mn(min4[0], min4[1], min4[2], min4[3]), mx(max4[0], max4[1], max4[2], max4[3]);
    for (uint32_t i = 0; i < count; i++) out[i] = ref_morton_code(vec3(pos3[3 * i], pos3[3 * i + 1], pos3[3 * i + 2]), mn, mx);
}
// per fragment: uv = frag_coord (x, y), depth; normals3 may be NULL (-> vec3(0)); out: key, view-space z
void refglsl_cluster_key(const float* uv2, const float* depth, const float* normals3, uint32_t count, uint32_t wg_x, uint32_t wg_y, uint32_t num_wg_y,
                         float camera_near, float camera_half_fov_y, const float* proj16, uint32_t* out_key, float* out_view_z)
{
    mat4 p; for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) p.c[c][r] = proj16[c * 4 + r];
    for (uint32_t i = 0; i < count; i++)
    {
        const vec3 n = normals3 ? vec3(normals3[3 * i], normals3[3 * i + 1], normals3[3 * i + 2]) : vec3(0);
        out_key[i] = ref_cluster_key(vec2(uv2[2 * i], uv2[2 * i + 1]), depth[i], uvec3(wg_x, wg_y, 0), uvec3(0, num_wg_y, 1), n, camera_near,
                                     camera_half_fov_y, p, out_view_z[i]);
    }
}
void refglsl_position_to_view_space(const float* view16, const float* pos4, uint32_t count, float* out4)
{
    mat4 v; for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) v.c[c][r] = view16[c * 4 + r];
    for (uint32_t i = 0; i < count; i++)
    {
        const vec4 o = ref_position_to_view_space(v, vec4(pos4[4 * i], pos4[4 * i + 1], pos4[4 * i + 2], pos4[4 * i + 3]));
        for (int k = 0; k < 4; k++) out4[4 * i + k] = o[k];
    }
}
void refglsl_light_leaf_box(const float* view_pos4, const float* intensity, uint32_t count, float* out_min3, float* out_max3)
{
    for (uint32_t i = 0; i < count; i++)
    {
        vec3 mn, mx;
        ref_light_leaf_box(vec4(view_pos4[4 * i], view_pos4[4 * i + 1], view_pos4[4 * i + 2], view_pos4[4 * i + 3]), intensity[i], mn, mx);
        for (int k = 0; k < 3; k++) { out_min3[3 * i + k] = mn[k]; out_max3[3 * i + k] = mx[k]; }
    }
}
// n2: one level of the depth pyramid (every invocation of a (to_w x to_h) dispatch)
void refglsl_depth_reduce(const float* from, int from_w, int from_h, float* to, int to_w, int to_h)
{
    from_image.data = const_cast<float*>(from); from_image.width = from_w; from_image.height = from_h;
    to_image.data = to; to_image.width = to_w; to_image.height = to_h;
    for (int y = 0; y < to_h; y++)
        for (int x = 0; x < to_w; x++) { gl_GlobalInvocationID = uvec3((uint) x, (uint) y, 0); ref_depth_reduce_invocation(); }
}
// n3: every light of one bounce_point_lights dispatch, in place
void refglsl_bounce_point_lights(float* positions4, float* directions4, uint32_t count, const float* lo3, const float* hi3, float speed_, float dt_)
{
    point_light_positions.data = reinterpret_cast<vec4*>(positions4); point_light_positions.count = count;
    point_lights_directions.data = reinterpret_cast<vec4*>(directions4); point_lights_directions.count = count;
    aabb_min = vec3(lo3[0], lo3[1], lo3[2]); aabb_max = vec3(hi3[0], hi3[1], hi3[2]); speed = speed_; dt = dt_;
    for (uint32_t i = 0; i < count; i++) { gl_GlobalInvocationID = uvec3(i, 0, 0); ref_bounce_invocation(); }
}
// n1: XOR of the light indices of a cluster (show_clusters.comp, LIGHT_ASSIGNMENT_INDICES mode)
void refglsl_light_list_xor(const uint32_t* counts, const uint32_t* offsets, const uint32_t* indices, uint32_t index_count, uint32_t cluster_count, uint32_t* out)
{
    assigned_light_counts.data = const_cast<uint32_t*>(counts); assigned_light_counts.count = cluster_count;
    assigned_light_offsets.data = const_cast<uint32_t*>(offsets); assigned_light_offsets.count = cluster_count;
    assigned_light_indices.data = const_cast<uint32_t*>(indices); assigned_light_indices.count = index_count;
    for (uint32_t c = 0; c < cluster_count; c++) out[c] = ref_light_list_xor(c);
}
}
"""


def build_glsl(force: bool) -> None:
    lib = OUT / "libvrenref_glsl.so"
    shim = Path(__file__).resolve().parent / "glsl_shim.hpp"
    if lib.exists() and not force and lib.stat().st_mtime >= max(shim.stat().st_mtime, Path(__file__).stat().st_mtime):
        return
    cs = SHADERS / "clustered_shading.glsl"
    al = SHADERS / "clustered_shading/assign_lights.comp"
    parts = ['#include "../glsl_shim.hpp"\nnamespace glsl {\n#define VREN_MAX_POINT_LIGHTS_BVH_DEPTH 4\n']
    parts.append(glsl(lines(cs, 7, 43, ["clustered_shading_discretize_normal", "disc_normal &= 0x3F"])))
    parts.append(glsl(lines(cs, 58, 110, ["clustered_shading_decode_cluster_key", "clustered_shading_calc_cluster_aabb", "cluster_far / d2.z * d2"])))
    parts.append(glsl(lines(al, 84, 119, ["test_aabb_aabb", "test_sphere_aabb", "get_node_address", "findLSB(g_level_overlaps[i])"])))
    parts.append("uint ref_morton_code(vec3 pos, vec4 _min, vec4 _max)\n{\n" +
                 glsl(lines(SHADERS / "clustered_shading/discretize_point_light_positions.comp", 31, 39, ["discretized_pos", "morton_code |="])) +
                 "\nreturn morton_code;\n}")
    fu = SHADERS / "clustered_shading/find_unique_clusters.comp"
    parts.append("uint ref_cluster_key(vec2 frag_coord, float frag_z, uvec3 gl_WorkGroupID, uvec3 gl_NumWorkGroups, vec3 frag_normal, float camera_near,\n"
                 "                     float camera_half_fov_y, mat4 camera_projection, float& out_view_z)\n{\n" +
                 glsl(lines(fu, 52, 65, ["inverse(camera_projection) * frag_pos", "frag_pos /= frag_pos.w", "log(frag_pos.z / camera_near)"])) +
                 "\nout_view_z = frag_pos.z;\n" +
                 glsl(lines(fu, 69, 76, ["clustered_shading_discretize_normal(frag_normal)", "frag_normal_discretized << 26"])) +
                 "\nreturn cluster_key;\n}")
    k9 = lines(SHADERS / "clustered_shading/point_light_position_to_view_space.comp", 30, 30, ["push_constants.camera_view * vec4("])
    k9 = k9.split("=", 1)[1].replace("push_constants.camera_view", "camera_view").replace("point_light_positions[gl_GlobalInvocationID.x]", "position")
    parts.append("vec4 ref_position_to_view_space(mat4 camera_view, vec4 position)\n{\nreturn " + glsl(k9) + "\n}")
    k11 = lines(SHADERS / "clustered_shading/init_light_array_bvh.comp", 52, 53, ["node._min =", "node._max ="])
    k11 = k11.replace("view_space_point_light_positions[point_light_idx]", "view_pos").replace("point_light.intensity", "intensity")
    parts.append("struct BvhNode { vec3 _min; uint next; vec3 _max; uint _pad; };\n"
                 "void ref_light_leaf_box(vec4 view_pos, float intensity, vec3& out_min, vec3& out_max)\n{\nBvhNode node;\n" + glsl(k11) +
                 "\nout_min = node._min; out_max = node._max;\n}")
    parts.append("uvec3 gl_GlobalInvocationID;\nimage2D from_image, to_image;\nvoid ref_depth_reduce_invocation()\n{\n" +
                 glsl(lines(SHADERS / "depth_buffer_reduce.comp", 12, 34, ["imageSize(from_image)", "max_depth = max(max_depth, depth)", "imageStore(to_image"])) + "\n}")
    parts.append("buffer_array<vec4> point_light_positions, point_lights_directions;\nvec3 aabb_min, aabb_max; float speed, dt;\n"
                 "#define MAX_BOUNCING_ITER 32\nvoid ref_bounce_invocation()\n{\n" +
                 glsl(lines(DEMO_SHADERS / "bounce_point_lights.comp", 37, 71, ["p = min(max(p, aabb_min), aabb_max)", "float step = min(min_t - EPS, rem_t)", "rem_t -= step", "= vec4(d, 0)"])) + "\n}")
    parts.append("buffer_array<uint> assigned_light_counts, assigned_light_offsets, assigned_light_indices;\n"
                 "uint ref_light_list_xor(uint cluster_key_idx)\n{\n" +
                 glsl(lines(DEMO_SHADERS / "show_clusters.comp", 111, 116, ["_hash ^= point_light_idx"])) + "\nreturn _hash;\n}")
    src = OUT / "extracted_glsl.cpp"
    src.write_text("\n".join(parts) + GLSL_WRAPPERS)
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-w", "-o", str(lib), str(src)], check=True)
    print(lib)


def main() -> int:
    if not REF.exists():
        print("ref_extract: /root/reference not present, nothing to do")
        return 0
    lib = OUT / "libvrenref.so"
    force = "--force" in sys.argv
    OUT.mkdir(parents=True, exist_ok=True)
    build_glsl(force)
    if lib.exists() and not force:
        return 0
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
