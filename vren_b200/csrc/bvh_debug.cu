// n4 — vren_demo::visualize_bvh::write (reference: vren_demo/vren_demo/visualize_bvh.cpp:59-94,
// vren_demo/resources/shaders/show_bvh.comp:62-78): every BVH node becomes the 12 edges (24 vertices of
// {vec3 position; uint color}) of its box in a debug-draw vertex buffer, one colour per level; INVALID nodes become 12
// degenerate black lines.  A consumer of the level layout build_bvh produces: level l (l = level_count .. 0, leaves first)
// holds 32^l nodes starting at the sum of the sizes of the levels before it.
//
// Reference: one dispatch per level.  Here one launch; 24 consecutive threads write the 24 vertices of a node, so a warp
// stores 512 contiguous bytes.  32 B read + 384 B written per node: a pure streaming store kernel.
#include "common.cuh"

namespace vrenb200 {
namespace {

struct debug_vertex
{
    float x, y, z;
    uint32_t color;
};

constexpr uint32_t kInvalidNode = 0xFFFFFFFEu;
// corner codes (bit0: x from max, bit1: y from max, bit2: z from max) of the 24 line endpoints, in the order of
// VREN_WRITE_DEBUG_DRAW_BUFFER_AABB (show_bvh.comp:46-60): three bits per vertex
constexpr uint64_t kCornersLo = 0ull | (1ull << 3) | (1ull << 6) | (5ull << 9) | (5ull << 12) | (4ull << 15) | (4ull << 18) | (0ull << 21) |
                                (2ull << 24) | (3ull << 27) | (3ull << 30) | (7ull << 33) | (7ull << 36) | (6ull << 39) | (6ull << 42) | (2ull << 45);
constexpr uint32_t kCornersHi = 0u | (2u << 3) | (1u << 6) | (3u << 9) | (5u << 12) | (7u << 15) | (4u << 18) | (6u << 21);   // vertices 16..23

struct level_table
{
    uint32_t first[8];   // first node index of level_count - i
    uint32_t color[8];
    uint32_t levels;     // level_count + 1
};

__global__ void __launch_bounds__(256)
visualize_bvh_kernel(const float4* __restrict__ bvh, uint64_t total_vertices, level_table lv, debug_vertex* __restrict__ vertices)
{
    const uint64_t v = (uint64_t) blockIdx.x * 256 + threadIdx.x;
    if (v >= total_vertices) return;
    const uint32_t node = (uint32_t) (v / 24), k = (uint32_t) (v % 24);
    const float4 lo = bvh[2 * (size_t) node], hi = bvh[2 * (size_t) node + 1];
    uint32_t color = 0;
#pragma unroll
    for (int i = 0; i < 8; i++)
        if (i < (int) lv.levels && node >= lv.first[i]) color = lv.color[i];
    const uint32_t corner = k < 16 ? (uint32_t) (kCornersLo >> (3 * k)) & 7u : (kCornersHi >> (3 * (k - 16))) & 7u;
    debug_vertex out;
    if (__float_as_uint(lo.w) != kInvalidNode)
    {
        out.x = (corner & 1u) ? hi.x : lo.x;
        out.y = (corner & 2u) ? hi.y : lo.y;
        out.z = (corner & 4u) ? hi.z : lo.z;
        out.color = color;
    }
    else
    {
        out.x = out.y = out.z = 0.0f;      // show_bvh.comp:75: AABB(vec3(0), vec3(0), 0x000000)
        out.color = 0u;
    }
    reinterpret_cast<float4*>(vertices)[v] = make_float4(out.x, out.y, out.z, __uint_as_float(out.color));
}

} // namespace
} // namespace vrenb200

using namespace vrenb200;

extern "C" uint64_t vrenb200_visualize_bvh_vertex_count(uint32_t level_count)
{
    uint64_t nodes = 0, n = 1;
    for (uint32_t l = 0; l <= level_count; l++, n *= 32) nodes += n;
    return nodes * 24;
}

extern "C" int vrenb200_visualize_bvh(vrenb200_stream_t stream, const void* bvh_nodes, uint32_t level_count, void* vertices)
{
    if (bvh_nodes == nullptr || vertices == nullptr) return VRENB200_EINVAL_ARG;
    if (level_count > 4) return VRENB200_ELIMIT;      // colors[level_count + 2] of a 7-entry table (visualize_bvh.cpp:74-79)
    if (((reinterpret_cast<uintptr_t>(bvh_nodes) | reinterpret_cast<uintptr_t>(vertices)) & 15) != 0) return VRENB200_EALIGN;
    static const uint32_t colors[7] = {0xff0000, 0xffff00, 0x00ff00, 0x0000ff, 0x00ffff, 0xff00ff, 0xffffff};
    level_table lv{};
    lv.levels = level_count + 1;
    uint32_t offset = 0;
    for (int32_t level = (int32_t) level_count, i = 0; level >= 0; level--, i++)
    {
        lv.first[i] = offset;
        lv.color[i] = colors[level_count - level + 2];
        offset += 1u << (5 * level);
    }
    const uint64_t total = vrenb200_visualize_bvh_vertex_count(level_count);
    const uint64_t blocks = (total + 255) / 256;
    visualize_bvh_kernel<<<(unsigned) blocks, 256, 0, as_stream(stream)>>>(static_cast<const float4*>(bvh_nodes), total, lv,
                                                                         static_cast<debug_vertex*>(vertices));
    return check_launch();
}
