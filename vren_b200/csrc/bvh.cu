// a5 — vren::build_bvh (reference: vren/vren/primitives/build_bvh.{hpp,cpp}, shaders/build_bvh.comp:32-55)
// and K11 init_light_array_bvh.comp:42-65 fused into the same launch for the light-BVH chain (a6).
//
// Layout kept bit-for-bit: implicit complete 32-ary tree, leaves [0,Lp), level k+1 right after level k, root last;
// parent = component-wise min/max over its VALID children, next = absolute index of child 0, or INVALID if all
// 32 children are invalid.
//
// Reference: one dispatch + barrier per level.  Here ONE launch builds every level bottom-up (Karras-style):
// a warp owns 32 consecutive children, reduces them with redux.sync on order-preserving integer images of the
// floats, writes the parent, then bumps an arrival counter kept in the grand-parent's `_pad` word; the 32nd
// arriver continues one level up.  Only the leaves are read from HBM once (32 B/leaf) — upper levels are L2 hits.
#include "common.cuh"

namespace vrenb200 {

namespace {

constexpr uint32_t kLeaf = VRENB200_BVH_LEAF_NODE;
constexpr uint32_t kInvalid = VRENB200_BVH_INVALID_NODE;

// order-preserving float <-> uint map (-0 < +0; NaN not supported, like the reference)
__device__ __forceinline__ uint32_t f2ord(float f)
{
    const uint32_t u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o)
{
    return __uint_as_float(o ^ ((o >> 31) ? 0x80000000u : 0xFFFFFFFFu));
}

struct node_regs
{
    float4 lo; // min.xyz, next
    float4 hi; // max.xyz, _pad
};

// parent of 32 children held one per lane. Canonical value for an all-invalid parent: empty box, _pad = 0.
__device__ __forceinline__ node_regs reduce_children(const node_regs& c, uint32_t first_child_index)
{
    const bool valid = __float_as_uint(c.lo.w) != kInvalid;
    const unsigned any = __ballot_sync(kFullMask, valid);
    uint32_t mn[3] = {f2ord(c.lo.x), f2ord(c.lo.y), f2ord(c.lo.z)};
    uint32_t mx[3] = {f2ord(c.hi.x), f2ord(c.hi.y), f2ord(c.hi.z)};
#pragma unroll
    for (int k = 0; k < 3; k++)
    {
        mn[k] = __reduce_min_sync(kFullMask, valid ? mn[k] : 0xFFFFFFFFu);
        mx[k] = __reduce_max_sync(kFullMask, valid ? mx[k] : 0u);
    }
    node_regs p;
    if (any)
    {
        p.lo = make_float4(ord2f(mn[0]), ord2f(mn[1]), ord2f(mn[2]), __uint_as_float(first_child_index));
        p.hi = make_float4(ord2f(mx[0]), ord2f(mx[1]), ord2f(mx[2]), __uint_as_float(0u));
    }
    else
    {
        p.lo = make_float4(1e35f, 1e35f, 1e35f, __uint_as_float(kInvalid));
        p.hi = make_float4(-1e35f, -1e35f, -1e35f, __uint_as_float(0u));
    }
    return p;
}

// leaf source for the fused light-BVH build: init_light_array_bvh.comp:46-64
struct light_leaf_source
{
    const uint2* sorted;     // {morton, light index}
    const float4* view_pos;
    const float4* lights;    // point_light {vec3 color; float intensity}
    uint32_t light_count;
    float4* leaf_spheres;    // out (may be null): {view_pos.xyz, T} per leaf with T = sq_threshold((max.x - min.x) / 2): the
                             // operands of the leaf test of assign_lights.comp:97-102,209-214 in 16 B instead of 56 B
};

template <bool FROM_LIGHTS>
__global__ void __launch_bounds__(256)
build_bvh_kernel(vrenb200_bvh_node* nodes, uint32_t padded_leaf_count, light_leaf_source src)
{
    const unsigned lane = threadIdx.x & 31;
    uint32_t j = (blockIdx.x * 256u + threadIdx.x) >> 5;   // index of the parent inside its level
    if (j >= padded_leaf_count / 32) return;               // warp-uniform
    uint32_t level_start = 0, count = padded_leaf_count;
    float4* raw = reinterpret_cast<float4*>(nodes);

    node_regs c;
    {
        const uint32_t i = j * 32 + lane;
        if (FROM_LIGHTS)
        {
            if (i < src.light_count)
            {
                const uint32_t l = src.sorted[i].y;
                const float4 p = src.view_pos[l];
                const float r = src.lights[l].w;
                c.lo = make_float4(__fsub_rn(p.x, r), __fsub_rn(p.y, r), __fsub_rn(p.z, r), __uint_as_float(kLeaf));
                c.hi = make_float4(__fadd_rn(p.x, r), __fadd_rn(p.y, r), __fadd_rn(p.z, r), __uint_as_float(0u));
                if (src.leaf_spheres)
                    src.leaf_spheres[i] = make_float4(p.x, p.y, p.z, sq_threshold(__fdiv_rn(__fsub_rn(c.hi.x, c.lo.x), 2.0f)));
            }
            else
            {
                c.lo = make_float4(1e35f, 1e35f, 1e35f, __uint_as_float(kInvalid));
                c.hi = make_float4(-1e35f, -1e35f, -1e35f, __uint_as_float(0u));
            }
            raw[2 * (size_t) i] = c.lo;
            raw[2 * (size_t) i + 1] = c.hi;
        }
        else
        {
            c.lo = raw[2 * (size_t) i];
            c.hi = raw[2 * (size_t) i + 1];
        }
    }

    while (true)
    {
        const node_regs p = reduce_children(c, level_start + j * 32);
        const uint32_t dst_start = level_start + count;
        count >>= 5;
        const size_t pi = (size_t) dst_start + j;
        if (lane == 0) { raw[2 * pi] = p.lo; raw[2 * pi + 1] = p.hi; }
        if (count == 1) break; // that was the root
        // arrival at the grand-parent: its _pad word (zeroed by the host) counts finished children
        const size_t gp = (size_t) dst_start + count + (j >> 5);
        uint32_t arrived = 0;
        if (lane == 0)
        {
            __threadfence();
            arrived = atomicAdd(&nodes[gp]._pad, 1u);
        }
        arrived = __shfl_sync(kFullMask, arrived, 0);
        if (arrived != 31) break;
        __threadfence();
        level_start = dst_start;
        j >>= 5;
        const size_t ci = (size_t) level_start + j * 32 + lane;
        c.lo = __ldcg(&raw[2 * ci]);       // written by other SMs: read through L2
        c.hi = __ldcg(&raw[2 * ci + 1]);
    }
}

int launch_build(cudaStream_t s, vrenb200_bvh_node* nodes, uint32_t padded, const light_leaf_source* src)
{
    if (nodes == nullptr) return VRENB200_EINVAL_ARG;
    if (padded < 32 || !vrenb200_is_power_of(padded, 32)) return VRENB200_EINVAL_LENGTH; // build_bvh.cpp:45
    if (reinterpret_cast<uintptr_t>(nodes) & 15) return VRENB200_EALIGN;
    // zero the arrival counters: every node above level 1
    const size_t first_counter = (size_t) padded + padded / 32;
    const size_t length = vrenb200_calc_bvh_buffer_length(padded);
    if (length > first_counter)
        VRENB200_TRY(check_cuda(cudaMemsetAsync(nodes + first_counter, 0, (length - first_counter) * sizeof(vrenb200_bvh_node), s)));
    const uint32_t warps = padded / 32;
    const uint32_t grid = (warps + 7) / 8;
    if (src)
        build_bvh_kernel<true><<<grid, 256, 0, s>>>(nodes, padded, *src);
    else
        build_bvh_kernel<false><<<grid, 256, 0, s>>>(nodes, padded, light_leaf_source{});
    return check_launch();
}

} // namespace

// used by light_bvh.cu
int build_light_bvh_fused(cudaStream_t s, vrenb200_bvh_node* nodes, uint32_t padded, const void* sorted_pairs,
                          const float* view_pos, const float* lights, uint32_t light_count, void* leaf_spheres)
{
    light_leaf_source src{static_cast<const uint2*>(sorted_pairs), reinterpret_cast<const float4*>(view_pos),
                          reinterpret_cast<const float4*>(lights), light_count, static_cast<float4*>(leaf_spheres)};
    return launch_build(s, nodes, padded, &src);
}

} // namespace vrenb200

using namespace vrenb200;

extern "C" int vrenb200_build_bvh(vrenb200_stream_t stream, vrenb200_bvh_node* nodes, uint32_t padded_leaf_count)
{
    return launch_build(as_stream(stream), nodes, padded_leaf_count, nullptr);
}
