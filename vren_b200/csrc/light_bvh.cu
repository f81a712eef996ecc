// a6 — clustered_shading::construct_point_light_bvh (reference: vren/vren/pipeline/clustered_shading.cpp:60-346;
// shaders K9 point_light_position_to_view_space.comp:26-32, K10 discretize_point_light_positions.comp:26-43,
// K11 init_light_array_bvh.comp:42-65).
//
// Reference chain: K9 -> reduce<vec4,max> -> reduce<vec4,min> -> copy -> K10 -> bucket_sort -> K11 -> build_bvh
// (~25 dispatches, two full tree reductions each writing next_pow2(L) vec4s).
// Here: [view transform + min/max (atomics on order-preserving integer images)] -> [Morton pairs] -> stable
// bucket sort (2 onesweep digits) -> [leaf init fused with the single-launch 32-ary build].
//
// fp32 contract: mat4*vec4 as ((m0*x + m1*y) + (m2*z + m3*w)) with separately rounded mul/add,
// (p - min) / (max - min) * 32 evaluated left to right with IEEE div; no FMA contraction (-fmad=false + _rn).
#include <algorithm>

#include "common.cuh"

namespace vrenb200 {

int build_light_bvh_fused(cudaStream_t s, vrenb200_bvh_node* nodes, uint32_t padded, const void* sorted_pairs,
                          const float* view_pos, const float* lights, uint32_t light_count, void* leaf_spheres);
size_t light_leaf_sphere_offset(uint32_t light_count);

namespace {

struct mat4 { float m[16]; }; // column-major: m[4*col + row] (glm)

__device__ __forceinline__ uint32_t f2ord(float f)
{
    const uint32_t u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o)
{
    return __uint_as_float(o ^ ((o >> 31) ? 0x80000000u : 0xFFFFFFFFu));
}

// K9 + the two reductions of clustered_shading.cpp:141-178. minmax_ord[0..2] = min xyz, [3..5] = max xyz (ordered ints)
__global__ void __launch_bounds__(256)
view_space_minmax_kernel(const float4* __restrict__ positions, float4* __restrict__ view_pos, uint32_t light_count,
                         mat4 view, uint32_t* minmax_ord)
{
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    uint32_t mn[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, mx[3] = {0u, 0u, 0u};
    if (i < light_count)
    {
        const float4 p = positions[i];
        float r[4];
#pragma unroll
        for (int c = 0; c < 4; c++)
        {
            // (m0*x + m1*y) + (m2*z + m3*1)
            const float a = __fadd_rn(__fmul_rn(view.m[0 + c], p.x), __fmul_rn(view.m[4 + c], p.y));
            const float b = __fadd_rn(__fmul_rn(view.m[8 + c], p.z), __fmul_rn(view.m[12 + c], 1.0f));
            r[c] = __fadd_rn(a, b);
        }
        view_pos[i] = make_float4(r[0], r[1], r[2], r[3]);
#pragma unroll
        for (int c = 0; c < 3; c++) mn[c] = mx[c] = f2ord(r[c]);
    }
#pragma unroll
    for (int c = 0; c < 3; c++)
    {
        mn[c] = __reduce_min_sync(kFullMask, mn[c]);
        mx[c] = __reduce_max_sync(kFullMask, mx[c]);
    }
    // one atomic per block and component (per-warp atomics on the same six words serialise in L2: 11 us at 65536 lights)
    __shared__ uint32_t s_part[6][8];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0)
    {
#pragma unroll
        for (int c = 0; c < 3; c++)
        {
            s_part[c][warp] = mn[c];
            s_part[3 + c][warp] = mx[c];
        }
    }
    __syncthreads();
    if (warp == 0 && lane < 6)
    {
        uint32_t v = s_part[lane][0];
#pragma unroll
        for (int w = 1; w < 8; w++) v = lane < 3 ? min(v, s_part[lane][w]) : max(v, s_part[lane][w]);
        if (lane < 3) atomicMin(&minmax_ord[lane], v);
        else atomicMax(&minmax_ord[lane], v);
    }
}

// K10: 5 bits per axis, Morton interleave of the LOW 5 bits (q = 32 at the max point wraps to bin 0)
__global__ void __launch_bounds__(256)
morton_pairs_kernel(const float4* __restrict__ view_pos, uint32_t light_count, const uint32_t* __restrict__ minmax_ord,
                    int clamp_identity, uint2* __restrict__ pairs)
{
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i >= light_count) return;
    float mn[3], mx[3];
#pragma unroll
    for (int c = 0; c < 3; c++)
    {
        mn[c] = ord2f(minmax_ord[c]);
        mx[c] = ord2f(minmax_ord[3 + c]);
        if (clamp_identity)
        {
            // the reference pads the reduction to next_pow2(L) with +-1e35 (VRen.cmake:74-75)
            mn[c] = 1e35f < mn[c] ? 1e35f : mn[c];
            mx[c] = mx[c] < -1e35f ? -1e35f : mx[c];
        }
    }
    const float4 p = view_pos[i];
    const float pc[3] = {p.x, p.y, p.z};
    uint32_t q[3];
#pragma unroll
    for (int c = 0; c < 3; c++)
    {
        const float t = __fmul_rn(__fdiv_rn(__fsub_rn(pc[c], mn[c]), __fsub_rn(mx[c], mn[c])), 32.0f);
        const float f = floorf(t);
        q[c] = (f >= 0.0f) ? (uint32_t) f : 0u; // NaN (max == min) -> bin 0 (canonical choice v)
    }
    uint32_t code = 0;
#pragma unroll
    for (int b = 0; b < 5; b++)
    {
        code |= ((q[0] >> b) & 1u) << (3 * b);
        code |= ((q[1] >> b) & 1u) << (3 * b + 1);
        code |= ((q[2] >> b) & 1u) << (3 * b + 2);
    }
    pairs[i] = make_uint2(code, i);
}

} // namespace
} // namespace vrenb200

using namespace vrenb200;

extern "C" size_t vrenb200_light_bvh_scratch_bytes(uint32_t light_count)
{
    // [min/max 256 B] [morton pairs 8L] [bucket sort scratch]
    return 256 + align_up((size_t) light_count * 8, 256) + vrenb200_bucket_sort_scratch_bytes(light_count);
}

extern "C" size_t vrenb200_light_bvh_buffer_bytes(uint32_t light_count)
{
    // clustered_shading.cpp:34-47, and large enough to double as the scratch of this call (as the reference does)
    const size_t P = next_pow2_u32(light_count);
    const size_t ref = std::max<size_t>(std::max<size_t>(vrenb200_calc_bvh_buffer_size(light_count), (size_t) light_count * 8),
                                        256 * 3 + align_up(P * 16, 256) + 2 * 16);
    return std::max<size_t>(ref, vrenb200_light_bvh_scratch_bytes(light_count));
}

// the index buffer = bucket-sort output {sorted pairs | 65536 END offsets} followed by one compact leaf sphere per light
size_t vrenb200::light_leaf_sphere_offset(uint32_t light_count)
{
    return align_up(vrenb200_bucket_sort_output_bytes(light_count), 256);
}

extern "C" size_t vrenb200_light_index_buffer_bytes(uint32_t light_count)
{
    // clustered_shading.cpp:54-58 is smaller than what bucket_sort itself asserts (bucket_sort.cpp:84): take the max
    return std::max<size_t>(std::max<size_t>((size_t) light_count * 8, ((size_t) light_count + 2) * 16),
                            light_leaf_sphere_offset(light_count) + (size_t) light_count * 16);
}

extern "C" int vrenb200_construct_point_light_bvh(vrenb200_stream_t stream,
                                                  const float* positions, const float* lights, uint32_t light_count,
                                                  const float* view_col_major,
                                                  float* view_pos, void* bvh_buffer, void* index_buffer,
                                                  void* scratch, size_t scratch_bytes)
{
    if (light_count == 0) return VRENB200_EINVAL_LENGTH; // the caller skips the stage (clustered_shading.cpp:979)
    if (!positions || !lights || !view_col_major || !view_pos || !bvh_buffer || !index_buffer) return VRENB200_EINVAL_ARG;
    if ((reinterpret_cast<uintptr_t>(positions) | reinterpret_cast<uintptr_t>(lights) | reinterpret_cast<uintptr_t>(view_pos) |
         reinterpret_cast<uintptr_t>(bvh_buffer) | reinterpret_cast<uintptr_t>(index_buffer) | reinterpret_cast<uintptr_t>(scratch)) & 15)
        return VRENB200_EALIGN;
    // scratch == NULL: alias the BVH buffer, which is dead until the final build (clustered_shading.cpp:84-85 does the same)
    if (scratch == nullptr)
    {
        scratch = bvh_buffer;
        scratch_bytes = vrenb200_light_bvh_buffer_bytes(light_count);
    }
    if (scratch_bytes < vrenb200_light_bvh_scratch_bytes(light_count)) return VRENB200_ESCRATCH;
    cudaStream_t s = as_stream(stream);
    char* p = static_cast<char*>(scratch);
    uint32_t* minmax = reinterpret_cast<uint32_t*>(p);
    uint2* pairs = reinterpret_cast<uint2*>(p + 256);
    void* sort_scratch = p + 256 + align_up((size_t) light_count * 8, 256);
    const size_t sort_scratch_bytes = vrenb200_bucket_sort_scratch_bytes(light_count);

    VRENB200_TRY(check_cuda(cudaMemsetAsync(minmax, 0xFF, 12, s)));
    VRENB200_TRY(check_cuda(cudaMemsetAsync(minmax + 3, 0x00, 12, s)));
    mat4 view;
    for (int i = 0; i < 16; i++) view.m[i] = view_col_major[i];
    const uint32_t grid = (light_count + 255) / 256;
    view_space_minmax_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const float4*>(positions), reinterpret_cast<float4*>(view_pos),
                                                  light_count, view, minmax);
    VRENB200_TRY(check_launch());
    const int clamp_identity = (light_count & (light_count - 1)) != 0;
    morton_pairs_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const float4*>(view_pos), light_count, minmax, clamp_identity, pairs);
    VRENB200_TRY(check_launch());
    VRENB200_TRY(vrenb200_bucket_sort(stream, pairs, light_count, index_buffer, sort_scratch, sort_scratch_bytes));
    const uint32_t padded = vrenb200_calc_bvh_padded_leaf_count(light_count);
    return build_light_bvh_fused(s, static_cast<vrenb200_bvh_node*>(bvh_buffer), padded, index_buffer, view_pos, lights, light_count,
                                 static_cast<char*>(index_buffer) + light_leaf_sphere_offset(light_count));
}
