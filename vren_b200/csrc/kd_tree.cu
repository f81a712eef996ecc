// n4 — vren::kd_tree_build / vren::kd_tree_search (reference: vren/vren/base/kd_tree.{hpp,cpp}:5-129, test
// vren_test/vren_test/kd_tree.cpp:29): a kd-tree over 3-D points for exact nearest-neighbour queries.
//
// The build is host code like the reference's (it is a one-off, pointer-chasing partition; the reference runs it on the
// CPU too); it is written as a loop over an explicit work list instead of the reference's recursion, and produces the
// same pre-order node array: an inner node is followed by its left subtree, the right subtree starts
// `right_child_distance` nodes further; a leaf is a run of nodes, the first of which carries the length of the run.
// The split axis is the one with the largest variance (Welford, one pass), the split value the mean on that axis, points
// strictly below go left.  What is new here is the search side: vrenb200_kd_tree_search_batch answers many queries at
// once on the GPU, one thread per query with a small explicit stack (the reference has only the recursive host search).
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace vrenb200 {
namespace {

constexpr uint32_t kLeafAxis = 3u;
constexpr uint32_t kLinkMask = 0x3FFFFFFFu;

__host__ __device__ inline uint32_t node_axis(const vrenb200_kd_tree_node& n) { return n.axis_and_link & 3u; }
__host__ __device__ inline uint32_t node_link(const vrenb200_kd_tree_node& n) { return n.axis_and_link >> 2; }
inline uint32_t pack(uint32_t axis, uint32_t link) { return (axis & 3u) | (link << 2); }

struct build_item
{
    size_t first, count;     // range of `indices`
    size_t node;             // where the subtree starts
    size_t* end_out;         // receives the index one past the subtree
};

// One subtree, iteratively: left subtrees are descended at once, right subtrees wait on a stack until their start (the
// end of the left sibling) is known.
size_t build_subtree(const float* points, size_t stride, uint32_t* indices, size_t first, size_t count, vrenb200_kd_tree_node* nodes,
                     size_t node_offset, size_t max_leaf)
{
    struct pending { size_t parent, first, count; int depth; };
    int depth = 0;
    std::vector<pending> right;
    size_t cursor = node_offset;
    while (true)
    {
        // descend along left children
        while (true)
        {
            uint32_t* idx = indices + first;
            bool leaf = count <= max_leaf;
            size_t middle = 0;
            uint32_t axis = 0;
            float split = 0.0f;
            if (!leaf)
            {
                float mean[3] = {0, 0, 0}, var[3] = {0, 0, 0}, k = 1.0f;
                for (size_t i = 0; i < count; i++, k += 1.0f)
                {
                    const float* p = points + (size_t) idx[i] * stride;
                    for (int a = 0; a < 3; a++)
                    {
                        const float d = p[a] - mean[a];
                        mean[a] += d / k;
                        var[a] += (p[a] - mean[a]) * d;
                    }
                }
                // the reference's selection, kd_tree.cpp:52 (x wins only over y; then y against z)
                axis = var[0] > var[1] ? 0u : (var[1] > var[2] ? 1u : 2u);
                split = mean[axis];
                for (size_t i = 0; i < count; i++)
                    if (points[(size_t) idx[i] * stride + axis] < split) std::swap(idx[i], idx[middle++]);
                // Past depth 32 (badly skewed data) or when everything falls on one side, split at the median instead: halving
                // bounds the depth by 32 + log2(count) < 64, the capacity of the search stack.  Points equal to the split
                // value may then sit on both sides, which the search's pruning rule allows (left <= split <= right).
                if (middle == 0 || middle == count || depth >= 32)
                {
                    middle = count / 2;
                    std::nth_element(idx, idx + middle, idx + count, [&](uint32_t a, uint32_t b) {
                        return points[(size_t) a * stride + axis] < points[(size_t) b * stride + axis];
                    });
                    split = points[(size_t) idx[middle] * stride + axis];
                }
            }
            if (leaf)
            {
                for (size_t i = 0; i < count; i++)
                {
                    nodes[cursor + i].index = idx[i];
                    nodes[cursor + i].axis_and_link = pack(kLeafAxis, i == 0 ? (uint32_t) count : kLinkMask);
                }
                cursor += count;
                break;
            }
            nodes[cursor].split = split;
            nodes[cursor].axis_and_link = pack(axis, 0);         // link patched when the left subtree is complete
            right.push_back({cursor, first + middle, count - middle, depth + 1});
            cursor += 1;
            count = middle;                                       // left child: same `first`
            depth += 1;
        }
        if (right.empty()) return cursor;
        const pending r = right.back();
        right.pop_back();
        nodes[r.parent].axis_and_link = pack(node_axis(nodes[r.parent]), (uint32_t) (cursor - r.parent));
        first = r.first;
        count = r.count;
        depth = r.depth;
    }
}

struct keep_all
{
    __host__ __device__ bool operator()(uint32_t) const { return true; }
};

struct keep_if
{
    int (*filter)(uint32_t, void*);
    void* user;
    __host__ __device__ bool operator()(uint32_t p) const { return filter(p, user) != 0; }
};

template <typename Filter>
__host__ __device__ inline void search_tree(const float* points, size_t stride, const vrenb200_kd_tree_node* nodes, const float sample[3],
                                            const Filter& keep, uint32_t& best_point, float& best_d2)
{
    // explicit stack of far children still worth a visit: {node, squared distance to its splitting plane}
    struct far_child { uint32_t node; float plane_d2; };
    far_child stack[64];
    int top = 0;
    uint32_t at = 0;
    while (true)
    {
        const vrenb200_kd_tree_node n = nodes[at];
        if (node_axis(n) == kLeafAxis)
        {
            const uint32_t run = node_link(n);
            for (uint32_t i = 0; i < run; i++)
            {
                const uint32_t pt = nodes[at + i].index;
                if (!keep(pt)) continue;
                const float* p = points + (size_t) pt * stride;
                const float dx = sample[0] - p[0], dy = sample[1] - p[1], dz = sample[2] - p[2];
                const float d2 = dx * dx + dy * dy + dz * dz;
                if (d2 < best_d2)
                {
                    best_d2 = d2;
                    best_point = pt;
                }
            }
            // next candidate: the most recent far child whose plane is still closer than the best point
            bool found = false;
            while (top > 0)
            {
                const far_child f = stack[--top];
                if (best_d2 > f.plane_d2)
                {
                    at = f.node;
                    found = true;
                    break;
                }
            }
            if (!found) return;
        }
        else
        {
            const float d = sample[node_axis(n)] - n.split;
            const uint32_t near_node = d <= 0.0f ? at + 1 : at + node_link(n);
            const uint32_t far_node = d <= 0.0f ? at + node_link(n) : at + 1;
            if (top < 64) stack[top++] = {far_node, d * d};
            at = near_node;
        }
    }
}

__global__ void __launch_bounds__(128)
kd_tree_search_kernel(const float* __restrict__ points, uint32_t stride, const vrenb200_kd_tree_node* __restrict__ nodes,
                      const float* __restrict__ samples, uint32_t sample_count, uint32_t* __restrict__ best_point, float* __restrict__ best_d2)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sample_count) return;
    const float s[3] = {samples[3 * i], samples[3 * i + 1], samples[3 * i + 2]};
    uint32_t bp = 0xFFFFFFFFu;
    float bd = __int_as_float(0x7f800000);   // +inf
    search_tree(points, stride, nodes, s, keep_all{}, bp, bd);
    best_point[i] = bp;
    best_d2[i] = bd;
}

} // namespace
} // namespace vrenb200

using namespace vrenb200;

extern "C" size_t vrenb200_kd_tree_build(const float* points, size_t point_stride, uint32_t* indices, size_t count,
                                         vrenb200_kd_tree_node* nodes, size_t max_leaf_point_count)
{
    if (points == nullptr || indices == nullptr || nodes == nullptr || count == 0 || count > kLinkMask) return 0;
    if (max_leaf_point_count == 0) max_leaf_point_count = 1;
    return build_subtree(points, point_stride, indices, 0, count, nodes, 0, max_leaf_point_count);
}

extern "C" void vrenb200_kd_tree_search(const float* points, size_t point_stride, const vrenb200_kd_tree_node* nodes, const float sample[3],
                                        int (*filter)(uint32_t point, void* user), void* user, uint32_t* best_point, float* best_distance_squared)
{
    uint32_t bp = *best_point;
    float bd = *best_distance_squared;
    if (filter)
        search_tree(points, point_stride, nodes, sample, keep_if{filter, user}, bp, bd);
    else
        search_tree(points, point_stride, nodes, sample, keep_all{}, bp, bd);
    *best_point = bp;
    *best_distance_squared = bd;
}

extern "C" int vrenb200_kd_tree_search_batch(vrenb200_stream_t stream, const float* points, uint32_t point_stride,
                                             const vrenb200_kd_tree_node* nodes, const float* samples, uint32_t sample_count,
                                             uint32_t* best_point, float* best_distance_squared)
{
    if (!points || !nodes || !samples || !best_point || !best_distance_squared) return VRENB200_EINVAL_ARG;
    if (sample_count == 0) return VRENB200_OK;
    kd_tree_search_kernel<<<(sample_count + 127) / 128, 128, 0, as_stream(stream)>>>(points, point_stride, nodes, samples, sample_count,
                                                                                   best_point, best_distance_squared);
    return check_launch();
}
