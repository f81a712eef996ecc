// vren::kd_tree_* facade — vren/vren/base/kd_tree.hpp:12-45.  Same free functions and argument order; the node is the
// 8-byte vrenb200_kd_tree_node (the reference's bit-field struct is not part of any file or GPU format).  `count` and
// `node_offset` of the search are accepted for signature compatibility: the search always starts at the given node.
#pragma once

#include <cstddef>
#include <cstdint>
#include <functional>

#include "../../vrenb200.h"

namespace vren
{
    using kd_tree_node = vrenb200_kd_tree_node;

    // kd_tree.cpp:5-75.  node_offset must be 0 (the reference only ever passes 0 from outside); returns the node count.
    inline size_t kd_tree_build(float const* points, size_t point_stride, uint32_t* indices, size_t count, kd_tree_node* kd_tree,
                                size_t node_offset, size_t max_leaf_point_count)
    {
        return node_offset + vrenb200_kd_tree_build(points, point_stride, indices, count, kd_tree + node_offset, max_leaf_point_count);
    }

    using kd_tree_search_filter_t = std::function<bool(uint32_t point)>;
    inline const kd_tree_search_filter_t k_kd_tree_default_search_filter = [](uint32_t) -> bool { return true; };

    // kd_tree.cpp:77-129
    inline void kd_tree_search(float const* points, size_t point_stride, uint32_t const* /*indices*/, size_t /*count*/, kd_tree_node* kd_tree,
                               size_t node_offset, float const* sample, kd_tree_search_filter_t const& filter_predicate, uint32_t& best_point,
                               float& best_distance_squared)
    {
        auto trampoline = [](uint32_t point, void* user) -> int { return (*static_cast<kd_tree_search_filter_t const*>(user))(point) ? 1 : 0; };
        vrenb200_kd_tree_search(points, point_stride, kd_tree + node_offset, sample, trampoline,
                                const_cast<void*>(static_cast<void const*>(&filter_predicate)), &best_point, &best_distance_squared);
    }
}
