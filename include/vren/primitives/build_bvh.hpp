// vren::build_bvh facade — vren/vren/primitives/build_bvh.hpp:6-56.
#pragma once

#include "../base/base.hpp"
#include "../vk_helpers/buffer.hpp"

namespace vren
{
    struct bvh_node // build_bvh.hpp:8-18, 32 bytes, identical layout
    {
        inline static const uint32_t k_leaf_node = 0xFFFFFFFFu;
        inline static const uint32_t k_invalid_node = 0xFFFFFFFEu;

        float m_min[3]; uint32_t m_next;
        float m_max[3]; uint32_t _pad;

        inline bool is_leaf() const { return m_next == k_leaf_node; }
        inline bool is_invalid() const { return m_next == k_invalid_node; }
    };
    static_assert(sizeof(bvh_node) == sizeof(vrenb200_bvh_node), "bvh_node layout");

    class build_bvh
    {
    public:
        inline static const uint32_t k_workgroup_size = 1024;

        explicit build_bvh(vren::context const&) {}

        static size_t get_required_buffer_size(uint32_t leaf_count) { return vrenb200_calc_bvh_buffer_size(leaf_count); }

        // build_bvh.cpp:38-97; leaf_count is the PADDED leaf count (>= 32, power of 32)
        void operator()(VkCommandBuffer command_buffer, vren::resource_container&, vren::vk_utils::buffer const& buffer, uint32_t leaf_count)
        {
            check_status(vrenb200_build_bvh((vrenb200_stream_t) command_buffer, buffer.ptr<vrenb200_bvh_node>(), leaf_count), "vren::build_bvh");
        }
    };

    inline uint32_t calc_bvh_padded_leaf_count(uint32_t leaf_count) { return vrenb200_calc_bvh_padded_leaf_count(leaf_count); }
    inline uint32_t calc_bvh_buffer_length(uint32_t leaf_count) { return vrenb200_calc_bvh_buffer_length(leaf_count); }
    inline size_t calc_bvh_buffer_size(uint32_t leaf_count) { return vrenb200_calc_bvh_buffer_size(leaf_count); }
    inline uint32_t calc_bvh_root_index(uint32_t leaf_count) { return vrenb200_calc_bvh_root_index(leaf_count); }
    inline uint32_t calc_bvh_level_count(uint32_t leaf_count) { return vrenb200_calc_bvh_level_count(leaf_count); }
}
