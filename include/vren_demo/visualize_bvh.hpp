// vren_demo::visualize_bvh facade — vren_demo/vren_demo/visualize_bvh.hpp:10-26.
// write() enqueues on the stream carried by the VkCommandBuffer instead of returning a render-graph node; the debug
// renderer's draw buffer is a plain vertex buffer of {float position[3]; uint32_t color} (common.glsl:93-97).
#pragma once

#include "../vren/context.hpp"

namespace vren_demo
{
    struct debug_draw_vertex
    {
        float position[3];
        uint32_t color;
    };

    class visualize_bvh
    {
    public:
        explicit visualize_bvh(vren::context const&) {}

        // bytes the draw buffer needs for a BVH with `level_count` levels above the root (24 vertices per node)
        static size_t get_required_vertex_buffer_size(uint32_t level_count)
        {
            return (size_t) vrenb200_visualize_bvh_vertex_count(level_count) * sizeof(debug_draw_vertex);
        }

        // visualize_bvh.cpp:59-94: every level, leaves first, one colour per level
        void write(VkCommandBuffer command_buffer, vren::vk_utils::buffer const& bvh, uint32_t level_count,
                   vren::vk_utils::buffer const& draw_vertex_buffer) const
        {
            vren::check_status(vrenb200_visualize_bvh((vrenb200_stream_t) command_buffer, bvh.ptr<void>(), level_count, draw_vertex_buffer.ptr<void>()),
                               "vren_demo::visualize_bvh::write");
        }
    };
}
