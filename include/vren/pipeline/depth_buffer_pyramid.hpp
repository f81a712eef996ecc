// vren::depth_buffer_pyramid / vren::depth_buffer_reductor facade — vren/vren/pipeline/depth_buffer_pyramid.hpp:19-115.
// The Vulkan mip-mapped R32F image becomes one flat float buffer with the levels back to back.
#pragma once

#include "../context.hpp"
#include "clustered_shading.hpp"

namespace vren
{
    class depth_buffer_pyramid
    {
    public:
        static const uint32_t k_max_depth_buffer_pyramid_level_count = 16;

    private:
        uint32_t m_base_width, m_base_height;
        uint32_t m_level_count;

    public:
        vren::vk_utils::buffer m_image; // float[sum of level sizes]

        depth_buffer_pyramid(vren::context const& context, uint32_t width, uint32_t height) :
            m_base_width(width), m_base_height(height), m_level_count(vrenb200_depth_pyramid_level_count(width, height)), // depth_buffer_pyramid.cpp:13-18
            m_image(vren::vk_utils::alloc_device_only_buffer(context, vrenb200_depth_pyramid_bytes(width, height)))
        {
        }

        inline uint32_t get_image_width(uint32_t level) const { return vrenb200_depth_pyramid_level_width(m_base_width, level); }
        inline uint32_t get_image_height(uint32_t level) const { return vrenb200_depth_pyramid_level_height(m_base_height, level); }
        inline uint32_t get_level_count() const { return m_level_count; }
        // replaces get_level_image_view(level): pointer to the first texel of a level
        inline float* get_level(uint32_t level) const { return m_image.ptr<float>() + vrenb200_depth_pyramid_level_offset(m_base_width, m_base_height, level); }
        inline uint32_t get_base_width() const { return m_base_width; }
        inline uint32_t get_base_height() const { return m_base_height; }
    };

    class depth_buffer_reductor
    {
    public:
        explicit depth_buffer_reductor(vren::context const&) {}

        // copy_and_reduce (depth_buffer_pyramid.cpp:177-305): level 0 = copy of the depth buffer, then the 2x2-max chain
        void copy_and_reduce(VkCommandBuffer command_buffer, vren::vk_utils::depth_buffer_t const& depth_buffer,
                             vren::depth_buffer_pyramid const& depth_buffer_pyramid) const
        {
            check_status(vrenb200_depth_pyramid_build((vrenb200_stream_t) command_buffer, depth_buffer.m_image.ptr<float>(),
                                                      depth_buffer_pyramid.get_base_width(), depth_buffer_pyramid.get_base_height(),
                                                      depth_buffer_pyramid.m_image.ptr<float>()),
                         "vren::depth_buffer_reductor::copy_and_reduce");
        }
    };
}
