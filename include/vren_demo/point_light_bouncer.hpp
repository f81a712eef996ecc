// vren_demo::point_light_bouncer facade — vren_demo/vren_demo/point_light_bouncer.hpp:12-35.
// Same class, method and argument order; the VkCommandBuffer carries the CUDA stream (vk_helpers/buffer.hpp), the
// resource_container is accepted and ignored (nothing is allocated per call), glm::vec3 becomes three floats.
#pragma once

#include "../vren/context.hpp"

namespace vren_demo
{
    struct vec3 { float x, y, z; };

    class point_light_bouncer
    {
    public:
        static const uint32_t k_workgroup_size = 1024;   // reference dispatch granularity (point_light_bouncer.hpp:15); informational

        explicit point_light_bouncer(vren::context const&) {}

        // point_light_bouncer.cpp bounce(): one dispatch over point_light_count lights (bounce_point_lights.comp:33-73)
        void bounce(uint32_t /*frame_idx*/, VkCommandBuffer command_buffer, vren::resource_container& /*resource_container*/,
                    vren::vk_utils::buffer const& point_light_position_buffer, vren::vk_utils::buffer const& point_light_direction_buffer,
                    uint32_t point_light_count, vec3 const& aabb_min, vec3 const& aabb_max, float speed, float dt) const
        {
            const float lo[3] = { aabb_min.x, aabb_min.y, aabb_min.z }, hi[3] = { aabb_max.x, aabb_max.y, aabb_max.z };
            vren::check_status(vrenb200_bounce_point_lights((vrenb200_stream_t) command_buffer, point_light_position_buffer.ptr<float>(),
                                                            point_light_direction_buffer.ptr<float>(), point_light_count, lo, hi, speed, dt),
                               "vren_demo::point_light_bouncer::bounce");
        }
    };
}
