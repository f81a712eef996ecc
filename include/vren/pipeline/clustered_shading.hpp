// clustered-shading facade — vren/vren/pipeline/clustered_shading.hpp:11-183 minus `shade`.
// Inputs keep the reference formats: camera (camera.hpp:14-23), light_array (light.hpp:21-37), gbuffer normals
// RGBA16F (gbuffer.cpp:7-13), depth D32 (config.hpp:13).
#pragma once

#include <cmath>

#include "../context.hpp"

#define VREN_MAX_SCREEN_WIDTH 1920                     // config.hpp:15-16 (defaults; cluster_and_shade takes runtime limits)
#define VREN_MAX_SCREEN_HEIGHT 1080
#define VREN_MAX_POINT_LIGHT_COUNT (1 << 20)           // config.hpp:18
#define VREN_MAX_UNIQUE_CLUSTER_KEY_COUNT (1 << 17)    // config.hpp:23
#define VREN_MAX_ASSIGNED_LIGHT_COUNT (1 << 23)        // config.hpp:24

namespace vren
{
    struct uvec2 { uint32_t x, y; };

    struct camera // camera.hpp:14-23; matrices are produced by the caller (camera.cpp:32-50) and passed as inputs
    {
        float m_position[3] = {0, 0, 0};
        float m_yaw = 0, m_pitch = 0;
        float m_fov_y = 0.78539816339744830962f; // glm::radians(45.0f)
        float m_aspect_ratio = 1.0f;
        float m_near_plane = 0.01f;
        float m_far_plane = 1000.0f;
        float m_view[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}; // get_view(), column-major (glm)

        vrenb200_camera abi() const { return vrenb200_camera{m_fov_y, m_aspect_ratio, m_near_plane, m_far_plane}; }
    };

    struct light_array // light.hpp:21-37 (point lights only: the path ignores directional lights)
    {
        vren::vk_utils::buffer m_point_light_position_buffer; // vec4[L]
        vren::vk_utils::buffer m_point_light_buffer;          // point_light{vec3 color; float intensity}[L] (gpu_repr.hpp:61-64)
        uint32_t m_point_light_count = 0;
    };

    struct gbuffer // gbuffer.hpp: only the attachment this path samples
    {
        uint32_t m_width = 0, m_height = 0;
        vren::vk_utils::buffer m_normal_buffer; // RGBA16F [H][W], may be empty (== cleared to 0)
    };

    namespace vk_utils
    {
        struct depth_buffer_t { vren::vk_utils::buffer m_image; };                 // D32 [H][W]
        struct combined_image_view { vren::vk_utils::buffer m_image; };            // R32_UINT [H][W]
    }

    namespace clustered_shading
    {
        class construct_point_light_bvh // clustered_shading.hpp:19-46
        {
        public:
            explicit construct_point_light_bvh(vren::context const&) {}

            static size_t get_required_bvh_buffer_size(uint32_t point_light_count) { return vrenb200_light_bvh_buffer_bytes(point_light_count); }
            static size_t get_required_point_light_index_buffer_size(uint32_t point_light_count) { return vrenb200_light_index_buffer_bytes(point_light_count); }

            // clustered_shading.cpp:60-346; like the reference, bvh_buffer doubles as scratch until the final build
            void operator()(VkCommandBuffer command_buffer, vren::resource_container&, vren::light_array const& light_array,
                            vren::vk_utils::buffer const& view_space_point_light_position_buffer, vren::camera const& camera,
                            vren::vk_utils::buffer const& bvh_buffer, vren::vk_utils::buffer const& point_light_index_buffer)
            {
                const uint32_t n = light_array.m_point_light_count;
                if (bvh_buffer.m_size < get_required_bvh_buffer_size(n) || point_light_index_buffer.m_size < get_required_point_light_index_buffer_size(n))
                    check_status(VRENB200_ESCRATCH, "construct_point_light_bvh: buffer too small");    // clustered_shading.cpp:72-73
                check_status(vrenb200_construct_point_light_bvh((vrenb200_stream_t) command_buffer,
                                                                light_array.m_point_light_position_buffer.ptr<float>(),
                                                                light_array.m_point_light_buffer.ptr<float>(), n, camera.m_view,
                                                                view_space_point_light_position_buffer.ptr<float>(), bvh_buffer.m_ptr,
                                                                point_light_index_buffer.m_ptr, nullptr, 0),
                             "construct_point_light_bvh");
            }
        };

        class find_unique_cluster_list // clustered_shading.hpp:52-74
        {
            vren::scratch_pool m_scratch;

        public:
            explicit find_unique_cluster_list(vren::context const&) {}
            void reserve_scratch(uint32_t width, uint32_t height, size_t frames_in_flight = 1)
            {
                m_scratch.reserve(vrenb200_find_unique_clusters_scratch_bytes(width, height), frames_in_flight);
            }

            // clustered_shading.cpp:363-448; cluster_key_dispatch_params = uvec4 {count, 1, 1, overflow flag}
            void operator()(uint32_t /*frame_idx*/, VkCommandBuffer command_buffer, vren::resource_container& resource_container, vren::uvec2 const& screen,
                            vren::camera const& camera, vren::gbuffer const& gbuffer, vren::vk_utils::depth_buffer_t const& depth_buffer,
                            vren::vk_utils::buffer const& cluster_key_buffer, vren::vk_utils::buffer const& cluster_key_dispatch_params_buffer,
                            vren::vk_utils::combined_image_view const& cluster_reference_buffer)
            {
                const size_t bytes = vrenb200_find_unique_clusters_scratch_bytes(screen.x, screen.y);
                std::shared_ptr<void> lease = m_scratch.acquire(bytes);
                void* scratch = lease.get();
                resource_container.add_resource(lease);
                const vrenb200_camera cam = camera.abi();
                check_status(vrenb200_find_unique_clusters((vrenb200_stream_t) command_buffer, depth_buffer.m_image.ptr<float>(),
                                                           gbuffer.m_normal_buffer.m_ptr, screen.x, screen.y, &cam,
                                                           cluster_key_buffer.ptr<uint32_t>(), (uint32_t) (cluster_key_buffer.m_size / 4),
                                                           cluster_key_dispatch_params_buffer.ptr<uint32_t>(),
                                                           cluster_reference_buffer.m_image.ptr<uint32_t>(), scratch, bytes),
                             "find_unique_cluster_list");
            }
        };

        class assign_lights // clustered_shading.hpp:80-108
        {
            vren::scratch_pool m_scratch;

        public:
            explicit assign_lights(vren::context const&) {}
            void reserve_scratch(uint32_t max_keys, uint32_t max_assigned, size_t frames_in_flight = 1)
            {
                m_scratch.reserve(vrenb200_assign_lights_scratch_bytes(max_keys, max_assigned), frames_in_flight);
            }

            // clustered_shading.cpp:473-690
            void operator()(uint32_t /*frame_idx*/, VkCommandBuffer command_buffer, vren::resource_container& resource_container, vren::uvec2 const& screen,
                            vren::camera const& camera, vren::vk_utils::buffer const& cluster_key_buffer,
                            vren::vk_utils::buffer const& cluster_key_dispatch_params_buffer, vren::vk_utils::buffer const& light_bvh_buffer,
                            uint32_t light_bvh_root_index, uint32_t light_count, vren::vk_utils::buffer const& light_index_buffer,
                            vren::vk_utils::buffer const& assigned_light_indices_buffer, vren::vk_utils::buffer const& assigned_light_counts_buffer,
                            vren::vk_utils::buffer const& assigned_light_offsets_buffer,
                            vren::vk_utils::buffer const& view_space_point_light_position_buffer, uint32_t* status_out = nullptr)
            {
                const uint32_t max_keys = (uint32_t) (assigned_light_counts_buffer.m_size / 4);
                const size_t bytes = vrenb200_assign_lights_scratch_bytes(max_keys, (uint32_t) (assigned_light_indices_buffer.m_size / 4));
                std::shared_ptr<void> lease = m_scratch.acquire(bytes);
                void* scratch = lease.get();
                resource_container.add_resource(lease);
                const vrenb200_camera cam = camera.abi();
                check_status(vrenb200_assign_lights((vrenb200_stream_t) command_buffer, screen.x, screen.y, &cam, cluster_key_buffer.ptr<uint32_t>(),
                                                    cluster_key_dispatch_params_buffer.ptr<uint32_t>(), max_keys, light_bvh_buffer.m_ptr,
                                                    light_bvh_root_index, light_count, light_index_buffer.m_ptr,
                                                    view_space_point_light_position_buffer.ptr<float>(),
                                                    assigned_light_indices_buffer.ptr<uint32_t>(), (uint32_t) (assigned_light_indices_buffer.m_size / 4),
                                                    assigned_light_counts_buffer.ptr<uint32_t>(), assigned_light_offsets_buffer.ptr<uint32_t>(),
                                                    status_out, scratch, bytes),
                             "assign_lights");
            }
        };
    }

    // cluster_and_shade (clustered_shading.hpp:147-182, clustered_shading.cpp:817-1172) without the shade step:
    // owns every intermediate buffer (sized by runtime limits whose defaults are the reference's compile-time maxima)
    // and sequences a6 -> a7 -> a8 on a stream.
    struct cluster_and_shade_limits
    {
        uint32_t max_screen_width = VREN_MAX_SCREEN_WIDTH, max_screen_height = VREN_MAX_SCREEN_HEIGHT;
        uint32_t max_point_light_count = VREN_MAX_POINT_LIGHT_COUNT;
        uint32_t max_unique_cluster_key_count = VREN_MAX_UNIQUE_CLUSTER_KEY_COUNT;
        uint32_t max_assigned_light_count = VREN_MAX_ASSIGNED_LIGHT_COUNT;
    };

    class cluster_and_shade
    {
    public:
        using limits = cluster_and_shade_limits;

        vren::clustered_shading::construct_point_light_bvh m_construct_point_light_bvh;
        vren::clustered_shading::find_unique_cluster_list m_find_unique_cluster_list;
        vren::clustered_shading::assign_lights m_assign_lights;

        vren::vk_utils::buffer m_view_space_point_light_position_buffer;
        vren::vk_utils::buffer m_point_light_bvh_buffer;
        vren::vk_utils::buffer m_point_light_index_buffer;
        vren::vk_utils::buffer m_cluster_key_buffer;
        vren::vk_utils::buffer m_cluster_key_dispatch_params_buffer;
        vren::vk_utils::combined_image_view m_cluster_reference_buffer;
        vren::vk_utils::buffer m_assigned_light_indices_buffer;
        vren::vk_utils::buffer m_assigned_light_counts_buffer;
        vren::vk_utils::buffer m_assigned_light_offsets_buffer;
        vren::vk_utils::buffer m_status_buffer; // uint[4]: {assigned total, overflow, node tests, leaf tests}

        explicit cluster_and_shade(vren::context const& c, limits const& l = limits()) :
            m_construct_point_light_bvh(c), m_find_unique_cluster_list(c), m_assign_lights(c),
            m_view_space_point_light_position_buffer(vk_utils::alloc_device_only_buffer(c, (size_t) l.max_point_light_count * 16)),
            m_point_light_bvh_buffer(vk_utils::alloc_device_only_buffer(c, clustered_shading::construct_point_light_bvh::get_required_bvh_buffer_size(l.max_point_light_count))),
            m_point_light_index_buffer(vk_utils::alloc_device_only_buffer(c, clustered_shading::construct_point_light_bvh::get_required_point_light_index_buffer_size(l.max_point_light_count))),
            m_cluster_key_buffer(vk_utils::alloc_device_only_buffer(c, (size_t) l.max_unique_cluster_key_count * 4)),
            m_cluster_key_dispatch_params_buffer(vk_utils::alloc_device_only_buffer(c, 16)),
            m_cluster_reference_buffer{vk_utils::alloc_device_only_buffer(c, (size_t) l.max_screen_width * l.max_screen_height * 4)},
            m_assigned_light_indices_buffer(vk_utils::alloc_device_only_buffer(c, (size_t) l.max_assigned_light_count * 4)),
            m_assigned_light_counts_buffer(vk_utils::alloc_device_only_buffer(c, (size_t) l.max_unique_cluster_key_count * 4)),
            m_assigned_light_offsets_buffer(vk_utils::alloc_device_only_buffer(c, (size_t) l.max_unique_cluster_key_count * 4)),
            m_status_buffer(vk_utils::alloc_device_only_buffer(c, 16))
        {
            // the stages' scratch for the largest frame is ready before the first frame is recorded (two frames in flight):
            // nothing is allocated while recording
            m_find_unique_cluster_list.reserve_scratch(l.max_screen_width, l.max_screen_height, 2);
            m_assign_lights.reserve_scratch(l.max_unique_cluster_key_count, l.max_assigned_light_count, 2);
        }

        // steps 1-3 of clustered_shading.cpp:975-1147 (step 4, shade, is out of scope)
        void operator()(VkCommandBuffer command_buffer, vren::resource_container& resource_container, vren::uvec2 const& screen,
                        vren::camera const& camera, vren::gbuffer const& gbuffer, vren::vk_utils::depth_buffer_t const& depth_buffer,
                        vren::light_array const& light_array)
        {
            if (light_array.m_point_light_count > 0) // :979
                m_construct_point_light_bvh(command_buffer, resource_container, light_array, m_view_space_point_light_position_buffer, camera,
                                            m_point_light_bvh_buffer, m_point_light_index_buffer);
            m_find_unique_cluster_list(0, command_buffer, resource_container, screen, camera, gbuffer, depth_buffer, m_cluster_key_buffer,
                                       m_cluster_key_dispatch_params_buffer, m_cluster_reference_buffer);
            m_assign_lights(0, command_buffer, resource_container, screen, camera, m_cluster_key_buffer, m_cluster_key_dispatch_params_buffer,
                            m_point_light_bvh_buffer, vren::calc_bvh_root_index(light_array.m_point_light_count), light_array.m_point_light_count,
                            m_point_light_index_buffer, m_assigned_light_indices_buffer, m_assigned_light_counts_buffer,
                            m_assigned_light_offsets_buffer, m_view_space_point_light_position_buffer, m_status_buffer.ptr<uint32_t>());
        }
    };
}
