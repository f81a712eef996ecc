// vren::bucket_sort facade — vren/vren/primitives/bucket_sort.hpp:9-46.
#pragma once

#include "../base/base.hpp"
#include "../vk_helpers/buffer.hpp"

namespace vren
{
    class bucket_sort
    {
    public:
        inline static const uint32_t k_workgroup_size = 1024;
        inline static const uint32_t k_max_items = 1;
        inline static const uint32_t k_key_size = 1 << 16;
        inline static const uint32_t k_key_mask = k_key_size - 1;

    private:
        vren::scratch_pool m_scratch; // ping-pong pair buffer + look-back state

    public:
        explicit bucket_sort(vren::context const&) {}

        static size_t get_required_output_buffer_size(uint32_t length) { return vrenb200_bucket_sort_output_bytes(length); } // bucket_sort.cpp:67-70

        // bucket_sort.cpp:72-161. Ties keep input order (the reference's atomics leave it unspecified).
        void operator()(VkCommandBuffer command_buffer, vren::resource_container& resource_container, vren::vk_utils::buffer const& input_buffer,
                        uint32_t input_buffer_length, size_t input_buffer_offset, vren::vk_utils::buffer const& output_buffer,
                        size_t output_buffer_offset)
        {
            if (input_buffer_offset % VREN_MIN_STORAGE_BUFFER_OFFSET_ALIGNMENT != 0 || output_buffer_offset % VREN_MIN_STORAGE_BUFFER_OFFSET_ALIGNMENT != 0)
                check_status(VRENB200_EALIGN, "vren::bucket_sort");                                      // bucket_sort.cpp:82-83
            if (output_buffer.m_size - output_buffer_offset < get_required_output_buffer_size(input_buffer_length))
                check_status(VRENB200_ESCRATCH, "vren::bucket_sort: output buffer too small");           // bucket_sort.cpp:84
            const size_t bytes = vrenb200_bucket_sort_scratch_bytes(input_buffer_length);
            std::shared_ptr<void> lease = m_scratch.acquire(bytes);
            void* scratch = lease.get();
            resource_container.add_resource(lease);
            check_status(vrenb200_bucket_sort((vrenb200_stream_t) command_buffer, input_buffer.ptr<>(input_buffer_offset), input_buffer_length,
                                              output_buffer.ptr<>(output_buffer_offset), scratch, bytes),
                         "vren::bucket_sort");
        }
    };
}
