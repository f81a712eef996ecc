// vren::radix_sort facade — vren/vren/primitives/radix_sort.hpp:9-53.
#pragma once

#include "blelloch_scan.hpp"

namespace vren
{
    class radix_sort
    {
    public:
        inline static const uint32_t k_workgroup_size = 1024;
        inline static const uint32_t k_max_items = 1;
        // the reference sorts 4 bits per pass (radix_sort.hpp:15-16); the CUDA path uses 8-bit onesweep digits
        inline static constexpr uint32_t k_radix_bits = 8;
        inline static constexpr uint32_t k_radix = 1 << k_radix_bits;

    private:
        vren::context const* m_context;

    public:
        explicit radix_sort(vren::context const& context) : m_context(&context) {}

        // radix_sort.cpp:124-147 — sizes differ from the reference (look-back state instead of per-workgroup counts)
        vren::vk_utils::buffer create_scratch_buffer_1(uint32_t length)
        {
            return vren::vk_utils::alloc_device_only_buffer(*m_context, vrenb200_radix_sort_scratch_buffer_1_bytes(length));
        }
        vren::vk_utils::buffer create_scratch_buffer_2(uint32_t length)
        {
            return vren::vk_utils::alloc_device_only_buffer(*m_context, vrenb200_radix_sort_scratch_buffer_2_bytes(length));
        }

        // radix_sort.cpp:149-337; throws std::invalid_argument unless length >= 1024 and a power of 2 (:158-161)
        void operator()(VkCommandBuffer command_buffer, vren::resource_container&, vren::vk_utils::buffer const& buffer, uint32_t length,
                        vren::vk_utils::buffer const& scratch_buffer_1, vren::vk_utils::buffer const& scratch_buffer_2)
        {
            const int st = vrenb200_radix_sort_compat((vrenb200_stream_t) command_buffer, buffer.ptr<uint32_t>(), length, scratch_buffer_1.m_ptr,
                                                      scratch_buffer_1.m_size, scratch_buffer_2.m_ptr, scratch_buffer_2.m_size);
            if (st == VRENB200_EINVAL_LENGTH) throw std::invalid_argument("Length must be higher than 1024 and a power of 2");
            check_status(st, "vren::radix_sort");
        }

        // key-value extension (not in the reference): stable ascending by key, scratch from scratch_bytes_pairs()
        static size_t scratch_bytes_pairs(uint32_t length) { return vrenb200_radix_sort_scratch_bytes(length, 1); }
        void sort_pairs(VkCommandBuffer command_buffer, vren::vk_utils::buffer const& keys, vren::vk_utils::buffer const& values, uint32_t length,
                        vren::vk_utils::buffer const& scratch)
        {
            check_status(vrenb200_radix_sort_pairs((vrenb200_stream_t) command_buffer, keys.ptr<uint32_t>(), values.ptr<uint32_t>(), length,
                                                   scratch.m_ptr, scratch.m_size),
                         "vren::radix_sort::sort_pairs");
        }
    };
}
