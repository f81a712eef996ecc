// vren::blelloch_scan facade — vren/vren/primitives/blelloch_scan.hpp:9-49.
#pragma once

#include "reduce.hpp"

namespace vren
{
    class blelloch_scan
    {
    public:
        inline static const uint32_t k_workgroup_size = 1024;
        inline static const uint32_t k_max_items = 1;

    private:
        vren::scratch_pool m_scratch; // look-back status words of the single-pass scan

    public:
        explicit blelloch_scan(vren::context const&) {}

        // blelloch_scan.cpp:57-139
        void downsweep(VkCommandBuffer command_buffer, vren::resource_container&, vren::vk_utils::buffer const& buffer, uint32_t length,
                       uint32_t offset, uint32_t blocks_num, bool clear_last)
        {
            check_status(vrenb200_blelloch_downsweep_u32((vrenb200_stream_t) command_buffer, buffer.ptr<uint32_t>(offset), length, blocks_num,
                                                         clear_last ? 1 : 0),
                         "vren::blelloch_scan::downsweep");
        }

        // blelloch_scan.cpp:141-166. The reference reduces block 0 only (blocks_num hard-coded to 1, :151) and then
        // down-sweeps every block; that quirk is kept: block 0 is scanned, further blocks only get the down-sweep.
        void operator()(VkCommandBuffer command_buffer, vren::resource_container& resource_container, vren::vk_utils::buffer const& buffer,
                        uint32_t length, uint32_t offset, uint32_t blocks_num)
        {
            if (!vren::is_power_of_2(length)) throw std::invalid_argument("vren::blelloch_scan: length must be a power of 2"); // :67
            const size_t bytes = vrenb200_scan_scratch_bytes(length);
            std::shared_ptr<void> lease = m_scratch.acquire(bytes);
            void* scratch = lease.get();
            resource_container.add_resource(lease);
            uint32_t* data = buffer.ptr<uint32_t>(offset);
            check_status(vrenb200_exclusive_scan_u32((vrenb200_stream_t) command_buffer, data, data, length, scratch, bytes), "vren::blelloch_scan");
            if (blocks_num > 1)
            {
                vren::vk_utils::buffer rest(data + length, (size_t) (blocks_num - 1) * length * sizeof(uint32_t));
                downsweep(command_buffer, resource_container, rest, length, 0, blocks_num - 1, true);
            }
        }
    };
}
