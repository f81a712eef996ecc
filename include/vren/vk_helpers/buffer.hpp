// vren::vk_utils::buffer facade (vren/vren/vk_helpers/buffer.hpp:24-35): a device allocation + its size.
// VkBuffer/VmaAllocation become a raw device pointer; ownership is RAII like the reference's vk_raii handles.
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <stdexcept>
#include <string>
#include <utility>

#include "../../vrenb200.h"

// VkCommandBuffer -> cudaStream_t: "recording" a primitive == enqueueing it on the stream
using VkCommandBuffer = cudaStream_t;

#define VREN_MIN_STORAGE_BUFFER_OFFSET_ALIGNMENT 256ull   // config.hpp:30

namespace vren
{
    class context; // context.hpp

    // VREN_CHECK (vk_helpers/misc.cpp:22-39): non-success -> std::runtime_error
    inline void check_status(int status, char const* what)
    {
        if (status == VRENB200_OK) return;
        std::string msg = std::string(what) + ": " + vrenb200_status_string(status);
        if (status == VRENB200_ECUDA) msg += std::string(" (") + cudaGetErrorString((cudaError_t) vrenb200_last_cuda_error()) + ")";
        if (status == VRENB200_EINVAL_LENGTH) throw std::invalid_argument(msg);
        throw std::runtime_error(msg);
    }

    namespace vk_utils
    {
        struct buffer
        {
            void* m_ptr = nullptr;     // device pointer (m_buffer.m_handle in the reference)
            size_t m_size = 0;         // m_allocation_info.size
            bool m_owned = false;

            buffer() = default;
            buffer(void* ptr, size_t size) : m_ptr(ptr), m_size(size), m_owned(false) {} // non-owning view
            buffer(buffer const&) = delete;
            buffer& operator=(buffer const&) = delete;
            buffer(buffer&& o) noexcept { *this = std::move(o); }
            buffer& operator=(buffer&& o) noexcept
            {
                if (this != &o)
                {
                    release();
                    m_ptr = o.m_ptr; m_size = o.m_size; m_owned = o.m_owned;
                    o.m_ptr = nullptr; o.m_size = 0; o.m_owned = false;
                }
                return *this;
            }
            ~buffer() { release(); }

            template <typename T = void> T* ptr(size_t byte_offset = 0) const { return reinterpret_cast<T*>(static_cast<char*>(m_ptr) + byte_offset); }

        private:
            void release()
            {
                if (m_owned && m_ptr) cudaFree(m_ptr);
                m_ptr = nullptr;
            }
        };

        // alloc_device_only_buffer (vk_helpers/buffer.cpp) — usage flags have no CUDA meaning and are dropped
        inline buffer alloc_device_only_buffer(vren::context const&, size_t size)
        {
            buffer b;
            if (cudaMalloc(&b.m_ptr, size == 0 ? 256 : size) != cudaSuccess) throw std::runtime_error("cudaMalloc failed");
            b.m_size = size;
            b.m_owned = true;
            return b;
        }
    }

    // grow-only device scratch owned by a primitive (the analogue of the reference's pooled descriptor sets)
    class scratch_arena
    {
        void* m_ptr = nullptr;
        size_t m_size = 0;

    public:
        scratch_arena() = default;
        scratch_arena(scratch_arena const&) = delete;
        ~scratch_arena() { if (m_ptr) cudaFree(m_ptr); }
        void* reserve(size_t bytes)
        {
            if (bytes > m_size)
            {
                if (m_ptr) { cudaDeviceSynchronize(); cudaFree(m_ptr); }
                if (cudaMalloc(&m_ptr, bytes) != cudaSuccess) { m_ptr = nullptr; m_size = 0; throw std::runtime_error("cudaMalloc failed"); }
                m_size = bytes;
            }
            return m_ptr;
        }
        size_t size() const { return m_size; }
    };
}
