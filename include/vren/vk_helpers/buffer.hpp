// vren::vk_utils::buffer facade (vren/vren/vk_helpers/buffer.hpp:24-35): a device allocation + its size.
// VkBuffer/VmaAllocation become a raw device pointer; ownership is RAII like the reference's vk_raii handles.
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

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

    // Device scratch of a primitive: a pool of blocks, the analogue of the reference's descriptor pools
    // (pool/object_pool.hpp:60-117).  operator() LEASES a block and parks the lease in the caller's resource_container; the
    // block returns to the pool when the container lets go of it (after the stream has retired).  Calls in flight never share
    // a block, nothing is freed or synchronised while recording, and a pool that was given its size up front
    // (reserve(), e.g. from the constructor) never allocates while recording either.
    class scratch_pool
    {
        struct block { void* ptr; size_t size; };
        struct state { std::vector<block> free_blocks; ~state() { for (block& b : free_blocks) cudaFree(b.ptr); } };
        std::shared_ptr<state> m_state = std::make_shared<state>();

        static block allocate(size_t bytes)
        {
            block b{nullptr, bytes < 256 ? 256 : bytes};
            if (cudaMalloc(&b.ptr, b.size) != cudaSuccess) throw std::runtime_error("cudaMalloc failed");
            return b;
        }

    public:
        scratch_pool() = default;
        scratch_pool(scratch_pool const&) = delete;

        // make sure `count` blocks of at least `bytes` are ready (constructors / set-up code)
        void reserve(size_t bytes, size_t count = 1)
        {
            size_t have = 0;
            for (block const& b : m_state->free_blocks) have += b.size >= bytes;
            for (; have < count; have++) m_state->free_blocks.push_back(allocate(bytes));
        }

        // a block of at least `bytes`, owned by the returned lease (park it in the resource_container of the call)
        std::shared_ptr<void> acquire(size_t bytes)
        {
            block got{nullptr, 0};
            std::vector<block>& fb = m_state->free_blocks;
            for (size_t i = 0; i < fb.size(); i++)
                if (fb[i].size >= bytes && (got.ptr == nullptr || fb[i].size < got.size)) got = fb[i];
            if (got.ptr != nullptr)
            {
                for (size_t i = 0; i < fb.size(); i++)
                    if (fb[i].ptr == got.ptr) { fb.erase(fb.begin() + i); break; }
            }
            else
                got = allocate(bytes);
            std::weak_ptr<state> pool = m_state;
            return std::shared_ptr<void>(got.ptr, [pool, got](void*) {
                if (auto p = pool.lock()) p->free_blocks.push_back(got);      // back to the pool
                else cudaFree(got.ptr);                                        // the primitive is gone
            });
        }
    };
}
