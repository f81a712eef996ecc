// vren::context / vren::toolbox facade (vren/vren/context.hpp:37-91, toolbox.hpp:43-55).
// The reference context owns the Vulkan device and a toolbox holding one instance of every primitive; here the
// context is a CUDA device ordinal and the toolbox keeps the same member names.
#pragma once

#include <memory>

#include "primitives/blelloch_scan.hpp"
#include "primitives/bucket_sort.hpp"
#include "primitives/build_bvh.hpp"
#include "primitives/radix_sort.hpp"
#include "primitives/reduce.hpp"

namespace vren
{
    class toolbox;

    class context
    {
    public:
        int m_device = 0;
        std::unique_ptr<vren::toolbox> m_toolbox;

        explicit context(int device = 0);
        ~context();
    };

    class toolbox // toolbox.hpp:43-55
    {
    public:
        vren::reduce<uint32_t, vren::ReduceOperationAdd> m_reduce_uint_add;
        vren::reduce<uint32_t, vren::ReduceOperationMin> m_reduce_uint_min;
        vren::reduce<uint32_t, vren::ReduceOperationMax> m_reduce_uint_max;
        vren::reduce<glm_compat::vec4, vren::ReduceOperationAdd> m_reduce_vec4_add;
        vren::reduce<glm_compat::vec4, vren::ReduceOperationMin> m_reduce_vec4_min;
        vren::reduce<glm_compat::vec4, vren::ReduceOperationMax> m_reduce_vec4_max;
        vren::blelloch_scan m_blelloch_scan;
        vren::radix_sort m_radix_sort;
        vren::bucket_sort m_bucket_sort;
        vren::build_bvh m_build_bvh;

        explicit toolbox(vren::context const& c) :
            m_reduce_uint_add(c), m_reduce_uint_min(c), m_reduce_uint_max(c), m_reduce_vec4_add(c), m_reduce_vec4_min(c),
            m_reduce_vec4_max(c), m_blelloch_scan(c), m_radix_sort(c), m_bucket_sort(c), m_build_bvh(c)
        {
        }
    };

    inline context::context(int device) : m_device(device)
    {
        if (cudaSetDevice(device) != cudaSuccess) throw std::runtime_error("cudaSetDevice failed");
        m_toolbox = std::make_unique<vren::toolbox>(*this);
    }
    inline context::~context() = default;

    namespace vk_utils
    {
        // immediate_graphics_queue_submit (vk_helpers/misc.cpp:72-124): record, submit, wait for the fence
        template <typename _fn_t> void immediate_graphics_queue_submit(vren::context const&, _fn_t&& record)
        {
            cudaStream_t stream;
            if (cudaStreamCreate(&stream) != cudaSuccess) throw std::runtime_error("cudaStreamCreate failed");
            vren::resource_container resources;
            try { record(stream, resources); }
            catch (...) { cudaStreamDestroy(stream); throw; }
            const cudaError_t e = cudaStreamSynchronize(stream);
            cudaStreamDestroy(stream);
            if (e != cudaSuccess) throw std::runtime_error(std::string("device error: ") + cudaGetErrorString(e));
        }
    }
}
