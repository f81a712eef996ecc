// vren::reduce<T, op> facade — same class, method and argument order as vren/vren/primitives/reduce.hpp:15-49.
#pragma once

#include "../base/base.hpp"
#include "../vk_helpers/buffer.hpp"

namespace glm_compat
{
    struct vec4 { float x, y, z, w; };
    using uint = uint32_t;
}

namespace vren
{
    enum reduce_operation { ReduceOperationAdd, ReduceOperationMin, ReduceOperationMax };

    template <typename _data_type_t> struct reduce_dtype;
    template <> struct reduce_dtype<uint32_t> { static constexpr int value = VRENB200_U32; };
    template <> struct reduce_dtype<float> { static constexpr int value = VRENB200_F32; };
    template <> struct reduce_dtype<glm_compat::vec4> { static constexpr int value = VRENB200_VEC4; };

    template <typename _data_type_t, vren::reduce_operation _operation_t>
    class reduce
    {
    public:
        inline static const uint32_t k_workgroup_size = 1024;

        explicit reduce(vren::context const&) {}

        // reduce.cpp:33-114: full up-sweep tree into output_buffer (next_pow2(length) elements per block)
        void operator()(VkCommandBuffer command_buffer, vren::resource_container&, vren::vk_utils::buffer const& input_buffer,
                        uint32_t input_buffer_length, size_t input_buffer_offset, vren::vk_utils::buffer const& output_buffer,
                        size_t output_buffer_offset, uint32_t blocks_num)
        {
            check_status(vrenb200_reduce((vrenb200_stream_t) command_buffer, reduce_dtype<_data_type_t>::value, (int) _operation_t,
                                         VRENB200_REDUCE_TREE, input_buffer.ptr<>(input_buffer_offset), input_buffer_length,
                                         output_buffer.ptr<>(output_buffer_offset), blocks_num, nullptr, 0),
                         "vren::reduce");
        }

        // reduce.cpp:116-128: in place
        void operator()(VkCommandBuffer command_buffer, vren::resource_container& resource_container, vren::vk_utils::buffer const& buffer,
                        uint32_t length, size_t offset, uint32_t blocks_num)
        {
            (*this)(command_buffer, resource_container, buffer, length, offset, buffer, offset, blocks_num);
        }
    };

    inline uint32_t calc_reduce_output_buffer_length(uint32_t count) { return vrenb200_calc_reduce_output_buffer_length(count); }
}
