// vren/base/base.hpp facade — integer helpers with the reference's names (vren/vren/base/base.hpp:32-79),
// evaluated in pure integer arithmetic by the C ABI.
#pragma once

#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>

#include "../../vrenb200.h"

namespace vren
{
    inline bool is_power_of_2(uint32_t value) { return vrenb200_is_power_of_2(value) != 0; }
    inline uint32_t round_to_next_power_of_2(uint32_t v) { return vrenb200_round_to_next_power_of_2(v); }
    template <typename _integer_t> _integer_t round_to_next_multiple_of(_integer_t value, _integer_t multiple)
    {
        return (_integer_t) vrenb200_round_to_next_multiple_of((uint64_t) value, (uint64_t) multiple);
    }
    template <typename _integer_t> bool is_power_of(_integer_t n, _integer_t base) { return vrenb200_is_power_of((uint32_t) n, (uint32_t) base) != 0; }
    template <typename _integer_t> _integer_t round_to_next_power_of(_integer_t n, _integer_t base)
    {
        return (_integer_t) vrenb200_round_to_next_power_of((uint32_t) n, (uint32_t) base);
    }
    inline uint32_t divide_and_ceil(uint32_t value, uint32_t divider) { return vrenb200_divide_and_ceil(value, divider); }

    // base/resource_container.hpp:8-38: per-call objects (pooled descriptor sets in the reference, pooled scratch blocks
    // here) are parked in it until the command buffer — the stream — they were recorded into has retired; the caller
    // clears or destroys the container after synchronising, exactly as with the reference (reduce.cpp:64-78).
    class resource_container
    {
        std::vector<std::shared_ptr<void>> m_resources;

    public:
        template <typename _t> void add_resource(std::shared_ptr<_t> resource) { m_resources.push_back(std::move(resource)); }
        template <typename... _t> void add_resources(std::shared_ptr<_t>... resources) { (add_resource(std::move(resources)), ...); }
        void clear() { m_resources.clear(); }
        size_t size() const { return m_resources.size(); }
    };
}
