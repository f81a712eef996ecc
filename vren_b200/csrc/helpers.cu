// Library-wide bits: version, status strings and the integer helpers of vren/vren/base/base.hpp:32-79 and
// vren/vren/primitives/build_bvh.cpp:99-136, restated in pure integer arithmetic (the reference goes through
// double log/pow, which is only exact for the small arguments its tests use).
#include "common.cuh"

namespace vrenb200 {
thread_local int g_last_cuda_error = 0;
}

using namespace vrenb200;

extern "C" const char* vrenb200_version(void) { return "vrenb200 0.1.0 (sm_100a)"; }

extern "C" const char* vrenb200_status_string(int status)
{
    switch (status)
    {
    case VRENB200_OK: return "ok";
    case VRENB200_EINVAL_LENGTH: return "invalid length";
    case VRENB200_EALIGN: return "misaligned pointer or offset";
    case VRENB200_ESCRATCH: return "scratch buffer missing or too small";
    case VRENB200_ECUDA: return "CUDA runtime error";
    case VRENB200_EINVAL_ARG: return "invalid argument";
    case VRENB200_ELIMIT: return "implementation limit exceeded";
    default: return "unknown status";
    }
}

extern "C" int vrenb200_last_cuda_error(void) { return g_last_cuda_error; }

// ---- base/base.hpp ---------------------------------------------------------------------------------------------
extern "C" int vrenb200_is_power_of_2(uint32_t v) { return v > 0 && (v & (v - 1)) == 0; }

extern "C" uint32_t vrenb200_round_to_next_power_of_2(uint32_t v) { return next_pow2_u32(v); }

extern "C" uint64_t vrenb200_round_to_next_multiple_of(uint64_t v, uint64_t multiple)
{
    const uint64_t r = v % multiple;
    return r == 0 ? v : v + multiple - r;
}

extern "C" uint32_t vrenb200_divide_and_ceil(uint32_t v, uint32_t d) { return (uint32_t) (((uint64_t) v + d - 1) / d); }

extern "C" int vrenb200_is_power_of(uint32_t n, uint32_t base)
{
    if (n == 0 || base < 2) return n == 1;
    while (n % base == 0) n /= base;
    return n == 1;
}

extern "C" uint32_t vrenb200_round_to_next_power_of(uint32_t n, uint32_t base)
{
    uint64_t p = 1;
    while (p < n) p *= base;
    return (uint32_t) p;
}

// ---- build_bvh.cpp:99-136 ----------------------------------------------------------------------------------
extern "C" uint32_t vrenb200_calc_bvh_padded_leaf_count(uint32_t leaf_count)
{
    return leaf_count <= 1 ? 32u : vrenb200_round_to_next_power_of(leaf_count, 32u);
}

extern "C" uint32_t vrenb200_calc_bvh_buffer_length(uint32_t leaf_count)
{
    uint32_t padded = vrenb200_calc_bvh_padded_leaf_count(leaf_count);
    uint64_t length = 0;
    while (padded != 0)
    {
        length += padded;
        padded >>= 5;
    }
    return (uint32_t) length;
}

extern "C" size_t vrenb200_calc_bvh_buffer_size(uint32_t leaf_count)
{
    return (size_t) vrenb200_calc_bvh_buffer_length(leaf_count) * sizeof(vrenb200_bvh_node);
}

extern "C" uint32_t vrenb200_calc_bvh_root_index(uint32_t leaf_count)
{
    return vrenb200_calc_bvh_buffer_length(leaf_count) - 1;
}

extern "C" uint32_t vrenb200_calc_bvh_level_count(uint32_t leaf_count)
{
    uint32_t padded = vrenb200_calc_bvh_padded_leaf_count(leaf_count);
    uint32_t levels = 0;
    while (padded > 1)
    {
        padded >>= 5;
        levels++;
    }
    return levels;
}
