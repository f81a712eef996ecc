// Shared device/host helpers for the sm_100a kernels behind include/vrenb200.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include "../../include/vrenb200.h"

namespace vrenb200 {

constexpr int kNumSMs = 148;           // B200: 2 dies x 74 SMs
constexpr unsigned kFullMask = 0xFFFFFFFFu;

extern thread_local int g_last_cuda_error;

inline int check_cuda(cudaError_t e)
{
    if (e != cudaSuccess)
    {
        g_last_cuda_error = (int) e;
        return VRENB200_ECUDA;
    }
    return VRENB200_OK;
}

// launch check: catches bad configurations without synchronising
inline int check_launch() { return check_cuda(cudaGetLastError()); }

#define VRENB200_TRY(expr)                       \
    do                                           \
    {                                            \
        int _st = (expr);                        \
        if (_st != VRENB200_OK) return _st;      \
    } while (0)

inline cudaStream_t as_stream(vrenb200_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline uint32_t next_pow2_u32(uint32_t v)
{
    // base/base.hpp:37-47 (0 -> 0, like the reference's wrap-around)
    v--;
    v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16;
    v++;
    return v;
}

// carve sub-buffers out of a scratch blob, 256-byte aligned like VREN_MIN_STORAGE_BUFFER_OFFSET_ALIGNMENT
struct scratch_carver
{
    char* base;
    size_t offset = 0;
    explicit scratch_carver(void* p) : base(static_cast<char*>(p)) {}
    template <typename T> T* take(size_t count)
    {
        offset = align_up(offset, 256);
        T* r = reinterpret_cast<T*>(base + offset);
        offset += count * sizeof(T);
        return r;
    }
    size_t used() const { return align_up(offset, 256); }
};

#ifdef __CUDACC__

__device__ __forceinline__ unsigned lane_id()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(r));
    return r;
}

__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%lanemask_lt;" : "=r"(r));
    return r;
}

__device__ __forceinline__ unsigned lanemask_ge()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%lanemask_ge;" : "=r"(r));
    return r;
}

// single-word flag+value protocol of the decoupled look-back: relaxed gpu-scope accesses are enough
// because status and payload travel in the same 32/64-bit word
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_relaxed_u32(uint32_t* p, uint32_t v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t* p)
{
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_relaxed_u64(uint64_t* p, uint64_t v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// streaming loads/stores: data touched exactly once per pass must not evict the look-back state from L2/L1
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4* p)
{
    uint4 v;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}

__device__ __forceinline__ uint32_t ldg_stream_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.global.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// smallest float s >= 0 with sqrt_rn(s) >= r.  sqrt_rn is monotone, so for every s >= 0:
//   sqrt_rn(s) < r  <=>  s < sq_threshold(r)      (r <= 0 or NaN: never true -> 0)
__device__ __forceinline__ float sq_threshold(float r)
{
    if (!(r > 0.0f)) return 0.0f;
    uint32_t t = __float_as_uint(__fmul_rn(r, r));          // non-negative floats order like their bit patterns
    while (t > 0 && __fsqrt_rn(__uint_as_float(t)) >= r) t--;
    while (t < 0x7F800000u && __fsqrt_rn(__uint_as_float(t)) < r) t++;
    return __uint_as_float(t);
}

#endif // __CUDACC__

} // namespace vrenb200
