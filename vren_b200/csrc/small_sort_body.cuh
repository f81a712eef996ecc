// Body of the single-CTA radix sort (small_sort.cu), kept in its own header so that the SAME source can be executed on the
// host: tests/cpp/cta_emulator.hpp provides threadIdx, the barriers and the warp intrinsics over 1024 OS threads, and
// tests/test_small_sort_emulation.py runs it there (also under ThreadSanitizer: a missing barrier is a data race).
//
// Reference semantics: vren::radix_sort (radix_sort.cpp:149-337) — ascending sort of uint32 keys — plus the key / value
// extension (stable by key).  The reference runs 8 passes x ~14 dispatches whatever the size; its own test sorts 1024 keys
// (vren_test radix_sort.cpp:130-143).  Here, for n <= kSmallSortMax, ONE launch of ONE CTA does all four 8-bit passes in
// shared memory: no histogram kernel, no look-back, no scratch.
//
// Element order.  `rows` = ceil(n / 1024) rows of 32 elements per warp are in use (1..8); element e = warp * rows * 32 +
// row * 32 + lane, so (warp, row, lane) order is element order and the ranking below is stable.  Slots e >= n are padding
// with key 0xFFFFFFFF: last in element order and never smaller than a real key, they stay behind every real element.
//
// One pass (digit = byte `pass` of the key):
//   1. every thread holds its <= 8 elements in registers; the (warp, digit) counters are cleared;
//   2. row by row, the lanes of a warp that hold the same digit find each other (match.any); the highest lane of a group
//      bumps the warp-private counter by the group's size and the group reads the old value from it by shuffle:
//      rank inside the (warp, digit) run = old value + number of lower lanes of the group (order by construction);
//   3. thread d < 256 turns column d of the counters into exclusive offsets over the warps and keeps the digit's total; a
//      block-wide exclusive scan of the totals gives the digit bases;
//   4. every element goes to base[digit] + offset[warp][digit] + rank in the SAME shared arrays (all elements were in
//      registers before step 3's barrier, so nothing is overwritten before it has been read).
#pragma once

#include <stdint.h>

namespace vrenb200 {

constexpr int kSmallSortThreads = 1024;
constexpr int kSmallSortWarps = kSmallSortThreads / 32;
constexpr int kSmallSortRows = 8;
constexpr uint32_t kSmallSortMax = (uint32_t) kSmallSortThreads * kSmallSortRows;   // 8192 elements

template <bool HAS_VALUES>
struct small_sort_smem
{
    uint32_t keys[kSmallSortMax];
    uint32_t vals[HAS_VALUES ? kSmallSortMax : 1];
    uint32_t count[kSmallSortWarps][256];   // per (warp, digit): count, then exclusive offset over the warps
    uint32_t base[256];                     // per digit: exclusive offset over the digits
    uint32_t warp_total[8];
};

template <bool HAS_VALUES>
__device__ __forceinline__ void single_cta_sort_body(unsigned char* smem_raw, uint32_t* keys, uint32_t* vals, uint32_t n,
                                                     int first_pass, int num_passes)
{
    small_sort_smem<HAS_VALUES>& sm = *reinterpret_cast<small_sort_smem<HAS_VALUES>*>(smem_raw);
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const unsigned lower_lanes = (1u << lane) - 1u;
    const uint32_t rows = (n + (uint32_t) kSmallSortThreads - 1u) / (uint32_t) kSmallSortThreads;   // 1..kSmallSortRows
    const uint32_t first = warp * rows * 32u + lane;                                                // this thread's element of row 0

    uint32_t key[kSmallSortRows], val[kSmallSortRows], rank[kSmallSortRows];
#pragma unroll
    for (int j = 0; j < kSmallSortRows; j++)
    {
        key[j] = 0xFFFFFFFFu;
        val[j] = 0u;
        rank[j] = 0u;
        if ((uint32_t) j < rows)
        {
            const uint32_t e = first + (uint32_t) j * 32u;
            if (e < n)
            {
                key[j] = keys[e];
                if (HAS_VALUES) val[j] = vals[e];
            }
        }
    }

    for (int pass = first_pass; pass < first_pass + num_passes; pass++)
    {
        const uint32_t shift = (uint32_t) pass * 8u;
        // 1. clear the counters
        for (uint32_t i = tid; i < (uint32_t) kSmallSortWarps * 256u; i += (uint32_t) kSmallSortThreads) (&sm.count[0][0])[i] = 0u;
        __syncthreads();
        // 2. rank inside the (warp, digit) runs
#pragma unroll
        for (int j = 0; j < kSmallSortRows; j++)
        {
            if ((uint32_t) j < rows)     // uniform over the CTA
            {
                const uint32_t d = (key[j] >> shift) & 0xFFu;
                const unsigned group = __match_any_sync(0xFFFFFFFFu, d);
                const unsigned leader = 31u - (unsigned) __clz((int) group);
                uint32_t old = 0u;
                if (lane == leader)
                {
                    old = sm.count[warp][d];
                    sm.count[warp][d] = old + (uint32_t) __popc(group);
                }
                __syncwarp();            // the counter update of this row is visible to the leaders of the next row
                old = __shfl_sync(0xFFFFFFFFu, old, (int) leader);
                rank[j] = old + (uint32_t) __popc(group & lower_lanes);
            }
        }
        __syncthreads();
        // 3. counts -> offsets
        uint32_t total = 0u, inclusive = 0u;
        if (tid < 256u)
        {
            for (int w = 0; w < kSmallSortWarps; w++)
            {
                const uint32_t c = sm.count[w][tid];
                sm.count[w][tid] = total;
                total += c;
            }
            inclusive = total;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1)
            {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inclusive, (unsigned) s);
                if (lane >= (unsigned) s) inclusive += t;
            }
            if (lane == 31u) sm.warp_total[warp] = inclusive;
        }
        __syncthreads();
        if (tid < 256u)
        {
            uint32_t before = 0u;
            for (unsigned w = 0; w < warp; w++) before += sm.warp_total[w];
            sm.base[tid] = before + inclusive - total;
        }
        __syncthreads();
        // 4. scatter (every element of the CTA is in registers: the arrays can be overwritten in place)
#pragma unroll
        for (int j = 0; j < kSmallSortRows; j++)
        {
            if ((uint32_t) j < rows)
            {
                const uint32_t d = (key[j] >> shift) & 0xFFu;
                const uint32_t pos = sm.base[d] + sm.count[warp][d] + rank[j];
                sm.keys[pos] = key[j];
                if (HAS_VALUES) sm.vals[pos] = val[j];
            }
        }
        __syncthreads();
        // the next pass (or the write-out) takes the elements in their new order
#pragma unroll
        for (int j = 0; j < kSmallSortRows; j++)
        {
            if ((uint32_t) j < rows)
            {
                const uint32_t e = first + (uint32_t) j * 32u;
                key[j] = sm.keys[e];
                if (HAS_VALUES) val[j] = sm.vals[e];
            }
        }
        // (the clearing loop and the barrier at the top of the next pass separate these reads from the next scatter)
    }

    // write-out: the first n slots hold the real elements (the padding sorted behind them)
#pragma unroll
    for (int j = 0; j < kSmallSortRows; j++)
    {
        if ((uint32_t) j < rows)
        {
            const uint32_t e = first + (uint32_t) j * 32u;
            if (e < n)
            {
                keys[e] = key[j];
                if (HAS_VALUES) vals[e] = val[j];
            }
        }
    }
}

} // namespace vrenb200
