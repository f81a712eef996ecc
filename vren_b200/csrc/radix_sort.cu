// a3 — vren::radix_sort   (reference: vren/vren/primitives/radix_sort.{hpp,cpp}, shaders/radix_sort_*.comp)
// a4 — vren::bucket_sort  (reference: vren/vren/primitives/bucket_sort.{hpp,cpp}, shaders/bucket_sort_*.comp)
//
// Reference radix sort: LSD, 4-bit digits, 8 passes, each pass = fill + local_count + reduce + global_offset +
// downsweep + reorder (~14 dispatches, 12n bytes of key traffic per pass, one key per thread, n pow2 >= 1024).
// Reference bucket sort: global atomics on 65536 counters (count), blelloch scan, atomics again (write): unstable.
//
// Here (B200-first): "onesweep" — 8-bit digits, one fused kernel per digit.
//   1. ONE histogram kernel reads the keys once (128-bit loads) and builds every 256-bin digit histogram.
//   2. One tiny kernel turns them into exclusive digit offsets.
//   3. Per digit ONE fused kernel (onesweep_pass_kernel below).  A CTA takes a tile, stages keys (and values) into shared
//      memory with single-thread TMA bulk copies (cp.async.bulk + mbarrier), counts the tile's digits, publishes them and
//      starts a decoupled look-back over a flag|count word per (tile, digit), ranks every pair into its final in-tile slot,
//      regroups the tile by digit in shared memory and writes it out coalesced.
// Radix sort = 4 digits: 4n (histogram) + 4 x 8n keys [+ 4 x 8n values] = 36 B/key, 68 B/pair of HBM traffic.
// Bucket sort = the same kernel over interleaved uvec2 pairs with 2 digits (16-bit key): 8n + 2 x 16n = 40 B/pair,
// deterministic and stable (the canonical tie-break), bucket END offsets from a fused 65536-bin count or a search.
// Stability: warp-striped order (warp, item, lane) == element order, so equal keys keep input order.
//
// The same kernel also serves the receive side of the multi-GPU sort (sharded_sort.cu): F_SEGMENTED sorts many independent
// tile-aligned segments in one launch (look-back restarts at every segment, digit offsets per segment).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "radix_internal.cuh"

namespace vrenb200 {

namespace {

constexpr uint32_t kLbFlagAggregate = 1u << 30;
constexpr uint32_t kLbFlagInclusive = 2u << 30;
constexpr uint32_t kLbValueMask = (1u << 30) - 1;

constexpr uint32_t kBucketKeys = 1u << 16; // bucket_sort.hpp:15-16

// ---- mbarrier / bulk-copy (TMA 1D) wrappers --------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src_gmem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- 1. histograms in one read of the keys -------------------------------------------------------------------
constexpr int kHistThreads = 512;

__global__ void __launch_bounds__(kHistThreads)
radix_histogram_kernel(const uint32_t* __restrict__ keys, uint32_t n, sort_control* ctl)
{
    __shared__ uint32_t s_hist[kPasses][kRadix];
    for (int i = threadIdx.x; i < kPasses * kRadix; i += kHistThreads) (&s_hist[0][0])[i] = 0;
    __syncthreads();

    const uint32_t n4 = n / 4;
    const uint4* keys4 = reinterpret_cast<const uint4*>(keys);
    const uint32_t stride = gridDim.x * kHistThreads;
    auto count = [&](uint32_t k) {
        atomicAdd(&s_hist[0][k & 0xFF], 1u);
        atomicAdd(&s_hist[1][(k >> 8) & 0xFF], 1u);
        atomicAdd(&s_hist[2][(k >> 16) & 0xFF], 1u);
        atomicAdd(&s_hist[3][k >> 24], 1u);
    };
    uint32_t i = blockIdx.x * kHistThreads + threadIdx.x;
    // four independent 128-bit loads in flight per thread per iteration (the kernel is latency-bound otherwise)
    for (; (uint64_t) i + 3ull * stride < n4; i += 4 * stride)
    {
        const uint4 a = ldg_stream_u4(keys4 + i);
        const uint4 b = ldg_stream_u4(keys4 + i + stride);
        const uint4 c = ldg_stream_u4(keys4 + i + 2 * stride);
        const uint4 d = ldg_stream_u4(keys4 + i + 3 * stride);
        count(a.x); count(a.y); count(a.z); count(a.w);
        count(b.x); count(b.y); count(b.z); count(b.w);
        count(c.x); count(c.y); count(c.z); count(c.w);
        count(d.x); count(d.y); count(d.z); count(d.w);
    }
    for (; i < n4; i += stride)
    {
        const uint4 a = ldg_stream_u4(keys4 + i);
        count(a.x); count(a.y); count(a.z); count(a.w);
    }
    if (blockIdx.x == 0)
        for (uint32_t t = n4 * 4 + threadIdx.x; t < n; t += kHistThreads) count(keys[t]);
    __syncthreads();
    for (int t = threadIdx.x; t < kPasses * kRadix; t += kHistThreads)
    {
        const uint32_t c = (&s_hist[0][0])[t];
        if (c != 0) atomicAdd(&(&ctl->hist[0][0])[t], c);
    }
}

// Conflict-free variant of the histogram: every lane owns a column of every counter (hist[pass][digit][lane], 128 KB of
// the SM's shared memory, one 1024-thread CTA per SM), so the 32 shared atomics of a warp instruction always hit 32
// different banks: one wavefront per instruction instead of ~3.4 for 32 random digits.  The shared-atomic pipe was the
// limiter of the plain version (0.40 ms at 2^28 keys).
constexpr int kHist2Threads = 1024;
constexpr size_t kHist2Smem = (size_t) kPasses * kRadix * 32 * sizeof(uint32_t);

__global__ void __launch_bounds__(kHist2Threads, 1)
radix_histogram_columns_kernel(const uint32_t* __restrict__ keys, uint32_t n, sort_control* ctl)
{
    extern __shared__ __align__(128) uint32_t s_cols[];   // [kPasses][kRadix][32]
    for (uint32_t i = threadIdx.x; i < kPasses * kRadix * 32; i += kHist2Threads) s_cols[i] = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31;
    uint32_t* col = s_cols + lane;
    auto count = [&](uint32_t k) {
        atomicAdd(&col[(0 * kRadix + (k & 0xFF)) * 32], 1u);
        atomicAdd(&col[(1 * kRadix + ((k >> 8) & 0xFF)) * 32], 1u);
        atomicAdd(&col[(2 * kRadix + ((k >> 16) & 0xFF)) * 32], 1u);
        atomicAdd(&col[(3 * kRadix + (k >> 24)) * 32], 1u);
    };
    const uint32_t n4 = n / 4;
    const uint4* keys4 = reinterpret_cast<const uint4*>(keys);
    const uint32_t stride = gridDim.x * kHist2Threads;
    uint32_t i = blockIdx.x * kHist2Threads + threadIdx.x;
    // four 128-bit loads in flight per thread (64 KB per SM): one CTA of 1024 threads per SM needs that much to cover the
    // HBM latency at full bandwidth (0.207 vs 0.217 ms at 2^28 keys with two)
    for (; (uint64_t) i + 3ull * stride < n4; i += 4 * stride)
    {
        const uint4 a = ldg_stream_u4(keys4 + i);
        const uint4 b = ldg_stream_u4(keys4 + i + stride);
        const uint4 c = ldg_stream_u4(keys4 + i + 2 * stride);
        const uint4 d = ldg_stream_u4(keys4 + i + 3 * stride);
        count(a.x); count(a.y); count(a.z); count(a.w);
        count(b.x); count(b.y); count(b.z); count(b.w);
        count(c.x); count(c.y); count(c.z); count(c.w);
        count(d.x); count(d.y); count(d.z); count(d.w);
    }
    for (; (uint64_t) i + stride < n4; i += 2 * stride)
    {
        const uint4 a = ldg_stream_u4(keys4 + i);
        const uint4 b = ldg_stream_u4(keys4 + i + stride);
        count(a.x); count(a.y); count(a.z); count(a.w);
        count(b.x); count(b.y); count(b.z); count(b.w);
    }
    for (; i < n4; i += stride)
    {
        const uint4 a = ldg_stream_u4(keys4 + i);
        count(a.x); count(a.y); count(a.z); count(a.w);
    }
    if (blockIdx.x == 0)
        for (uint32_t t = n4 * 4 + threadIdx.x; t < n; t += kHist2Threads) count(keys[t]);
    __syncthreads();
    // thread t sums the 32 columns of counter t, starting at a rotated column so that a warp reads 32 different banks
    const uint32_t t = threadIdx.x; // == pass * 256 + digit
    uint32_t sum = 0;
#pragma unroll 8
    for (uint32_t k = 0; k < 32; k++) sum += s_cols[t * 32 + ((k + t) & 31)];
    if (sum != 0) atomicAdd(&(&ctl->hist[0][0])[t], sum);
}

// bucket sort: uvec2 pairs, key = x & 0xFFFF -> two digit histograms + the 65536 bucket counters
// (bucket_sort_count.comp:27-34) in the same read
__global__ void __launch_bounds__(kHistThreads)
bucket_histogram_kernel(const uint2* __restrict__ pairs, uint32_t n, sort_control* ctl, uint32_t* bucket_counters)
{
    __shared__ uint32_t s_hist[2][kRadix];
    for (int i = threadIdx.x; i < 2 * kRadix; i += kHistThreads) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t n2 = n / 2;
    const uint4* pairs2 = reinterpret_cast<const uint4*>(pairs);
    auto count = [&](uint32_t x) {
        atomicAdd(&s_hist[0][x & 0xFF], 1u);
        atomicAdd(&s_hist[1][(x >> 8) & 0xFF], 1u);
        atomicAdd(&bucket_counters[x & (kBucketKeys - 1)], 1u);
    };
    for (uint32_t i = blockIdx.x * kHistThreads + threadIdx.x; i < n2; i += gridDim.x * kHistThreads)
    {
        const uint4 a = ldg_stream_u4(pairs2 + i);
        count(a.x); count(a.z);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && (n & 1)) count(pairs[n - 1].x);
    __syncthreads();
    for (int t = threadIdx.x; t < 2 * kRadix; t += kHistThreads)
    {
        const uint32_t c = (&s_hist[0][0])[t];
        if (c != 0) atomicAdd(&(&ctl->hist[0][0])[t], c);
    }
}

// Large inputs: the two digit histograms only (shared atomics, no global ones).  The 65536 global atomics-per-key of the
// kernel above cost more than both scatter passes together at 2^26 pairs (0.58 of 1.13 ms); the END offsets are then
// read off the sorted output by bucket_end_offsets_search_kernel.  Two pairs per 128-bit load, two loads in flight.
__global__ void __launch_bounds__(kHistThreads)
bucket_digit_histogram_kernel(const uint2* __restrict__ pairs, uint32_t n, sort_control* ctl)
{
    __shared__ uint32_t s_hist[2][kRadix];
    for (int i = threadIdx.x; i < 2 * kRadix; i += kHistThreads) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t n2 = n / 2;
    const uint4* pairs2 = reinterpret_cast<const uint4*>(pairs);
    auto count = [&](uint32_t x) {
        atomicAdd(&s_hist[0][x & 0xFF], 1u);
        atomicAdd(&s_hist[1][(x >> 8) & 0xFF], 1u);
    };
    const uint32_t stride = gridDim.x * kHistThreads;
    uint32_t i = blockIdx.x * kHistThreads + threadIdx.x;
    for (; i + stride < n2; i += 2 * stride)
    {
        const uint4 a = ldg_stream_u4(pairs2 + i);
        const uint4 b = ldg_stream_u4(pairs2 + i + stride);
        count(a.x); count(a.z); count(b.x); count(b.z);
    }
    if (i < n2)
    {
        const uint4 a = ldg_stream_u4(pairs2 + i);
        count(a.x); count(a.z);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && (n & 1)) count(pairs[n - 1].x);
    __syncthreads();
    for (int t = threadIdx.x; t < 2 * kRadix; t += kHistThreads)
    {
        const uint32_t c = (&s_hist[0][0])[t];
        if (c != 0) atomicAdd(&(&ctl->hist[0][0])[t], c);
    }
}

// END offset of bucket b = number of pairs whose 16-bit key is <= b = upper bound of b in the SORTED output
// (bucket_sort_write.comp:32 leaves exactly that behind; an empty bucket repeats its predecessor's END).  One thread per
// bucket, log2(n) dependent reads each; the top levels of the search are shared by all threads and stay in L2.
__global__ void __launch_bounds__(256)
bucket_end_offsets_search_kernel(const uint2* __restrict__ sorted, uint32_t n, uint32_t* __restrict__ counters)
{
    const uint32_t b = blockIdx.x * 256u + threadIdx.x;
    uint32_t lo = 0, hi = n;    // first index whose key is > b lies in [lo, hi]
    while (lo < hi)
    {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        const uint32_t key = __ldg(&sorted[mid].x) & (kBucketKeys - 1);
        if (key <= b)
            lo = mid + 1;
        else
            hi = mid;
    }
    counters[b] = lo;
}

// counts -> bucket END offsets (inclusive prefix): what bucket_sort_write.comp:32 leaves behind.  64 CTAs x 1024
// counters: every CTA first sums the raw counts of the buckets before its slice (coalesced 128-bit reads, at most 252 KB
// from L2), then scans its own 1024.  Out of place (raw counts live in the scratch), so CTAs never read what another one
// has already rewritten.  (One CTA with 64 strided counters per thread took 22 us of the 0.4 ms light-assignment chain.)
constexpr int kEndOffsetCtas = kBucketKeys / 1024;

__global__ void __launch_bounds__(1024)
bucket_end_offsets_kernel(const uint32_t* __restrict__ raw, uint32_t* __restrict__ counters)
{
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_base;
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // sum of all the buckets before this CTA's slice
    uint32_t before = 0;
    const uint4* raw4 = reinterpret_cast<const uint4*>(raw);
    for (uint32_t i = tid; i < blockIdx.x * 256u; i += 1024)
    {
        const uint4 v = raw4[i];
        before += v.x + v.y + v.z + v.w;
    }
    before = __reduce_add_sync(kFullMask, before);
    if (lane == 0) s_warp[warp] = before;
    __syncthreads();
    if (warp == 0)
    {
        const uint32_t t = __reduce_add_sync(kFullMask, s_warp[lane]);
        if (lane == 0) s_base = t;
    }
    __syncthreads();
    const uint32_t base = s_base;
    const uint32_t mine = raw[blockIdx.x * 1024u + tid];
    uint32_t inc = mine;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1)
    {
        const uint32_t t = __shfl_up_sync(kFullMask, inc, s);
        if (lane >= (unsigned) s) inc += t;
    }
    __syncthreads();   // s_warp is free again
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t wp = 0;
#pragma unroll
    for (int w = 0; w < 32; w++)
        if (w < (int) warp) wp += s_warp[w];
    counters[blockIdx.x * 1024u + tid] = base + wp + inc;
}

// ---- 2. exclusive scan of each 256-bin histogram (grid = passes, block = 256) ----------------------------
__global__ void __launch_bounds__(kRadix)
radix_scan_histograms_kernel(sort_control* ctl)
{
    __shared__ uint32_t s_warp[kRadix / 32];
    uint32_t* h = ctl->hist[blockIdx.x];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t c = h[threadIdx.x];
    uint32_t inc = c;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1)
    {
        const uint32_t t = __shfl_up_sync(kFullMask, inc, s);
        if (lane >= (unsigned) s) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t wp = 0;
#pragma unroll
    for (int w = 0; w < kRadix / 32; w++)
        if (w < (int) warp) wp += s_warp[w];
    h[threadIdx.x] = wp + inc - c;
}

// ---- 3. the fused onesweep pass ------------------------------------------------------------------------------
// Option bits of the kernel template (F):
enum : uint32_t
{
    F_LB_INTERLEAVED = 1u << 1,   // the look-back advances in non-blocking steps between ranking rows
    F_LB_STEP8 = 1u << 2,         // ... every 8 rows instead of every 4
    F_EARLY_TMA = 1u << 3,        // the staging copies are issued before the counters are cleared
    F_PREFETCH_L2 = 1u << 4,      // a CTA asks L2 for the tile of the CTA that will take its place on the SM
    F_RANK_LEADER = 1u << 5,      // ranking by ballot match + ONE returning shared atomic by the highest lane of every match
                                  // group, read by the group through a shuffle: order by construction
    F_RANK_ATOMIC = 1u << 6,      // ranking by one returning shared atomic per lane, no match: right only if the lanes of one
                                  // instruction that hit the same address are served in ascending lane order (unspecified in PTX)
    F_VERIFY_ALL = 1u << 7,       // F_RANK_ATOMIC: every row's ranks are checked against the ballot match, off the critical path;
    F_VERIFY_SAMPLED = 1u << 8,   // ... or one row in eight.  A mismatch raises ctl->order_violation: the pass is then repeated
                                  // by the F_REDO form of the kernel (order by construction), see radix_sort_impl
    F_REG_COUNTS = 1u << 9,       // the per-warp digit counts stay in registers between the publish and the offset step
    F_KEYS_CHUNKED = 1u << 10,    // the key staging copy is split in four, every warp waits only for its own quarter
    F_MATCH_SPLIT4 = 1u << 11,    // ballot match with four accumulators (shorter dependent chains)
    F_SEGMENTED = 1u << 12,       // many independent tile-aligned segments (multi-GPU receive side), see radix_internal.cuh
    F_REDO = 1u << 13,            // repeat of a finished pass: persistent CTAs, no look-back (the predecessors' inclusive
                                  // prefixes are final in the look-back plane), ranking by match; returns at once unless
                                  // ctl->order_violation has the plane's bit
};

template <int THREADS, int ITEMS, int LAYOUT>
struct onesweep_smem
{
    static constexpr int WARPS = THREADS / 32;
    static constexpr int TILE = THREADS * ITEMS;
    static constexpr int KV_WORDS = LAYOUT == LAYOUT_KEYS ? TILE : 2 * TILE;
    // KEYS: keys[TILE] | SOA: keys[TILE] then values[TILE] | AOS: uint2[TILE]; afterwards the regroup area
    alignas(128) uint32_t kv[KV_WORDS];
    uint32_t warp_hist[WARPS][kRadix];
    uint32_t digit_base[kRadix];
    uint32_t scan_warp[kRadix / 32];
    uint32_t lane_dummy[WARPS][32];   // F_RANK_LEADER: where the lanes that do not lead a match group add
    alignas(8) uint64_t bar_keys;
    alignas(8) uint64_t bar_vals;
    alignas(8) uint64_t bar_chunk[4];   // F_KEYS_CHUNKED: one barrier per quarter of the staged keys
    uint32_t tile;                      // ticket mode: the tile id taken by thread 0
};

// lanes of the warp holding the same 8-bit digit: AND of the ballots of my set bits, minus OR of the ballots of my clear
// bits — two independent accumulators, each updated by ONE predicated LOP3 per bit (vote + 2 instructions per bit)
template <bool SPLIT4>
__device__ __forceinline__ unsigned match_digit(uint32_t d)
{
    if (SPLIT4)
    {
        // four accumulators (bits 0-3 and 4-7 apart): dependent chains of 4 instead of 8, one more instruction per key
        unsigned ones_lo = kFullMask, zeros_lo = 0u, ones_hi = kFullMask, zeros_hi = 0u;
#pragma unroll
        for (int b = 0; b < kRadixBits / 2; b++)
        {
            asm("{\n"
                ".reg .pred p, q;\n"
                ".reg .b32 t, u, bal, bal2;\n"
                "and.b32 t, %4, %5;\n"
                "setp.ne.u32 p, t, 0;\n"
                "and.b32 u, %4, %6;\n"
                "setp.ne.u32 q, u, 0;\n"
                "vote.sync.ballot.b32 bal, p, 0xffffffff;\n"
                "vote.sync.ballot.b32 bal2, q, 0xffffffff;\n"
                "@p and.b32 %0, %0, bal;\n"
                "@!p or.b32 %1, %1, bal;\n"
                "@q and.b32 %2, %2, bal2;\n"
                "@!q or.b32 %3, %3, bal2;\n"
                "}\n"
                : "+r"(ones_lo), "+r"(zeros_lo), "+r"(ones_hi), "+r"(zeros_hi) : "r"(d), "r"(1u << b), "r"(16u << b));
        }
        return (ones_lo & ones_hi) & ~(zeros_lo | zeros_hi);
    }
    unsigned ones = kFullMask, zeros = 0u;
#pragma unroll
    for (int b = 0; b < kRadixBits; b++)
    {
        asm("{\n"
            ".reg .pred p;\n"
            ".reg .b32 t, bal;\n"
            "and.b32 t, %2, %3;\n"
            "setp.ne.u32 p, t, 0;\n"
            "vote.sync.ballot.b32 bal, p, 0xffffffff;\n"
            "@p and.b32 %0, %0, bal;\n"
            "@!p or.b32 %1, %1, bal;\n"
            "}\n"
            : "+r"(ones), "+r"(zeros) : "r"(d), "r"(1u << b));
    }
    return ones & ~zeros;
}

__device__ __forceinline__ uint32_t digit_of(uint32_t key, uint32_t prmt_sel) { return __byte_perm(key, 0u, prmt_sel); }

// F_PREFETCH_L2: how many tiles ahead a CTA prefetches (74-222 tiles ahead is a plateau for the default tile,
// profiles/r1x_prefetch_distance.log)
constexpr uint32_t kPrefetchTiles = kNumSMs;

// Order of work inside a tile ("count first"):
//   1. the warp-private digit counters are filled (one non-returning shared atomic per key),
//   2. the tile's digit counts are published and the look-back loads of the first predecessors are issued,
//   3. a per-digit cross-warp scan turns the counters into the final in-tile offset of every (warp, digit) run,
//   4. ranking: every pair gets its final in-tile slot and is stored to the regroup buffer at once; the look-back advances
//      by one non-blocking step every few rows, so its L2 round trips overlap the ranking,
//   5. the look-back finishes and the tile is written out coalesced through the per-digit base.
// Tile ids: the block index (CTAs of a 1-D grid start in index order on every CUDA GPU so far — the assumption CUB's
// decoupled-look-back scan makes too — so every tile a CTA waits for has started) or, with pass_params::ticket, the value of an
// atomic counter taken at the start of the CTA (start order by construction: forward progress without the assumption).
template <int THREADS, int ITEMS, int LAYOUT, uint32_t F, int MIN_BLOCKS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
onesweep_pass_kernel(const __grid_constant__ pass_params p)
{
    using smem_t = onesweep_smem<THREADS, ITEMS, LAYOUT>;
    constexpr int WARPS = smem_t::WARPS;
    constexpr int TILE = smem_t::TILE;
    constexpr bool HAS_VALUES = LAYOUT != LAYOUT_KEYS;
    constexpr int KSTRIDE = LAYOUT == LAYOUT_AOS ? 2 : 1;       // words between consecutive staged keys
    constexpr int VOFF = LAYOUT == LAYOUT_AOS ? 1 : TILE;       // word offset key -> its value
    constexpr uint32_t ELEM_BYTES = LAYOUT == LAYOUT_AOS ? 8 : 4;
    constexpr bool SEG = (F & F_SEGMENTED) != 0;
    constexpr bool REDO = (F & F_REDO) != 0;
    constexpr bool EARLY = (F & F_EARLY_TMA) != 0;
    constexpr bool CHUNKED = (F & F_KEYS_CHUNKED) != 0 && WARPS % 4 == 0;
    constexpr bool KEEP_COUNTS = (F & F_REG_COUNTS) != 0;
    constexpr bool INTERLEAVED = (F & F_LB_INTERLEAVED) != 0 && !REDO;
    constexpr bool ATOMIC = (F & F_RANK_ATOMIC) != 0 && !REDO;
    constexpr bool LEADER = !ATOMIC;
    constexpr bool VERIFY_ALL = ATOMIC && (F & F_VERIFY_ALL) != 0;
    constexpr bool VERIFY_SAMPLED = ATOMIC && (F & F_VERIFY_SAMPLED) != 0;
    constexpr bool SPLIT4 = (F & F_MATCH_SPLIT4) != 0;
    static_assert(THREADS >= kRadix, "one thread per digit needed");
    static_assert(LEADER || (F & F_RANK_ATOMIC), "a ranking mode is needed");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    smem_t& sm = *reinterpret_cast<smem_t*>(smem_raw);

    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // ---- what this launch covers (uniform over the grid) ----
    int pass = p.pass;
    if (p.dyn_pass != nullptr) pass = (int) *p.dyn_pass;
    uint32_t tile_begin = 0, tile_end = p.num_tiles;
    uint32_t* keys_out = p.keys_out;
    uint32_t* vals_out = p.vals_out;
    const bool ticket = p.ticket != nullptr && !REDO;
    bool final_pass = false;
    if (SEG)
    {
        const uint32_t pstar = p.plan->pstar;
        if (p.plan->error != 0 || (uint32_t) pass >= pstar) return;
        final_pass = (uint32_t) pass + 1 == pstar;
        if (final_pass)
        {
            keys_out = p.final_keys;
            vals_out = p.final_vals;
        }
        tile_begin = p.plan->round_tile[p.round];
        tile_end = p.plan->round_tile[p.round + 1];
    }
    if (REDO && ((ld_relaxed_u32(&p.ctl->order_violation) >> p.lb_plane) & 1u) == 0) return;
    const uint32_t prmt_sel = 0x4440u | (uint32_t) pass;   // byte `pass` of the key -> one PRMT per digit extraction
    const uint32_t n = p.n;

    if (tid == 0)
    {
        mbar_init(&sm.bar_keys, 1);
        mbar_init(&sm.bar_vals, 1);
        if (CHUNKED)
        {
#pragma unroll
            for (int c = 0; c < 4; c++) mbar_init(&sm.bar_chunk[c], 1);
        }
        mbar_fence_init();
    }
    // staging copies of a whole tile (one thread): keys (whole or in quarters), then the values
    auto issue_copies = [&](uint64_t base) {
        if (CHUNKED)
        {
            constexpr uint32_t Q = TILE / 4;
#pragma unroll
            for (int c = 0; c < 4; c++)
            {
                mbar_arrive_expect_tx(&sm.bar_chunk[c], Q * ELEM_BYTES);
                bulk_copy_g2s(sm.kv + c * Q * KSTRIDE, p.keys_in + (base + c * Q) * KSTRIDE, Q * ELEM_BYTES, &sm.bar_chunk[c]);
            }
        }
        else
        {
            mbar_arrive_expect_tx(&sm.bar_keys, TILE * ELEM_BYTES);
            bulk_copy_g2s(sm.kv, p.keys_in + base * KSTRIDE, TILE * ELEM_BYTES, &sm.bar_keys);
        }
        if (LAYOUT == LAYOUT_SOA)
        {
            mbar_arrive_expect_tx(&sm.bar_vals, TILE * 4);
            bulk_copy_g2s(sm.kv + TILE, p.vals_in + base, TILE * 4, &sm.bar_vals);
        }
    };
    // a segmented launch always stages whole tiles (the padded layout keeps them inside the buffer)
    auto whole_tile = [&](uint32_t t) { return SEG || (uint64_t) t * TILE + TILE <= (uint64_t) n; };

    for (uint32_t it = 0;; it++)   // one iteration per CTA, except F_REDO (persistent CTAs striding over the tiles)
    {
        const uint32_t par = it & 1u;
        uint32_t tile = tile_begin + blockIdx.x + (REDO ? it * gridDim.x : 0u);
        if (!ticket && tile >= tile_end) return;
        if (tid == 0)
        {
            if (ticket)
            {
                tile = tile_begin + atomicAdd(p.ticket, 1u);
                sm.tile = tile;
            }
            // the copies start before the counters are cleared (nobody else touches the staging area yet)
            if (EARLY && tile < tile_end && whole_tile(tile)) issue_copies((uint64_t) tile * TILE);
        }
        // global offset of this thread's digit: needed at the very end, fetched now
        uint32_t pass_base = 0;
        if (!SEG && tid < kRadix) pass_base = p.ctl->hist[pass][tid];
#pragma unroll
        for (int i = lane; i < kRadix; i += 32) sm.warp_hist[warp][i] = 0;
        __syncthreads();
        if (ticket)
        {
            tile = sm.tile;
            if (tile >= tile_end) return;
        }
        // this tile's row of the next look-back plane is cleared here (the host clears only the first plane of a sort)
        if (!REDO && p.clear_next_plane && tid < kRadix)
            p.lookback[((size_t) (p.lb_plane + 1) * p.num_tiles + tile) * kRadix + tid] = 0u;

        const uint64_t tile_base = (uint64_t) tile * TILE;
        const bool whole = whole_tile(tile);
        uint32_t valid, local_tile = tile;   // local_tile: index of the tile inside its segment (look-back restarts there)
        uint32_t seg = 0, seg_out = 0;
        if (SEG)
        {
            seg = p.tile_seg[tile];
            const seg_desc sd = p.plan->seg[seg];
            local_tile = tile - sd.first_tile;
            const uint32_t left = sd.len - local_tile * (uint32_t) TILE;
            valid = left < (uint32_t) TILE ? left : (uint32_t) TILE;
            seg_out = final_pass ? sd.out_start : sd.first_tile * (uint32_t) TILE;
        }
        else
            valid = (n - tile_base) < (uint64_t) TILE ? (uint32_t) (n - tile_base) : (uint32_t) TILE;
        const bool full = valid == (uint32_t) TILE;
        if ((F & F_PREFETCH_L2) && !REDO && tid == 32)
        {
            // the CTA that takes this one's place on the SM is about (resident CTAs) tiles ahead: have its input
            // waiting in L2 by the time it starts
            const uint64_t next_base = tile_base + (uint64_t) kPrefetchTiles * TILE;
            if (next_base + TILE <= (uint64_t) n)
            {
                bulk_prefetch_l2(p.keys_in + next_base * KSTRIDE, TILE * ELEM_BYTES);
                if (LAYOUT == LAYOUT_SOA) bulk_prefetch_l2(p.vals_in + next_base, TILE * 4);
            }
        }

        const uint32_t warp_off = warp * (ITEMS * 32) + lane;
        uint32_t* my_hist = sm.warp_hist[warp];
        uint32_t key[ITEMS];
        uint32_t val[HAS_VALUES ? ITEMS : 1];
        if (whole)
        {
            if (!EARLY && tid == 0) issue_copies(tile_base);
            mbar_wait(CHUNKED ? &sm.bar_chunk[warp / (WARPS / 4)] : &sm.bar_keys, par);
        }
        else
        {
            // ragged last tile: guarded loads, padding keys 0xFFFFFFFF sort behind every real key of the tile
            for (uint32_t i = tid; i < (uint32_t) TILE; i += THREADS)
            {
                const bool in = i < valid;
                sm.kv[i * KSTRIDE] = in ? p.keys_in[(tile_base + i) * KSTRIDE] : 0xFFFFFFFFu;
                if (LAYOUT == LAYOUT_SOA) sm.kv[TILE + i] = in ? p.vals_in[tile_base + i] : 0u;
                if (LAYOUT == LAYOUT_AOS) sm.kv[i * 2 + 1] = in ? p.keys_in[(tile_base + i) * 2 + 1] : 0u;
            }
            __syncthreads();
        }
#pragma unroll
        for (int j = 0; j < ITEMS; j++) key[j] = sm.kv[(warp_off + j * 32) * KSTRIDE];
        if (SEG && !full)
        {
            // last tile of a segment: what lies behind the segment's end in the staged tile is not part of it
#pragma unroll
            for (int j = 0; j < ITEMS; j++)
                if (warp_off + j * 32 >= valid) key[j] = 0xFFFFFFFFu;
        }
        // 1. count
#pragma unroll
        for (int j = 0; j < ITEMS; j++) atomicAdd(&my_hist[digit_of(key[j], prmt_sel)], 1u);
        __syncthreads();

        // 2. publish the aggregate, start the look-back
        constexpr int K = 4;
        uint32_t* lb = p.lookback + ((size_t) p.lb_plane * p.num_tiles + tile) * kRadix;
        uint32_t cnt = 0, inc = 0, real_cnt = 0, lb_pre[K];
        uint32_t cw[KEEP_COUNTS ? WARPS : 1];   // F_REG_COUNTS: this thread's digit, count per warp
        // look-back state of this thread's digit: window of K predecessors starting at tile lb_t, words in lb_pre[]
        const uint32_t* lb_p = lb - kRadix + tid;
        int32_t lb_t = (int32_t) local_tile - 1;
        bool lb_done = local_tile == 0 || tid >= kRadix || REDO;
        uint32_t exclusive = 0;
        auto lb_load = [&]() {
#pragma unroll
            for (int k = 0; k < K; k++) lb_pre[k] = (lb_t - k >= 0) ? ld_relaxed_u32(lb_p - k * kRadix) : kLbFlagInclusive;
        };
        // non-blocking step: consume the window if all of it has been published (and open the next one), else re-poll the
        // missing words; the loads complete while the ranking goes on
        auto lb_try = [&]() {
            if (lb_done) return;
            bool ready = true;
#pragma unroll
            for (int k = 0; k < K; k++) ready = ready && (lb_pre[k] >> 30) != 0;
            if (ready)
            {
#pragma unroll
                for (int k = 0; k < K; k++)
                    if (!lb_done)
                    {
                        exclusive += lb_pre[k] & kLbValueMask;
                        lb_done = (lb_pre[k] >> 30) == 2;
                    }
                if (!lb_done)
                {
                    lb_t -= K;
                    lb_p -= K * kRadix;
                    lb_load();
                }
            }
            else
            {
#pragma unroll
                for (int k = 0; k < K; k++)
                    if ((lb_pre[k] >> 30) == 0) lb_pre[k] = ld_relaxed_u32(lb_p - k * kRadix);
            }
        };
        if (tid < kRadix)
        {
#pragma unroll
            for (int w = 0; w < WARPS; w++)
            {
                const uint32_t c = sm.warp_hist[w][tid];
                if (KEEP_COUNTS) cw[w] = c;
                cnt += c;
            }
            real_cnt = cnt - ((tid == kRadix - 1) ? (uint32_t) TILE - valid : 0u);
            if (!REDO)
            {
                st_relaxed_u32(&lb[tid], (local_tile == 0 ? kLbFlagInclusive : kLbFlagAggregate) | real_cnt);
                lb_load();
            }
            inc = cnt;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1)
            {
                const uint32_t t = __shfl_up_sync(kFullMask, inc, s);
                if (lane >= (unsigned) s) inc += t;
            }
            if (lane == 31) sm.scan_warp[warp] = inc;
        }
        __syncthreads();
        // 3. counters -> in-tile offset of every (warp, digit) run
        uint32_t tile_off = 0;
        if (tid < kRadix)
        {
            uint32_t wp = 0;
#pragma unroll
            for (int w = 0; w < kRadix / 32; w++)
                if (w < (int) warp) wp += sm.scan_warp[w];
            tile_off = wp + inc - cnt;
            uint32_t run = tile_off;
#pragma unroll
            for (int w = 0; w < WARPS; w++)
            {
                const uint32_t c = KEEP_COUNTS ? cw[w] : sm.warp_hist[w][tid];
                sm.warp_hist[w][tid] = run;
                run += c;
            }
        }
        if (HAS_VALUES)
        {
            if (LAYOUT == LAYOUT_SOA && whole) mbar_wait(&sm.bar_vals, par);
#pragma unroll
            for (int j = 0; j < ITEMS; j++) val[j] = sm.kv[(warp_off + j * 32) * KSTRIDE + VOFF];
        }
        __syncthreads(); // offsets visible; every warp has consumed the staged inputs: the buffers become the regroup area

        // 4. rank and regroup
        const unsigned lt = lanemask_lt();
        // F_RANK_LEADER: the highest lane of every match group adds the group's size to the run counter with ONE returning
        // shared atomic (predicated, no branch) and the group reads the old value from it by shuffle; the atomic of row j+1
        // is issued before row j's result is consumed, so its latency hides behind a ranking.  Rows stay ordered: same
        // warp, program order, __syncwarp() between the rows.
        const uint32_t hist_addr = smem_u32(my_hist);
        // (lanes that are not the leader add to a private dummy word instead: ptxas turns a predicated atom into a
        // divergent branch, which costs more than the ~2 extra lanes per row)
        const uint32_t dummy_addr = smem_u32(&sm.lane_dummy[warp][lane]);
        auto leader_atomic = [&](uint32_t d, unsigned mask) {
            uint32_t old;
            const bool leader = lane == 31u - (uint32_t) __clz(mask);
            asm volatile("atom.shared.add.u32 %0, [%1], %2;"
                         : "=r"(old) : "r"(leader ? hist_addr + d * 4u : dummy_addr), "r"((uint32_t) __popc(mask)) : "memory");
            return old;
        };
        uint32_t d_nxt = 0, old_nxt = 0, bad = 0;
        unsigned mask_nxt = 0;
        if (LEADER)
        {
            d_nxt = digit_of(key[0], prmt_sel);
            mask_nxt = match_digit<SPLIT4>(d_nxt);
            old_nxt = leader_atomic(d_nxt, mask_nxt);
        }
#pragma unroll
        for (int j = 0; j < ITEMS; j++)
        {
            uint32_t r;
            if (ATOMIC)
            {
                // no match on the way to the slot: every lane takes it with a returning shared atomic.  Stable iff the lanes
                // of ONE instruction that hit the same address are served in ascending lane order; the check below (all
                // rows or one in eight) compares with what the ballot match says and feeds nothing but the violation flag,
                // so it runs in the issue slots the shared-memory pipe leaves idle.
                const uint32_t d = digit_of(key[j], prmt_sel);
                r = atomicAdd(&my_hist[d], 1u);
                __syncwarp();   // row j's adds are performed before row j+1's (different lanes may hit the same counter)
                if (VERIFY_ALL || (VERIFY_SAMPLED && (j & 7) == 5))
                {
                    // ascending lane order <=> (slot - number of lower lanes of my group) is the same for the whole group
                    const unsigned mask = match_digit<SPLIT4>(d);
                    const uint32_t first_slot = r - (uint32_t) __popc(mask & lt);
                    bad |= first_slot ^ __shfl_sync(kFullMask, first_slot, __ffs(mask) - 1);
                }
            }
            else
            {
                const unsigned mask = mask_nxt;
                const uint32_t old = old_nxt;
                if (j + 1 < ITEMS)
                {
                    d_nxt = digit_of(key[j + 1], prmt_sel);
                    mask_nxt = match_digit<SPLIT4>(d_nxt);
                    __syncwarp();
                    old_nxt = leader_atomic(d_nxt, mask_nxt);
                }
                const uint32_t prior = __shfl_sync(kFullMask, old, 31 - __clz(mask));
                r = prior + __popc(mask & lt);
            }
            if (HAS_VALUES)
                reinterpret_cast<uint2*>(sm.kv)[r] = make_uint2(key[j], val[j]);
            else
                sm.kv[r] = key[j];
            constexpr int LB_EVERY = (F & F_LB_STEP8) ? 8 : 4;
            if (INTERLEAVED && (j % LB_EVERY) == LB_EVERY - 1 && j + 1 < ITEMS) lb_try();
        }
        if (VERIFY_ALL || VERIFY_SAMPLED)
        {
            if (p.selftest) bad = 1;   // tests: pretend the check failed (and write nothing below): the redo must rebuild the pass
            if (__any_sync(kFullMask, bad != 0) && lane == 0) atomicOr(&p.ctl->order_violation, 1u << p.lb_plane);
        }

        // 5. finish the look-back
        if (tid < kRadix)
        {
            if (REDO)
            {
                // the pass has run: the predecessor's word holds its final inclusive prefix
                if (local_tile > 0) exclusive = ld_relaxed_u32(lb - kRadix + tid) & kLbValueMask;
            }
            else if (local_tile > 0)
            {
                while (!lb_done)
                {
#pragma unroll
                    for (int k = 0; k < K; k++)
                        if (!lb_done)
                        {
                            while ((lb_pre[k] >> 30) == 0) lb_pre[k] = ld_relaxed_u32(lb_p - k * kRadix);
                            exclusive += lb_pre[k] & kLbValueMask;
                            lb_done = (lb_pre[k] >> 30) == 2;
                        }
                    if (!lb_done)
                    {
                        lb_t -= K;
                        lb_p -= K * kRadix;
                        lb_load();
                    }
                }
                st_relaxed_u32(&lb[tid], kLbFlagInclusive | (exclusive + real_cnt));
            }
            if (SEG) pass_base = seg_out + p.seg_hist[((size_t) seg * (kPasses - 1) + pass) * kRadix + tid];
            sm.digit_base[tid] = pass_base + exclusive - tile_off;
        }
        __syncthreads();

        if ((VERIFY_ALL || VERIFY_SAMPLED) && p.selftest)
        {
            // nothing is written: see above
        }
        else if (full)
        {
#pragma unroll
            for (int j = 0; j < ITEMS; j++)
            {
                const uint32_t q = j * THREADS + tid;
                if (HAS_VALUES)
                {
                    const uint2 e = reinterpret_cast<const uint2*>(sm.kv)[q];
                    const uint32_t g = sm.digit_base[digit_of(e.x, prmt_sel)] + q;
                    if (LAYOUT == LAYOUT_AOS)
                        reinterpret_cast<uint2*>(keys_out)[g] = e;
                    else
                    {
                        keys_out[g] = e.x;
                        vals_out[g] = e.y;
                    }
                }
                else
                {
                    const uint32_t k = sm.kv[q];
                    keys_out[sm.digit_base[digit_of(k, prmt_sel)] + q] = k;
                }
            }
        }
        else
        {
            for (uint32_t q = tid; q < valid; q += THREADS)
            {
                const uint32_t k = sm.kv[q * (HAS_VALUES ? 2 : 1)];
                const uint32_t g = sm.digit_base[digit_of(k, prmt_sel)] + q;
                if (LAYOUT == LAYOUT_AOS)
                    reinterpret_cast<uint2*>(keys_out)[g] = make_uint2(k, sm.kv[q * 2 + 1]);
                else
                {
                    keys_out[g] = k;
                    if (HAS_VALUES) vals_out[g] = sm.kv[q * 2 + 1];
                }
            }
        }
        if (!REDO) return;
        __syncthreads();   // the staging area and the counters are reused by the next tile of this CTA
    }
}

// ---- variants ----------------------------------------------------------------------------------------------------------------
template <int THREADS, int ITEMS, int LAYOUT, uint32_t F, int MIN_BLOCKS>
int launch_kernel(cudaStream_t s, const pass_params& p, uint32_t grid)
{
    auto kern = onesweep_pass_kernel<THREADS, ITEMS, LAYOUT, F, MIN_BLOCKS>;
    constexpr size_t smem = sizeof(onesweep_smem<THREADS, ITEMS, LAYOUT>);
    // per launch: function attributes belong to the current device's context, a process may use several
    VRENB200_TRY(check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)));
    if (grid == 0) return VRENB200_OK;   // preload only
    kern<<<grid, THREADS, smem, s>>>(p);
    return check_launch();
}

// grid == 0: only load the kernel into the current context (CUDA loads modules lazily, and loading may wait for running
// kernels: a multi-GPU sort whose kernels wait for each other must not meet its first load in the middle of a call)
template <int THREADS, int ITEMS, uint32_t F, int MIN_BLOCKS>
int launch_pass(cudaStream_t s, const pass_params& p, uint32_t grid, int layout, bool segmented)
{
    if (segmented)   // the multi-GPU receive side sorts key / value arrays
        return layout == LAYOUT_SOA ? launch_kernel<THREADS, ITEMS, LAYOUT_SOA, F | F_SEGMENTED, MIN_BLOCKS>(s, p, grid) : VRENB200_EINVAL_ARG;
    switch (layout)
    {
    case LAYOUT_KEYS: return launch_kernel<THREADS, ITEMS, LAYOUT_KEYS, F, MIN_BLOCKS>(s, p, grid);
    case LAYOUT_SOA:  return launch_kernel<THREADS, ITEMS, LAYOUT_SOA, F, MIN_BLOCKS>(s, p, grid);
    default:          return launch_kernel<THREADS, ITEMS, LAYOUT_AOS, F, MIN_BLOCKS>(s, p, grid);
    }
}

// the by-construction repeat of a pass whose ranking check failed: same tile, ranking by match, persistent CTAs, one per SM
template <int THREADS, int ITEMS, uint32_t F>
int launch_redo(cudaStream_t s, const pass_params& p, int layout, bool segmented)
{
    constexpr uint32_t R = (F & (F_EARLY_TMA | F_REG_COUNTS | F_KEYS_CHUNKED)) | F_RANK_LEADER | F_REDO;
    return launch_pass<THREADS, ITEMS, R, 1>(s, p, (uint32_t) kNumSMs, layout, segmented);
}

template <int THREADS, int ITEMS, uint32_t F, int MIN_BLOCKS, bool VERIFIED>
int preload_variant(int layout, bool segmented)
{
    const pass_params none{};
    VRENB200_TRY((launch_pass<THREADS, ITEMS, F, MIN_BLOCKS>(nullptr, none, 0u, layout, segmented)));
    if (VERIFIED)
    {
        constexpr uint32_t R = (F & (F_EARLY_TMA | F_REG_COUNTS | F_KEYS_CHUNKED)) | F_RANK_LEADER | F_REDO;
        VRENB200_TRY((launch_pass<THREADS, ITEMS, R, 1>(nullptr, none, 0u, layout, segmented)));
    }
    return VRENB200_OK;
}

constexpr uint32_t kBase = F_LB_INTERLEAVED | F_EARLY_TMA;
constexpr uint32_t kBig = kBase | F_PREFETCH_L2;
#define VARIANT(T, I, F, B) { #T "x" #I "/" #F "/occ" #B, (T) * (I), launch_pass<T, I, F, B>, nullptr, preload_variant<T, I, F, B, false> }
#define VVARIANT(T, I, F, B) { #T "x" #I "/" #F "/occ" #B, (T) * (I), launch_pass<T, I, F, B>, launch_redo<T, I, F>, preload_variant<T, I, F, B, true> }
const sort_variant g_variants[] = {
    // ballot-match ranking (order by construction)
    VARIANT(256, 16, kBase | F_RANK_LEADER, 3),                                                   // 0: below 2^21 elements
    VARIANT(256, 64, kBig | F_RANK_LEADER, 2),                                                    // 1: keys only
    VARIANT(256, 46, kBig | F_RANK_LEADER, 2),                                                    // 2: pairs
    // atomic ranking, every row checked against the match, by-construction redo of a pass that fails the check
    VVARIANT(256, 16, kBase | F_RANK_ATOMIC | F_VERIFY_ALL, 3),                                   // 3
    VVARIANT(256, 64, kBig | F_LB_STEP8 | F_RANK_ATOMIC | F_VERIFY_ALL | F_REG_COUNTS | F_KEYS_CHUNKED, 2),   // 4
    VVARIANT(256, 48, kBig | F_LB_STEP8 | F_RANK_ATOMIC | F_VERIFY_ALL | F_REG_COUNTS | F_KEYS_CHUNKED, 2),   // 5
    // ... one row in eight checked
    VVARIANT(256, 16, kBase | F_RANK_ATOMIC | F_VERIFY_SAMPLED, 3),                               // 6
    VVARIANT(256, 64, kBig | F_LB_STEP8 | F_RANK_ATOMIC | F_VERIFY_SAMPLED | F_REG_COUNTS | F_KEYS_CHUNKED, 2),   // 7
    VVARIANT(256, 48, kBig | F_LB_STEP8 | F_RANK_ATOMIC | F_VERIFY_SAMPLED | F_REG_COUNTS | F_KEYS_CHUNKED, 2),   // 8
    // ... unchecked (explicit opt-in: relies on the lane order of same-address shared atomics)
    VARIANT(256, 16, kBase | F_RANK_ATOMIC, 3),                                                   // 9
    VARIANT(256, 64, kBig | F_LB_STEP8 | F_RANK_ATOMIC | F_REG_COUNTS | F_KEYS_CHUNKED, 2),       // 10
    VARIANT(256, 48, kBig | F_LB_STEP8 | F_RANK_ATOMIC | F_REG_COUNTS | F_KEYS_CHUNKED, 2),       // 11
#ifdef VRENB200_TUNING
    VVARIANT(256, 46, kBig | F_LB_STEP8 | F_RANK_ATOMIC | F_VERIFY_ALL | F_REG_COUNTS | F_KEYS_CHUNKED, 2),   // 12
    VVARIANT(256, 44, kBig | F_LB_STEP8 | F_RANK_ATOMIC | F_VERIFY_ALL | F_REG_COUNTS | F_KEYS_CHUNKED, 2),   // 13
    VVARIANT(256, 48, kBig | F_LB_STEP8 | F_RANK_ATOMIC | F_VERIFY_ALL | F_MATCH_SPLIT4 | F_REG_COUNTS | F_KEYS_CHUNKED, 2),   // 14
    VARIANT(256, 46, kBig | F_RANK_LEADER | F_MATCH_SPLIT4, 2),                                   // 15
    VARIANT(256, 44, kBig | F_RANK_LEADER, 2),                                                    // 16
    VARIANT(384, 24, kBig | F_RANK_LEADER, 2),                                                    // 17
#endif
};
constexpr int kNumVariants = sizeof(g_variants) / sizeof(g_variants[0]);
// the scratch layout must not depend on the variant: size the look-back for the smallest tile
constexpr uint32_t kMinTile = 256 * 16;
// inputs below 2^21 elements take the 4096-element tile (more CTAs than SMs from 2^19 on, shorter per-tile steps: 44 vs 63 us
// at 2^16 pairs, 64.5 vs 71.6 us at 2^20, equal at 2^21-2^22, profiles/r1x_sort_size_variants.log); keys-only sorts stage
// 4 B per key, so 64 rows per thread fit the same shared memory and registers (16 384-key tiles)
constexpr uint32_t kSmallTileBelow = 1u << 21;

} // namespace

sort_options resolve_options(const vrenb200_sort_config* cfg)
{
    // the defaults can be overridden through the environment (read once, never written again)
    static const sort_options env = []() {
        sort_options o{VRENB200_RANKING_AUTO, VRENB200_TILE_IDS_AUTO, 0, false};
        if (const char* e = std::getenv("VRENB200_SORT_RANKING"))
            o.ranking = !std::strcmp(e, "match") ? VRENB200_RANKING_MATCH
                      : !std::strcmp(e, "verified") ? VRENB200_RANKING_ATOMIC_VERIFIED
                      : !std::strcmp(e, "sampled") ? VRENB200_RANKING_ATOMIC_SAMPLED
                      : !std::strcmp(e, "atomic") ? VRENB200_RANKING_ATOMIC_UNVERIFIED : VRENB200_RANKING_AUTO;
        if (const char* e = std::getenv("VRENB200_SORT_TILE_IDS"))
            o.tile_ids = !std::strcmp(e, "ticket") ? VRENB200_TILE_IDS_TICKET : !std::strcmp(e, "block") ? VRENB200_TILE_IDS_BLOCK_INDEX : VRENB200_TILE_IDS_AUTO;
        return o;
    }();
    sort_options o = env;
    if (cfg != nullptr)
    {
        if (cfg->ranking != VRENB200_RANKING_AUTO) o.ranking = cfg->ranking;
        if (cfg->tile_ids != VRENB200_TILE_IDS_AUTO) o.tile_ids = cfg->tile_ids;
        o.variant = cfg->variant;
    }
    o.ranking_auto = o.ranking == VRENB200_RANKING_AUTO;
    if (o.ranking == VRENB200_RANKING_AUTO) o.ranking = VRENB200_RANKING_DEFAULT;
    if (o.tile_ids == VRENB200_TILE_IDS_AUTO) o.tile_ids = VRENB200_TILE_IDS_DEFAULT;
    return o;
}

const sort_variant& pick_variant(uint32_t n, int layout, const sort_options& opt)
{
    if (opt.variant > 0 && opt.variant <= kNumVariants) return g_variants[opt.variant - 1];
    int group;
    switch (opt.ranking)
    {
    case VRENB200_RANKING_MATCH: group = 0; break;
    case VRENB200_RANKING_ATOMIC_UNVERIFIED: group = 3; break;
    case VRENB200_RANKING_ATOMIC_VERIFIED: group = 1; break;
    default: group = 2; break;   // ATOMIC_SAMPLED (the default), SELFTEST
    }
    const int size = n < kSmallTileBelow ? 0 : (layout == LAYOUT_KEYS ? 1 : 2);
    // Small inputs are launch-bound (profiles/r1z_size_sweep.log: 41-67 us from 2^10 to 2^20 elements, the kernels themselves a
    // few microseconds each), so when the ranking is left to the library they take the ballot-match kernel: ordered by
    // construction, hence no repeat kernel behind every pass (4 launches fewer per radix sort, 2 per bucket sort)
    if (size == 0 && opt.ranking_auto) group = 0;
    return g_variants[group * 3 + size];
}

int preload_sort_kernels(int layout, bool segmented, const sort_options& opt)
{
    // the small-tile and the large-tile entry the options select (a sort picks between them by its size)
    for (uint32_t n : {1u, kSmallTileBelow})
        VRENB200_TRY(pick_variant(n, layout, opt).preload(layout, segmented));
    cudaFuncAttributes attr;
    VRENB200_TRY(check_cuda(cudaFuncGetAttributes(&attr, radix_histogram_kernel)));
    VRENB200_TRY(check_cuda(cudaFuncSetAttribute(radix_histogram_columns_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kHist2Smem)));
    return check_cuda(cudaFuncGetAttributes(&attr, radix_scan_histograms_kernel));
}

size_t lookback_words(uint32_t n)
{
    const size_t tiles = ((size_t) n + kMinTile - 1) / kMinTile;
    return (size_t) kPasses * (tiles > 0 ? tiles : 1) * kRadix;
}

size_t control_bytes(uint32_t n)
{
    return align_up(sizeof(sort_control) + lookback_words(n) * sizeof(uint32_t), 256);
}

int launch_digit_histograms(cudaStream_t s, const uint32_t* keys, uint32_t n, sort_control* ctl)
{
    if (n >= (1u << 20))
    {
        VRENB200_TRY(check_cuda(cudaFuncSetAttribute(radix_histogram_columns_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kHist2Smem)));
        radix_histogram_columns_kernel<<<kNumSMs, kHist2Threads, kHist2Smem, s>>>(keys, n, ctl);
    }
    else
        radix_histogram_kernel<<<kNumSMs * 4, kHistThreads, 0, s>>>(keys, n, ctl);
    return check_launch();
}

int launch_scan_histograms(cudaStream_t s, sort_control* ctl, int passes)
{
    radix_scan_histograms_kernel<<<passes, kRadix, 0, s>>>(ctl);
    return check_launch();
}

} // namespace vrenb200

// per-kernel device timing for bench.py's roofline: events recorded on the launching stream between launches
struct vrenb200_sort_profile
{
    cudaEvent_t ev[vrenb200::kPasses + 3]; // start | after hist | after scan | after pass 0..3
};

namespace vrenb200 {
namespace {

// control + look-back live in `ctl_mem`; alt buffers given explicitly
int radix_sort_impl(cudaStream_t s, uint32_t* keys, uint32_t* vals, uint32_t n, uint32_t* alt_keys, uint32_t* alt_vals,
                    void* ctl_mem, const vrenb200_sort_config* cfg, vrenb200_sort_profile* prof = nullptr, int first_pass = 0,
                    int num_passes = kPasses)
{
    if (n == 0) return VRENB200_OK;
    if (n >= (1u << 30)) return VRENB200_ELIMIT;
    if ((reinterpret_cast<uintptr_t>(keys) | reinterpret_cast<uintptr_t>(alt_keys) |
         reinterpret_cast<uintptr_t>(vals) | reinterpret_cast<uintptr_t>(alt_vals) |
         reinterpret_cast<uintptr_t>(ctl_mem)) & 15)
        return VRENB200_EALIGN;
    const int layout = vals != nullptr ? LAYOUT_SOA : LAYOUT_KEYS;
    const sort_options opt = resolve_options(cfg);
    const sort_variant& var = pick_variant(n, layout, opt);
    const uint32_t tiles = (uint32_t) (((size_t) n + var.tile - 1) / var.tile);
    sort_control* ctl = static_cast<sort_control*>(ctl_mem);
    uint32_t* lookback = reinterpret_cast<uint32_t*>(ctl + 1);
    // the host clears the control block and the first look-back plane; every pass clears its successor's rows
    VRENB200_TRY(check_cuda(cudaMemsetAsync(ctl_mem, 0, sizeof(sort_control) + (size_t) tiles * kRadix * sizeof(uint32_t), s)));
    if (prof) cudaEventRecord(prof->ev[0], s);
    VRENB200_TRY(launch_digit_histograms(s, keys, n, ctl));
    if (prof) cudaEventRecord(prof->ev[1], s);
    VRENB200_TRY(launch_scan_histograms(s, ctl, kPasses));
    if (prof) cudaEventRecord(prof->ev[2], s);
    // ping-pong: the i-th executed pass reads (keys, vals) when i is even; with 4 passes the result is back in `keys`
    for (int i = 0; i < num_passes; i++)
    {
        const bool even = (i & 1) == 0;
        pass_params p{};
        p.keys_in = even ? keys : alt_keys;
        p.keys_out = even ? alt_keys : keys;
        p.vals_in = even ? vals : alt_vals;
        p.vals_out = even ? alt_vals : vals;
        p.n = n;
        p.pass = first_pass + i;
        p.lb_plane = i;
        p.clear_next_plane = i + 1 < num_passes;
        p.ctl = ctl;
        p.lookback = lookback;
        p.num_tiles = tiles;
        p.ticket = opt.tile_ids == VRENB200_TILE_IDS_TICKET ? &ctl->tickets[i] : nullptr;
        p.selftest = opt.ranking == VRENB200_RANKING_SELFTEST_REDO;
        VRENB200_TRY(var.launch(s, p, tiles, layout, false));
        if (var.redo) VRENB200_TRY(var.redo(s, p, layout, false));
        if (prof && num_passes == kPasses) cudaEventRecord(prof->ev[3 + i], s);
    }
    return VRENB200_OK;
}

} // namespace
} // namespace vrenb200

using namespace vrenb200;

// name of the pass configuration a sort of n elements would use with the given configuration (reporting)
extern "C" const char* vrenb200_radix_sort_selected_variant_name(uint32_t n, int with_values, const vrenb200_sort_config* cfg)
{
    return pick_variant(n, with_values ? LAYOUT_SOA : LAYOUT_KEYS, resolve_options(cfg)).name;
}
extern "C" int vrenb200_radix_sort_num_variants(void) { return kNumVariants; }
extern "C" const char* vrenb200_radix_sort_variant_name(int v)
{
    return (v < 1 || v > kNumVariants) ? "" : g_variants[v - 1].name;
}

extern "C" size_t vrenb200_radix_sort_scratch_bytes(uint32_t n, int with_values)
{
    const size_t alt = align_up((size_t) n * 4, 256);
    return alt * (with_values ? 2 : 1) + control_bytes(n);
}

extern "C" int vrenb200_radix_sort_keys(vrenb200_stream_t stream, uint32_t* keys, uint32_t n,
                                        void* scratch, size_t scratch_bytes)
{
    return vrenb200_radix_sort_ex(stream, keys, nullptr, n, scratch, scratch_bytes, nullptr, nullptr);
}

extern "C" int vrenb200_radix_sort_pairs(vrenb200_stream_t stream, uint32_t* keys, uint32_t* values, uint32_t n,
                                         void* scratch, size_t scratch_bytes)
{
    if (n > 0 && values == nullptr) return VRENB200_EINVAL_ARG;
    return vrenb200_radix_sort_ex(stream, keys, values, n, scratch, scratch_bytes, nullptr, nullptr);
}

extern "C" vrenb200_sort_profile* vrenb200_sort_profile_create(void)
{
    vrenb200_sort_profile* p = new vrenb200_sort_profile();
    for (auto& e : p->ev)
        if (cudaEventCreate(&e) != cudaSuccess) { delete p; return nullptr; }
    return p;
}
extern "C" void vrenb200_sort_profile_destroy(vrenb200_sort_profile* p)
{
    if (!p) return;
    for (auto& e : p->ev) cudaEventDestroy(e);
    delete p;
}
// ms_out[6] = {histogram, histogram scan, pass0, pass1, pass2, pass3}; call after the stream has been synchronised
extern "C" int vrenb200_sort_profile_read(vrenb200_sort_profile* p, float* ms_out)
{
    if (!p || !ms_out) return VRENB200_EINVAL_ARG;
    for (int i = 0; i < kPasses + 2; i++)
        VRENB200_TRY(check_cuda(cudaEventElapsedTime(&ms_out[i], p->ev[i], p->ev[i + 1])));
    return VRENB200_OK;
}
// values may be NULL (keys only); cfg may be NULL (defaults); prof may be NULL
extern "C" int vrenb200_radix_sort_ex(vrenb200_stream_t stream, uint32_t* keys, uint32_t* values, uint32_t n,
                                      void* scratch, size_t scratch_bytes, const vrenb200_sort_config* cfg, vrenb200_sort_profile* prof)
{
    if (n == 0) return VRENB200_OK;
    if (keys == nullptr) return VRENB200_EINVAL_ARG;
    const int kv = values != nullptr;
    if (scratch == nullptr || scratch_bytes < vrenb200_radix_sort_scratch_bytes(n, kv)) return VRENB200_ESCRATCH;
    // opt-in: the whole sort in one CTA, one launch (small_sort.cu); larger inputs and profiled calls take the tiled path
    if (cfg != nullptr && cfg->variant == VRENB200_SORT_VARIANT_SINGLE_CTA && n <= single_cta_sort_max() && prof == nullptr)
    {
        if ((reinterpret_cast<uintptr_t>(keys) | reinterpret_cast<uintptr_t>(values)) & 3) return VRENB200_EALIGN;
        return launch_single_cta_sort(as_stream(stream), keys, values, n, 0, kPasses);
    }
    char* p = static_cast<char*>(scratch);
    const size_t alt = align_up((size_t) n * 4, 256);
    return radix_sort_impl(as_stream(stream), keys, values, n, reinterpret_cast<uint32_t*>(p),
                           kv ? reinterpret_cast<uint32_t*>(p + alt) : nullptr, p + (kv ? 2 : 1) * alt, cfg, prof);
}
// 1 if the ranking check of a sort that used `scratch` failed in some pass (the pass was then repeated by the match kernel);
// device word, read it after the stream has been synchronised
extern "C" const uint32_t* vrenb200_radix_sort_violation_word(void* scratch, uint32_t n, int with_values)
{
    const size_t alt = align_up((size_t) n * 4, 256);
    sort_control* ctl = reinterpret_cast<sort_control*>(static_cast<char*>(scratch) + (with_values ? 2 : 1) * alt);
    return &ctl->order_violation;
}

// ---- building blocks of the NCCL form of the multi-GPU sort (vren_b200/dist.py): digit histograms and a digit-range sort ----
// hist_out: device uint32[4][256], counts of every 8-bit digit of the keys (digit 3 = most significant byte)
extern "C" int vrenb200_radix_digit_histograms(vrenb200_stream_t stream, const uint32_t* keys, uint32_t n, uint32_t* hist_out)
{
    if (hist_out == nullptr || (n > 0 && keys == nullptr)) return VRENB200_EINVAL_ARG;
    if ((reinterpret_cast<uintptr_t>(keys) & 15) || (reinterpret_cast<uintptr_t>(hist_out) & 3)) return VRENB200_EALIGN;
    cudaStream_t s = as_stream(stream);
    VRENB200_TRY(check_cuda(cudaMemsetAsync(hist_out, 0, sizeof(uint32_t) * kPasses * kRadix, s)));
    if (n == 0) return VRENB200_OK;
    // the kernel addresses its output through sort_control::hist
    sort_control* fake = reinterpret_cast<sort_control*>(reinterpret_cast<char*>(hist_out) - offsetof(sort_control, hist));
    return launch_digit_histograms(s, keys, n, fake);
}

extern "C" size_t vrenb200_radix_sort_range_scratch_bytes(uint32_t n) { return control_bytes(n); }

// stable sort by the digits [first_pass, first_pass + num_passes) only (8 bits each, pass 0 = least significant).
// Ping-pongs between (keys, values) and (alt_keys, alt_values); *result_in_alt = num_passes & 1. values may be NULL.
extern "C" int vrenb200_radix_sort_pairs_range(vrenb200_stream_t stream, uint32_t* keys, uint32_t* values, uint32_t* alt_keys,
                                               uint32_t* alt_values, uint32_t n, int first_pass, int num_passes,
                                               void* scratch, size_t scratch_bytes, int* result_in_alt)
{
    if (first_pass < 0 || num_passes < 1 || first_pass + num_passes > kPasses) return VRENB200_EINVAL_ARG;
    if (result_in_alt) *result_in_alt = num_passes & 1;
    if (n == 0) return VRENB200_OK;
    if (keys == nullptr || alt_keys == nullptr || ((values == nullptr) != (alt_values == nullptr))) return VRENB200_EINVAL_ARG;
    if (scratch == nullptr || scratch_bytes < control_bytes(n)) return VRENB200_ESCRATCH;
    return radix_sort_impl(as_stream(stream), keys, values, n, alt_keys, alt_values, scratch, nullptr, nullptr, first_pass, num_passes);
}

extern "C" size_t vrenb200_radix_sort_scratch_buffer_1_bytes(uint32_t n) { return control_bytes(n); }
extern "C" size_t vrenb200_radix_sort_scratch_buffer_2_bytes(uint32_t n) { return align_up((size_t) n * 4, 256); }

extern "C" int vrenb200_radix_sort_compat(vrenb200_stream_t stream, uint32_t* keys, uint32_t n,
                                          void* scratch_1, size_t scratch_1_bytes,
                                          void* scratch_2, size_t scratch_2_bytes)
{
    // radix_sort.cpp:158-161: "Length must be higher than 1024 and a power of 2"
    if (!(n >= 1024 && (n & (n - 1)) == 0)) return VRENB200_EINVAL_LENGTH;
    if (keys == nullptr) return VRENB200_EINVAL_ARG;
    if (scratch_1 == nullptr || scratch_1_bytes < control_bytes(n)) return VRENB200_ESCRATCH;
    if (scratch_2 == nullptr || scratch_2_bytes < (size_t) n * 4) return VRENB200_ESCRATCH;
    return radix_sort_impl(as_stream(stream), keys, nullptr, n, static_cast<uint32_t*>(scratch_2), nullptr, scratch_1, nullptr);
}

extern "C" size_t vrenb200_radix_sort_host_work_bytes(uint32_t n, int with_values)
{
    const size_t buf = align_up((size_t) n * 4, 256);
    return buf * (with_values ? 2 : 1) + vrenb200_radix_sort_scratch_bytes(n, with_values);
}

// Enqueue-only form with separate source and destination host buffers (equal pointers = in place): H2D, sort, D2H on
// `stream`, no synchronisation.  With pinned host memory, calls on different streams (each with its own device work
// buffer) overlap one call's upload with another call's download: PCIe is full duplex, a single call is not.
extern "C" int vrenb200_radix_sort_pairs_host_async(vrenb200_stream_t stream, const uint32_t* keys_in_host, const uint32_t* values_in_host,
                                                    uint32_t* keys_out_host, uint32_t* values_out_host, uint32_t n,
                                                    void* dev_work, size_t dev_work_bytes)
{
    if (n == 0) return VRENB200_OK;
    if (keys_in_host == nullptr || keys_out_host == nullptr || dev_work == nullptr) return VRENB200_EINVAL_ARG;
    if ((values_in_host == nullptr) != (values_out_host == nullptr)) return VRENB200_EINVAL_ARG;
    const int kv = values_in_host != nullptr;
    if (dev_work_bytes < vrenb200_radix_sort_host_work_bytes(n, kv)) return VRENB200_ESCRATCH;
    cudaStream_t s = as_stream(stream);
    char* p = static_cast<char*>(dev_work);
    const size_t buf = align_up((size_t) n * 4, 256);
    uint32_t* dk = reinterpret_cast<uint32_t*>(p);
    uint32_t* dv = kv ? reinterpret_cast<uint32_t*>(p + buf) : nullptr;
    char* scratch = p + buf * (kv ? 2 : 1);
    const size_t scratch_bytes = dev_work_bytes - buf * (kv ? 2 : 1);
    VRENB200_TRY(check_cuda(cudaMemcpyAsync(dk, keys_in_host, (size_t) n * 4, cudaMemcpyHostToDevice, s)));
    if (kv) VRENB200_TRY(check_cuda(cudaMemcpyAsync(dv, values_in_host, (size_t) n * 4, cudaMemcpyHostToDevice, s)));
    VRENB200_TRY(vrenb200_radix_sort_ex(stream, dk, dv, n, scratch, scratch_bytes, nullptr, nullptr));
    VRENB200_TRY(check_cuda(cudaMemcpyAsync(keys_out_host, dk, (size_t) n * 4, cudaMemcpyDeviceToHost, s)));
    if (kv) VRENB200_TRY(check_cuda(cudaMemcpyAsync(values_out_host, dv, (size_t) n * 4, cudaMemcpyDeviceToHost, s)));
    return VRENB200_OK;
}

extern "C" int vrenb200_radix_sort_pairs_host(vrenb200_stream_t stream, uint32_t* keys_host, uint32_t* values_host,
                                              uint32_t n, void* dev_work, size_t dev_work_bytes)
{
    if (n == 0) return VRENB200_OK;
    VRENB200_TRY(vrenb200_radix_sort_pairs_host_async(stream, keys_host, values_host, keys_host, values_host, n, dev_work, dev_work_bytes));
    return check_cuda(cudaStreamSynchronize(as_stream(stream)));
}

// ---- a4: bucket sort ------------------------------------------------------------------------------------------
constexpr uint32_t kBucketSearchMin = 1u << 20;   // from this many pairs on, END offsets come from a search in the sorted output

extern "C" size_t vrenb200_bucket_sort_output_bytes(uint32_t n)
{
    // bucket_sort.cpp:67-70
    return align_up((size_t) n * 8, 256) + (size_t) kBucketKeys * sizeof(uint32_t);
}

extern "C" size_t vrenb200_bucket_sort_scratch_bytes(uint32_t n)
{
    return align_up((size_t) n * 8, 256) + control_bytes(n) + (size_t) kBucketKeys * sizeof(uint32_t);   // tmp | control | raw bucket counts
}

extern "C" int vrenb200_bucket_sort(vrenb200_stream_t stream, const void* in_pairs, uint32_t n, void* out,
                                    void* scratch, size_t scratch_bytes)
{
    return vrenb200_bucket_sort_ex(stream, in_pairs, n, out, scratch, scratch_bytes, nullptr, -1);
}

// end_offsets: -1 automatic, 0 per-key global atomics in the histogram read, 1 search in the sorted output (identical results)
extern "C" int vrenb200_bucket_sort_ex(vrenb200_stream_t stream, const void* in_pairs, uint32_t n, void* out,
                                       void* scratch, size_t scratch_bytes, const vrenb200_sort_config* cfg, int end_offsets)
{
    if (out == nullptr || (n > 0 && in_pairs == nullptr)) return VRENB200_EINVAL_ARG;
    if (n >= (1u << 30)) return VRENB200_ELIMIT;
    if ((reinterpret_cast<uintptr_t>(in_pairs) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(scratch)) & 15)
        return VRENB200_EALIGN;
    cudaStream_t s = as_stream(stream);
    uint32_t* counters = reinterpret_cast<uint32_t*>(static_cast<char*>(out) + align_up((size_t) n * 8, 256)); // bucket_sort.cpp:86
    if (n == 0) return check_cuda(cudaMemsetAsync(counters, 0, kBucketKeys * sizeof(uint32_t), s));           // bucket_sort.cpp:104
    if (scratch == nullptr || scratch_bytes < vrenb200_bucket_sort_scratch_bytes(n)) return VRENB200_ESCRATCH;
    char* sp = static_cast<char*>(scratch);
    uint32_t* tmp = reinterpret_cast<uint32_t*>(sp);
    void* ctl_mem = sp + align_up((size_t) n * 8, 256);
    uint32_t* raw_counts = reinterpret_cast<uint32_t*>(sp + align_up((size_t) n * 8, 256) + control_bytes(n));
    const sort_options opt = resolve_options(cfg);
    const sort_variant& var = pick_variant(n, LAYOUT_AOS, opt);
    const uint32_t tiles = (uint32_t) (((size_t) n + var.tile - 1) / var.tile);
    sort_control* ctl = static_cast<sort_control*>(ctl_mem);
    uint32_t* lookback = reinterpret_cast<uint32_t*>(ctl + 1);
    VRENB200_TRY(check_cuda(cudaMemsetAsync(ctl_mem, 0, sizeof(sort_control) + (size_t) tiles * kRadix * sizeof(uint32_t), s)));
    // small inputs (the light Morton sort of the clustered chain): bucket counts by global atomics in the histogram read,
    // prefix by bucket_end_offsets_kernel; large inputs: no global atomics, END offsets by search in the sorted output
    const bool by_search = end_offsets < 0 ? n >= kBucketSearchMin : end_offsets != 0;
    if (!by_search) VRENB200_TRY(check_cuda(cudaMemsetAsync(raw_counts, 0, kBucketKeys * sizeof(uint32_t), s)));
    const uint32_t hist_grid = (uint32_t) std::min<size_t>(kNumSMs * (by_search ? 4 : 2), ((size_t) n / 2 + kHistThreads - 1) / kHistThreads + 1);
    if (by_search)
        bucket_digit_histogram_kernel<<<hist_grid, kHistThreads, 0, s>>>(static_cast<const uint2*>(in_pairs), n, ctl);
    else
        bucket_histogram_kernel<<<hist_grid, kHistThreads, 0, s>>>(static_cast<const uint2*>(in_pairs), n, ctl, raw_counts);
    VRENB200_TRY(check_launch());
    VRENB200_TRY(launch_scan_histograms(s, ctl, 2));
    for (int i = 0; i < 2; i++)
    {
        pass_params p{};
        p.keys_in = i == 0 ? static_cast<const uint32_t*>(in_pairs) : tmp;
        p.keys_out = i == 0 ? tmp : static_cast<uint32_t*>(out);
        p.n = n;
        p.pass = i;
        p.lb_plane = i;
        p.clear_next_plane = i == 0;
        p.ctl = ctl;
        p.lookback = lookback;
        p.num_tiles = tiles;
        p.ticket = opt.tile_ids == VRENB200_TILE_IDS_TICKET ? &ctl->tickets[i] : nullptr;
        p.selftest = opt.ranking == VRENB200_RANKING_SELFTEST_REDO;
        VRENB200_TRY(var.launch(s, p, tiles, LAYOUT_AOS, false));
        if (var.redo) VRENB200_TRY(var.redo(s, p, LAYOUT_AOS, false));
    }
    if (by_search)
        bucket_end_offsets_search_kernel<<<kBucketKeys / 256, 256, 0, s>>>(static_cast<const uint2*>(out), n, counters);
    else
        bucket_end_offsets_kernel<<<kEndOffsetCtas, 1024, 0, s>>>(raw_counts, counters);
    return check_launch();
}
