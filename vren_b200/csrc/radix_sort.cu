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
//   3. Per digit ONE fused kernel.  A CTA takes a tile (dynamic ticket), stages keys (and values) into shared
//      memory with a single-thread TMA bulk copy (cp.async.bulk + mbarrier: no register staging, values land
//      while keys are being ranked), ranks keys stably with warp-ballot digit matching against warp-private
//      histograms, resolves the tile's global digit offsets with a decoupled look-back over a flag|count word
//      per (tile, digit), regroups keys/values by digit in shared memory and writes them out coalesced.
// Radix sort = 4 digits: 4n (histogram) + 4 x 8n keys [+ 4 x 8n values] = 36 B/key, 68 B/pair of HBM traffic.
// Bucket sort = the same kernel over interleaved uvec2 pairs with 2 digits (16-bit key): 8n + 2 x 16n = 40 B/pair,
// deterministic and stable (the canonical tie-break), bucket END offsets from a fused 65536-bin count.
// Stability: warp-striped order (warp, item, lane) == element order, so equal keys keep input order.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace vrenb200 {

namespace {

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kPasses = 32 / kRadixBits;

constexpr uint32_t kLbFlagAggregate = 1u << 30;
constexpr uint32_t kLbFlagInclusive = 2u << 30;
constexpr uint32_t kLbValueMask = (1u << 30) - 1;

constexpr uint32_t kBucketKeys = 1u << 16; // bucket_sort.hpp:15-16
constexpr int kMaxRanks = 32;              // fused exchange: destinations per call

// device table of the fused partition + exchange (vrenb200_radix_partition_scatter)
struct p2p_table
{
    unsigned long long kptr[kMaxRanks];   // where THIS rank's block starts in every destination's key buffer
    unsigned long long vptr[kMaxRanks];   // ... value buffer
    uint8_t rank_of[kRadix];              // destination rank of a key, indexed by its most significant byte
};

// ---- control block carved from scratch ----------------------------------------------------------------------
struct sort_control
{
    uint32_t tickets[kPasses];            // dynamic tile ids per pass
    uint32_t _pad[60];
    uint32_t hist[kPasses][kRadix];       // global digit counts, then exclusive offsets
    // followed by look-back words: [passes][tiles][kRadix]
};

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
__device__ uint32_t g_hist_loads_in_flight = 4;   // tuning hook (vrenb200_radix_sort_set_hist_loads): 2 or 4 (0.217 vs 0.207 ms at 2^28 keys)
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
    if (g_hist_loads_in_flight >= 4)
    {
        // four 128-bit loads in flight per thread (64 KB per SM): one CTA of 1024 threads per SM needs that much to cover
        // the HBM latency at full bandwidth
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
enum { LAYOUT_KEYS = 0, LAYOUT_SOA = 1, LAYOUT_AOS = 2 };
constexpr int kClearNextPassRow = 0x100;   // flag in the `pass` argument of the count-first kernel

template <int THREADS, int ITEMS, int LAYOUT, bool P2P = false>
struct onesweep_smem
{
    static constexpr int WARPS = THREADS / 32;
    static constexpr int TILE = THREADS * ITEMS;
    static constexpr int KV_WORDS = LAYOUT == LAYOUT_KEYS ? TILE : 2 * TILE;
    // KEYS: keys[TILE] | SOA: keys[TILE] then values[TILE] | AOS: uint2[TILE]
    alignas(128) uint32_t kv[KV_WORDS];
    uint32_t warp_hist[WARPS][kRadix];
    uint32_t digit_base[kRadix];
    uint32_t tile_hist[kRadix];       // EARLY_HIST: digit counts of the tile, known before the ranking
    // P2P_DEST: per-digit destination base pointers (keys, values), possibly in a peer GPU's memory
    alignas(8) unsigned long long dst_ptr[P2P ? 2 : 1][P2P ? kMaxRanks : 1];
    uint8_t rank_of[P2P ? kRadix : 4];    // P2P_DEST: destination rank of every most-significant byte
    uint32_t run_start[P2P ? kMaxRanks : 1], run_len[P2P ? kMaxRanks : 1], run_g[P2P ? kMaxRanks : 1];   // P2P_DEST: per-destination run of the tile
    uint32_t scan_warp[kRadix / 32];
    uint32_t lane_dummy[WARPS][32];   // RANK_LEADER_ATOMIC: where the lanes that do not lead a match group add
    alignas(8) uint64_t bar_keys;
    alignas(8) uint64_t bar_vals;
    alignas(8) uint64_t bar_chunk[4];   // KEYS_CHUNKED: one barrier per quarter of the staged keys
    uint32_t tile;
};

enum { MATCH_BALLOT = 0, MATCH_BALLOT_C = 1, TILE_BY_BLOCKIDX = 2, EARLY_HIST = 4, P2P_DEST = 8, DEPHASE = 16, LEADER_ATOMIC = 32, SPLIT_KV = 64,
       FAKE_LOOKBACK = 128 /* timing experiment: no chain, approximate destinations (WRONG results) */,
       DIRECT_LOAD = 256 /* count-first kernel: keys / values go from global memory straight to registers (no staging copy) */,
       LB_INTERLEAVED = 512 /* count-first kernel: the look-back advances in non-blocking steps between ranking rows */,
       LB_STEP2 = 1024, LB_STEP8 = 2048 /* ... every 2 / every 8 rows instead of every 4 */,
       PREFETCH_L2 = 4096 /* count-first kernel: a CTA asks L2 for the tile of the CTA that will take its place on the SM */,
       RANK_LEADER_ATOMIC = 8192 /* count-first kernel: one returning shared atomic by the leader of every match group + shuffle,
                                    instead of a counter load and store by every lane */,
       EARLY_TMA = 32768 /* count-first kernel: the staging copies are issued before the counters are cleared */,
       VALS_DIRECT = 65536 /* count-first kernel, SOA: the values go from global memory straight to registers (issued at the
                              start of the tile, consumed by the regroup stores); only the keys are staged */,
       REG_COUNTS = 131072 /* count-first kernel: the per-warp digit counts stay in registers between the publish and the
                              offset step instead of being read from shared memory twice */,
       RANK_ATOMIC_ORDER = 2097152 /* rank = returning shared atomic per lane, no match (relies on the lane order of
                                      same-address shared atomics: guarded by the device probe, see pick_variant) */,
       LB_STEP16 = 4194304 /* look-back step every 16 ranking rows */,
       MATCH_SPLIT4 = 1048576 /* ballot match with four accumulators (shorter dependent chains) */,
       VALS_LATE = 524288 /* VALS_DIRECT: the value loads are issued after the counting step instead of before the wait for the keys */,
       KEYS_CHUNKED = 262144 /* count-first kernel: the key staging copy is split in four, every warp waits only for the
                                quarter that holds its own keys */
     }; // option bits of the MATCH template argument

// lanes of the warp holding the same 8-bit digit.
// MATCH_BALLOT: hand-scheduled, 4 instructions per bit (bit test -> predicate, vote, two predicated LOP3);
// MATCH_BALLOT_C: the plain C++ form (the compiler spends 6 per bit), kept for A/B runs.
template <int MATCH>
__device__ __forceinline__ unsigned match_digit(uint32_t d)
{
    unsigned mask = kFullMask;
    if (MATCH & MATCH_SPLIT4)
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
        mask = (ones_lo & ones_hi) & ~(zeros_lo | zeros_hi);
    }
    else if ((MATCH & MATCH_BALLOT_C) == 0)
    {
        // peers = AND of the ballots of my set bits, minus OR of the ballots of my clear bits: two independent
        // accumulators, each updated by ONE predicated LOP3 per bit (vote + 2 instructions per bit)
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
        mask = ones & ~zeros;
    }
    else
    {
#pragma unroll
        for (int b = 0; b < kRadixBits; b++)
        {
            const bool bit = (d >> b) & 1u;
            const unsigned bal = __ballot_sync(kFullMask, bit);
            mask &= bit ? bal : ~bal;
        }
    }
    return mask;
}

__device__ __forceinline__ uint32_t digit_of(uint32_t key, uint32_t prmt_sel) { return __byte_perm(key, 0u, prmt_sel); }

template <int THREADS, int ITEMS, int LAYOUT, int MATCH, int MIN_BLOCKS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
onesweep_pass_kernel(const uint32_t* __restrict__ keys_in, uint32_t* __restrict__ keys_out,
                     const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ vals_out,
                     uint32_t n, int pass, sort_control* ctl, uint32_t* lookback, uint32_t num_tiles)
{
    using smem_t = onesweep_smem<THREADS, ITEMS, LAYOUT, (MATCH & P2P_DEST) != 0>;
    constexpr int WARPS = smem_t::WARPS;
    constexpr int TILE = smem_t::TILE;
    constexpr bool HAS_VALUES = LAYOUT != LAYOUT_KEYS;
    constexpr int KSTRIDE = LAYOUT == LAYOUT_AOS ? 2 : 1;       // words between consecutive staged keys
    constexpr int VOFF = LAYOUT == LAYOUT_AOS ? 1 : TILE;       // word offset key -> its value
    constexpr uint32_t ELEM_BYTES = LAYOUT == LAYOUT_AOS ? 8 : 4;
    static_assert(THREADS >= kRadix, "one thread per digit needed");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    smem_t& sm = *reinterpret_cast<smem_t*>(smem_raw);

    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t prmt_sel = 0x4440u | (uint32_t) pass;   // byte `pass` of the key -> one PRMT per digit extraction
    // P2P_DEST: the "digit" is the destination rank of the key (looked up by its most significant byte), so a tile
    // produces one long run per destination instead of 256 short ones: long contiguous stores over NVLink
    auto digit_of = [&](uint32_t k, uint32_t sel) -> uint32_t {
        if (MATCH & P2P_DEST) return sm.rank_of[k >> 24];
        return __byte_perm(k, 0u, sel);
    };
    if (MATCH & P2P_DEST)
    {
        const p2p_table* table = reinterpret_cast<const p2p_table*>(keys_out);
        if (tid < kRadix) sm.rank_of[tid] = table->rank_of[tid];
    }

    if (tid == 0)
    {
        // tile id: dynamic ticket (a tile can only wait on tiles that already started), or the block index when the
        // variant relies on in-order CTA dispatch (saves one L2 atomic round trip at the head of every tile)
        sm.tile = (MATCH & TILE_BY_BLOCKIDX) ? blockIdx.x : atomicAdd(&ctl->tickets[pass], 1u);
        mbar_init(&sm.bar_keys, 1);
        mbar_init(&sm.bar_vals, 1);
        mbar_fence_init();
    }
    // warp-private digit counters
#pragma unroll
    for (int i = lane; i < kRadix; i += 32) sm.warp_hist[warp][i] = 0;
    if ((MATCH & EARLY_HIST) && tid < kRadix) sm.tile_hist[tid] = 0;
    __syncthreads();

    const uint32_t tile = sm.tile;
    const uint64_t tile_base = (uint64_t) tile * TILE;
    const uint32_t valid = (n - tile_base) < (uint64_t) TILE ? (uint32_t) (n - tile_base) : (uint32_t) TILE;
    const bool full = valid == (uint32_t) TILE;

    if (full)
    {
        if (tid == 0)
        {
            mbar_arrive_expect_tx(&sm.bar_keys, TILE * ELEM_BYTES);
            bulk_copy_g2s(sm.kv, keys_in + tile_base * KSTRIDE, TILE * ELEM_BYTES, &sm.bar_keys);
            if (LAYOUT == LAYOUT_SOA)
            {
                mbar_arrive_expect_tx(&sm.bar_vals, TILE * 4);
                bulk_copy_g2s(sm.kv + TILE, vals_in + tile_base, TILE * 4, &sm.bar_vals);
            }
        }
        mbar_wait(&sm.bar_keys, 0);
    }
    else
    {
        // ragged last tile: guarded loads, padding keys 0xFFFFFFFF sort behind every real key of the tile
        for (uint32_t i = tid; i < (uint32_t) TILE; i += THREADS)
        {
            const bool in = i < valid;
            sm.kv[i * KSTRIDE] = in ? keys_in[(tile_base + i) * KSTRIDE] : 0xFFFFFFFFu;
            if (LAYOUT == LAYOUT_SOA) sm.kv[TILE + i] = in ? vals_in[tile_base + i] : 0u;
            if (LAYOUT == LAYOUT_AOS) sm.kv[i * 2 + 1] = in ? keys_in[(tile_base + i) * 2 + 1] : 0u;
        }
        __syncthreads();
    }

    // warp-striped register tile: element (warp, j, lane)
    const uint32_t warp_off = warp * (ITEMS * 32) + lane;
    uint32_t key[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; j++) key[j] = sm.kv[(warp_off + j * 32) * KSTRIDE];

    // EARLY_HIST: count the tile's digits with shared atomics right away, publish the aggregate and put the first
    // look-back loads in flight BEFORE the (long) ranking phase: successors never find an empty word, and this
    // tile's own look-back latency hides behind its ranking
    constexpr int LBK = 4;
    uint32_t* lb = lookback + ((size_t) pass * num_tiles + tile) * kRadix;
    uint32_t early_cnt = 0, lb_pre[LBK];
    if (MATCH & EARLY_HIST)
    {
#pragma unroll
        for (int j = 0; j < ITEMS; j++) atomicAdd(&sm.tile_hist[digit_of(key[j], prmt_sel)], 1u);
        __syncthreads();
        if (tid < kRadix)
        {
            early_cnt = sm.tile_hist[tid];
            const uint32_t real = early_cnt - ((tid == kRadix - 1) ? (uint32_t) TILE - valid : 0u);
            st_relaxed_u32(&lb[tid], (tile == 0 ? kLbFlagInclusive : kLbFlagAggregate) | real);
#pragma unroll
            for (int k = 0; k < LBK; k++)
                lb_pre[k] = ((int64_t) tile - 1 - k >= 0) ? ld_relaxed_u32(lb - (k + 1) * kRadix + tid) : kLbFlagInclusive;
        }
    }

    // stable in-warp ranking: ballot match + warp-private running digit counters.  Every lane of a match group
    // reads the counter, then every lane writes back the same new value (no leader election, no divergence).
    uint32_t rank[ITEMS];
    uint32_t* my_hist = sm.warp_hist[warp];
    const unsigned lt = lanemask_lt();
#pragma unroll
    for (int j = 0; j < ITEMS; j++)
    {
        const uint32_t d = digit_of(key[j], prmt_sel);
        const unsigned mask = match_digit<MATCH>(d);
        if (MATCH & LEADER_ATOMIC)
        {
            // one shared atomic by the group's leader + a shuffle instead of a load and a store by every lane: fewer
            // shared-memory wavefronts (the limiter), a few more instructions
            const unsigned leader = __ffs(mask) - 1;
            uint32_t prior = 0;
            if (lane == leader) prior = atomicAdd(&my_hist[d], (uint32_t) __popc(mask));
            prior = __shfl_sync(kFullMask, prior, leader);
            rank[j] = prior + __popc(mask & lt);
        }
        else
        {
            const uint32_t prior = my_hist[d];
            rank[j] = prior + __popc(mask & lt);
            __syncwarp();
            my_hist[d] = prior + __popc(mask);
            __syncwarp();
        }
    }
    __syncthreads();

    // per digit: tile count, publish aggregate, tile-local exclusive offsets
    uint32_t cnt = 0, inc = 0, real_cnt = 0;
    if (tid < kRadix)
    {
        if (MATCH & EARLY_HIST)
            cnt = early_cnt;
        else
        {
#pragma unroll
            for (int w = 0; w < WARPS; w++) cnt += sm.warp_hist[w][tid];
        }
        // the padding keys (0xFFFFFFFF) of a ragged last tile carry the last digit; in exchange mode their "digit" is the
        // destination rank of the top byte 0xFF
        const uint32_t pad_digit = (MATCH & P2P_DEST) ? (uint32_t) sm.rank_of[kRadix - 1] : (uint32_t) kRadix - 1;
        real_cnt = cnt - ((tid == pad_digit) ? (uint32_t) TILE - valid : 0u);
        if (!(MATCH & EARLY_HIST)) st_relaxed_u32(&lb[tid], (tile == 0 ? kLbFlagInclusive : kLbFlagAggregate) | real_cnt);
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
            const uint32_t c = sm.warp_hist[w][tid];
            sm.warp_hist[w][tid] = run;
            run += c;
        }
    }
    __syncthreads();

    // in-tile destination of every item; fetch the staged values with the same striping
#pragma unroll
    for (int j = 0; j < ITEMS; j++) rank[j] += my_hist[digit_of(key[j], prmt_sel)];
    // SPLIT_KV (SOA only): keys and values are regrouped one after the other in their own halves of the staging buffer,
    // so the key registers are dead before the values are fetched (a third fewer live registers: more resident warps)
    constexpr bool SPLIT = (MATCH & SPLIT_KV) != 0 && LAYOUT == LAYOUT_SOA;
    if (SPLIT)
    {
        // every warp passed the barriers above after loading its keys: the key half can be overwritten in place
#pragma unroll
        for (int j = 0; j < ITEMS; j++) sm.kv[rank[j]] = key[j];
        if (full) mbar_wait(&sm.bar_vals, 0);
        uint32_t val[ITEMS];
#pragma unroll
        for (int j = 0; j < ITEMS; j++) val[j] = sm.kv[TILE + warp_off + j * 32];
        __syncthreads();
#pragma unroll
        for (int j = 0; j < ITEMS; j++) sm.kv[TILE + rank[j]] = val[j];
    }
    else
    {
        uint32_t val[HAS_VALUES ? ITEMS : 1];
        if (HAS_VALUES)
        {
            if (LAYOUT == LAYOUT_SOA && full) mbar_wait(&sm.bar_vals, 0);
#pragma unroll
            for (int j = 0; j < ITEMS; j++) val[j] = sm.kv[(warp_off + j * 32) * KSTRIDE + VOFF];
        }
        __syncthreads(); // every warp has consumed the staged inputs: the buffers become the regroup area

#pragma unroll
        for (int j = 0; j < ITEMS; j++)
        {
            // pairs are regrouped interleaved (one 64-bit shared store / load per pair) whatever the global layout
            if (HAS_VALUES)
                reinterpret_cast<uint2*>(sm.kv)[rank[j]] = make_uint2(key[j], val[j]);
            else
                sm.kv[rank[j]] = key[j];
        }
    }

    // decoupled look-back, one thread per digit
    if (tid < kRadix)
    {
        uint32_t exclusive = 0;
        if (MATCH & FAKE_LOOKBACK)
            exclusive = tile * (TILE / kRadix);
        else if (tile > 0)
        {
            // K predecessors are fetched per round trip (independent loads), then consumed in order: with hundreds of
            // tiles in flight a one-at-a-time walk spends most of the tile's life in dependent L2 round trips
            constexpr int K = 4;
            const uint32_t* p = lb - kRadix + tid;
            int64_t t = (int64_t) tile - 1;
            bool done = false;
            while (!done)
            {
                uint32_t s[K];
                static_assert(K == LBK, "prefetch depth");
#pragma unroll
                for (int k = 0; k < K; k++)
                {
                    if ((MATCH & EARLY_HIST) && t == (int64_t) tile - 1)
                        s[k] = lb_pre[k];   // fetched before the ranking; anything still empty is re-polled below
                    else
                        s[k] = (t - k >= 0) ? ld_relaxed_u32(p - k * kRadix) : kLbFlagInclusive;
                }
#pragma unroll
                for (int k = 0; k < K; k++)
                {
                    if (done) break;
                    while ((s[k] >> 30) == 0) s[k] = ld_relaxed_u32(p - k * kRadix);
                    exclusive += s[k] & kLbValueMask;
                    done = (s[k] >> 30) == 2;
                }
                t -= K;
                p -= K * kRadix;
            }
            st_relaxed_u32(&lb[tid], kLbFlagInclusive | (exclusive + real_cnt));
        }
        if (MATCH & P2P_DEST)
        {
            // exchange mode: keys_out is a device table [2][256] of per-digit destination pointers (keys, values);
            // a digit's run starts at its pointer, this tile's slice of it at the look-back prefix
            const p2p_table* table = reinterpret_cast<const p2p_table*>(keys_out);
            if (tid < kMaxRanks)
            {
                sm.dst_ptr[0][(MATCH & P2P_DEST) ? tid : 0] = table->kptr[tid];
                sm.dst_ptr[(MATCH & P2P_DEST) ? 1 : 0][(MATCH & P2P_DEST) ? tid : 0] = table->vptr[tid];
            }
            sm.digit_base[tid] = exclusive - tile_off;
            if (tid < kMaxRanks)
            {
                sm.run_start[(MATCH & P2P_DEST) ? tid : 0] = tile_off;
                sm.run_len[(MATCH & P2P_DEST) ? tid : 0] = real_cnt;
                sm.run_g[(MATCH & P2P_DEST) ? tid : 0] = exclusive;
            }
        }
        else
            sm.digit_base[tid] = ctl->hist[pass][tid] + exclusive - tile_off;
    }
    __syncthreads();

    if ((MATCH & P2P_DEST) && LAYOUT == LAYOUT_SOA)
    {
        // exchange mode: one run per destination rank, written in chunks that start on 128-byte lines of the DESTINATION
        // (a warp store that straddles two lines becomes two partial NVLink write packets; the flat position loop below
        // reached only ~55 % of the measured peer-store bandwidth, profiles/r1s_*)
        constexpr int NR = (MATCH & P2P_DEST) ? kMaxRanks : 1;
        for (int d = 0; d < NR; d++)
        {
            const uint32_t len = sm.run_len[d];
            if (len == 0) continue;
            const uint32_t s0 = sm.run_start[d], g0 = sm.run_g[d];
            uint32_t* kdst = reinterpret_cast<uint32_t*>(sm.dst_ptr[0][d]) + g0;
            uint32_t* vdst = reinterpret_cast<uint32_t*>(sm.dst_ptr[(MATCH & P2P_DEST) ? 1 : 0][d]) + g0;
            const int32_t mis = (int32_t) ((reinterpret_cast<uintptr_t>(kdst) >> 2) & 31);   // elements past a 128-byte line
            for (int32_t i = -mis + 32 * (int32_t) warp; i < (int32_t) len; i += 32 * WARPS)
            {
                const int32_t e = i + (int32_t) lane;
                if (e >= 0 && e < (int32_t) len)
                {
                    const uint2 kv = reinterpret_cast<const uint2*>(sm.kv)[s0 + e];
                    kdst[e] = kv.x;
                    vdst[e] = kv.y;
                }
            }
        }
    }
    else if (full)
    {
#pragma unroll
        for (int j = 0; j < ITEMS; j++)
        {
            const uint32_t p = j * THREADS + tid;
            if (HAS_VALUES)
            {
                const uint2 e = SPLIT ? make_uint2(sm.kv[p], sm.kv[TILE + p]) : reinterpret_cast<const uint2*>(sm.kv)[p];
                uint32_t g = sm.digit_base[digit_of(e.x, prmt_sel)] + p;
                if ((MATCH & FAKE_LOOKBACK) && g >= n) g = n - 1;
                if (LAYOUT == LAYOUT_AOS)
                    reinterpret_cast<uint2*>(keys_out)[g] = e;
                else if (MATCH & P2P_DEST)
                {
                    const uint32_t d = digit_of(e.x, prmt_sel);
                    reinterpret_cast<uint32_t*>(sm.dst_ptr[0][d])[g] = e.x;          // st.global, peer or local
                    reinterpret_cast<uint32_t*>(sm.dst_ptr[(MATCH & P2P_DEST) ? 1 : 0][d])[g] = e.y;
                }
                else
                {
                    keys_out[g] = e.x;
                    vals_out[g] = e.y;
                }
            }
            else
            {
                const uint32_t k = sm.kv[p];
                keys_out[sm.digit_base[digit_of(k, prmt_sel)] + p] = k;
            }
        }
    }
    else
    {
        for (uint32_t p = tid; p < valid; p += THREADS)
        {
            const uint32_t k = sm.kv[SPLIT ? p : p * (HAS_VALUES ? 2 : 1)];
            const uint32_t v = HAS_VALUES ? sm.kv[SPLIT ? TILE + p : p * 2 + 1] : 0u;
            const uint32_t g = sm.digit_base[digit_of(k, prmt_sel)] + p;
            if (LAYOUT == LAYOUT_AOS)
                reinterpret_cast<uint2*>(keys_out)[g] = make_uint2(k, v);
            else if (MATCH & P2P_DEST)
            {
                const uint32_t d = digit_of(k, prmt_sel);
                reinterpret_cast<uint32_t*>(sm.dst_ptr[0][d])[g] = k;
                reinterpret_cast<uint32_t*>(sm.dst_ptr[(MATCH & P2P_DEST) ? 1 : 0][d])[g] = v;
            }
            else
            {
                keys_out[g] = k;
                if (HAS_VALUES) vals_out[g] = v;
            }
        }
    }
}

// PREFETCH_L2: how many tiles ahead a CTA prefetches (half of the CTAs resident at 2 per SM); tuning hook below
__device__ uint32_t g_prefetch_tiles = kNumSMs;   // 74-148 tiles ahead measured best for the default tile (profiles/r1x_prefetch_distance.log)
// DEPHASE (count-first kernel): the CTAs of the first wave that arrive second on their SM start g_dephase_ns late, so that
// the two CTAs of an SM do not run the same step (counting / ranking / write-out) at the same time.  The offset is
// inherited by the CTAs that replace them.  rule 0: CTA i shares its SM with CTA i + 148; rule 1: with CTA i ^ 1.
__device__ uint32_t g_dephase_ns = 5000, g_dephase_rule = 0;

// ---- 3a. count-first onesweep pass ------------------------------------------------------------------------------
// Same contract as onesweep_pass_kernel, different order of work inside the tile:
//   1. the warp-private digit counters are filled FIRST (one non-returning shared atomic per key),
//   2. the tile's digit counts are published and the look-back loads of the first predecessors are issued,
//   3. the counters are turned into the final in-tile offset of every (warp, digit) run,
//   4. the ballot ranking then yields final in-tile positions, so every pair is stored to the regroup buffer as soon
//      as it is ranked (no rank array in registers, no second counter lookup per item),
//   5. the look-back finishes (its first round trip has been in flight during the whole ranking) and the tile is
//      written out.
// Successors see this tile's aggregate a ranking phase earlier, and this tile's own look-back latency hides behind
// its ranking.
template <int THREADS, int ITEMS, int LAYOUT, int MATCH, int MIN_BLOCKS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
onesweep_count_first_kernel(const uint32_t* __restrict__ keys_in, uint32_t* __restrict__ keys_out,
                            const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ vals_out,
                            uint32_t n, int pass_arg, sort_control* ctl, uint32_t* lookback, uint32_t num_tiles)
{
    // bit 8 of the pass argument: this tile also clears its look-back row of the NEXT pass (the host then clears only the
    // first pass's rows: 22 MB instead of 89 MB of memset per 2^28-pair sort)
    const int pass = pass_arg & 0xFF;
    if ((pass_arg & kClearNextPassRow) && threadIdx.x < kRadix)
        lookback[((size_t) (pass + 1) * num_tiles + blockIdx.x) * kRadix + threadIdx.x] = 0u;
    using smem_t = onesweep_smem<THREADS, ITEMS, LAYOUT, false>;
    constexpr int WARPS = smem_t::WARPS;
    constexpr int TILE = smem_t::TILE;
    constexpr bool HAS_VALUES = LAYOUT != LAYOUT_KEYS;
    constexpr int KSTRIDE = LAYOUT == LAYOUT_AOS ? 2 : 1;
    constexpr int VOFF = LAYOUT == LAYOUT_AOS ? 1 : TILE;
    constexpr uint32_t ELEM_BYTES = LAYOUT == LAYOUT_AOS ? 8 : 4;
    static_assert(THREADS >= kRadix, "one thread per digit needed");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    smem_t& sm = *reinterpret_cast<smem_t*>(smem_raw);

    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t prmt_sel = 0x4440u | (uint32_t) pass;
    const uint32_t tile = blockIdx.x;
    if ((MATCH & DEPHASE) && tile < (uint32_t) (MIN_BLOCKS * kNumSMs))
    {
        const uint32_t slot = g_dephase_rule == 0 ? tile / (uint32_t) kNumSMs : tile % (uint32_t) MIN_BLOCKS;
        if (slot) __nanosleep(slot * g_dephase_ns);
    }
    const uint64_t tile_base = (uint64_t) tile * TILE;
    const uint32_t valid = (n - tile_base) < (uint64_t) TILE ? (uint32_t) (n - tile_base) : (uint32_t) TILE;
    const bool full = valid == (uint32_t) TILE;
    constexpr bool DIRECT = (MATCH & DIRECT_LOAD) != 0;
    constexpr bool EARLY = (MATCH & EARLY_TMA) != 0 && !DIRECT;
    constexpr bool VDIRECT = (MATCH & VALS_DIRECT) != 0 && LAYOUT == LAYOUT_SOA && !DIRECT;
    constexpr bool VLATE = (MATCH & VALS_LATE) != 0;
    constexpr bool CHUNKED = (MATCH & KEYS_CHUNKED) != 0 && !DIRECT && WARPS % 4 == 0;
    constexpr bool KEEP_COUNTS = (MATCH & REG_COUNTS) != 0;
    // staging copies of a full tile (one thread): keys (whole or in quarters), then the values unless they are loaded directly
    auto issue_copies = [&]() {
        if (CHUNKED)
        {
            constexpr uint32_t Q = TILE / 4;
#pragma unroll
            for (int c = 0; c < 4; c++)
            {
                mbar_arrive_expect_tx(&sm.bar_chunk[c], Q * ELEM_BYTES);
                bulk_copy_g2s(sm.kv + c * Q * KSTRIDE, keys_in + (tile_base + c * Q) * KSTRIDE, Q * ELEM_BYTES, &sm.bar_chunk[c]);
            }
        }
        else
        {
            mbar_arrive_expect_tx(&sm.bar_keys, TILE * ELEM_BYTES);
            bulk_copy_g2s(sm.kv, keys_in + tile_base * KSTRIDE, TILE * ELEM_BYTES, &sm.bar_keys);
        }
        if (LAYOUT == LAYOUT_SOA && !VDIRECT)
        {
            mbar_arrive_expect_tx(&sm.bar_vals, TILE * 4);
            bulk_copy_g2s(sm.kv + TILE, vals_in + tile_base, TILE * 4, &sm.bar_vals);
        }
    };
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
        // EARLY: the copies start before the counters are cleared (the staging area is not touched by anyone else yet)
        if (EARLY && full) issue_copies();
    }
    if ((MATCH & PREFETCH_L2) && tid == 32)
    {
        // the CTA that takes this one's place on the SM is about (resident CTAs) tiles ahead: have its input
        // waiting in L2 by the time it starts
        const uint64_t next_base = tile_base + (uint64_t) g_prefetch_tiles * TILE;
        if (next_base + TILE <= (uint64_t) n)
        {
            bulk_prefetch_l2(keys_in + next_base * KSTRIDE, TILE * ELEM_BYTES);
            if (LAYOUT == LAYOUT_SOA) bulk_prefetch_l2(vals_in + next_base, TILE * 4);
        }
    }
    // global offset of this thread's digit: needed at the very end, fetched now
    uint32_t pass_base = 0;
    if ((MATCH & EARLY_TMA) && tid < kRadix) pass_base = ctl->hist[pass][tid];
#pragma unroll
    for (int i = lane; i < kRadix; i += 32) sm.warp_hist[warp][i] = 0;
    __syncthreads();

    const uint32_t warp_off = warp * (ITEMS * 32) + lane;
    uint32_t* my_hist = sm.warp_hist[warp];
    uint32_t key[ITEMS];
    uint32_t val[HAS_VALUES ? ITEMS : 1];
    // VALS_DIRECT: warp-striped value loads into the register tile (no staging copy, no shared load)
    auto load_vals_direct = [&]() {
#pragma unroll
        for (int j = 0; j < ITEMS; j++)
        {
            const uint32_t i = warp_off + j * 32;
            val[j] = full ? ldg_stream_u32(vals_in + tile_base + i) : (i < valid ? vals_in[tile_base + i] : 0u);
        }
    };
    if (DIRECT)
    {
        // warp-striped loads straight into the register tile: one L1 wavefront per 32 keys instead of a staging write
        // plus a shared load, no barrier word, and the shared buffer is only ever the regroup area
        if (full)
        {
#pragma unroll
            for (int j = 0; j < ITEMS; j++)
            {
                if (LAYOUT == LAYOUT_AOS)
                {
                    const uint2 e = reinterpret_cast<const uint2*>(keys_in)[tile_base + warp_off + j * 32];
                    key[j] = e.x;
                    val[j] = e.y;
                }
                else
                    key[j] = ldg_stream_u32(keys_in + tile_base + warp_off + j * 32);
            }
        }
        else
        {
#pragma unroll
            for (int j = 0; j < ITEMS; j++)
            {
                const uint32_t i = warp_off + j * 32;
                key[j] = i < valid ? keys_in[(tile_base + i) * KSTRIDE] : 0xFFFFFFFFu;
                if (LAYOUT == LAYOUT_AOS) val[j] = i < valid ? keys_in[(tile_base + i) * 2 + 1] : 0u;
            }
        }
    }
    else
    {
        if (full)
        {
            if (!EARLY && tid == 0) issue_copies();
            if (VDIRECT && !VLATE) load_vals_direct();   // before the wait for the keys
            mbar_wait(CHUNKED ? &sm.bar_chunk[warp / (WARPS / 4)] : &sm.bar_keys, 0);
        }
        else
        {
            for (uint32_t i = tid; i < (uint32_t) TILE; i += THREADS)
            {
                const bool in = i < valid;
                sm.kv[i * KSTRIDE] = in ? keys_in[(tile_base + i) * KSTRIDE] : 0xFFFFFFFFu;
                if (LAYOUT == LAYOUT_SOA && !VDIRECT) sm.kv[TILE + i] = in ? vals_in[tile_base + i] : 0u;
                if (LAYOUT == LAYOUT_AOS) sm.kv[i * 2 + 1] = in ? keys_in[(tile_base + i) * 2 + 1] : 0u;
            }
            if (VDIRECT && !VLATE) load_vals_direct();
            __syncthreads();
        }
#pragma unroll
        for (int j = 0; j < ITEMS; j++) key[j] = sm.kv[(warp_off + j * 32) * KSTRIDE];
    }
    // 1. count
#pragma unroll
    for (int j = 0; j < ITEMS; j++) atomicAdd(&my_hist[digit_of(key[j], prmt_sel)], 1u);
    __syncthreads();
    if (VDIRECT && VLATE) load_vals_direct();   // the loads complete under the publish and offset steps

    // 2. publish the aggregate, start the look-back
    constexpr int K = 4;
    uint32_t* lb = lookback + ((size_t) pass * num_tiles + tile) * kRadix;
    uint32_t cnt = 0, inc = 0, real_cnt = 0, lb_pre[K];
    uint32_t cw[KEEP_COUNTS ? WARPS : 1];   // REG_COUNTS: this thread's digit, count per warp
    // look-back state of this thread's digit: window of K predecessors starting at tile lb_t, words in lb_pre[]
    constexpr bool INTERLEAVED = (MATCH & LB_INTERLEAVED) != 0;
    const uint32_t* lb_p = lb - kRadix + tid;
    int32_t lb_t = (int32_t) tile - 1;
    bool lb_done = tile == 0 || tid >= kRadix;
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
        st_relaxed_u32(&lb[tid], (tile == 0 ? kLbFlagInclusive : kLbFlagAggregate) | real_cnt);
        lb_load();
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
    if (HAS_VALUES && !DIRECT && !VDIRECT)
    {
        if (LAYOUT == LAYOUT_SOA && full) mbar_wait(&sm.bar_vals, 0);
#pragma unroll
        for (int j = 0; j < ITEMS; j++) val[j] = sm.kv[(warp_off + j * 32) * KSTRIDE + VOFF];
    }
    if (DIRECT && LAYOUT == LAYOUT_SOA)
    {
        // issued here, consumed by the regroup stores: the loads fly during the ranking of the first rows
#pragma unroll
        for (int j = 0; j < ITEMS; j++)
        {
            const uint32_t i = warp_off + j * 32;
            val[j] = full ? ldg_stream_u32(vals_in + tile_base + i) : (i < valid ? vals_in[tile_base + i] : 0u);
        }
    }
    __syncthreads(); // offsets visible; every warp has consumed the staged inputs: the buffers become the regroup area

    // 4. rank and regroup
    const unsigned lt = lanemask_lt();
    // RANK_LEADER_ATOMIC: the highest lane of every match group adds the group's size to the run counter with ONE
    // returning shared atomic (predicated, no branch) and the group reads the old value from it by shuffle; the atomic
    // of row j+1 is issued before row j's result is consumed, so its latency hides behind a ranking.  Rows stay
    // ordered: same warp, program order, __syncwarp() between the rows.
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
    uint32_t d_nxt = 0, old_nxt = 0;
    unsigned mask_nxt = 0;
    if (MATCH & RANK_LEADER_ATOMIC)
    {
        d_nxt = digit_of(key[0], prmt_sel);
        mask_nxt = match_digit<MATCH>(d_nxt);
        old_nxt = leader_atomic(d_nxt, mask_nxt);
    }
#pragma unroll
    for (int j = 0; j < ITEMS; j++)
    {
        uint32_t r;
        if (MATCH & RANK_ATOMIC_ORDER)
        {
            // no match at all: every lane takes its slot with a returning shared atomic.  Stable only if the hardware serves
            // the lanes of ONE instruction that hit the same address in ascending lane order, which PTX leaves unspecified:
            // selected by pick_variant only after ranking_order_probe_kernel has verified it on the device.
            r = atomicAdd(&my_hist[digit_of(key[j], prmt_sel)], 1u);
            __syncwarp();   // row j's adds are performed before row j+1's (different lanes may hit the same counter)
        }
        else if (MATCH & RANK_LEADER_ATOMIC)
        {
            const unsigned mask = mask_nxt;
            const uint32_t old = old_nxt;
            if (j + 1 < ITEMS)
            {
                d_nxt = digit_of(key[j + 1], prmt_sel);
                mask_nxt = match_digit<MATCH>(d_nxt);
                __syncwarp();
                old_nxt = leader_atomic(d_nxt, mask_nxt);
            }
            const uint32_t prior = __shfl_sync(kFullMask, old, 31 - __clz(mask));
            r = prior + __popc(mask & lt);
        }
        else
        {
            const uint32_t d = digit_of(key[j], prmt_sel);
            const unsigned mask = match_digit<MATCH>(d);
            const uint32_t prior = my_hist[d];
            r = prior + __popc(mask & lt);
            __syncwarp();
            my_hist[d] = prior + __popc(mask);
            __syncwarp();
        }
        if (HAS_VALUES)
            reinterpret_cast<uint2*>(sm.kv)[r] = make_uint2(key[j], val[j]);
        else
            sm.kv[r] = key[j];
        constexpr int LB_EVERY = (MATCH & LB_STEP2) ? 2 : ((MATCH & LB_STEP8) ? 8 : ((MATCH & LB_STEP16) ? 16 : 4));
        if (INTERLEAVED && (j % LB_EVERY) == LB_EVERY - 1 && j + 1 < ITEMS) lb_try();
    }

    // 5. finish the look-back
    if (tid < kRadix)
    {
        if (tile > 0)
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
        if (!(MATCH & EARLY_TMA)) pass_base = ctl->hist[pass][tid];
        sm.digit_base[tid] = pass_base + exclusive - tile_off;
    }
    __syncthreads();

    if (full)
    {
#pragma unroll
        for (int j = 0; j < ITEMS; j++)
        {
            const uint32_t p = j * THREADS + tid;
            if (HAS_VALUES)
            {
                const uint2 e = reinterpret_cast<const uint2*>(sm.kv)[p];
                const uint32_t g = sm.digit_base[digit_of(e.x, prmt_sel)] + p;
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
                const uint32_t k = sm.kv[p];
                keys_out[sm.digit_base[digit_of(k, prmt_sel)] + p] = k;
            }
        }
    }
    else
    {
        for (uint32_t p = tid; p < valid; p += THREADS)
        {
            const uint32_t k = sm.kv[p * (HAS_VALUES ? 2 : 1)];
            const uint32_t g = sm.digit_base[digit_of(k, prmt_sel)] + p;
            if (LAYOUT == LAYOUT_AOS)
                reinterpret_cast<uint2*>(keys_out)[g] = make_uint2(k, sm.kv[p * 2 + 1]);
            else
            {
                keys_out[g] = k;
                if (HAS_VALUES) vals_out[g] = sm.kv[p * 2 + 1];
            }
        }
    }
}

// ---- 3b. persistent onesweep pass: CTAs loop over tiles, the next tile's keys are prefetched ---------------------
// Same algorithm as onesweep_pass_kernel (LAYOUT_KEYS / LAYOUT_SOA only), restructured so that no tile waits for its
// input: a CTA takes the ticket of its NEXT tile and starts the TMA bulk copy of that tile's keys into a second
// shared buffer while the current tile is being scanned, regrouped and written out; values are fetched at the start
// of a tile and land during its ranking.  Tickets keep the look-back deadlock-free whatever the residency: a tile is
// only ever waited on after a running CTA has taken it.
template <int THREADS, int ITEMS, int LAYOUT>
struct persistent_smem
{
    static constexpr int WARPS = THREADS / 32;
    static constexpr int TILE = THREADS * ITEMS;
    static constexpr int KV_WORDS = LAYOUT == LAYOUT_KEYS ? TILE : 2 * TILE;
    alignas(128) uint32_t kv[KV_WORDS];   // values staging (SOA: second half), then interleaved regroup area
    alignas(128) uint32_t pre[TILE];      // keys of the current tile (prefetched during the previous one)
    uint32_t warp_hist[WARPS][kRadix];
    uint32_t digit_base[kRadix];
    uint32_t scan_warp[kRadix / 32];
    alignas(8) uint64_t bar_keys;
    alignas(8) uint64_t bar_vals;
    uint32_t tile_next;
};

template <int THREADS, int ITEMS, int LAYOUT, int MATCH, int MIN_BLOCKS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
onesweep_persistent_kernel(const uint32_t* __restrict__ keys_in, uint32_t* __restrict__ keys_out,
                           const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ vals_out,
                           uint32_t n, int pass, sort_control* ctl, uint32_t* lookback, uint32_t num_tiles)
{
    using smem_t = persistent_smem<THREADS, ITEMS, LAYOUT>;
    constexpr int WARPS = smem_t::WARPS;
    constexpr int TILE = smem_t::TILE;
    constexpr bool HAS_VALUES = LAYOUT == LAYOUT_SOA;
    static_assert(LAYOUT != LAYOUT_AOS, "persistent kernel: keys or keys+values arrays");
    static_assert(THREADS >= kRadix, "one thread per digit needed");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    smem_t& sm = *reinterpret_cast<smem_t*>(smem_raw);

    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t prmt_sel = 0x4440u | (uint32_t) pass;
    const unsigned lt = lanemask_lt();
    uint32_t* my_hist = sm.warp_hist[warp];
    const uint32_t warp_off = warp * (ITEMS * 32) + lane;
    const uint32_t last_full_tiles = n / TILE;   // tiles [0, last_full_tiles) are full

    if (tid == 0)
    {
        mbar_init(&sm.bar_keys, 1);
        mbar_init(&sm.bar_vals, 1);
        mbar_fence_init();
        // static striding (TILE_BY_BLOCKIDX: needs every CTA of the grid co-resident, guaranteed by the launcher) keeps
        // consecutive tiles in lock-step; tickets are safe under any residency but start a tile one tile-time late
        const uint32_t t0 = (MATCH & TILE_BY_BLOCKIDX) ? blockIdx.x : atomicAdd(&ctl->tickets[pass], 1u);
        sm.tile_next = t0;
        if (t0 < last_full_tiles)
        {
            mbar_arrive_expect_tx(&sm.bar_keys, TILE * 4);
            bulk_copy_g2s(sm.pre, keys_in + (uint64_t) t0 * TILE, TILE * 4, &sm.bar_keys);
        }
    }
    __syncthreads();
    uint32_t kphase = 0, vphase = 0;
    // DEPHASE: statically strided CTAs start in lock-step, so the two CTAs of an SM would rank at the same time and write
    // at the same time; delaying every other CTA by about half a tile keeps one of them in its ALU phase while the
    // other is in its memory phase
    if ((MATCH & DEPHASE) && (blockIdx.x & 1)) __nanosleep(4000);

    while (true)
    {
        const uint32_t tile = sm.tile_next;
        if (tile >= num_tiles) break;
        const uint64_t tile_base = (uint64_t) tile * TILE;
        const uint32_t valid = (n - tile_base) < (uint64_t) TILE ? (uint32_t) (n - tile_base) : (uint32_t) TILE;
        const bool full = valid == (uint32_t) TILE;

        if (HAS_VALUES && full && tid == 0)
        {
            mbar_arrive_expect_tx(&sm.bar_vals, TILE * 4);
            bulk_copy_g2s(sm.kv + TILE, vals_in + tile_base, TILE * 4, &sm.bar_vals);
        }
#pragma unroll
        for (int i = lane; i < kRadix; i += 32) my_hist[i] = 0;
        if (full)
        {
            mbar_wait(&sm.bar_keys, kphase);
            kphase ^= 1;
        }
        else
        {
            for (uint32_t i = tid; i < (uint32_t) TILE; i += THREADS)
            {
                const bool in = i < valid;
                sm.pre[i] = in ? keys_in[tile_base + i] : 0xFFFFFFFFu;
                if (HAS_VALUES) sm.kv[TILE + i] = in ? vals_in[tile_base + i] : 0u;
            }
            __syncthreads();
        }
        __syncwarp();

        uint32_t key[ITEMS];
#pragma unroll
        for (int j = 0; j < ITEMS; j++) key[j] = sm.pre[warp_off + j * 32];

        uint32_t rank[ITEMS];
#pragma unroll
        for (int j = 0; j < ITEMS; j++)
        {
            const uint32_t d = digit_of(key[j], prmt_sel);
            const unsigned mask = match_digit<MATCH>(d);
            const uint32_t prior = my_hist[d];
            rank[j] = prior + __popc(mask & lt);
            __syncwarp();
            my_hist[d] = prior + __popc(mask);
            __syncwarp();
        }
        __syncthreads();   // S1: every warp has its keys in registers -> `pre` can take the next tile

        if (tid == 0)
        {
            const uint32_t tn = (MATCH & TILE_BY_BLOCKIDX) ? tile + gridDim.x : atomicAdd(&ctl->tickets[pass], 1u);
            sm.tile_next = tn;
            if (tn < last_full_tiles)
            {
                mbar_arrive_expect_tx(&sm.bar_keys, TILE * 4);
                bulk_copy_g2s(sm.pre, keys_in + (uint64_t) tn * TILE, TILE * 4, &sm.bar_keys);
            }
        }

        uint32_t cnt = 0, inc = 0, real_cnt = 0;
        uint32_t* lb = lookback + ((size_t) pass * num_tiles + tile) * kRadix;
        if (tid < kRadix)
        {
#pragma unroll
            for (int w = 0; w < WARPS; w++) cnt += sm.warp_hist[w][tid];
            real_cnt = cnt - ((tid == kRadix - 1) ? (uint32_t) TILE - valid : 0u);
            st_relaxed_u32(&lb[tid], (tile == 0 ? kLbFlagInclusive : kLbFlagAggregate) | real_cnt);
            inc = cnt;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1)
            {
                const uint32_t t = __shfl_up_sync(kFullMask, inc, s);
                if (lane >= (unsigned) s) inc += t;
            }
            if (lane == 31) sm.scan_warp[warp] = inc;
        }
        __syncthreads();   // S2
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
                const uint32_t c = sm.warp_hist[w][tid];
                sm.warp_hist[w][tid] = run;
                run += c;
            }
        }
        __syncthreads();   // S3

#pragma unroll
        for (int j = 0; j < ITEMS; j++) rank[j] += my_hist[digit_of(key[j], prmt_sel)];
        uint32_t val[HAS_VALUES ? ITEMS : 1];
        if (HAS_VALUES)
        {
            if (full)
            {
                mbar_wait(&sm.bar_vals, vphase);
                vphase ^= 1;
            }
#pragma unroll
            for (int j = 0; j < ITEMS; j++) val[j] = sm.kv[TILE + warp_off + j * 32];
        }
        __syncthreads();   // S4: staged values consumed, kv becomes the regroup area

#pragma unroll
        for (int j = 0; j < ITEMS; j++)
        {
            if (HAS_VALUES)
                reinterpret_cast<uint2*>(sm.kv)[rank[j]] = make_uint2(key[j], val[j]);
            else
                sm.kv[rank[j]] = key[j];
        }

        if (tid < kRadix)
        {
            uint32_t exclusive = 0;
            if (tile > 0)
            {
                constexpr int K = 4;
                const uint32_t* p = lb - kRadix + tid;
                int64_t t = (int64_t) tile - 1;
                bool done = false;
                while (!done)
                {
                    uint32_t s[K];
#pragma unroll
                    for (int k = 0; k < K; k++) s[k] = (t - k >= 0) ? ld_relaxed_u32(p - k * kRadix) : kLbFlagInclusive;
#pragma unroll
                    for (int k = 0; k < K; k++)
                    {
                        if (done) break;
                        while ((s[k] >> 30) == 0) s[k] = ld_relaxed_u32(p - k * kRadix);
                        exclusive += s[k] & kLbValueMask;
                        done = (s[k] >> 30) == 2;
                    }
                    t -= K;
                    p -= K * kRadix;
                }
                st_relaxed_u32(&lb[tid], kLbFlagInclusive | (exclusive + real_cnt));
            }
            sm.digit_base[tid] = ctl->hist[pass][tid] + exclusive - tile_off;
        }
        __syncthreads();   // S5

        if (full)
        {
#pragma unroll
            for (int j = 0; j < ITEMS; j++)
            {
                const uint32_t p = j * THREADS + tid;
                if (HAS_VALUES)
                {
                    const uint2 e = reinterpret_cast<const uint2*>(sm.kv)[p];
                    const uint32_t g = sm.digit_base[digit_of(e.x, prmt_sel)] + p;
                    keys_out[g] = e.x;
                    vals_out[g] = e.y;
                }
                else
                {
                    const uint32_t k = sm.kv[p];
                    keys_out[sm.digit_base[digit_of(k, prmt_sel)] + p] = k;
                }
            }
        }
        else
        {
            for (uint32_t p = tid; p < valid; p += THREADS)
            {
                const uint32_t k = sm.kv[p * (HAS_VALUES ? 2 : 1)];
                const uint32_t g = sm.digit_base[digit_of(k, prmt_sel)] + p;
                keys_out[g] = k;
                if (HAS_VALUES) vals_out[g] = sm.kv[p * 2 + 1];
            }
        }
        __syncthreads();   // S6: kv and warp_hist are free again, tile_next is visible
    }
}

// ---- host side ------------------------------------------------------------------------------------------------
using launch_fn = int (*)(cudaStream_t, const uint32_t*, uint32_t*, const uint32_t*, uint32_t*, uint32_t, int,
                          sort_control*, uint32_t*, uint32_t, int);

struct sort_variant
{
    const char* name;
    uint32_t tile;
    launch_fn launch;
    bool clears_next_pass;   // count-first kernels: a pass clears the look-back rows of its successor (kClearNextPassRow)
};

template <int THREADS, int ITEMS, int LAYOUT, int MATCH, int MIN_BLOCKS>
int launch_one(cudaStream_t s, const uint32_t* kin, uint32_t* kout, const uint32_t* vin, uint32_t* vout,
               uint32_t n, int pass, sort_control* ctl, uint32_t* lookback, uint32_t tiles)
{
    auto kern = onesweep_pass_kernel<THREADS, ITEMS, LAYOUT, MATCH, MIN_BLOCKS>;
    constexpr size_t smem = sizeof(onesweep_smem<THREADS, ITEMS, LAYOUT, (MATCH & P2P_DEST) != 0>);
    VRENB200_TRY(check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)));
    kern<<<tiles, THREADS, smem, s>>>(kin, kout, vin, vout, n, pass, ctl, lookback, tiles);
    return check_launch();
}

template <int THREADS, int ITEMS, int MATCH, int MIN_BLOCKS>
int launch_onesweep(cudaStream_t s, const uint32_t* kin, uint32_t* kout, const uint32_t* vin, uint32_t* vout,
                    uint32_t n, int pass, sort_control* ctl, uint32_t* lookback, uint32_t tiles, int layout)
{
    switch (layout)
    {
    case LAYOUT_KEYS: return launch_one<THREADS, ITEMS, LAYOUT_KEYS, MATCH, MIN_BLOCKS>(s, kin, kout, vin, vout, n, pass, ctl, lookback, tiles);
    case LAYOUT_SOA:  return launch_one<THREADS, ITEMS, LAYOUT_SOA, MATCH, MIN_BLOCKS>(s, kin, kout, vin, vout, n, pass, ctl, lookback, tiles);
    default:          return launch_one<THREADS, ITEMS, LAYOUT_AOS, MATCH, MIN_BLOCKS>(s, kin, kout, vin, vout, n, pass, ctl, lookback, tiles);
    }
}

template <int THREADS, int ITEMS, int MATCH, int MIN_BLOCKS>
int launch_persistent(cudaStream_t s, const uint32_t* kin, uint32_t* kout, const uint32_t* vin, uint32_t* vout,
                      uint32_t n, int pass, sort_control* ctl, uint32_t* lookback, uint32_t tiles, int layout)
{
    if (layout == LAYOUT_AOS)   // interleaved pairs (bucket sort) keep the one-tile-per-CTA kernel
        return launch_one<THREADS, ITEMS, LAYOUT_AOS, MATCH | TILE_BY_BLOCKIDX, MIN_BLOCKS>(s, kin, kout, vin, vout, n, pass, ctl, lookback, tiles);
    uint32_t grid = std::min<uint32_t>(tiles, (uint32_t) kNumSMs * MIN_BLOCKS);
    if (layout == LAYOUT_SOA)
    {
        auto kern = onesweep_persistent_kernel<THREADS, ITEMS, LAYOUT_SOA, MATCH, MIN_BLOCKS>;
        constexpr size_t smem = sizeof(persistent_smem<THREADS, ITEMS, LAYOUT_SOA>);
        VRENB200_TRY(check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)));
        if (MATCH & TILE_BY_BLOCKIDX)
        {
            // static striding spins on tiles of other resident CTAs: never launch more CTAs than fit at once
            int per_sm = 0, sms = 0, devid = 0;
            VRENB200_TRY(check_cuda(cudaGetDevice(&devid)));
            VRENB200_TRY(check_cuda(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, devid)));
            VRENB200_TRY(check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem)));
            if (per_sm < 1) return VRENB200_ELIMIT;
            grid = std::min<uint32_t>(tiles, (uint32_t) (per_sm * sms));
        }
        kern<<<grid, THREADS, smem, s>>>(kin, kout, vin, vout, n, pass, ctl, lookback, tiles);
    }
    else
    {
        auto kern = onesweep_persistent_kernel<THREADS, ITEMS, LAYOUT_KEYS, MATCH, MIN_BLOCKS>;
        constexpr size_t smem = sizeof(persistent_smem<THREADS, ITEMS, LAYOUT_KEYS>);
        VRENB200_TRY(check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)));
        kern<<<grid, THREADS, smem, s>>>(kin, kout, vin, vout, n, pass, ctl, lookback, tiles);
    }
    return check_launch();
}

template <int THREADS, int ITEMS, int LAYOUT, int MATCH, int MIN_BLOCKS>
int launch_count_first_one(cudaStream_t s, const uint32_t* kin, uint32_t* kout, const uint32_t* vin, uint32_t* vout,
                           uint32_t n, int pass, sort_control* ctl, uint32_t* lookback, uint32_t tiles)
{
    auto kern = onesweep_count_first_kernel<THREADS, ITEMS, LAYOUT, MATCH, MIN_BLOCKS>;
    constexpr size_t smem = sizeof(onesweep_smem<THREADS, ITEMS, LAYOUT, false>);
    VRENB200_TRY(check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)));
    kern<<<tiles, THREADS, smem, s>>>(kin, kout, vin, vout, n, pass, ctl, lookback, tiles);
    return check_launch();
}

template <int THREADS, int ITEMS, int MATCH, int MIN_BLOCKS>
int launch_count_first(cudaStream_t s, const uint32_t* kin, uint32_t* kout, const uint32_t* vin, uint32_t* vout,
                       uint32_t n, int pass, sort_control* ctl, uint32_t* lookback, uint32_t tiles, int layout)
{
    switch (layout)
    {
    case LAYOUT_KEYS: return launch_count_first_one<THREADS, ITEMS, LAYOUT_KEYS, MATCH, MIN_BLOCKS>(s, kin, kout, vin, vout, n, pass, ctl, lookback, tiles);
    case LAYOUT_SOA:  return launch_count_first_one<THREADS, ITEMS, LAYOUT_SOA, MATCH, MIN_BLOCKS>(s, kin, kout, vin, vout, n, pass, ctl, lookback, tiles);
    default:          return launch_count_first_one<THREADS, ITEMS, LAYOUT_AOS, MATCH, MIN_BLOCKS>(s, kin, kout, vin, vout, n, pass, ctl, lookback, tiles);
    }
}

#define VARIANT(T, I, M, B) { #T "x" #I "/" #M "/occ" #B, (T) * (I), launch_onesweep<T, I, M, B>, false }
#define PVARIANT(T, I, M, B) { #T "x" #I "/persistent/occ" #B, (T) * (I), launch_persistent<T, I, M, B>, false }
#define CVARIANT(T, I, M, B) { #T "x" #I "/count-first/" #M "/occ" #B, (T) * (I), launch_count_first<T, I, M, B>, true }
// retired entries: measured once (name and number appear in profiles/), no longer compiled; not selectable
#define RVARIANT(T, I, M, B) { #T "x" #I "/" #M "/occ" #B " [retired]", (T) * (I), nullptr, false }
#define RPVARIANT(T, I, M, B) { #T "x" #I "/persistent/occ" #B " [retired]", (T) * (I), nullptr, false }
#define RCVARIANT(T, I, M, B) { #T "x" #I "/count-first/" #M "/occ" #B " [retired]", (T) * (I), nullptr, false }
const sort_variant g_variants[] = {
    // 0: "automatic" (pick_variant); as a table entry, the ballot-match default for pairs: 11776-pair tiles (256 threads x
    // 46 rows, 2 CTAs = 16 warps per SM, 128 registers per thread: the per-tile steps are amortised over more pairs),
    // staging copies issued first, L2 prefetch for the successor CTA, leader-atomic ranking, interleaved look-back.
    // The atomic-order defaults are 79 (pairs), 80 (keys only) and 68 (below 2^21 elements).
    CVARIANT(256, 46, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_LEADER_ATOMIC, 2),
    RVARIANT(256, 32, TILE_BY_BLOCKIDX | SPLIT_KV, 2),
    PVARIANT(256, 32, TILE_BY_BLOCKIDX, 2),  // persistent CTAs + key prefetch, static tile striding
    PVARIANT(256, 32, MATCH_BALLOT, 2),      // persistent, tickets
    RVARIANT(512, 16, TILE_BY_BLOCKIDX | SPLIT_KV, 2),   // 32 warps/SM instead of 16: slower (profiles/r1k_*)
    VARIANT(256, 32, MATCH_BALLOT, 2),       // ticket instead of block index
    RVARIANT(256, 32, MATCH_BALLOT_C | TILE_BY_BLOCKIDX, 2),
    RVARIANT(256, 32, TILE_BY_BLOCKIDX | LEADER_ATOMIC, 2),
    RCVARIANT(256, 32, TILE_BY_BLOCKIDX, 2),  // 8
    VARIANT(256, 32, TILE_BY_BLOCKIDX, 2),   // 9: rank-then-count order, the default until r1m
    VARIANT(256, 32, TILE_BY_BLOCKIDX | FAKE_LOOKBACK, 2),  // 10: ceiling without the look-back chain (wrong results)
    RCVARIANT(256, 32, TILE_BY_BLOCKIDX | DIRECT_LOAD, 2),   // 11
    RCVARIANT(256, 24, TILE_BY_BLOCKIDX | DIRECT_LOAD, 3),   // 12
    RCVARIANT(256, 24, TILE_BY_BLOCKIDX | LB_INTERLEAVED, 3),   // 13
    CVARIANT(256, 24, TILE_BY_BLOCKIDX, 3),   // 14: default of r1m
    RCVARIANT(256, 32, TILE_BY_BLOCKIDX | LB_INTERLEAVED | LB_STEP2, 2),   // 15
    RCVARIANT(256, 32, TILE_BY_BLOCKIDX | LB_INTERLEAVED | LB_STEP8, 2),   // 16
    RCVARIANT(256, 32, TILE_BY_BLOCKIDX | LB_INTERLEAVED | PREFETCH_L2, 2),   // 17
    RCVARIANT(256, 32, TILE_BY_BLOCKIDX | LB_INTERLEAVED | RANK_LEADER_ATOMIC, 2),   // 18
    RCVARIANT(256, 32, TILE_BY_BLOCKIDX | LB_INTERLEAVED | PREFETCH_L2 | EARLY_TMA, 2),   // 19
    RCVARIANT(256, 32, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA, 2),   // 20
    RCVARIANT(256, 24, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2, 3),   // 21
    RCVARIANT(256, 32, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_LEADER_ATOMIC, 2),   // 22
    RCVARIANT(256, 24, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_LEADER_ATOMIC, 3),   // 23
    RCVARIANT(384, 24, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2, 2),   // 24
    CVARIANT(256, 32, TILE_BY_BLOCKIDX | LB_INTERLEAVED, 2),   // 25: default until r1w
    RCVARIANT(512, 16, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_LEADER_ATOMIC, 2),   // 26
    RCVARIANT(320, 32, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_LEADER_ATOMIC, 2),   // 27
    CVARIANT(384, 24, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_LEADER_ATOMIC | VALS_DIRECT, 2),   // 28
    RCVARIANT(384, 24, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_LEADER_ATOMIC | REG_COUNTS, 2),   // 29
    CVARIANT(384, 24, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_LEADER_ATOMIC | KEYS_CHUNKED, 2),   // 30
    RCVARIANT(384, 24, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_LEADER_ATOMIC | VALS_DIRECT | KEYS_CHUNKED, 2),   // 31
    RCVARIANT(384, 24, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_LEADER_ATOMIC | VALS_DIRECT | KEYS_CHUNKED | REG_COUNTS, 2),   // 32
    RCVARIANT(256, 24, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_LEADER_ATOMIC | VALS_DIRECT | KEYS_CHUNKED, 3),   // 33
    CVARIANT(384, 24, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_LEADER_ATOMIC | DEPHASE, 2),   // 34
    RCVARIANT(256, 24, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_LEADER_ATOMIC | DEPHASE, 3),   // 35
    RCVARIANT(256, 36, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_LEADER_ATOMIC, 2),   // 36: 16 warps/SM, more rows per thread
    RCVARIANT(256, 40, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_LEADER_ATOMIC, 2),   // 37
    CVARIANT(256, 44, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_LEADER_ATOMIC, 2),   // 38
    RCVARIANT(320, 28, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_LEADER_ATOMIC, 2),   // 39
    RCVARIANT(512, 40, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_LEADER_ATOMIC, 1),   // 40: one CTA per SM, 20480-pair tiles
    RCVARIANT(512, 44, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_LEADER_ATOMIC, 1),   // 41
    CVARIANT(384, 24, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_LEADER_ATOMIC, 2),   // 42: default of r1w
    RCVARIANT(288, 40, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_LEADER_ATOMIC, 2),   // 43
    RCVARIANT(256, 44, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_LEADER_ATOMIC | LB_STEP8, 2),   // 44
    RCVARIANT(256, 44, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2, 2),   // 45: without the leader atomic
    RCVARIANT(256, 48, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_LEADER_ATOMIC, 2),   // 46: the largest tile two CTAs fit
    CVARIANT(256, 16, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | RANK_LEADER_ATOMIC, 3),   // 47: 4096-pair tiles for mid-size inputs
    CVARIANT(256, 46, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_LEADER_ATOMIC | MATCH_SPLIT4, 2),   // 48
    RCVARIANT(256, 44, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_LEADER_ATOMIC | MATCH_SPLIT4, 2),   // 49
    RCVARIANT(256, 72, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_LEADER_ATOMIC, 2),   // 50: keys-only tiles (4 B/key of shared memory)
    RCVARIANT(256, 80, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_LEADER_ATOMIC, 2),   // 51
    RCVARIANT(256, 88, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_LEADER_ATOMIC, 2),   // 52
    RCVARIANT(384, 40, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_LEADER_ATOMIC, 2),   // 53: keys only, 24 warps/SM
    RCVARIANT(256, 56, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_LEADER_ATOMIC, 2),   // 54
    CVARIANT(256, 64, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_LEADER_ATOMIC, 2),   // 55
    CVARIANT(256, 46, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_ATOMIC_ORDER, 2),   // 56: first atomic-order measurement
    RCVARIANT(256, 48, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_ATOMIC_ORDER, 2),   // 57
    RCVARIANT(256, 50, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_ATOMIC_ORDER, 2),   // 58
    RCVARIANT(384, 28, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_ATOMIC_ORDER, 2),   // 59
    RCVARIANT(384, 30, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_ATOMIC_ORDER, 2),   // 60
    RCVARIANT(384, 32, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_ATOMIC_ORDER, 2),   // 61
    RCVARIANT(512, 24, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | PREFETCH_L2 | RANK_ATOMIC_ORDER, 2),   // 62
    CVARIANT(256, 46, TILE_BY_BLOCKIDX | LB_INTERLEAVED | LB_STEP8 | EARLY_TMA | PREFETCH_L2 | RANK_ATOMIC_ORDER, 2),   // 63
    RCVARIANT(256, 46, TILE_BY_BLOCKIDX | LB_INTERLEAVED | LB_STEP16 | EARLY_TMA | PREFETCH_L2 | RANK_ATOMIC_ORDER, 2),   // 64
    CVARIANT(256, 46, TILE_BY_BLOCKIDX | EARLY_TMA | PREFETCH_L2 | RANK_ATOMIC_ORDER, 2),   // 65: look-back after the ranking
    RCVARIANT(384, 28, TILE_BY_BLOCKIDX | LB_INTERLEAVED | LB_STEP8 | EARLY_TMA | PREFETCH_L2 | RANK_ATOMIC_ORDER, 2),   // 66
    CVARIANT(256, 48, TILE_BY_BLOCKIDX | LB_INTERLEAVED | LB_STEP8 | EARLY_TMA | PREFETCH_L2 | RANK_ATOMIC_ORDER, 2),   // 67
    CVARIANT(256, 16, TILE_BY_BLOCKIDX | LB_INTERLEAVED | EARLY_TMA | RANK_ATOMIC_ORDER, 3),   // 68: 4096-pair tiles
    CVARIANT(256, 64, TILE_BY_BLOCKIDX | LB_INTERLEAVED | LB_STEP8 | EARLY_TMA | PREFETCH_L2 | RANK_ATOMIC_ORDER, 2),   // 69: keys-only tiles
    RCVARIANT(256, 80, TILE_BY_BLOCKIDX | LB_INTERLEAVED | LB_STEP8 | EARLY_TMA | PREFETCH_L2 | RANK_ATOMIC_ORDER, 2),   // 70
    RCVARIANT(256, 96, TILE_BY_BLOCKIDX | LB_INTERLEAVED | LB_STEP8 | EARLY_TMA | PREFETCH_L2 | RANK_ATOMIC_ORDER, 2),   // 71
    RCVARIANT(384, 56, TILE_BY_BLOCKIDX | LB_INTERLEAVED | LB_STEP8 | EARLY_TMA | PREFETCH_L2 | RANK_ATOMIC_ORDER, 2),   // 72: keys only, 24 warps
    RCVARIANT(256, 16, TILE_BY_BLOCKIDX | LB_INTERLEAVED | LB_STEP8 | EARLY_TMA | RANK_ATOMIC_ORDER, 4),   // 73
    RCVARIANT(256, 48, TILE_BY_BLOCKIDX | LB_INTERLEAVED | LB_STEP8 | EARLY_TMA | PREFETCH_L2 | RANK_ATOMIC_ORDER | REG_COUNTS, 2),   // 74
    RCVARIANT(256, 48, TILE_BY_BLOCKIDX | LB_INTERLEAVED | LB_STEP8 | EARLY_TMA | PREFETCH_L2 | RANK_ATOMIC_ORDER | KEYS_CHUNKED, 2),   // 75
    RCVARIANT(256, 50, TILE_BY_BLOCKIDX | LB_INTERLEAVED | LB_STEP8 | EARLY_TMA | PREFETCH_L2 | RANK_ATOMIC_ORDER, 2),   // 76
    RCVARIANT(384, 32, TILE_BY_BLOCKIDX | LB_INTERLEAVED | LB_STEP8 | EARLY_TMA | PREFETCH_L2 | RANK_ATOMIC_ORDER, 2),   // 77
    RCVARIANT(256, 48, TILE_BY_BLOCKIDX | LB_INTERLEAVED | LB_STEP8 | EARLY_TMA | PREFETCH_L2 | RANK_ATOMIC_ORDER | VALS_DIRECT | VALS_LATE, 2),   // 78
    CVARIANT(256, 48, TILE_BY_BLOCKIDX | LB_INTERLEAVED | LB_STEP8 | EARLY_TMA | PREFETCH_L2 | RANK_ATOMIC_ORDER | REG_COUNTS | KEYS_CHUNKED, 2),   // 79
    CVARIANT(256, 64, TILE_BY_BLOCKIDX | LB_INTERLEAVED | LB_STEP8 | EARLY_TMA | PREFETCH_L2 | RANK_ATOMIC_ORDER | REG_COUNTS | KEYS_CHUNKED, 2),   // 80: keys only
    RCVARIANT(256, 48, TILE_BY_BLOCKIDX | LB_INTERLEAVED | LB_STEP16 | EARLY_TMA | PREFETCH_L2 | RANK_ATOMIC_ORDER | REG_COUNTS | KEYS_CHUNKED, 2),   // 81
    RCVARIANT(320, 36, TILE_BY_BLOCKIDX | LB_INTERLEAVED | LB_STEP8 | EARLY_TMA | PREFETCH_L2 | RANK_ATOMIC_ORDER | KEYS_CHUNKED, 2),   // 82
    RCVARIANT(320, 38, TILE_BY_BLOCKIDX | LB_INTERLEAVED | LB_STEP8 | EARLY_TMA | PREFETCH_L2 | RANK_ATOMIC_ORDER | KEYS_CHUNKED, 2),   // 83
    RCVARIANT(256, 28, TILE_BY_BLOCKIDX | LB_INTERLEAVED | LB_STEP8 | EARLY_TMA | PREFETCH_L2 | RANK_ATOMIC_ORDER | KEYS_CHUNKED, 3),   // 84
    RCVARIANT(256, 32, TILE_BY_BLOCKIDX | LB_INTERLEAVED | LB_STEP8 | EARLY_TMA | PREFETCH_L2 | RANK_ATOMIC_ORDER | KEYS_CHUNKED, 3),   // 85
};
constexpr int kNumVariants = sizeof(g_variants) / sizeof(g_variants[0]);
// the scratch layout must not depend on the variant: size the look-back for the smallest tile
constexpr uint32_t kMinTile = 256 * 16;

int g_variant = 0;
// Variant 0 means "automatic" (pick_variant below):
//  * inputs below 2^21 pairs take the 4096-pair tile (more CTAs than SMs from 2^19 on, shorter per-tile steps: 44 vs 63 us
//    at 2^16 pairs, 64.5 vs 71.6 us at 2^20, equal at 2^21-2^22, profiles/r1x_sort_size_variants.log, r1z_*);
//  * keys-only sorts stage 4 B per key, so 64 rows per thread fit the same shared memory and registers (16 384-key tiles);
//  * the ranking step is RANK_ATOMIC_ORDER where the device passes the probe below, else the ballot match.
constexpr uint32_t kSmallTileBelow = 1u << 21;
constexpr int kMatchSmall = 47, kMatchKeys = 55, kMatchPairs = 0;          // ballot-match ranking (order by construction)
constexpr int kAtomicSmall = 68, kAtomicKeys = 80, kAtomicPairs = 79;      // RANK_ATOMIC_ORDER ranking
static_assert(kNumVariants > 80, "variant table changed");

// ---- ranking mode --------------------------------------------------------------------------------------------------
// RANK_ATOMIC_ORDER kernels take every pair's slot with ONE returning shared atomic per lane and no match at all
// (3.72 vs 4.34 ms per 2^28-pair sort).  The sort stays stable only if the lanes of one warp instruction that hit the
// SAME shared address are served in ascending lane order.  PTX leaves that order unspecified; the B200 does it that
// way, and the library does not take it on trust: the first sort on a device runs ranking_order_probe_kernel (every SM,
// 16 warps per SM, ~4.7 M warp instructions over collision patterns from "all 32 lanes on one counter" to "256 random
// counters", checked against the ballot match) and falls back to the ballot-match kernels if a single lane disagrees.
// vrenb200_radix_sort_set_ranking / VRENB200_SORT_RANKING=match|atomic|auto override the choice.
enum { RANKING_AUTO = 0, RANKING_MATCH = 1, RANKING_ATOMIC_ORDER = 2 };
int g_ranking_mode = -1;                    // -1: not read from the environment yet
constexpr int kMaxDevices = 64;
int g_probe_state[kMaxDevices] = {};        // 0 unknown, 1 passed, 2 failed
std::mutex g_probe_mutex;
__device__ uint32_t g_probe_failures;

__global__ void __launch_bounds__(256) ranking_order_probe_kernel(uint32_t rounds)
{
    __shared__ uint32_t cnt[8][kRadix];
    const unsigned tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 8 * kRadix; i += 256) (&cnt[0][0])[i] = 0;
    __syncthreads();
    const unsigned lt = lanemask_lt();
    uint32_t bad = 0;
    uint32_t h = (blockIdx.x * 256u + tid) * 2654435761u + 12345u;
    for (uint32_t r = 0; r < rounds; r++)
    {
        h ^= h << 13; h ^= h >> 17; h ^= h << 5;                       // xorshift32 per lane
        const uint32_t distinct = 1u << ((r % 6u) * 2u > 8u ? 8u : (r % 6u) * 2u);   // 1, 4, 16, 64, 256, 256 counters in play
        uint32_t d = (h >> 7) & (distinct - 1u);
        if (r & 8u) d = (d * 32u + (d >> 3)) & 255u;                    // same-bank / different-address collisions as well
        const uint32_t prior = cnt[warp][d];
        __syncwarp();
        const uint32_t old = atomicAdd(&cnt[warp][d], 1u);
        __syncwarp();
        const unsigned mask = __match_any_sync(kFullMask, d);
        bad += old != prior + (uint32_t) __popc(mask & lt);
    }
    if (bad) atomicAdd(&g_probe_failures, bad);
}

// runs the probe once per device, on its own stream; never while the caller's stream is being captured
bool atomic_order_ranking_ok(cudaStream_t user_stream)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return false;
    std::lock_guard<std::mutex> lock(g_probe_mutex);
    if (g_probe_state[dev]) return g_probe_state[dev] == 1;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(user_stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone)
    {
        cudaGetLastError();
        return false;   // not cached: the next call outside a capture probes
    }
    cudaStream_t s = nullptr;
    uint32_t failures = 0xFFFFFFFFu, zero = 0;
    bool ran = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) == cudaSuccess;
    ran = ran && cudaMemcpyToSymbolAsync(g_probe_failures, &zero, sizeof(zero), 0, cudaMemcpyHostToDevice, s) == cudaSuccess;
    if (ran)
    {
        ranking_order_probe_kernel<<<kNumSMs * 2, 256, 0, s>>>(2000);
        ran = cudaGetLastError() == cudaSuccess;
    }
    ran = ran && cudaMemcpyFromSymbolAsync(&failures, g_probe_failures, sizeof(failures), 0, cudaMemcpyDeviceToHost, s) == cudaSuccess;
    ran = ran && cudaStreamSynchronize(s) == cudaSuccess;
    if (s) cudaStreamDestroy(s);
    if (!ran) cudaGetLastError();
    g_probe_state[dev] = (ran && failures == 0) ? 1 : 2;
    return g_probe_state[dev] == 1;
}

int ranking_mode()
{
    if (g_ranking_mode < 0)
    {
        const char* e = std::getenv("VRENB200_SORT_RANKING");
        g_ranking_mode = (e && std::strcmp(e, "match") == 0) ? RANKING_MATCH : ((e && std::strcmp(e, "atomic") == 0) ? RANKING_ATOMIC_ORDER : RANKING_AUTO);
    }
    return g_ranking_mode;
}

const sort_variant& pick_variant(uint32_t n, int layout, cudaStream_t s)
{
    if (g_variant != 0) return g_variants[g_variant];
    const int mode = ranking_mode();
    const bool atomic = mode == RANKING_ATOMIC_ORDER || (mode == RANKING_AUTO && atomic_order_ranking_ok(s));
    if (n < kSmallTileBelow) return g_variants[atomic ? kAtomicSmall : kMatchSmall];
    if (layout == LAYOUT_KEYS) return g_variants[atomic ? kAtomicKeys : kMatchKeys];
    return g_variants[atomic ? kAtomicPairs : kMatchPairs];
}
uint32_t g_bucket_search_min = 1u << 20;   // bucket sort: from this many pairs on, END offsets come from a search in the sorted output
int g_partition_shape = 0;   // 0: 256x32 (2 CTAs/SM), 1: 256x16 (4 CTAs/SM), 2: 512x16 (2 CTAs/SM)

size_t lookback_words(uint32_t n)
{
    const size_t tiles = ((size_t) n + kMinTile - 1) / kMinTile;
    return (size_t) kPasses * (tiles > 0 ? tiles : 1) * kRadix;
}

size_t control_bytes(uint32_t n)
{
    return align_up(sizeof(sort_control) + lookback_words(n) * sizeof(uint32_t), 256);
}

} // namespace
} // namespace vrenb200

// per-kernel device timing for bench.py's roofline: events recorded on the launching stream between launches
struct vrenb200_sort_profile
{
    cudaEvent_t ev[vrenb200::kPasses + 3]; // start | after hist | after scan | after pass 0..3
};

namespace vrenb200 {
namespace {

int launch_histogram(cudaStream_t s, const uint32_t* keys, uint32_t n, sort_control* ctl)
{
    if (n >= (1u << 20))
    {
        static bool configured = false;
        if (!configured)
        {
            VRENB200_TRY(check_cuda(cudaFuncSetAttribute(radix_histogram_columns_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kHist2Smem)));
            configured = true;
        }
        radix_histogram_columns_kernel<<<kNumSMs, kHist2Threads, kHist2Smem, s>>>(keys, n, ctl);
    }
    else
        radix_histogram_kernel<<<kNumSMs * 4, kHistThreads, 0, s>>>(keys, n, ctl);
    return check_launch();
}

// control + look-back live in `ctl_mem`; alt buffers given explicitly
int radix_sort_impl(cudaStream_t s, uint32_t* keys, uint32_t* vals, uint32_t n, uint32_t* alt_keys, uint32_t* alt_vals,
                    void* ctl_mem, vrenb200_sort_profile* prof = nullptr, int first_pass = 0, int num_passes = kPasses)
{
    if (n == 0) return VRENB200_OK;
    if (n >= (1u << 30)) return VRENB200_ELIMIT;
    if ((reinterpret_cast<uintptr_t>(keys) | reinterpret_cast<uintptr_t>(alt_keys) |
         reinterpret_cast<uintptr_t>(vals) | reinterpret_cast<uintptr_t>(alt_vals) |
         reinterpret_cast<uintptr_t>(ctl_mem)) & 15)
        return VRENB200_EALIGN;
    const int layout = vals != nullptr ? LAYOUT_SOA : LAYOUT_KEYS;
    const sort_variant& var = pick_variant(n, layout, s);
    const uint32_t tiles = (uint32_t) (((size_t) n + var.tile - 1) / var.tile);
    sort_control* ctl = static_cast<sort_control*>(ctl_mem);
    uint32_t* lookback = reinterpret_cast<uint32_t*>(ctl + 1);
    // only the part of the look-back this variant touches needs clearing; count-first passes clear their successor's rows
    if (var.clears_next_pass)
    {
        VRENB200_TRY(check_cuda(cudaMemsetAsync(ctl_mem, 0, sizeof(sort_control), s)));
        VRENB200_TRY(check_cuda(cudaMemsetAsync(lookback + (size_t) first_pass * tiles * kRadix, 0, (size_t) tiles * kRadix * sizeof(uint32_t), s)));
    }
    else
    {
        const size_t clear = sizeof(sort_control) + (size_t) kPasses * tiles * kRadix * sizeof(uint32_t);
        VRENB200_TRY(check_cuda(cudaMemsetAsync(ctl_mem, 0, clear, s)));
    }
    if (prof) cudaEventRecord(prof->ev[0], s);
    VRENB200_TRY(launch_histogram(s, keys, n, ctl));
    if (prof) cudaEventRecord(prof->ev[1], s);
    radix_scan_histograms_kernel<<<kPasses, kRadix, 0, s>>>(ctl);
    VRENB200_TRY(check_launch());
    if (prof) cudaEventRecord(prof->ev[2], s);
    // ping-pong: the i-th executed pass reads (keys, vals) when i is even; with 4 passes the result is back in `keys`
    for (int i = 0; i < num_passes; i++)
    {
        const int pass = first_pass + i;
        const bool even = (i & 1) == 0;
        const int pass_arg = pass | ((var.clears_next_pass && i + 1 < num_passes) ? kClearNextPassRow : 0);
        VRENB200_TRY(var.launch(s, even ? keys : alt_keys, even ? alt_keys : keys, even ? vals : alt_vals,
                                even ? alt_vals : vals, n, pass_arg, ctl, lookback, tiles, layout));
        if (prof && num_passes == kPasses) cudaEventRecord(prof->ev[3 + pass], s);
    }
    return VRENB200_OK;
}

} // namespace
} // namespace vrenb200

using namespace vrenb200;

// tuning hook (not part of the reference surface): select the kernel configuration used by subsequent calls
extern "C" int vrenb200_radix_sort_set_variant(int v)
{
    if (v < 0 || v >= kNumVariants || g_variants[v].launch == nullptr) return VRENB200_EINVAL_ARG;
    g_variant = v;
    return VRENB200_OK;
}
// ranking step of the default kernels: 0 = automatic (probe the device once), 1 = ballot match, 2 = atomic order
extern "C" int vrenb200_radix_sort_set_ranking(int mode)
{
    if (mode < RANKING_AUTO || mode > RANKING_ATOMIC_ORDER) return VRENB200_EINVAL_ARG;
    g_ranking_mode = mode;
    return VRENB200_OK;
}
// 1 if the current device serves same-address lanes of a shared atomic in ascending lane order (runs the probe if needed)
extern "C" int vrenb200_radix_sort_ranking_probe(void) { return atomic_order_ranking_ok(nullptr) ? 1 : 0; }
// name of the pass configuration a sort of n elements would use now (reporting)
extern "C" const char* vrenb200_radix_sort_selected_variant_name(uint32_t n, int with_values)
{
    return pick_variant(n, with_values ? LAYOUT_SOA : LAYOUT_KEYS, nullptr).name;
}
extern "C" int vrenb200_radix_sort_set_dephase(uint32_t ns, uint32_t rule)
{
    VRENB200_TRY(check_cuda(cudaMemcpyToSymbol(g_dephase_ns, &ns, sizeof(ns))));
    return check_cuda(cudaMemcpyToSymbol(g_dephase_rule, &rule, sizeof(rule)));
}

extern "C" int vrenb200_radix_sort_set_hist_loads(uint32_t loads)
{
    return check_cuda(cudaMemcpyToSymbol(g_hist_loads_in_flight, &loads, sizeof(loads)));
}

extern "C" int vrenb200_radix_sort_set_prefetch_tiles(uint32_t tiles)
{
    return check_cuda(cudaMemcpyToSymbol(g_prefetch_tiles, &tiles, sizeof(tiles)));
}

extern "C" int vrenb200_radix_sort_num_variants(void) { return kNumVariants; }
extern "C" const char* vrenb200_radix_sort_variant_name(int v)
{
    return (v < 0 || v >= kNumVariants) ? "" : g_variants[v].name;
}

extern "C" size_t vrenb200_radix_sort_scratch_bytes(uint32_t n, int with_values)
{
    const size_t alt = align_up((size_t) n * 4, 256);
    return alt * (with_values ? 2 : 1) + control_bytes(n);
}

extern "C" int vrenb200_radix_sort_keys(vrenb200_stream_t stream, uint32_t* keys, uint32_t n,
                                        void* scratch, size_t scratch_bytes)
{
    if (n == 0) return VRENB200_OK;
    if (keys == nullptr) return VRENB200_EINVAL_ARG;
    if (scratch == nullptr || scratch_bytes < vrenb200_radix_sort_scratch_bytes(n, 0)) return VRENB200_ESCRATCH;
    char* p = static_cast<char*>(scratch);
    const size_t alt = align_up((size_t) n * 4, 256);
    return radix_sort_impl(as_stream(stream), keys, nullptr, n, reinterpret_cast<uint32_t*>(p), nullptr, p + alt);
}

extern "C" int vrenb200_radix_sort_pairs(vrenb200_stream_t stream, uint32_t* keys, uint32_t* values, uint32_t n,
                                         void* scratch, size_t scratch_bytes)
{
    if (n == 0) return VRENB200_OK;
    if (keys == nullptr || values == nullptr) return VRENB200_EINVAL_ARG;
    if (scratch == nullptr || scratch_bytes < vrenb200_radix_sort_scratch_bytes(n, 1)) return VRENB200_ESCRATCH;
    char* p = static_cast<char*>(scratch);
    const size_t alt = align_up((size_t) n * 4, 256);
    return radix_sort_impl(as_stream(stream), keys, values, n, reinterpret_cast<uint32_t*>(p),
                           reinterpret_cast<uint32_t*>(p + alt), p + 2 * alt);
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
extern "C" int vrenb200_radix_sort_pairs_profiled(vrenb200_stream_t stream, uint32_t* keys, uint32_t* values, uint32_t n,
                                                  void* scratch, size_t scratch_bytes, vrenb200_sort_profile* prof)
{
    if (n == 0) return VRENB200_OK;
    if (keys == nullptr) return VRENB200_EINVAL_ARG;
    const int kv = values != nullptr;
    if (scratch == nullptr || scratch_bytes < vrenb200_radix_sort_scratch_bytes(n, kv)) return VRENB200_ESCRATCH;
    char* p = static_cast<char*>(scratch);
    const size_t alt = align_up((size_t) n * 4, 256);
    return radix_sort_impl(as_stream(stream), keys, values, n, reinterpret_cast<uint32_t*>(p),
                           kv ? reinterpret_cast<uint32_t*>(p + alt) : nullptr, p + (kv ? 2 : 1) * alt, prof);
}

// ---- building blocks of the multi-GPU sort (vren_b200/dist.py): digit histograms and a digit-range sort --------
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
    return launch_histogram(s, keys, n, fake);
}

extern "C" size_t vrenb200_radix_sort_range_scratch_bytes(uint32_t n) { return control_bytes(n); }

// 256-bin histogram of the most significant byte only (one shared atomic per key instead of four)
__global__ void __launch_bounds__(kHistThreads)
radix_top_digit_histogram_kernel(const uint32_t* __restrict__ keys, uint32_t n, uint32_t* hist_out)
{
    __shared__ uint32_t s_hist[kRadix];
    if (threadIdx.x < kRadix) s_hist[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t n4 = n / 4;
    const uint4* keys4 = reinterpret_cast<const uint4*>(keys);
    for (uint32_t i = blockIdx.x * kHistThreads + threadIdx.x; i < n4; i += gridDim.x * kHistThreads)
    {
        const uint4 a = ldg_stream_u4(keys4 + i);
        atomicAdd(&s_hist[a.x >> 24], 1u); atomicAdd(&s_hist[a.y >> 24], 1u);
        atomicAdd(&s_hist[a.z >> 24], 1u); atomicAdd(&s_hist[a.w >> 24], 1u);
    }
    if (blockIdx.x == 0)
        for (uint32_t t = n4 * 4 + threadIdx.x; t < n; t += kHistThreads) atomicAdd(&s_hist[keys[t] >> 24], 1u);
    __syncthreads();
    if (threadIdx.x < kRadix && s_hist[threadIdx.x] != 0) atomicAdd(&hist_out[threadIdx.x], s_hist[threadIdx.x]);
}

extern "C" int vrenb200_radix_top_digit_histogram(vrenb200_stream_t stream, const uint32_t* keys, uint32_t n, uint32_t* hist_out)
{
    if (hist_out == nullptr || (n > 0 && keys == nullptr)) return VRENB200_EINVAL_ARG;
    if (reinterpret_cast<uintptr_t>(keys) & 15) return VRENB200_EALIGN;
    cudaStream_t s = as_stream(stream);
    VRENB200_TRY(check_cuda(cudaMemsetAsync(hist_out, 0, sizeof(uint32_t) * kRadix, s)));
    if (n == 0) return VRENB200_OK;
    radix_top_digit_histogram_kernel<<<kNumSMs * 4, kHistThreads, 0, s>>>(keys, n, hist_out);
    return check_launch();
}

// Fused partition + exchange of the multi-GPU sort: one onesweep pass on the most significant byte whose write-out
// stores every pair straight into its destination rank's receive buffer (plain st.global on peer-mapped pointers over
// NVLink, or local memory for the rank's own range).  The pass partitions by DESTINATION RANK (a lookup on the most
// significant byte), so every tile emits one long contiguous run per destination.  dest_table: device struct
// { uint64 kptr[32]; uint64 vptr[32]; uint8 rank_of[256]; } — where THIS rank's block starts in every destination's
// key / value buffer.  No histogram kernel is needed: the offsets come from the all-gathered histograms (dist.py).
extern "C" int vrenb200_radix_partition_scatter(vrenb200_stream_t stream, const uint32_t* keys, const uint32_t* values, uint32_t n,
                                                const uint64_t* dest_table, void* scratch, size_t scratch_bytes)
{
    if (n == 0) return VRENB200_OK;
    if (keys == nullptr || values == nullptr || dest_table == nullptr) return VRENB200_EINVAL_ARG;
    if (n >= (1u << 30)) return VRENB200_ELIMIT;
    if (scratch == nullptr || scratch_bytes < control_bytes(n)) return VRENB200_ESCRATCH;
    if ((reinterpret_cast<uintptr_t>(keys) | reinterpret_cast<uintptr_t>(values) | reinterpret_cast<uintptr_t>(scratch)) & 15) return VRENB200_EALIGN;
    cudaStream_t s = as_stream(stream);
    // tile shape of the exchange pass (tuning hook): remote stores back-pressure the CTAs, so more, smaller CTAs per SM
    // keep more loads in flight while some CTAs drain into NVLink
    const int shape = g_partition_shape;
    const uint32_t tile = shape == 1 ? 256 * 16 : (shape == 2 ? 512 * 16 : 256 * 32);
    const uint32_t tiles = (uint32_t) (((size_t) n + tile - 1) / tile);
    sort_control* ctl = static_cast<sort_control*>(scratch);
    uint32_t* lookback = reinterpret_cast<uint32_t*>(ctl + 1);
    // the exchange pass runs as "pass kPasses - 1": only that region of the look-back words is touched
    VRENB200_TRY(check_cuda(cudaMemsetAsync(scratch, 0, sizeof(sort_control), s)));
    VRENB200_TRY(check_cuda(cudaMemsetAsync(lookback + (size_t) (kPasses - 1) * tiles * kRadix, 0, (size_t) tiles * kRadix * sizeof(uint32_t), s)));
    uint32_t* table = reinterpret_cast<uint32_t*>(const_cast<uint64_t*>(dest_table));
    if (shape == 1)
        return launch_one<256, 16, LAYOUT_SOA, TILE_BY_BLOCKIDX | P2P_DEST, 4>(s, keys, table, values, nullptr, n, kPasses - 1, ctl, lookback, tiles);
    if (shape == 2)
        return launch_one<512, 16, LAYOUT_SOA, TILE_BY_BLOCKIDX | P2P_DEST, 2>(s, keys, table, values, nullptr, n, kPasses - 1, ctl, lookback, tiles);
    return launch_one<256, 32, LAYOUT_SOA, TILE_BY_BLOCKIDX | P2P_DEST, 2>(s, keys, table, values, nullptr, n, kPasses - 1, ctl, lookback, tiles);
}

extern "C" int vrenb200_radix_partition_set_shape(int shape)
{
    if (shape < 0 || shape > 2) return VRENB200_EINVAL_ARG;
    g_partition_shape = shape;
    return VRENB200_OK;
}

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
    return radix_sort_impl(as_stream(stream), keys, values, n, alt_keys, alt_values, scratch, nullptr, first_pass, num_passes);
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
    return radix_sort_impl(as_stream(stream), keys, nullptr, n, static_cast<uint32_t*>(scratch_2), nullptr, scratch_1);
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
    int st = kv ? vrenb200_radix_sort_pairs(stream, dk, dv, n, scratch, scratch_bytes)
                : vrenb200_radix_sort_keys(stream, dk, n, scratch, scratch_bytes);
    if (st != VRENB200_OK) return st;
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
    if (out == nullptr || (n > 0 && in_pairs == nullptr)) return VRENB200_EINVAL_ARG;
    if (n >= (1u << 30)) return VRENB200_ELIMIT;
    if ((reinterpret_cast<uintptr_t>(in_pairs) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(scratch)) & 15)
        return VRENB200_EALIGN;
    cudaStream_t s = as_stream(stream);
    uint32_t* counters = reinterpret_cast<uint32_t*>(static_cast<char*>(out) + align_up((size_t) n * 8, 256)); // bucket_sort.cpp:86
    if (n == 0) return check_cuda(cudaMemsetAsync(counters, 0, kBucketKeys * sizeof(uint32_t), s));           // bucket_sort.cpp:104
    if (scratch == nullptr || scratch_bytes < vrenb200_bucket_sort_scratch_bytes(n)) return VRENB200_ESCRATCH;
    char* p = static_cast<char*>(scratch);
    uint32_t* tmp = reinterpret_cast<uint32_t*>(p);
    void* ctl_mem = p + align_up((size_t) n * 8, 256);
    uint32_t* raw_counts = reinterpret_cast<uint32_t*>(p + align_up((size_t) n * 8, 256) + control_bytes(n));
    const sort_variant& var = pick_variant(n, LAYOUT_AOS, s);
    const uint32_t tiles = (uint32_t) (((size_t) n + var.tile - 1) / var.tile);
    sort_control* ctl = static_cast<sort_control*>(ctl_mem);
    uint32_t* lookback = reinterpret_cast<uint32_t*>(ctl + 1);
    const size_t clear = sizeof(sort_control) + (size_t) (var.clears_next_pass ? 1 : 2) * tiles * kRadix * sizeof(uint32_t);
    VRENB200_TRY(check_cuda(cudaMemsetAsync(ctl_mem, 0, clear, s)));
    // small inputs (the light Morton sort of the clustered chain): bucket counts by global atomics in the histogram read,
    // prefix by bucket_end_offsets_kernel; large inputs: no global atomics, END offsets by search in the sorted output
    const bool by_search = n >= g_bucket_search_min;
    if (!by_search) VRENB200_TRY(check_cuda(cudaMemsetAsync(raw_counts, 0, kBucketKeys * sizeof(uint32_t), s)));
    const uint32_t hist_grid = (uint32_t) std::min<size_t>(kNumSMs * (by_search ? 4 : 2), ((size_t) n / 2 + kHistThreads - 1) / kHistThreads + 1);
    if (by_search)
        bucket_digit_histogram_kernel<<<hist_grid, kHistThreads, 0, s>>>(static_cast<const uint2*>(in_pairs), n, ctl);
    else
        bucket_histogram_kernel<<<hist_grid, kHistThreads, 0, s>>>(static_cast<const uint2*>(in_pairs), n, ctl, raw_counts);
    VRENB200_TRY(check_launch());
    radix_scan_histograms_kernel<<<2, kRadix, 0, s>>>(ctl);
    VRENB200_TRY(check_launch());
    VRENB200_TRY(var.launch(s, static_cast<const uint32_t*>(in_pairs), tmp, nullptr, nullptr, n, var.clears_next_pass ? kClearNextPassRow : 0,
                            ctl, lookback, tiles, LAYOUT_AOS));
    VRENB200_TRY(var.launch(s, tmp, static_cast<uint32_t*>(out), nullptr, nullptr, n, 1, ctl, lookback, tiles, LAYOUT_AOS));
    if (by_search)
        bucket_end_offsets_search_kernel<<<kBucketKeys / 256, 256, 0, s>>>(static_cast<const uint2*>(out), n, counters);
    else
        bucket_end_offsets_kernel<<<kEndOffsetCtas, 1024, 0, s>>>(raw_counts, counters);
    return check_launch();
}

extern "C" int vrenb200_bucket_sort_set_search_min(uint32_t n_min)
{
    g_bucket_search_min = n_min;
    return VRENB200_OK;
}
