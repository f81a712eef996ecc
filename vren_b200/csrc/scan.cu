// a2 — vren::blelloch_scan (reference: vren/vren/primitives/blelloch_scan.{hpp,cpp},
// shaders/blelloch_scan_downsweep.comp:34-125).
//
// operator(): exclusive add-scan of uint32 (mod 2^32).  The reference runs reduce (up-sweep) + one strided
// dispatch per level above 2^10 + a workgroup pass (~22 dispatches and ~3 round trips over the data at 2^28).
// Here: ONE single-pass chained scan with decoupled look-back: each CTA scans an 8192-element tile held in
// registers (8 x 128-bit loads per thread, warp-shuffle scans), publishes {aggregate | inclusive prefix} in a 64-bit
// status word, and a warp-wide look-back window resolves its exclusive prefix.  8 B/elt of HBM traffic.
//
// downsweep(): kept for API parity (radix_sort.cpp:281-289 calls it directly): takes an up-sweep TREE and
// turns it into the scan, 10 levels per launch in shared memory, top-down.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace vrenb200 {

namespace {

// [[chained-scan-defs-begin]] (this block and the kernel body below are also compiled for the host: tests/test_scan_emulation.py)
constexpr int kScanVecs = 8;                                        // uint4 per thread
constexpr uint32_t kScanMinTile = 256 * kScanVecs * 4;              // smallest tile of any variant: sizes the status array
constexpr uint32_t kWarpChunk = 32 * kScanVecs * 4;                 // 1024 contiguous elements per warp

constexpr uint64_t kFlagAggregate = 1ull << 32;
constexpr uint64_t kFlagInclusive = 2ull << 32;

struct scan_state
{
    uint32_t ticket;      // safe mode only (TICKET kernels): next tile id.  By default tile id = block index (CTAs are dispatched
                          // in index order, so a tile only waits on tiles that already started; saves an L2 atomic round trip)
    uint32_t _pad[63];
    uint64_t status[1];   // [tiles]
};
// [[chained-scan-defs-end]]

// THREADS sets the tile (THREADS x 32 elements).  The look-back walk is as long as the number of older tiles still in
// flight, so for a fixed number of bytes in flight larger tiles mean proportionally fewer L2 round trips per tile.
// TICKET (the "safe mode", vrenb200_exclusive_scan_u32_ex): the tile id is the value of an atomic counter taken when the
// CTA starts instead of the block index, so a tile only ever waits for tiles whose CTAs HAVE started — forward progress of
// the look-back without the assumption that CTAs are dispatched in index order (MPS, time slicing, debuggers).
// (The text between the two marker comments is also compiled for the host: tests/test_scan_emulation.py.)
template <int THREADS, bool TICKET>
__global__ void __launch_bounds__(THREADS)
exclusive_scan_u32_kernel(const uint32_t* in, uint32_t* out, uint32_t n, scan_state* state, uint32_t base)
{
    // [[chained-scan-body-begin]]
    constexpr int kScanWarps = THREADS / 32;
    constexpr uint32_t kScanTile = THREADS * kScanVecs * 4;
    __shared__ uint32_t s_warp_total[kScanWarps];
    __shared__ uint32_t s_tile_prefix;

    uint32_t tile = blockIdx.x;
    if (TICKET)
    {
        __shared__ uint32_t s_ticket;
        if (threadIdx.x == 0) s_ticket = atomicAdd(&state->ticket, 1u);
        __syncthreads();
        tile = s_ticket;
    }
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t warp_base = (uint64_t) tile * kScanTile + warp * kWarpChunk;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;

    uint4 x[kScanVecs];
#pragma unroll
    for (int v = 0; v < kScanVecs; v++)
    {
        const uint64_t idx = warp_base + v * 128 + lane * 4;
        if (vec_ok && idx + 4 <= n)
            x[v] = ldg_stream_u4(reinterpret_cast<const uint4*>(in + idx));
        else
        {
            x[v].x = idx + 0 < n ? in[idx + 0] : 0u;
            x[v].y = idx + 1 < n ? in[idx + 1] : 0u;
            x[v].z = idx + 2 < n ? in[idx + 2] : 0u;
            x[v].w = idx + 3 < n ? in[idx + 3] : 0u;
        }
    }

    // in-thread exclusive scan of each vector + warp scan of the vector totals, carried across v
    uint32_t carry = 0;
#pragma unroll
    for (int v = 0; v < kScanVecs; v++)
    {
        const uint32_t a = x[v].x, b = x[v].y, c = x[v].z, d = x[v].w;
        const uint32_t total = a + b + c + d;
        uint32_t inc = total;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1)
        {
            const uint32_t t = __shfl_up_sync(kFullMask, inc, s);
            if (lane >= (unsigned) s) inc += t;
        }
        const uint32_t base = carry + inc - total;
        carry += __shfl_sync(kFullMask, inc, 31);
        x[v].x = base;
        x[v].y = base + a;
        x[v].z = base + a + b;
        x[v].w = base + a + b + c;
    }
    if (lane == 31) s_warp_total[warp] = carry;
    __syncthreads();

    uint32_t warp_prefix = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < kScanWarps; w++)
    {
        const uint32_t t = s_warp_total[w];
        if (w < (int) warp) warp_prefix += t;
        tile_total += t;
    }

    // decoupled look-back by warp 0: four 32-wide windows (128 predecessors) are fetched per L2 round trip and then
    // consumed nearest first.  (A CTA-wide 256-lane window and a persistent TMA-prefetch variant were measured slower:
    // profiles/r1g_*.)
    if (warp == 0)
    {
        uint64_t* status = state->status;
        if (lane == 0)
            st_relaxed_u64(&status[tile], (tile == 0 ? kFlagInclusive : kFlagAggregate) | tile_total);
        uint32_t exclusive = 0;
        if (tile > 0)
        {
            constexpr int K = 2;
            int64_t window = (int64_t) tile - 1;
            bool done = false;
            while (!done)
            {
                uint64_t s[K];
#pragma unroll
                for (int k = 0; k < K; k++)
                {
                    const int64_t t = window - 32 * k - lane;
                    s[k] = t >= 0 ? ld_relaxed_u64(&status[t]) : kFlagInclusive; // virtual tiles before tile 0: inclusive 0
                }
#pragma unroll
                for (int k = 0; k < K; k++)
                {
                    if (done) break;
                    const int64_t t = window - 32 * k - lane;
                    while ((s[k] >> 32) == 0)
                    {
                        __nanosleep(100);   // back off: hundreds of tiles poll the same few status lines in L2
                        s[k] = ld_relaxed_u64(&status[t]);
                    }
                    const unsigned incl = __ballot_sync(kFullMask, (s[k] >> 32) == 2);
                    uint32_t val = (uint32_t) s[k];
                    if (incl != 0 && lane > (unsigned) (__ffs(incl) - 1)) val = 0;
                    exclusive += __reduce_add_sync(kFullMask, val);
                    done = incl != 0;
                }
                window -= 32 * K;
            }
            if (lane == 0) st_relaxed_u64(&status[tile], kFlagInclusive | (uint32_t) (exclusive + tile_total));
        }
        if (lane == 0) s_tile_prefix = exclusive;
    }
    __syncthreads();
    const uint32_t exclusive = s_tile_prefix;
    const uint32_t prefix = exclusive + warp_prefix + base;

#pragma unroll
    for (int v = 0; v < kScanVecs; v++)
    {
        const uint64_t idx = warp_base + v * 128 + lane * 4;
        x[v].x += prefix; x[v].y += prefix; x[v].z += prefix; x[v].w += prefix;
        if (vec_ok && idx + 4 <= n)
            *reinterpret_cast<uint4*>(out + idx) = x[v];
        else
        {
            if (idx + 0 < n) out[idx + 0] = x[v].x;
            if (idx + 1 < n) out[idx + 1] = x[v].y;
            if (idx + 2 < n) out[idx + 2] = x[v].z;
            if (idx + 3 < n) out[idx + 3] = x[v].w;
        }
    }
    // [[chained-scan-body-end]]
}

// ---- staged variant: the tile lives in shared memory, not in registers -------------------------------------------
// ncu on the register-tile kernel above (profiles/r1l_ncu_scan.md): 63 % of the warp samples sit at the barrier behind
// warp 0's look-back, and with 63 registers x 32 elements per thread only 4 CTAs (128 KB of loads) fit an SM, so
// little is in flight while a CTA waits.  Here ONE bulk copy (TMA) brings a 64 KB tile (16384 elements) into shared
// memory the moment the CTA starts; the tile is read twice from there (warp totals, then the scan itself), the
// registers stay small, three CTAs = 192 KB of loads are in flight per SM, and the status chain is half as long.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
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
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int kStagedThreads = 256;
constexpr int kStagedWarps = kStagedThreads / 32;
constexpr int kStagedVecs = 16;                                            // uint4 per thread
constexpr uint32_t kStagedTile = kStagedThreads * kStagedVecs * 4;         // 16384 elements = 64 KB
constexpr uint32_t kStagedWarpChunk = 32 * kStagedVecs * 4;                // 2048 contiguous elements per warp

// MODE 0: the real thing.  Timing experiments only (wrong results, reachable through the tuning hook, never the default):
// 1: no look-back, 2: only wait for the 32 nearest predecessors to publish, 3: full walk that never waits on a flag
template <int MODE>
__global__ void __launch_bounds__(kStagedThreads, 3)
exclusive_scan_staged_kernel(const uint32_t* in, uint32_t* out, uint32_t n, scan_state* state, uint32_t base)
{
    extern __shared__ __align__(128) uint32_t s_tile[];
    __shared__ uint32_t s_warp_total[kStagedWarps];
    __shared__ uint32_t s_tile_prefix;
    __shared__ __align__(8) uint64_t s_bar;

    const uint32_t tile = blockIdx.x;
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t tile_base = (uint64_t) tile * kStagedTile;
    const uint32_t valid = (n - tile_base) < (uint64_t) kStagedTile ? (uint32_t) (n - tile_base) : kStagedTile;
    const bool bulk = valid == kStagedTile && (reinterpret_cast<uintptr_t>(in) & 15) == 0;

    if (bulk)
    {
        if (tid == 0)
        {
            mbar_init(&s_bar, 1);
            mbar_arrive_expect_tx(&s_bar, kStagedTile * 4);
            bulk_copy_g2s(s_tile, in + tile_base, kStagedTile * 4, &s_bar);
        }
        __syncthreads();      // the barrier word is initialised before anyone polls it
        mbar_wait(&s_bar, 0);
    }
    else
    {
        for (uint32_t i = tid; i < kStagedTile; i += kStagedThreads) s_tile[i] = i < valid ? in[tile_base + i] : 0u;
        __syncthreads();
    }

    // first read: warp totals
    const uint4* mine = reinterpret_cast<const uint4*>(s_tile + warp * kStagedWarpChunk) + lane;
    uint32_t sum = 0;
#pragma unroll
    for (int v = 0; v < kStagedVecs; v++)
    {
        const uint4 x = mine[v * 32];
        sum += x.x + x.y + x.z + x.w;
    }
    sum = __reduce_add_sync(kFullMask, sum);
    if (lane == 0) s_warp_total[warp] = sum;
    __syncthreads();
    uint32_t warp_prefix = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < kStagedWarps; w++)
    {
        const uint32_t t = s_warp_total[w];
        if (w < (int) warp) warp_prefix += t;
        tile_total += t;
    }

    if (warp == 0)
    {
        uint64_t* status = state->status;
        if (lane == 0)
            st_relaxed_u64(&status[tile], (tile == 0 ? kFlagInclusive : kFlagAggregate) | tile_total);
        uint32_t exclusive = 0;
        if (MODE != 1 && tile > 0)
        {
            // 128 predecessors per L2 round trip, consumed nearest first
            constexpr int K = 4;
            int64_t window = (int64_t) tile - 1;
            bool done = false;
            while (!done)
            {
                uint64_t s0, s1, s2, s3;
                {
                    const int64_t t = window - lane;
                    s0 = t >= 0 ? ld_relaxed_u64(&status[t]) : kFlagInclusive;
                    s1 = t - 32 >= 0 ? ld_relaxed_u64(&status[t - 32]) : kFlagInclusive;
                    s2 = t - 64 >= 0 ? ld_relaxed_u64(&status[t - 64]) : kFlagInclusive;
                    s3 = t - 96 >= 0 ? ld_relaxed_u64(&status[t - 96]) : kFlagInclusive;
                }
                auto consume = [&](uint64_t sv, int k) {
                    if (done) return;
                    const int64_t t = window - 32 * k - lane;
                    while (MODE != 3 && (sv >> 32) == 0)
                    {
                        __nanosleep(40);
                        sv = ld_relaxed_u64(&status[t]);
                    }
                    if (MODE == 3 && window - 32 * k - 31 <= 0) sv = kFlagInclusive;   // the walk must end somewhere
                    if (MODE == 2) sv = kFlagInclusive;
                    const unsigned incl = __ballot_sync(kFullMask, (sv >> 32) == 2);
                    uint32_t val = (uint32_t) sv;
                    if (incl != 0 && lane > (unsigned) (__ffs(incl) - 1)) val = 0;
                    exclusive += __reduce_add_sync(kFullMask, val);
                    done = incl != 0;
                };
                consume(s0, 0);
                consume(s1, 1);
                consume(s2, 2);
                consume(s3, 3);
                window -= 32 * K;
            }
            if (lane == 0) st_relaxed_u64(&status[tile], kFlagInclusive | (uint32_t) (exclusive + tile_total));
        }
        if (lane == 0) s_tile_prefix = exclusive;
    }
    __syncthreads();

    // second read: the scan itself, carried across the warp's 16 rounds
    uint32_t carry = s_tile_prefix + warp_prefix + base;
    const uint64_t warp_base = tile_base + warp * kStagedWarpChunk;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
#pragma unroll 4
    for (int v = 0; v < kStagedVecs; v++)
    {
        const uint4 x = mine[v * 32];
        const uint32_t total = x.x + x.y + x.z + x.w;
        uint32_t inc = total;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1)
        {
            const uint32_t t = __shfl_up_sync(kFullMask, inc, s);
            if (lane >= (unsigned) s) inc += t;
        }
        uint4 y;
        y.x = carry + inc - total;
        y.y = y.x + x.x;
        y.z = y.y + x.y;
        y.w = y.z + x.z;
        carry += __shfl_sync(kFullMask, inc, 31);
        const uint64_t idx = warp_base + v * 128 + lane * 4;
        if (vec_ok && idx + 4 <= n)
            *reinterpret_cast<uint4*>(out + idx) = y;
        else
        {
            if (idx + 0 < n) out[idx + 0] = y.x;
            if (idx + 1 < n) out[idx + 1] = y.y;
            if (idx + 2 < n) out[idx + 2] = y.z;
            if (idx + 3 < n) out[idx + 3] = y.w;
        }
    }
}

// ---- run-ahead variant --------------------------------------------------------------------------------------------
// Timing experiments on the staged kernel (profiles/r1l_scan_lookback_experiments.md): with the look-back removed the
// kernel runs at the HBM roofline (0.32 ms at 2^28), and merely WAITING for the 32 nearest predecessors to publish
// their aggregates costs 0.18 ms: a tile holds 64 KB of shared memory while the slowest of its predecessors' loads
// arrives, so the loads in flight drop.  The chain cannot be made faster; it is moved off the critical path instead.
//
// CTA c (512 threads, CTAs dispatched in index order) does three unrelated things:
//   * warps 8-15 REDUCE tile c: read it from HBM (which also pulls it into the 126 MB L2) and publish its aggregate;
//   * warp 8 then FINALIZES tile c - A: a look-back walk over aggregates published >= A CTAs ago (no waiting on
//     stragglers), publishing the tile's inclusive prefix;
//   * warps 0-7 SCAN tile c - A - B: one bulk copy re-reads the tile (an L2 hit: (A+B) x 64 KB is a fraction of the
//     L2), the inclusive prefix of the previous tile was finalized >= B CTAs ago, two passes over shared memory, store.
// Every dependency points to a CTA launched earlier.  HBM traffic stays 8 B/element; the second read is L2 traffic.
constexpr int kRunaheadThreads = 512;

// L2 residency hints: the tile must survive in L2 between the reduce role's read and the scan role's re-read
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint4 ldg_u4_hint(const uint4* p, uint64_t policy)
{
    uint4 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(policy));
    return v;
}
__device__ __forceinline__ void stg_u4_hint(uint4* p, uint4 v, uint64_t policy)
{
    asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;"
                 ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(policy) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

__device__ __forceinline__ void named_barrier(int id, int threads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

template <bool HINTS>
__global__ void __launch_bounds__(kRunaheadThreads, 3)
exclusive_scan_runahead_kernel(const uint32_t* in, uint32_t* out, uint32_t n, scan_state* state, uint32_t base,
                               uint32_t tiles, uint32_t lag_finalize, uint32_t lag_scan)
{
    extern __shared__ __align__(128) uint32_t s_tile[];
    __shared__ uint32_t s_warp_total[kStagedWarps];
    __shared__ uint32_t s_reduce_total[kStagedWarps];
    __shared__ uint32_t s_tile_prefix;
    __shared__ __align__(8) uint64_t s_bar;

    const unsigned lane = threadIdx.x & 31;
    uint64_t* status = state->status;
    const int64_t c = blockIdx.x;
    const bool in_aligned = (reinterpret_cast<uintptr_t>(in) & 15) == 0;

    if (threadIdx.x >= kStagedThreads)
    {
        // ---- reduce role: tile c ----
        const unsigned tid = threadIdx.x - kStagedThreads, warp = tid >> 5;
        if (c < (int64_t) tiles)
        {
            const uint64_t tile_base = (uint64_t) c * kStagedTile;
            const uint32_t valid = (n - tile_base) < (uint64_t) kStagedTile ? (uint32_t) (n - tile_base) : kStagedTile;
            uint32_t sum = 0;
            if (valid == kStagedTile && in_aligned)
            {
                const uint4* src = reinterpret_cast<const uint4*>(in + tile_base) + tid;
                const uint64_t keep = HINTS ? l2_policy_evict_last() : 0;
#pragma unroll 4
                for (int v = 0; v < kStagedVecs; v++)
                {
                    const uint4 x = HINTS ? ldg_u4_hint(src + v * kStagedThreads, keep) : ldg_stream_u4(src + v * kStagedThreads);
                    sum += x.x + x.y + x.z + x.w;
                }
            }
            else
                for (uint32_t i = tid; i < valid; i += kStagedThreads) sum += in[tile_base + i];
            sum = __reduce_add_sync(kFullMask, sum);
            if (lane == 0) s_reduce_total[warp] = sum;
        }
        named_barrier(1, kStagedThreads);
        if (warp != 0) return;
        if (c < (int64_t) tiles && lane == 0)
        {
            uint32_t tile_total = 0;
#pragma unroll
            for (int w = 0; w < kStagedWarps; w++) tile_total += s_reduce_total[w];
            st_relaxed_u64(&status[c], kFlagAggregate | tile_total);
        }
        // ---- finalize role: tile f = c - lag_finalize ----
        const int64_t f = c - (int64_t) lag_finalize;
        if (f < 0 || f >= (int64_t) tiles) return;
        uint32_t exclusive = 0;
        constexpr int K = 4;
        int64_t window = f - 1;
        bool done = f == 0;
        while (!done)
        {
            uint64_t s0, s1, s2, s3;
            {
                const int64_t t = window - lane;
                s0 = t >= 0 ? ld_relaxed_u64(&status[t]) : kFlagInclusive;
                s1 = t - 32 >= 0 ? ld_relaxed_u64(&status[t - 32]) : kFlagInclusive;
                s2 = t - 64 >= 0 ? ld_relaxed_u64(&status[t - 64]) : kFlagInclusive;
                s3 = t - 96 >= 0 ? ld_relaxed_u64(&status[t - 96]) : kFlagInclusive;
            }
            auto consume = [&](uint64_t sv, int k) {
                if (done) return;
                const int64_t t = window - 32 * k - lane;
                while ((sv >> 32) == 0)
                {
                    __nanosleep(40);
                    sv = ld_relaxed_u64(&status[t]);
                }
                const unsigned incl = __ballot_sync(kFullMask, (sv >> 32) == 2);
                uint32_t val = (uint32_t) sv;
                if (incl != 0 && lane > (unsigned) (__ffs(incl) - 1)) val = 0;
                exclusive += __reduce_add_sync(kFullMask, val);
                done = incl != 0;
            };
            consume(s0, 0);
            consume(s1, 1);
            consume(s2, 2);
            consume(s3, 3);
            window -= 32 * K;
        }
        if (lane == 0)
        {
            // the tile's own aggregate is in its status word (published >= lag_finalize CTAs ago)
            uint64_t own = ld_relaxed_u64(&status[f]);
            while ((own >> 32) == 0)
            {
                __nanosleep(40);
                own = ld_relaxed_u64(&status[f]);
            }
            st_relaxed_u64(&status[f], kFlagInclusive | (uint32_t) (exclusive + (uint32_t) own));
        }
        return;
    }

    // ---- scan role: tile q = c - lag_finalize - lag_scan ----
    const unsigned tid = threadIdx.x, warp = tid >> 5;
    const int64_t q = c - (int64_t) lag_finalize - (int64_t) lag_scan;
    if (q < 0 || q >= (int64_t) tiles) return;
    const uint64_t tile_base = (uint64_t) q * kStagedTile;
    const uint32_t valid = (n - tile_base) < (uint64_t) kStagedTile ? (uint32_t) (n - tile_base) : kStagedTile;
    const bool bulk = valid == kStagedTile && in_aligned;
    if (bulk && tid == 0)
    {
        mbar_init(&s_bar, 1);
        mbar_arrive_expect_tx(&s_bar, kStagedTile * 4);
        if (HINTS)
            bulk_copy_g2s_hint(s_tile, in + tile_base, kStagedTile * 4, &s_bar, l2_policy_evict_first());
        else
            bulk_copy_g2s(s_tile, in + tile_base, kStagedTile * 4, &s_bar);
    }
    // the previous tile's inclusive prefix: one poller, its latency hides behind the copy
    if (tid == 32)
    {
        uint32_t exclusive = 0;
        if (q > 0)
        {
            uint64_t sv = ld_relaxed_u64(&status[q - 1]);
            while ((sv >> 32) != 2)
            {
                __nanosleep(40);
                sv = ld_relaxed_u64(&status[q - 1]);
            }
            exclusive = (uint32_t) sv;
        }
        // in-place scans (out == in): this CTA's stores must not land before the reduce role of CTA q has read the
        // tile; its aggregate being published proves it has.  (Nothing else orders the two: with fewer than
        // lag_finalize + lag_scan CTAs of distance they can be resident together.)
        while ((ld_relaxed_u64(&status[q]) >> 32) == 0) __nanosleep(40);
        s_tile_prefix = exclusive;
    }
    if (bulk)
    {
        named_barrier(2, kStagedThreads);   // the barrier word is initialised before anyone polls it
        mbar_wait(&s_bar, 0);
    }
    else
    {
        for (uint32_t i = tid; i < kStagedTile; i += kStagedThreads) s_tile[i] = i < valid ? in[tile_base + i] : 0u;
        named_barrier(2, kStagedThreads);
    }

    const uint4* mine = reinterpret_cast<const uint4*>(s_tile + warp * kStagedWarpChunk) + lane;
    uint32_t sum = 0;
#pragma unroll
    for (int v = 0; v < kStagedVecs; v++)
    {
        const uint4 x = mine[v * 32];
        sum += x.x + x.y + x.z + x.w;
    }
    sum = __reduce_add_sync(kFullMask, sum);
    if (lane == 0) s_warp_total[warp] = sum;
    named_barrier(2, kStagedThreads);
    uint32_t warp_prefix = 0;
#pragma unroll
    for (int w = 0; w < kStagedWarps; w++)
        if (w < (int) warp) warp_prefix += s_warp_total[w];

    uint32_t carry = s_tile_prefix + warp_prefix + base;
    const uint64_t warp_base = tile_base + warp * kStagedWarpChunk;
    const uint64_t drop = HINTS ? l2_policy_evict_first() : 0;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
#pragma unroll 4
    for (int v = 0; v < kStagedVecs; v++)
    {
        const uint4 x = mine[v * 32];
        const uint32_t total = x.x + x.y + x.z + x.w;
        uint32_t inc = total;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1)
        {
            const uint32_t t = __shfl_up_sync(kFullMask, inc, s);
            if (lane >= (unsigned) s) inc += t;
        }
        uint4 y;
        y.x = carry + inc - total;
        y.y = y.x + x.x;
        y.z = y.y + x.y;
        y.w = y.z + x.z;
        carry += __shfl_sync(kFullMask, inc, 31);
        const uint64_t idx = warp_base + v * 128 + lane * 4;
        if (vec_ok && idx + 4 <= n)
        {
            if (HINTS) stg_u4_hint(reinterpret_cast<uint4*>(out + idx), y, drop);
            else *reinterpret_cast<uint4*>(out + idx) = y;
        }
        else
        {
            if (idx + 0 < n) out[idx + 0] = y.x;
            if (idx + 1 < n) out[idx + 1] = y.y;
            if (idx + 2 < n) out[idx + 2] = y.z;
            if (idx + 3 < n) out[idx + 3] = y.w;
        }
    }
}

// blelloch_scan_downsweep.comp:59-125 generalised to a strided logical array: logical element i lives at
// buf[(i+1)*stride-1]; `levels` = min(10, log2(count)) down-sweep levels in shared memory.
__global__ void __launch_bounds__(1024)
downsweep_strided_kernel(uint32_t* buf, uint32_t count, uint64_t stride, uint64_t row_stride,
                         int clear_last, int levels)
{
    __shared__ uint32_t s_data[1024];
    uint32_t* row = buf + blockIdx.y * row_stride;
    const uint32_t gi = blockIdx.x * 1024u + threadIdx.x;
    const uint64_t idx = ((uint64_t) gi + 1) * stride - 1;
    uint32_t v = gi < count ? row[idx] : 0u;
    if (clear_last && gi == count - 1) v = 0u;
    s_data[threadIdx.x] = v;
    __syncthreads();
    for (int level = levels - 1; level >= 0; level--)
    {
        const uint32_t m = (1u << (level + 1)) - 1;
        if ((threadIdx.x & m) == m)
        {
            const uint32_t a = threadIdx.x - (1u << level);
            const uint32_t t = s_data[threadIdx.x];
            s_data[threadIdx.x] = s_data[a] + t;
            s_data[a] = t;
        }
        __syncthreads();
    }
    if (gi < count) row[idx] = s_data[threadIdx.x];
}

inline int ilog2(uint32_t v) { int r = 0; while (v >>= 1) r++; return r; }

} // namespace
} // namespace vrenb200

using namespace vrenb200;

namespace {
#ifdef VRENB200_TUNING
int g_scan_variant = 0;      // tuning builds only: selected through vrenb200_scan_set_variant
int g_lag_finalize = 0, g_lag_scan = 0;
#else
constexpr int g_scan_variant = 0;   // release builds: no process-global selection state
constexpr int g_lag_finalize = 0, g_lag_scan = 0;
#endif
constexpr int kScanAuto = 9999;
constexpr int kScanVariantThreads[] = { kScanAuto, 256, 512, 1024, 0 /* staged 64 KB tile */, 1, 2, 3 /* staged, timing experiments (MODE) */,
                                        -64, -128, -256, -512 /* run-ahead, both lags = -value tiles */,
                                        -10064, -10128, -10256, -10512 /* same with L2 residency hints */ };
}

#ifdef VRENB200_TUNING
// tuning hook: explicit run-ahead distances (tiles of 16384 elements); selects the run-ahead kernel
extern "C" int vrenb200_scan_set_runahead(int finalize_lag, int scan_lag)
{
    if (finalize_lag < 1 || scan_lag < 1 || finalize_lag > 4096 || scan_lag > 4096) return VRENB200_EINVAL_ARG;
    g_lag_finalize = finalize_lag;
    g_lag_scan = scan_lag;
    g_scan_variant = 15;   // hinted run-ahead; the lags above override the table value
    return VRENB200_OK;
}

// tuning hook (bench.py / tools): 0: default (auto), 1-3: register tile with 256 / 512 / 1024 threads, 4: staged 64 KB tile,
// 5-7: timing experiments (wrong results), 8-11: run-ahead with distances 64..512, 12-15: same with L2 residency hints
extern "C" int vrenb200_scan_set_variant(int v)
{
    if (v < 0 || v >= (int) (sizeof(kScanVariantThreads) / sizeof(int))) return VRENB200_EINVAL_ARG;
    g_scan_variant = v;
    g_lag_finalize = g_lag_scan = 0;
    return VRENB200_OK;
}
#endif

extern "C" size_t vrenb200_scan_scratch_bytes(uint32_t n)
{
    const size_t tiles = ((size_t) n + kScanMinTile - 1) / kScanMinTile;
    return align_up(offsetof(scan_state, status) + (tiles > 0 ? tiles : 1) * sizeof(uint64_t), 256);
}

extern "C" int vrenb200_exclusive_scan_u32(vrenb200_stream_t stream, const uint32_t* in, uint32_t* out, uint32_t n,
                                           void* scratch, size_t scratch_bytes)
{
    return vrenb200_exclusive_scan_u32_base(stream, in, out, n, 0u, scratch, scratch_bytes);
}

// out[i] = base + sum_{k<i} in[k]: the local step of the sharded scan (base = total of the lower ranks)
extern "C" int vrenb200_exclusive_scan_u32_base(vrenb200_stream_t stream, const uint32_t* in, uint32_t* out, uint32_t n,
                                                uint32_t base, void* scratch, size_t scratch_bytes)
{
    // the environment can make the safe mode the default of a process (read once, never written again)
    static const uint32_t env_flags = []() {
        const char* e = std::getenv("VRENB200_SCAN_TILE_IDS");
        return (e != nullptr && !std::strcmp(e, "ticket")) ? (uint32_t) VRENB200_SCAN_TILE_IDS_TICKET : 0u;
    }();
    return vrenb200_exclusive_scan_u32_ex(stream, in, out, n, base, scratch, scratch_bytes, env_flags);
}

extern "C" int vrenb200_exclusive_scan_u32_ex(vrenb200_stream_t stream, const uint32_t* in, uint32_t* out, uint32_t n,
                                              uint32_t base, void* scratch, size_t scratch_bytes, uint32_t flags)
{
    if (in == nullptr || out == nullptr) return VRENB200_EINVAL_ARG;
    if ((flags & ~(uint32_t) VRENB200_SCAN_TILE_IDS_TICKET) != 0) return VRENB200_EINVAL_ARG;
    if (n == 0) return VRENB200_EINVAL_LENGTH;
    const size_t need = vrenb200_scan_scratch_bytes(n);
    if (scratch == nullptr || scratch_bytes < need) return VRENB200_ESCRATCH;
    if ((reinterpret_cast<uintptr_t>(scratch) & 7) != 0) return VRENB200_EALIGN;
    cudaStream_t s = as_stream(stream);
    VRENB200_TRY(check_cuda(cudaMemsetAsync(scratch, 0, need, s)));
    if (flags & VRENB200_SCAN_TILE_IDS_TICKET)
    {
        // safe mode: the chained register-tile kernel with ticket tile ids at every size (the run-ahead kernel's roles are tied
        // to the block index); the memset above has zeroed the ticket
        const uint32_t tile = (n >= (1u << 22) ? 1024u : 256u) * kScanVecs * 4;
        const uint32_t tiles = (uint32_t) (((size_t) n + tile - 1) / tile);
        scan_state* st = static_cast<scan_state*>(scratch);
        if (n >= (1u << 22)) exclusive_scan_u32_kernel<1024, true><<<tiles, 1024, 0, s>>>(in, out, n, st, base);
        else exclusive_scan_u32_kernel<256, true><<<tiles, 256, 0, s>>>(in, out, n, st, base);
        return check_launch();
    }
    int threads = kScanVariantThreads[g_scan_variant];
    // default (profiles/r1v_scan_mid_sizes.log): the run-ahead kernel (256 + 256 tiles of distance, L2 residency hints)
    // from 2^24 elements, the register-tile kernel below that (no idle prologue CTAs), with 1024-thread CTAs from 2^22
    if (threads == kScanAuto) threads = n >= (1u << 24) ? -10256 : (n >= (1u << 22) ? 1024 : 256);
    if (threads < 0)
    {
        // per call: function attributes belong to the current device's context, and a process may use several devices
        VRENB200_TRY(check_cuda(cudaFuncSetAttribute(exclusive_scan_runahead_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     (int) (kStagedTile * 4))));
        VRENB200_TRY(check_cuda(cudaFuncSetAttribute(exclusive_scan_runahead_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     (int) (kStagedTile * 4))));
        const uint32_t staged_tiles = (uint32_t) (((size_t) n + kStagedTile - 1) / kStagedTile);
        const bool hints = -threads > 10000;
        const uint32_t lag = (uint32_t) (-threads % 10000);    // finalize lag = scan lag = lag
        const uint32_t la = g_lag_finalize > 0 ? (uint32_t) g_lag_finalize : lag, lb = g_lag_scan > 0 ? (uint32_t) g_lag_scan : lag;
        scan_state* st = static_cast<scan_state*>(scratch);
        if (hints)
            exclusive_scan_runahead_kernel<true><<<staged_tiles + la + lb, kRunaheadThreads, kStagedTile * 4, s>>>(in, out, n, st, base, staged_tiles, la, lb);
        else
            exclusive_scan_runahead_kernel<false><<<staged_tiles + la + lb, kRunaheadThreads, kStagedTile * 4, s>>>(in, out, n, st, base, staged_tiles, la, lb);
        return check_launch();
    }
    if (threads <= 3)
    {
        const int bytes = (int) (kStagedTile * 4);
        VRENB200_TRY(check_cuda(cudaFuncSetAttribute(exclusive_scan_staged_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)));
        VRENB200_TRY(check_cuda(cudaFuncSetAttribute(exclusive_scan_staged_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)));
        VRENB200_TRY(check_cuda(cudaFuncSetAttribute(exclusive_scan_staged_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)));
        VRENB200_TRY(check_cuda(cudaFuncSetAttribute(exclusive_scan_staged_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)));
        const uint32_t staged_tiles = (uint32_t) (((size_t) n + kStagedTile - 1) / kStagedTile);
        scan_state* st = static_cast<scan_state*>(scratch);
        const size_t smem = kStagedTile * 4;
        switch (threads)
        {
        case 0: exclusive_scan_staged_kernel<0><<<staged_tiles, kStagedThreads, smem, s>>>(in, out, n, st, base); break;
        case 1: exclusive_scan_staged_kernel<1><<<staged_tiles, kStagedThreads, smem, s>>>(in, out, n, st, base); break;
        case 2: exclusive_scan_staged_kernel<2><<<staged_tiles, kStagedThreads, smem, s>>>(in, out, n, st, base); break;
        default: exclusive_scan_staged_kernel<3><<<staged_tiles, kStagedThreads, smem, s>>>(in, out, n, st, base); break;
        }
        return check_launch();
    }
    const uint32_t tile = (uint32_t) threads * kScanVecs * 4;
    const uint32_t tiles = (uint32_t) (((size_t) n + tile - 1) / tile);
    scan_state* st = static_cast<scan_state*>(scratch);
    if (threads == 256) exclusive_scan_u32_kernel<256, false><<<tiles, 256, 0, s>>>(in, out, n, st, base);
    else if (threads == 512) exclusive_scan_u32_kernel<512, false><<<tiles, 512, 0, s>>>(in, out, n, st, base);
    else exclusive_scan_u32_kernel<1024, false><<<tiles, 1024, 0, s>>>(in, out, n, st, base);
    return check_launch();
}

extern "C" int vrenb200_blelloch_downsweep_u32(vrenb200_stream_t stream, uint32_t* buf, uint32_t n, uint32_t blocks,
                                               int clear_last)
{
    if (buf == nullptr) return VRENB200_EINVAL_ARG;
    if (n == 0 || (n & (n - 1)) != 0 || blocks == 0 || blocks > 65535) return VRENB200_EINVAL_LENGTH;
    cudaStream_t s = as_stream(stream);
    // n < 1024: the reference's zero-filled 1024-wide tile swallows the root, i.e. behaves as clear_last
    // (blelloch_scan_downsweep.comp:68-75,83-97)
    if (n < 1024) clear_last = 1;
    const int total_levels = ilog2(n);
    // top-down groups of <=10 levels; the top group takes the remainder so lower groups are full 1024-blocks
    int top_levels = total_levels % 10;
    int done = 0;
    bool first = true;
    if (total_levels == 0)
    {
        downsweep_strided_kernel<<<dim3(1, blocks), 1024, 0, s>>>(buf, 1, 1, n, clear_last, 0);
        return check_launch();
    }
    while (done < total_levels)
    {
        const int levels = (first && top_levels != 0) ? top_levels : 10;
        const int remaining_below = total_levels - done - levels;     // levels handled by later launches
        const uint64_t stride = 1ull << remaining_below;
        const uint32_t count = (uint32_t) ((uint64_t) n >> remaining_below);
        const dim3 grid((count + 1023) / 1024, blocks);
        downsweep_strided_kernel<<<grid, 1024, 0, s>>>(buf, count, stride, n, first ? clear_last : 0, levels);
        VRENB200_TRY(check_launch());
        done += levels;
        first = false;
    }
    return VRENB200_OK;
}
