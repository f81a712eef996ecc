// Multi-GPU radix sort of key / value pairs (the reference is single-device — this is the north-star's sharded
// form of vren::radix_sort, reference entry: vren/vren/primitives/radix_sort.cpp:149-337).
//
// One context per rank (one process per GPU, or several ranks in one process); every rank owns a "symmetric" region that all
// ranks can address (peer-mapped over NVLink: torch symmetric memory, cudaIpc, or plain peer access), holding its receive
// buffers, the all-gathered digit histograms, per-segment digit histograms and a few flag words.  No NCCL on the data path and
// no host synchronisation: a call only enqueues work on the caller's stream and two private streams.
//
//   main stream   digit histograms (one read of the keys) -> PUBLISH them into every peer -> PLAN (one CTA per rank, same
//                 result on every rank): the partition digit p* = highest byte in which the keys differ at all, contiguous
//                 ranges of its 256 values per rank balanced by count, where every (source, digit) block lands in its owner's
//                 receive buffer, rounds -> local PARTITION of the shard by digit p* (one onesweep pass)
//   copy stream   per round: TRANSFER kernel — a few one-warp CTAs, each driving a ring of TMA bulk copies: global -> shared
//                 (local HBM) and shared -> global (the owner's receive buffer, a peer-mapped address over NVLink).  It needs
//                 no registers or load/store slots to speak of, so the SMs stay with the passes it overlaps; the local
//                 partition is laid out so that every block is 16-byte co-aligned with its destination; the last CTA raises
//                 the round's flag in every peer
//   sort stream   per round: wait for the round's flags of all sources -> digit histograms of the round's segments (one read of
//                 the received keys) -> scan -> p* SEGMENTED onesweep passes (radix_sort.cu, F_SEGMENTED) over the round's
//                 tile-aligned segments; the last pass writes the segments back to back into the output.
// The receive buffer is laid out by digit, inside a digit by source rank, inside a source in input order: concatenating the
// ranks' outputs gives the stable sort of the concatenated input.  Transfers of round k+1 overlap the sorting of round k
// (NVLink-bound against HBM-bound work).  Per pair and GPU: 4 (histogram) + 16 (partition) + 16 (transfer) + 4 (segment
// histograms) + 16 p* bytes of HBM traffic and 8 (G-1)/G bytes each way over NVLink.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>

#include "radix_internal.cuh"

namespace vrenb200 {
namespace {

constexpr int kMaxRanks = 32;      // [[plan-defs-a]] (this line, the structs below and the body of plan_kernel are also compiled for the host: tests/test_plan_emulation.py)
// The transfer kernel is NVLink-bound and must leave the SMs to the segmented passes it overlaps with: one warp per CTA drives a
// ring of TMA bulk copies (stages x 2 x kXferChunk x 4 bytes of shared memory).
constexpr int kXferChunk = 2048;          // pairs per stage: 8 KB of keys + 8 KB of values
constexpr int kXferMaxStages = 13;        // 13 x 16 KB = 208 KB: the CTA then has its SM to itself
// measured at N = 4 (profiles/r2j_bench_n4_exclusive.log): 24 CTAs with the whole shared memory of their SM each (nothing else
// fits next to them: the passes see uniform SMs, and a tile that crawls next to a copy no longer holds up the look-back chain
// of all the tiles behind it) 5.94 ms per sort; 32 CTAs of 96 KB next to one pass CTA each 6.14; 16 / 32 exclusive CTAs 6.36 / 6.29
constexpr int kXferDefaultStages = 13;
constexpr int kXferDefaultCtas = 24;
constexpr size_t xfer_smem_bytes(uint32_t stages) { return (size_t) stages * 2 * kXferChunk * sizeof(uint32_t) + 256; }
constexpr uint32_t kPartSlack = 1024;     // the local partition leaves up to 3 pairs of padding in front of every digit

// [[plan-defs-b-begin]]
// ---- symmetric region ---------------------------------------------------------------------------------------------------------
struct sym_header
{
    uint32_t hist_ready[kMaxRanks];               // [source]: epoch of the histograms found in hist_all[source]
    uint32_t round_ready[kMaxRounds][kMaxRanks];  // [round][source]: epoch of the last completed transfer
    uint32_t done[kMaxRanks];                     // [rank]: that rank has finished the call of this epoch (its buffers are free)
    uint32_t _pad[kMaxRanks];
    uint32_t hist_all[kMaxRanks][kPasses][kRadix];      // digit counts of every rank's shard
    uint32_t seg_hist[kRadix][kPasses - 1][kRadix];     // per value of the partition digit: counts, then offsets, of the lower digits
};

size_t sym_recv_offset() { return align_up(sizeof(sym_header), 256); }
size_t sym_bytes(uint32_t capacity) { return sym_recv_offset() + 2 * align_up((size_t) capacity * 4, 256); }

// ---- per-call tables written by the plan kernel (device memory) ---------------------------------------------------------------
struct xfer_plan
{
    uint32_t src_off[kRadix];      // where the pairs of a digit start in the locally partitioned shard
    uint32_t len[kRadix];          // how many this rank has
    uint32_t dst_off[kRadix];      // where this rank's block of the digit starts in the owner's receive buffer
    uint8_t owner[kRadix];
    uint8_t round_of[kRadix];
    uint32_t cum_pairs[kMaxRounds][kRadix];   // per round: inclusive prefix over the digit values of the pairs this rank sends
    uint32_t finished[kMaxRounds];            // CTAs of the round's transfer kernel that are done (reset by the last one)
};

struct shard_params
{
    uint32_t rank, world, rounds;
    uint32_t key_digits;       // bytes of the key that take part (4; 2 for the 16-bit bucket-sort key)
    uint32_t tile;             // tile of the segmented passes
    uint32_t cap_tiles;        // capacity of a receive buffer in tiles
    uint32_t round_bound;      // tiles one round may span (grid of the segmented launches)
    uint32_t epoch;
};

struct peer_table
{
    sym_header* hdr[kMaxRanks];
    uint32_t* recv_keys[kMaxRanks];
    uint32_t* recv_vals[kMaxRanks];
};
// [[plan-defs-b-end]]

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// epochs only grow; the comparison survives a wrap of the 32-bit counter
__device__ __forceinline__ bool epoch_reached(uint32_t seen, uint32_t want) { return (int32_t) (seen - want) >= 0; }

__device__ __forceinline__ void wait_epoch(const uint32_t* flag, uint32_t want)
{
    while (!epoch_reached(ld_acquire_sys(flag), want)) __nanosleep(200);
}

// ---- PUBLISH: this rank's raw digit counts into every peer's table, then the flag ------------------------------------------------
__global__ void __launch_bounds__(1024)
publish_histograms_kernel(const sort_control* ctl, peer_table peers, shard_params sp)
{
    const uint32_t c = (&ctl->hist[0][0])[threadIdx.x];
    for (uint32_t r = 0; r < sp.world; r++) (&peers.hdr[r]->hist_all[sp.rank][0][0])[threadIdx.x] = c;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < sp.world) st_release_sys(&peers.hdr[threadIdx.x]->hist_ready[sp.rank], sp.epoch);
}

// ---- PLAN -------------------------------------------------------------------------------------------------------------------------
// One CTA of 256 threads, thread d owns the value d of every digit.  Every rank runs it on the same all-gathered
// histograms and gets the same ranges, owners, segment layouts and rounds.
__global__ void __launch_bounds__(kRadix)
plan_kernel(sym_header* mine, sort_control* ctl_part, shard_params sp, seg_plan* plan, xfer_plan* xp, uint16_t* tile_seg,
            uint32_t* status)
{
    // [[plan-body-begin]]
    __shared__ unsigned long long s_cum[kRadix];     // inclusive prefix of the counts of the partition digit
    __shared__ uint32_t s_tiles_ex[kRadix + 1];      // exclusive prefix of the tiles per digit value
    __shared__ uint32_t s_total[kRadix];
    __shared__ uint32_t s_bounds[kMaxRanks + 1];
    __shared__ uint32_t s_round_digit[kMaxRanks][kMaxRounds + 1];
    __shared__ uint32_t s_warp[kRadix / 32];
    __shared__ unsigned long long s_warp64[kRadix / 32];
    __shared__ uint32_t s_pstar, s_error;
    const unsigned d = threadIdx.x, lane = d & 31, warp = d >> 5;

    if (d < sp.world) wait_epoch(&mine->hist_ready[d], sp.epoch);
    if (d == 0) s_error = 0;
    __syncthreads();

    // the partition digit: the highest byte (of those that take part) in which the keys differ at all
    uint32_t pstar = 0;
    for (int p = (int) sp.key_digits - 1; p >= 0; p--)
    {
        uint32_t t = 0;
        for (uint32_t s = 0; s < sp.world; s++) t += mine->hist_all[s][p][d];
        if (__syncthreads_count(t != 0) >= 2)
        {
            pstar = (uint32_t) p;
            break;
        }
    }
    unsigned long long total64 = 0;
    for (uint32_t s = 0; s < sp.world; s++) total64 += mine->hist_all[s][pstar][d];
    // a digit value holds at most capacity pairs if the plan is to fit at all; saturate (the capacity check below fails then)
    const uint32_t total = total64 > 0xFFFFFFF0ull ? 0xFFFFFFF0u : (uint32_t) total64;
    s_total[d] = total;
    // inclusive prefix (64-bit: up to 32 x 2^30 pairs)
    unsigned long long inc = total64;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1)
    {
        const unsigned long long t = __shfl_up_sync(kFullMask, inc, s);
        if (lane >= (unsigned) s) inc += t;
    }
    if (lane == 31) s_warp64[warp] = inc;
    __syncthreads();
    for (unsigned w = 0; w < warp; w++) inc += s_warp64[w];
    s_cum[d] = inc;
    __syncthreads();
    const unsigned long long grand = s_cum[kRadix - 1];

    // contiguous ranges of digit values per rank, balanced by count: boundary r is the digit boundary nearest to r/world of
    // the pairs (same rule as dist.py::plan_digit_ranges)
    if (d == 0)
    {
        s_bounds[0] = 0;
        for (uint32_t r = 1; r < sp.world; r++)
        {
            const unsigned long long target = (grand * r + sp.world - 1) / sp.world;
            uint32_t lo = 0, hi = kRadix;                    // first digit whose inclusive prefix reaches the target
            while (lo < hi)
            {
                const uint32_t mid = (lo + hi) / 2;
                if (s_cum[mid] < target) lo = mid + 1; else hi = mid;
            }
            uint32_t dd = lo < s_bounds[r - 1] ? s_bounds[r - 1] : lo;
            if (dd > kRadix - 1) dd = kRadix - 1;
            const unsigned long long before = dd > 0 ? s_cum[dd - 1] : 0ull, after = s_cum[dd];
            // (a range boundary inside digit dd goes to the nearer end of the digit)
            uint32_t cut = (target >= before ? target - before : 0ull) <= (after >= target ? after - target : 0ull) ? dd : dd + 1;
            if (cut < s_bounds[r - 1]) cut = s_bounds[r - 1];
            if (cut > kRadix) cut = kRadix;
            s_bounds[r] = cut;
        }
        s_bounds[sp.world] = kRadix;
    }
    __syncthreads();
    uint32_t owner = 0;
    for (uint32_t r = 1; r < sp.world; r++) owner += s_bounds[r] <= d;

    // padded layout: every segment starts on a tile boundary of its owner's receive buffer
    const uint32_t tiles = (uint32_t) (((unsigned long long) total + sp.tile - 1) / sp.tile);
    uint32_t tinc = tiles;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1)
    {
        const uint32_t t = __shfl_up_sync(kFullMask, tinc, s);
        if (lane >= (unsigned) s) tinc += t;
    }
    if (lane == 31) s_warp[warp] = tinc;
    __syncthreads();
    for (unsigned w = 0; w < warp; w++) tinc += s_warp[w];
    s_tiles_ex[d + 1] = tinc;
    if (d == 0) s_tiles_ex[0] = 0;
    __syncthreads();
    const uint32_t own_lo = s_bounds[owner], own_hi = s_bounds[owner + 1];
    const uint32_t first_tile = s_tiles_ex[d] - s_tiles_ex[own_lo];
    if (d < sp.world && s_tiles_ex[s_bounds[d + 1]] - s_tiles_ex[s_bounds[d]] > sp.cap_tiles) atomicOr(&s_error, 1u);   // capacity

    // rounds of every owner: `rounds` groups of consecutive segments with about the same number of tiles
    if (d < sp.world)
    {
        const uint32_t lo = s_bounds[d], hi = s_bounds[d + 1];
        const uint32_t base = s_tiles_ex[lo], all = s_tiles_ex[hi] - base;
        uint32_t dd = lo;
        s_round_digit[d][0] = lo;
        for (uint32_t k = 1; k < sp.rounds; k++)
        {
            // round k starts at the first segment that begins at or after k / rounds of the tiles (a segment that straddles
            // the mark stays with the round before: a round is at most one segment larger than its share)
            const uint32_t target = (uint32_t) (((unsigned long long) all * k + sp.rounds - 1) / sp.rounds);
            while (dd < hi && s_tiles_ex[dd] - base < target) dd++;
            s_round_digit[d][k] = dd;
        }
        s_round_digit[d][sp.rounds] = hi;
        for (uint32_t k = 0; k < sp.rounds; k++)
            if (s_tiles_ex[s_round_digit[d][k + 1]] - s_tiles_ex[s_round_digit[d][k]] > sp.round_bound) atomicOr(&s_error, 2u);
    }
    __syncthreads();
    uint32_t round_of = 0;
    for (uint32_t k = 1; k < sp.rounds; k++) round_of += s_round_digit[owner][k] <= d;

    // sender side: where my block of digit d goes.  The local partition pass scatters digit d to ctl_part->hist[pstar][d];
    // those offsets are re-laid here with up to 3 pairs of padding in front of every digit, so that a block starts at the
    // same offset modulo 16 bytes in the partitioned shard and in its owner's receive buffer (bulk copies need both aligned)
    __shared__ uint32_t s_len[kRadix], s_dst[kRadix];
    uint32_t before_me = 0;
    for (uint32_t s = 0; s < sp.rank; s++) before_me += mine->hist_all[s][pstar][d];
    const uint32_t my_len = mine->hist_all[sp.rank][pstar][d];
    const uint32_t dst_off = first_tile * sp.tile + before_me;
    s_len[d] = my_len;
    s_dst[d] = dst_off;
    xp->len[d] = my_len;
    xp->dst_off[d] = dst_off;
    xp->owner[d] = (uint8_t) owner;
    xp->round_of[d] = (uint8_t) round_of;
    __syncthreads();
    if (d == 0)
    {
        uint32_t cur = 0;
        for (uint32_t i = 0; i < kRadix; i++)
        {
            cur += (s_dst[i] - cur) & 3u;
            xp->src_off[i] = cur;
            ctl_part->hist[pstar][i] = cur;
            cur += s_len[i];
        }
    }
    for (uint32_t k = 0; k < sp.rounds; k++)
    {
        uint32_t linc = round_of == k ? my_len : 0u;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1)
        {
            const uint32_t t = __shfl_up_sync(kFullMask, linc, s);
            if (lane >= (unsigned) s) linc += t;
        }
        __syncthreads();
        if (lane == 31) s_warp[warp] = linc;
        __syncthreads();
        for (unsigned w = 0; w < warp; w++) linc += s_warp[w];
        xp->cum_pairs[k][d] = linc;
    }

    // receiver side: my segments
    const uint32_t my_lo = s_bounds[sp.rank], my_hi = s_bounds[sp.rank + 1];
    seg_desc sd;
    sd.first_tile = first_tile;
    sd.len = total;
    sd.out_start = (uint32_t) (s_cum[d] - total64 - (my_lo > 0 ? s_cum[my_lo - 1] : 0ull));   // meaningful for my digits only
    sd._pad = 0;
    plan->seg[d] = sd;
    __syncthreads();
    if (d == 0)
    {
        s_pstar = pstar;
        const uint32_t my_tiles = s_tiles_ex[my_hi] - s_tiles_ex[my_lo];
        plan->pstar = pstar;
        plan->error = s_error;
        plan->num_tiles = my_tiles;
        plan->out_count = (uint32_t) ((my_hi > 0 ? s_cum[my_hi - 1] : 0ull) - (my_lo > 0 ? s_cum[my_lo - 1] : 0ull));
        plan->digit_lo = my_lo;
        plan->digit_hi = my_hi;
        plan->rounds = sp.rounds;
        for (uint32_t k = 0; k <= sp.rounds; k++)
        {
            plan->round_digit[k] = s_round_digit[sp.rank][k];
            plan->round_tile[k] = s_tiles_ex[s_round_digit[sp.rank][k]] - s_tiles_ex[my_lo];
        }
        status[0] = s_error;
        status[1] = s_error ? 0u : plan->out_count;
        status[2] = my_lo;
        status[3] = my_hi;
        status[4] = pstar;
    }
    // which segment every tile of my receive buffer belongs to
    if (s_error == 0)
    {
        const uint32_t my_tiles = s_tiles_ex[my_hi] - s_tiles_ex[my_lo], base = s_tiles_ex[my_lo];
        for (uint32_t t = d; t < my_tiles; t += kRadix)
        {
            uint32_t lo = my_lo, hi = my_hi;        // last digit whose first tile is <= t and that has tiles
            while (hi - lo > 1)
            {
                const uint32_t mid = (lo + hi) / 2;
                if (s_tiles_ex[mid] - base <= t) lo = mid; else hi = mid;
            }
            tile_seg[t] = (uint16_t) lo;
        }
    }
    // [[plan-body-end]]
}

// ---- TRANSFER of one round --------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_addr(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void mbar_init_(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}

// One warp per CTA.  The pairs this rank sends in the round are split evenly, in digit order, over the CTAs.  Inside a
// (digit, source) block the 16-byte-aligned body moves in chunks through a ring in shared memory: lane 0 keeps
// kXferStages - 2 chunk loads in flight (cp.async.bulk global -> shared, local HBM) and stores every landed chunk with
// cp.async.bulk shared -> global into the owner's receive buffer (local, or a peer's over NVLink); the up to 3 pairs before
// and after the aligned body go through registers.
__global__ void __launch_bounds__(32, 1)
transfer_round_kernel(const uint32_t* __restrict__ part_keys, const uint32_t* __restrict__ part_vals, const seg_plan* plan,
                      xfer_plan* xp, peer_table peers, shard_params sp, uint32_t round, uint32_t kXferStages)
{
    extern __shared__ __align__(128) unsigned char xfer_smem[];
    uint32_t* ring_k = reinterpret_cast<uint32_t*>(xfer_smem);                       // [stages][chunk]
    uint32_t* ring_v = ring_k + kXferStages * kXferChunk;
    uint64_t* full = reinterpret_cast<uint64_t*>(ring_v + kXferStages * kXferChunk); // one mbarrier per stage
    __shared__ uint32_t s_cum[kRadix];
    __shared__ uint32_t* s_dst_k[kXferMaxStages];
    __shared__ uint32_t* s_dst_v[kXferMaxStages];
    __shared__ uint32_t s_count[kXferMaxStages];
    const unsigned lane = threadIdx.x;
    if (plan->error == 0)
    {
        for (uint32_t i = lane; i < kRadix; i += 32) s_cum[i] = xp->cum_pairs[round][i];
        if (lane == 0)
        {
            for (uint32_t s = 0; s < kXferStages; s++) mbar_init_(&full[s], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        const uint32_t total = s_cum[kRadix - 1];
        const uint32_t begin = (uint32_t) ((unsigned long long) total * blockIdx.x / gridDim.x);
        const uint32_t end = (uint32_t) ((unsigned long long) total * (blockIdx.x + 1) / gridDim.x);
        uint32_t issued = 0, stored = 0;      // chunks (lane 0)
        auto store_one = [&]() {
            const uint32_t st = stored % kXferStages;
            mbar_wait_(&full[st], (stored / kXferStages) & 1u);
            bulk_s2g(s_dst_k[st], ring_k + st * kXferChunk, s_count[st] * 4u);
            bulk_s2g(s_dst_v[st], ring_v + st * kXferChunk, s_count[st] * 4u);
            bulk_commit();
            stored++;
        };
        uint32_t pos = begin;
        while (pos < end)
        {
            uint32_t lo = 0, hi = kRadix - 1;            // the digit of pair `pos`: first value whose inclusive prefix exceeds it
            while (lo < hi)
            {
                const uint32_t mid = (lo + hi) / 2;
                if (s_cum[mid] <= pos) lo = mid + 1; else hi = mid;
            }
            const uint32_t dgt = lo;
            const uint32_t first = dgt > 0 ? s_cum[dgt - 1] : 0u;
            const uint32_t piece_end = s_cum[dgt] < end ? s_cum[dgt] : end;
            const uint32_t e0 = pos - first, e1 = piece_end - first;      // pairs [e0, e1) of this rank's block of the digit
            const uint32_t owner = xp->owner[dgt], dst_off = xp->dst_off[dgt], src_off = xp->src_off[dgt];
            const uint32_t* src_k = part_keys + src_off;
            const uint32_t* src_v = part_vals + src_off;
            uint32_t* dst_k = peers.recv_keys[owner] + dst_off;
            uint32_t* dst_v = peers.recv_vals[owner] + dst_off;
            // aligned body [b0, b1): (dst_off + e) % 4 == 0 at b0 (and (src_off + e) % 4 == 0 too: the plan made them congruent)
            uint32_t b0 = e0 + ((4u - ((dst_off + e0) & 3u)) & 3u);
            if (b0 > e1) b0 = e1;
            const uint32_t b1 = b0 + ((e1 - b0) & ~3u);
            // head and tail through registers
            for (uint32_t e = e0 + lane; e < b0; e += 32) { dst_k[e] = src_k[e]; dst_v[e] = src_v[e]; }
            for (uint32_t e = b1 + lane; e < e1; e += 32) { dst_k[e] = src_k[e]; dst_v[e] = src_v[e]; }
            if (lane == 0)
            {
                for (uint32_t c0 = b0; c0 < b1; c0 += kXferChunk)
                {
                    const uint32_t count = b1 - c0 < (uint32_t) kXferChunk ? b1 - c0 : (uint32_t) kXferChunk;
                    // the stage is free once the store that last used it has read it: stores complete in order, and all but
                    // the two most recent chunks have been stored before a new load is issued
                    while (stored + (kXferStages - 2) < issued) store_one();
                    if (issued >= kXferStages) bulk_wait_read<1>();
                    const uint32_t st = issued % kXferStages;
                    s_dst_k[st] = dst_k + c0;
                    s_dst_v[st] = dst_v + c0;
                    s_count[st] = count;
                    mbar_expect_tx_(&full[st], count * 8u);
                    bulk_g2s(ring_k + st * kXferChunk, src_k + c0, count * 4u, &full[st]);
                    bulk_g2s(ring_v + st * kXferChunk, src_v + c0, count * 4u, &full[st]);
                    issued++;
                }
            }
            pos = piece_end;
        }
        if (lane == 0)
        {
            while (stored < issued) store_one();
            bulk_wait_all();
            asm volatile("fence.proxy.async;" ::: "memory");   // the bulk stores went through the async proxy: order them before the flag
        }
        __syncwarp();
    }
    // completion: the last CTA of the round tells every peer that this rank's part of the round has arrived
    __threadfence_system();
    __syncwarp();
    uint32_t last = 0;
    if (lane == 0) last = atomicAdd(&xp->finished[round], 1u) == gridDim.x - 1;
    last = __shfl_sync(kFullMask, last, 0);
    if (last)
    {
        __threadfence_system();
        if (lane < sp.world) st_release_sys(&peers.hdr[lane]->round_ready[round][sp.rank], sp.epoch);
        if (lane == 0) xp->finished[round] = 0;
    }
}

// ---- digit histograms of the round's segments (receiver side) --------------------------------------------------------------------
// One read of the received keys: per segment (value of the partition digit) the 256-bin histograms of the digits below it.
// Conflict-free columns as in radix_histogram_columns_kernel (every lane owns a column of every counter); a CTA takes a
// contiguous range of the round's tiles and flushes its counters to the segment's table whenever the segment changes.
constexpr int kSegHistThreads = 1024;
constexpr size_t kSegHistSmem = (size_t) (kPasses - 1) * kRadix * 32 * sizeof(uint32_t);

__global__ void __launch_bounds__(kSegHistThreads, 1)
segment_histograms_kernel(const uint32_t* __restrict__ recv_keys, sym_header* mine, const seg_plan* plan, const uint16_t* __restrict__ tile_seg,
                          uint32_t tile, uint32_t round)
{
    extern __shared__ __align__(128) uint32_t s_cols[];   // [3][256][32]
    const uint32_t pstar = plan->pstar;
    if (plan->error != 0 || pstar == 0) return;
    const uint32_t t0 = plan->round_tile[round], t1 = plan->round_tile[round + 1];
    const uint32_t begin = t0 + (uint32_t) ((unsigned long long) (t1 - t0) * blockIdx.x / gridDim.x);
    const uint32_t end = t0 + (uint32_t) ((unsigned long long) (t1 - t0) * (blockIdx.x + 1) / gridDim.x);
    if (begin >= end) return;
    const unsigned tid = threadIdx.x, lane = tid & 31;
    for (uint32_t i = tid; i < (kPasses - 1) * kRadix * 32; i += kSegHistThreads) s_cols[i] = 0;
    __syncthreads();
    uint32_t* col = s_cols + lane;
    auto count = [&](uint32_t k) {
        atomicAdd(&col[(0 * kRadix + (k & 0xFF)) * 32], 1u);
        if (pstar > 1) atomicAdd(&col[(1 * kRadix + ((k >> 8) & 0xFF)) * 32], 1u);
        if (pstar > 2) atomicAdd(&col[(2 * kRadix + ((k >> 16) & 0xFF)) * 32], 1u);
    };
    auto flush = [&](uint32_t seg) {
        __syncthreads();
        if (tid < (kPasses - 1) * kRadix)
        {
            uint32_t sum = 0;
#pragma unroll 8
            for (uint32_t k = 0; k < 32; k++)
            {
                const uint32_t idx = tid * 32 + ((k + tid) & 31);
                sum += s_cols[idx];
                s_cols[idx] = 0;
            }
            if (sum != 0) atomicAdd(&mine->seg_hist[seg][0][0] + tid, sum);
        }
        __syncthreads();
    };
    uint32_t cur_seg = tile_seg[begin];
    for (uint32_t t = begin; t < end; t++)
    {
        const uint32_t seg = tile_seg[t];
        if (seg != cur_seg)
        {
            flush(cur_seg);
            cur_seg = seg;
        }
        const seg_desc sd = plan->seg[seg];
        const uint32_t left = sd.len - (t - sd.first_tile) * tile;
        const uint32_t valid = left < tile ? left : tile;
        const uint4* k4 = reinterpret_cast<const uint4*>(recv_keys + (size_t) t * tile);
        const uint32_t n4 = valid / 4;
        uint32_t i = tid;
        for (; i + kSegHistThreads < n4; i += 2 * kSegHistThreads)
        {
            const uint4 a = ldg_stream_u4(k4 + i), b = ldg_stream_u4(k4 + i + kSegHistThreads);
            count(a.x); count(a.y); count(a.z); count(a.w);
            count(b.x); count(b.y); count(b.z); count(b.w);
        }
        for (; i < n4; i += kSegHistThreads)
        {
            const uint4 a = ldg_stream_u4(k4 + i);
            count(a.x); count(a.y); count(a.z); count(a.w);
        }
        for (uint32_t e = n4 * 4 + tid; e < valid; e += kSegHistThreads) count(recv_keys[(size_t) t * tile + e]);
    }
    flush(cur_seg);
}

// ---- small stream-ordering kernels ----------------------------------------------------------------------------------------------
__global__ void wait_round_kernel(const sym_header* mine, shard_params sp, uint32_t round)
{
    if (threadIdx.x < sp.world) wait_epoch(&mine->round_ready[round][threadIdx.x], sp.epoch);
}
__global__ void wait_peers_done_kernel(const sym_header* mine, shard_params sp)
{
    if (threadIdx.x < sp.world) wait_epoch(&mine->done[threadIdx.x], sp.epoch - 1);
}
__global__ void signal_done_kernel(peer_table peers, shard_params sp)
{
    __threadfence_system();
    if (threadIdx.x < sp.world) st_release_sys(&peers.hdr[threadIdx.x]->done[sp.rank], sp.epoch);
}

// exclusive scan of the lower-digit counts of the round's segments (grid: 256 segments x 3 digits)
__global__ void __launch_bounds__(kRadix)
scan_segment_histograms_kernel(sym_header* mine, const seg_plan* plan, uint32_t round)
{
    __shared__ uint32_t s_warp[kRadix / 32];
    const uint32_t seg = blockIdx.x / (kPasses - 1), p = blockIdx.x % (kPasses - 1);
    if (plan->error != 0 || p >= plan->pstar || seg < plan->round_digit[round] || seg >= plan->round_digit[round + 1]) return;
    uint32_t* h = mine->seg_hist[seg][p];
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

// p* == 0 (all keys agree above their lowest byte, or there is nothing to sort): the exchange has sorted; compact the segments
__global__ void __launch_bounds__(256)
compact_segments_kernel(const uint32_t* __restrict__ recv_keys, const uint32_t* __restrict__ recv_vals, const seg_plan* plan,
                        uint32_t tile, uint32_t* out_keys, uint32_t* out_vals)
{
    if (plan->error != 0 || plan->pstar != 0) return;
    for (uint32_t seg = plan->digit_lo; seg < plan->digit_hi; seg++)
    {
        const seg_desc sd = plan->seg[seg];
        const size_t src = (size_t) sd.first_tile * tile;
        for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < sd.len; i += gridDim.x * 256u)
        {
            out_keys[sd.out_start + i] = recv_keys[src + i];
            out_vals[sd.out_start + i] = recv_vals[src + i];
        }
    }
}

} // namespace
} // namespace vrenb200

using namespace vrenb200;

struct vrenb200_sharded_sort
{
    // releases whatever create() has made so far (null handles are skipped): used by its error paths and by destroy()
    void release()
    {
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (sort_stream) cudaStreamDestroy(sort_stream);
        for (cudaEvent_t* e : {&ev_start, &ev_part, &ev_copy, &ev_sort})
            if (*e) cudaEventDestroy(*e);
        copy_stream = sort_stream = nullptr;
        ev_start = ev_part = ev_copy = ev_sort = nullptr;
    }
    shard_params sp;
    peer_table peers;
    uint32_t max_n, capacity;
    sort_options opt;
    const sort_variant* var_seg;
    uint32_t xfer_ctas, xfer_stages;
    cudaStream_t copy_stream = nullptr, sort_stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_part = nullptr, ev_copy = nullptr, ev_sort = nullptr;
    int device;
    // carved from the caller's local scratch
    uint32_t *part_keys, *part_vals, *alt_keys, *alt_vals, *out_keys, *out_vals;
    sort_control* ctl_part;
    sort_control* ctl_seg;
    uint32_t* tickets;       // [rounds][3]
    seg_plan* plan;
    xfer_plan* xp;
    uint16_t* tile_seg;
    uint32_t* status;
};

namespace {

// usable tiles of a receive buffer and the tile itself follow from the capacity (the kernel variant a sort of that size takes)
const sort_variant& segment_variant(uint32_t capacity, const sort_options& opt) { return pick_variant(capacity, LAYOUT_SOA, opt); }

size_t seg_control_bytes(uint32_t cap_tiles)
{
    return align_up(sizeof(sort_control) + (size_t) (kPasses - 1) * cap_tiles * kRadix * sizeof(uint32_t), 256);
}

} // namespace

extern "C" size_t vrenb200_sharded_sort_symmetric_bytes(uint32_t capacity) { return sym_bytes(capacity); }

extern "C" size_t vrenb200_sharded_sort_local_bytes(uint32_t max_n, uint32_t capacity, const vrenb200_sort_config* cfg)
{
    const sort_options opt = resolve_options(cfg);
    const uint32_t tile = segment_variant(capacity, opt).tile;
    const uint32_t cap_tiles = capacity / tile;
    size_t b = 0;
    b += 2 * align_up(((size_t) max_n + kPartSlack) * 4, 256);  // locally partitioned shard (digits padded to their destinations' alignment)
    b += 4 * align_up((size_t) capacity * 4, 256);            // ping-pong partner of the receive buffers, output
    b += control_bytes(max_n) + seg_control_bytes(cap_tiles);
    b += align_up(sizeof(uint32_t) * kMaxRounds * (kPasses - 1), 256) + align_up(sizeof(seg_plan), 256) + align_up(sizeof(xfer_plan), 256);
    b += align_up((size_t) cap_tiles * sizeof(uint16_t) + 2, 256) + 256;
    return b;
}

extern "C" int vrenb200_sharded_sort_create(vrenb200_sharded_sort** out, uint32_t rank, uint32_t world, uint32_t max_n,
                                            uint32_t capacity, uint32_t rounds, void* const* peer_regions, void* local,
                                            size_t local_bytes, const vrenb200_sort_config* cfg)
{
    if (out == nullptr || peer_regions == nullptr || local == nullptr) return VRENB200_EINVAL_ARG;
    if (world == 0 || world > (uint32_t) kMaxRanks || rank >= world || rounds == 0 || rounds > (uint32_t) kMaxRounds) return VRENB200_EINVAL_ARG;
    if (max_n >= (1u << 30) || capacity >= (1u << 30)) return VRENB200_ELIMIT;
    if (local_bytes < vrenb200_sharded_sort_local_bytes(max_n, capacity, cfg)) return VRENB200_ESCRATCH;
    if (reinterpret_cast<uintptr_t>(local) & 255) return VRENB200_EALIGN;
    vrenb200_sharded_sort* c = new (std::nothrow) vrenb200_sharded_sort();
    if (c == nullptr) return VRENB200_ECUDA;
    c->opt = resolve_options(cfg);
    c->var_seg = &segment_variant(capacity, c->opt);
    c->max_n = max_n;
    c->capacity = capacity;
    c->sp.rank = rank;
    c->sp.world = world;
    c->sp.rounds = rounds;
    c->sp.key_digits = kPasses;
    c->sp.tile = c->var_seg->tile;
    c->sp.cap_tiles = capacity / c->var_seg->tile;
    // a round is about 1/rounds of the received tiles plus one segment; the segmented launches get this many CTAs
    c->sp.round_bound = rounds == 1 ? c->sp.cap_tiles : std::min(c->sp.cap_tiles, (c->sp.cap_tiles + rounds - 1) / rounds + c->sp.cap_tiles / 16 + 1);
    c->sp.epoch = 0;
    for (uint32_t r = 0; r < world; r++)
    {
        char* base = static_cast<char*>(peer_regions[r]);
        if (base == nullptr || (reinterpret_cast<uintptr_t>(base) & 255)) { delete c; return VRENB200_EALIGN; }
        c->peers.hdr[r] = reinterpret_cast<sym_header*>(base);
        c->peers.recv_keys[r] = reinterpret_cast<uint32_t*>(base + sym_recv_offset());
        c->peers.recv_vals[r] = reinterpret_cast<uint32_t*>(base + sym_recv_offset() + align_up((size_t) capacity * 4, 256));
    }
    scratch_carver carve(local);
    c->part_keys = carve.take<uint32_t>((size_t) max_n + kPartSlack);
    c->part_vals = carve.take<uint32_t>((size_t) max_n + kPartSlack);
    c->alt_keys = carve.take<uint32_t>(capacity);
    c->alt_vals = carve.take<uint32_t>(capacity);
    c->out_keys = carve.take<uint32_t>(capacity);
    c->out_vals = carve.take<uint32_t>(capacity);
    c->ctl_part = reinterpret_cast<sort_control*>(carve.take<char>(control_bytes(max_n)));
    c->ctl_seg = reinterpret_cast<sort_control*>(carve.take<char>(seg_control_bytes(c->sp.cap_tiles)));
    c->tickets = carve.take<uint32_t>(kMaxRounds * (kPasses - 1));
    c->plan = carve.take<seg_plan>(1);
    c->xp = carve.take<xfer_plan>(1);
    c->tile_seg = carve.take<uint16_t>(c->sp.cap_tiles + 1);
    c->status = carve.take<uint32_t>(8);
    c->xfer_ctas = kXferDefaultCtas;   // one warp each; 32 keep ~2 MB of copies in flight
    if (const char* e = std::getenv("VRENB200_XFER_CTAS"))     // measurement knob, read when the context is created
    {
        const long v = std::strtol(e, nullptr, 10);
        if (v >= 1 && v <= 4 * kNumSMs) c->xfer_ctas = (uint32_t) v;
    }
    c->xfer_stages = kXferDefaultStages;
    if (const char* e = std::getenv("VRENB200_XFER_STAGES"))
    {
        const long v = std::strtol(e, nullptr, 10);
        if (v >= 3 && v <= kXferMaxStages) c->xfer_stages = (uint32_t) v;
    }
    // the copy stream outranks the sort stream: when an SM frees room, a waiting transfer CTA takes it before the next pass CTA
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaGetDevice(&c->device) != cudaSuccess || cudaStreamCreateWithPriority(&c->copy_stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaStreamCreateWithPriority(&c->sort_stream, cudaStreamNonBlocking, prio_lo) != cudaSuccess ||
        cudaEventCreate(&c->ev_start) != cudaSuccess || cudaEventCreate(&c->ev_part) != cudaSuccess ||
        cudaEventCreate(&c->ev_copy) != cudaSuccess || cudaEventCreate(&c->ev_sort) != cudaSuccess)
    {
        c->release();
        delete c;
        return VRENB200_ECUDA;
    }
    // every kernel a call may launch is loaded now: CUDA loads modules lazily and a load can wait for running kernels, which
    // must not happen between kernels that wait for each other (several ranks in one process would deadlock)
    {
        cudaFuncAttributes attr;
        int st = preload_sort_kernels(LAYOUT_SOA, false, c->opt);
        if (st == VRENB200_OK) st = preload_sort_kernels(LAYOUT_SOA, true, c->opt);
        const void* kernels[] = {(const void*) publish_histograms_kernel, (const void*) plan_kernel, (const void*) transfer_round_kernel,
                                 (const void*) wait_round_kernel, (const void*) wait_peers_done_kernel, (const void*) signal_done_kernel,
                                 (const void*) scan_segment_histograms_kernel, (const void*) compact_segments_kernel,
                                 (const void*) segment_histograms_kernel};
        for (const void* k : kernels)
            if (st == VRENB200_OK) st = check_cuda(cudaFuncGetAttributes(&attr, k));
        if (st == VRENB200_OK) st = check_cuda(cudaFuncSetAttribute(transfer_round_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) xfer_smem_bytes(kXferMaxStages)));
        if (st == VRENB200_OK) st = check_cuda(cudaFuncSetAttribute(segment_histograms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kSegHistSmem));
        if (st != VRENB200_OK)
        {
            c->release();
            delete c;
            return st;
        }
    }
    // the flag words, histograms and completion counters start at zero (epoch 0 = "nothing yet"); the caller makes sure every
    // rank has returned from create before any rank sorts (one barrier)
    if (cudaMemset(c->peers.hdr[rank], 0, sizeof(sym_header)) != cudaSuccess || cudaMemset(c->xp, 0, sizeof(xfer_plan)) != cudaSuccess)
    {
        c->release();
        delete c;
        return VRENB200_ECUDA;
    }
    *out = c;
    return VRENB200_OK;
}

extern "C" void vrenb200_sharded_sort_destroy(vrenb200_sharded_sort* c)
{
    if (c == nullptr) return;
    cudaStreamSynchronize(c->copy_stream);
    cudaStreamSynchronize(c->sort_stream);
    c->release();
    delete c;
}

extern "C" const uint32_t* vrenb200_sharded_sort_out_keys(const vrenb200_sharded_sort* c) { return c ? c->out_keys : nullptr; }
extern "C" const uint32_t* vrenb200_sharded_sort_out_values(const vrenb200_sharded_sort* c) { return c ? c->out_vals : nullptr; }
extern "C" const uint32_t* vrenb200_sharded_sort_status(const vrenb200_sharded_sort* c) { return c ? c->status : nullptr; }

// phases of the LAST call on this rank, from CUDA events on the three streams (call after the stream has been synchronised):
// ms_out[3] = {histograms + publish + plan + local partition, first transfer -> last transfer done, first -> last segmented
// pass}; the last two start together and overlap
extern "C" int vrenb200_sharded_sort_phases(const vrenb200_sharded_sort* c, float* ms_out)
{
    if (c == nullptr || ms_out == nullptr) return VRENB200_EINVAL_ARG;
    VRENB200_TRY(check_cuda(cudaEventElapsedTime(&ms_out[0], c->ev_start, c->ev_part)));
    VRENB200_TRY(check_cuda(cudaEventElapsedTime(&ms_out[1], c->ev_part, c->ev_copy)));
    return check_cuda(cudaEventElapsedTime(&ms_out[2], c->ev_part, c->ev_sort));
}

extern "C" int vrenb200_sharded_sort_pairs(vrenb200_sharded_sort* c, vrenb200_stream_t stream, const uint32_t* keys,
                                           const uint32_t* values, uint32_t n, int key_bits)
{
    if (c == nullptr || (n > 0 && (keys == nullptr || values == nullptr))) return VRENB200_EINVAL_ARG;
    if (n > c->max_n) return VRENB200_ELIMIT;
    if (key_bits != 32 && key_bits != 24 && key_bits != 16 && key_bits != 8) return VRENB200_EINVAL_ARG;
    if ((reinterpret_cast<uintptr_t>(keys) | reinterpret_cast<uintptr_t>(values)) & 15) return VRENB200_EALIGN;
    cudaStream_t s = as_stream(stream);
    shard_params sp = c->sp;
    sp.epoch = ++c->sp.epoch;
    sp.key_digits = (uint32_t) key_bits / 8;
    sym_header* mine = c->peers.hdr[sp.rank];
    const sort_variant& var_part = pick_variant(n, LAYOUT_SOA, c->opt);
    const sort_variant& var_seg = *c->var_seg;
    const uint32_t part_tiles = (uint32_t) (((size_t) n + var_part.tile - 1) / var_part.tile);
    uint32_t* lb_part = reinterpret_cast<uint32_t*>(c->ctl_part + 1);
    uint32_t* lb_seg = reinterpret_cast<uint32_t*>(c->ctl_seg + 1);
    const bool ticket = c->opt.tile_ids == VRENB200_TILE_IDS_TICKET;
    const int selftest = c->opt.ranking == VRENB200_RANKING_SELFTEST_REDO;

    // ---- main stream: histograms -> publish -> plan -> local partition ----
    VRENB200_TRY(check_cuda(cudaEventRecord(c->ev_start, s)));
    VRENB200_TRY(check_cuda(cudaMemsetAsync(c->ctl_part, 0, sizeof(sort_control) + (size_t) std::max(part_tiles, 1u) * kRadix * sizeof(uint32_t), s)));
    VRENB200_TRY(check_cuda(cudaMemsetAsync(c->ctl_seg, 0, sizeof(sort_control) + (size_t) sp.cap_tiles * kRadix * sizeof(uint32_t), s)));
    VRENB200_TRY(check_cuda(cudaMemsetAsync(c->tickets, 0, sizeof(uint32_t) * kMaxRounds * (kPasses - 1), s)));
    if (n > 0) VRENB200_TRY(launch_digit_histograms(s, keys, n, c->ctl_part));
    publish_histograms_kernel<<<1, kPasses * kRadix, 0, s>>>(c->ctl_part, c->peers, sp);
    VRENB200_TRY(check_launch());
    VRENB200_TRY(launch_scan_histograms(s, c->ctl_part, kPasses));
    plan_kernel<<<1, kRadix, 0, s>>>(mine, c->ctl_part, sp, c->plan, c->xp, c->tile_seg, c->status);
    VRENB200_TRY(check_launch());
    if (n > 0)
    {
        pass_params p{};
        p.keys_in = keys;
        p.vals_in = values;
        p.keys_out = c->part_keys;
        p.vals_out = c->part_vals;
        p.n = n;
        p.pass = 0;
        p.dyn_pass = &c->plan->pstar;
        p.lb_plane = 0;
        p.ctl = c->ctl_part;
        p.lookback = lb_part;
        p.num_tiles = part_tiles;
        p.ticket = ticket ? &c->ctl_part->tickets[0] : nullptr;
        p.selftest = selftest;
        VRENB200_TRY(var_part.launch(s, p, part_tiles, LAYOUT_SOA, false));
        if (var_part.redo) VRENB200_TRY(var_part.redo(s, p, LAYOUT_SOA, false));
    }
    VRENB200_TRY(check_cuda(cudaEventRecord(c->ev_part, s)));

    // ---- copy stream: the rounds' transfers, as soon as every peer is done with the previous call ----
    VRENB200_TRY(check_cuda(cudaStreamWaitEvent(c->copy_stream, c->ev_part, 0)));
    wait_peers_done_kernel<<<1, 32, 0, c->copy_stream>>>(mine, sp);
    VRENB200_TRY(check_launch());
    for (uint32_t k = 0; k < sp.rounds; k++)
    {
        transfer_round_kernel<<<c->xfer_ctas, 32, xfer_smem_bytes(c->xfer_stages), c->copy_stream>>>(c->part_keys, c->part_vals, c->plan, c->xp, c->peers, sp, k,
                                                                                                        c->xfer_stages);
        VRENB200_TRY(check_launch());
    }
    VRENB200_TRY(check_cuda(cudaEventRecord(c->ev_copy, c->copy_stream)));

    // ---- sort stream: per round, the segmented passes over what has arrived ----
    VRENB200_TRY(check_cuda(cudaStreamWaitEvent(c->sort_stream, c->ev_part, 0)));
    for (uint32_t k = 0; k < sp.rounds; k++)
    {
        wait_round_kernel<<<1, 32, 0, c->sort_stream>>>(mine, sp, k);
        VRENB200_TRY(check_launch());
        segment_histograms_kernel<<<kNumSMs, kSegHistThreads, kSegHistSmem, c->sort_stream>>>(c->peers.recv_keys[sp.rank], mine, c->plan, c->tile_seg,
                                                                                              sp.tile, k);
        VRENB200_TRY(check_launch());
        scan_segment_histograms_kernel<<<kRadix * (kPasses - 1), kRadix, 0, c->sort_stream>>>(mine, c->plan, k);
        VRENB200_TRY(check_launch());
        for (int i = 0; i < kPasses - 1; i++)
        {
            const bool even = (i & 1) == 0;
            pass_params p{};
            p.keys_in = even ? c->peers.recv_keys[sp.rank] : c->alt_keys;
            p.vals_in = even ? c->peers.recv_vals[sp.rank] : c->alt_vals;
            p.keys_out = even ? c->alt_keys : c->peers.recv_keys[sp.rank];
            p.vals_out = even ? c->alt_vals : c->peers.recv_vals[sp.rank];
            p.final_keys = c->out_keys;
            p.final_vals = c->out_vals;
            p.n = sp.cap_tiles * sp.tile;
            p.pass = i;
            p.lb_plane = i;
            p.clear_next_plane = i + 1 < kPasses - 1;
            p.ctl = c->ctl_seg;
            p.lookback = lb_seg;
            p.num_tiles = sp.cap_tiles;
            p.ticket = ticket ? &c->tickets[k * (kPasses - 1) + i] : nullptr;
            p.selftest = selftest;
            p.plan = c->plan;
            p.tile_seg = c->tile_seg;
            p.seg_hist = &mine->seg_hist[0][0][0];
            p.round = (int) k;
            VRENB200_TRY(var_seg.launch(c->sort_stream, p, sp.round_bound, LAYOUT_SOA, true));
            if (var_seg.redo) VRENB200_TRY(var_seg.redo(c->sort_stream, p, LAYOUT_SOA, true));
        }
    }
    compact_segments_kernel<<<kNumSMs * 4, 256, 0, c->sort_stream>>>(c->peers.recv_keys[sp.rank], c->peers.recv_vals[sp.rank], c->plan, sp.tile,
                                                                       c->out_keys, c->out_vals);
    VRENB200_TRY(check_launch());
    // my receive buffers and segment histograms are free again: clear the histograms, then tell every peer
    VRENB200_TRY(check_cuda(cudaMemsetAsync(&mine->seg_hist[0][0][0], 0, sizeof(mine->seg_hist), c->sort_stream)));
    signal_done_kernel<<<1, 32, 0, c->sort_stream>>>(c->peers, sp);
    VRENB200_TRY(check_launch());
    VRENB200_TRY(check_cuda(cudaEventRecord(c->ev_sort, c->sort_stream)));

    VRENB200_TRY(check_cuda(cudaStreamWaitEvent(s, c->ev_copy, 0)));
    VRENB200_TRY(check_cuda(cudaStreamWaitEvent(s, c->ev_sort, 0)));
    return VRENB200_OK;
}
