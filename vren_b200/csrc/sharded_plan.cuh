// The device-side PLAN of the multi-GPU sort (sharded_sort.cu) and the tables it fills, in a header of their own so that the
// same source also runs on the host: tests/cpp/plan_emulation.cpp executes plan_body() with tests/cpp/cta_emulator.hpp (256 OS
// threads for the 256 CUDA threads) and tests/test_plan_emulation.py compares every output with the numpy mirror
// vren_b200/dist.py::exchange_plan, which in turn is checked against the global stable sort by simulation (tests/test_dist_cpu.py).
#pragma once

#include "radix_internal.cuh"

namespace vrenb200 {

constexpr int kMaxRanks = 32;

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

inline size_t sym_recv_offset() { return align_up(sizeof(sym_header), 256); }
inline size_t sym_bytes(uint32_t capacity) { return sym_recv_offset() + 2 * align_up((size_t) capacity * 4, 256); }

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

// ---- PLAN -------------------------------------------------------------------------------------------------------------------------
// One CTA of 256 threads, thread d owns the value d of every digit.  Every rank runs it on the same all-gathered
// histograms and gets the same ranges, owners, segment layouts and rounds.
// (the kernel, sharded_sort.cu::plan_kernel, first waits for the histograms of all sources, then runs this body)
__device__ __forceinline__ void plan_body(sym_header* mine, sort_control* ctl_part, shard_params sp, seg_plan* plan, xfer_plan* xp,
                                          uint16_t* tile_seg, uint32_t* status)
{
    __shared__ unsigned long long s_cum[kRadix];     // inclusive prefix of the counts of the partition digit
    __shared__ uint32_t s_tiles_ex[kRadix + 1];      // exclusive prefix of the tiles per digit value
    __shared__ uint32_t s_total[kRadix];
    __shared__ uint32_t s_bounds[kMaxRanks + 1];
    __shared__ uint32_t s_round_digit[kMaxRanks][kMaxRounds + 1];
    __shared__ uint32_t s_warp[kRadix / 32];
    __shared__ unsigned long long s_warp64[kRadix / 32];
    __shared__ uint32_t s_pstar, s_error;
    const unsigned d = threadIdx.x, lane = d & 31, warp = d >> 5;

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
}

} // namespace vrenb200
