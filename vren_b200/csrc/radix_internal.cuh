// What radix_sort.cu shares with sharded_sort.cu: the control block of a sort, the parameter block of the onesweep
// pass kernel, the segment plan of a segmented (multi-GPU) sort and the launch helpers.
#pragma once

#include "common.cuh"

namespace vrenb200 {

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kPasses = 32 / kRadixBits;

enum { LAYOUT_KEYS = 0, LAYOUT_SOA = 1, LAYOUT_AOS = 2 };

// ---- control block carved from scratch --------------------------------------------------------------------------------------
struct sort_control
{
    uint32_t tickets[kPasses];            // TILE_IDS_TICKET: next tile id per look-back plane
    uint32_t order_violation;             // bit p: a lane-order check of the atomic ranking failed in look-back plane p
    uint32_t _pad[59];
    uint32_t hist[kPasses][kRadix];       // global digit counts, then exclusive offsets
    // followed by look-back words: [planes][tiles][kRadix]
};

// ---- segmented sort (the receive side of the multi-GPU sort) ----------------------------------------------------------------
// The receive buffer holds one segment per value of the partition digit, every segment starting on a tile boundary
// (start == first_tile * TILE), so that a tile never straddles two segments and its input address follows from its id alone.
// The segments are sorted independently on their `pstar` low digits; the last of those passes writes them back to back
// (out_start) into the output buffer.
constexpr int kMaxRounds = 8;

struct seg_desc
{
    uint32_t first_tile;   // start / TILE
    uint32_t len;          // pairs in the segment
    uint32_t out_start;    // where the segment starts in the compact output
    uint32_t _pad;
};

struct seg_plan
{
    uint32_t pstar;                        // index of the partition digit == number of low digits left to sort (0..3)
    uint32_t error;                        // != 0: the plan does not fit (capacity); every later kernel of the call returns at once
    uint32_t num_tiles;                    // tiles of the padded layout in this rank's receive buffer
    uint32_t out_count;                    // pairs this rank ends up with
    uint32_t digit_lo, digit_hi;           // this rank owns the values [digit_lo, digit_hi) of the partition digit
    uint32_t rounds;
    uint32_t _pad;
    uint32_t round_tile[kMaxRounds + 1];   // round k sorts the tiles [round_tile[k], round_tile[k + 1])
    uint32_t round_digit[kMaxRounds + 1];  // ... which hold the segments of the digits [round_digit[k], round_digit[k + 1])
    seg_desc seg[kRadix];                  // indexed by the value of the partition digit
};

// ---- parameter block of one launch of the onesweep pass kernel -----------------------------------------------------------------
struct pass_params
{
    const uint32_t* keys_in;
    uint32_t* keys_out;
    const uint32_t* vals_in;
    uint32_t* vals_out;
    uint32_t n;                  // elements of keys_in (segmented: the padded extent, num_tiles * TILE at most)
    int pass;                    // digit index, 0 = least significant byte
    const uint32_t* dyn_pass;    // != nullptr: the digit index is read from here (device memory) instead
    int lb_plane;                // which [num_tiles][256] plane of the look-back words (and which violation bit) this launch uses
    int clear_next_plane;        // every tile also clears its row of plane lb_plane + 1
    uint32_t* ticket;            // != nullptr: tile ids from this zero-initialised counter (start order) instead of the block index
    int selftest;                // verified-ranking kernels: report a failed check and write nothing (tests of the redo path)
    sort_control* ctl;
    uint32_t* lookback;
    uint32_t num_tiles;          // tiles per look-back plane
    // segmented launches only
    const seg_plan* plan;
    const uint16_t* tile_seg;    // partition-digit value of the segment every tile belongs to
    const uint32_t* seg_hist;    // [256 segments][3 digits][256]: exclusive digit offsets inside every segment
    uint32_t* final_keys;        // where the LAST pass (pass == pstar - 1) writes, compact
    uint32_t* final_vals;
    int round;
};

struct sort_variant
{
    const char* name;
    uint32_t tile;
    // launch of the main pass; redo != nullptr: by-construction repeat of the same pass, executed only if the ranking check of
    // the main pass raised ctl->order_violation
    int (*launch)(cudaStream_t, const pass_params&, uint32_t grid, int layout, bool segmented);
    int (*redo)(cudaStream_t, const pass_params&, int layout, bool segmented);
    // loads the kernels of this entry (main and redo) into the current context without launching anything
    int (*preload)(int layout, bool segmented);
};

// what VRENB200_RANKING_AUTO / VRENB200_TILE_IDS_AUTO resolve to (measured, profiles/r2d_sort_variant_sweep.log, 2^28 pairs):
// atomic ranking unchecked 71.1 Gpairs/s, one row in eight checked 70.1, every row checked 54.3 (its match work spills and
// fills the ALU pipe), ballot match 61.7; tickets instead of block indices cost 1.3 %
#define VRENB200_RANKING_DEFAULT VRENB200_RANKING_ATOMIC_SAMPLED
#define VRENB200_TILE_IDS_DEFAULT VRENB200_TILE_IDS_TICKET

struct sort_options
{
    int ranking;    // VRENB200_RANKING_*
    int tile_ids;   // VRENB200_TILE_IDS_*
    int variant;    // tuning builds: index into the variant table (0 = automatic)
    bool ranking_auto;   // the caller (and the environment) left the ranking to the library: pick_variant may choose by size
};

sort_options resolve_options(const vrenb200_sort_config* cfg);
const sort_variant& pick_variant(uint32_t n, int layout, const sort_options& opt);
size_t lookback_words(uint32_t n);
size_t control_bytes(uint32_t n);
int launch_digit_histograms(cudaStream_t s, const uint32_t* keys, uint32_t n, sort_control* ctl);
int launch_scan_histograms(cudaStream_t s, sort_control* ctl, int passes);
int preload_sort_kernels(int layout, bool segmented, const sort_options& opt);   // every kernel a sort with these options may launch
// small_sort.cu: the whole sort in one CTA (in place, vals may be nullptr), for 1 <= n <= single_cta_sort_max()
uint32_t single_cta_sort_max();
int launch_single_cta_sort(cudaStream_t s, uint32_t* keys, uint32_t* vals, uint32_t n, int first_pass, int num_passes);

} // namespace vrenb200
