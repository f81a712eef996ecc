// The device-side plan of the multi-GPU sort (the body of vren_b200/csrc/sharded_sort.cu::plan_kernel) executed on the host by
// cta_emulator.hpp.  The kernel's text and the tables it fills are taken from sharded_sort.cu itself: tests/test_plan_emulation.py
// extracts what lies between the [[plan-*]] markers into plan_defs.inc / plan_body.inc.  Built as a shared library and driven
// through ctypes by that test, which compares every output with the numpy mirror (vren_b200/dist.py::exchange_plan).
#include <cuda_runtime.h>      // host types only (cudaStream_t in radix_internal.cuh); nothing of the CUDA runtime is called

#include "cta_emulator.hpp"

#include "../../vren_b200/csrc/radix_internal.cuh"

#include <cstring>
#include <memory>

namespace vrenb200 {
namespace {

#include "plan_defs.inc"

// the kernel first waits for the histograms of all sources (an acquire load in a spin loop on the device); here they are in place
inline void wait_epoch(const uint32_t*, uint32_t) {}

void plan_body(sym_header* mine, sort_control* ctl_part, shard_params sp, seg_plan* plan, xfer_plan* xp, uint16_t* tile_seg, uint32_t* status)
{
#include "plan_body.inc"
}

} // namespace
} // namespace vrenb200

using namespace vrenb200;

// hist_all: uint32 [world][4][256] — digit counts of every rank's shard.  Outputs (caller-allocated):
//   scalars[8]      = {pstar, error, num_tiles, out_count, digit_lo, digit_hi, status[0], status[1]}
//   seg[256][3]     = {first_tile, len, out_start}
//   xfer[5][256]    = {src_off, len, dst_off, owner, round_of}
//   cum_pairs[rounds][256], round_digit[rounds + 1], round_tile[rounds + 1], tile_seg[cap_tiles], part_offsets[256] (= ctl_part->hist[pstar])
extern "C" int emu_plan(const uint32_t* hist_all, uint32_t world, uint32_t rank, uint32_t key_digits, uint32_t tile, uint32_t rounds,
                        uint32_t cap_tiles, uint32_t round_bound, uint32_t* scalars, uint32_t* seg, uint32_t* xfer, uint32_t* cum_pairs,
                        uint32_t* round_digit, uint32_t* round_tile, uint16_t* tile_seg, uint32_t* part_offsets)
{
    if (world == 0 || world > (uint32_t) kMaxRanks || rank >= world || rounds == 0 || rounds > (uint32_t) kMaxRounds) return 1;
    auto hdr = std::make_unique<sym_header>();
    std::memset(hdr.get(), 0, sizeof(sym_header));
    std::memcpy(&hdr->hist_all[0][0][0], hist_all, sizeof(uint32_t) * world * kPasses * kRadix);
    auto ctl = std::make_unique<sort_control>();
    std::memset(ctl.get(), 0, sizeof(sort_control));
    auto plan = std::make_unique<seg_plan>();
    std::memset(plan.get(), 0, sizeof(seg_plan));
    auto xp = std::make_unique<xfer_plan>();
    std::memset(xp.get(), 0, sizeof(xfer_plan));
    uint32_t status[8] = {0};
    shard_params sp{};
    sp.rank = rank;
    sp.world = world;
    sp.rounds = rounds;
    sp.key_digits = key_digits;
    sp.tile = tile;
    sp.cap_tiles = cap_tiles;
    sp.round_bound = round_bound;
    sp.epoch = 1;
    sym_header* h = hdr.get();
    sort_control* c = ctl.get();
    seg_plan* pl = plan.get();
    xfer_plan* x = xp.get();
    uint32_t* st = status;
    if (!cta_emu::run_cta(kRadix, [=]() { plan_body(h, c, sp, pl, x, tile_seg, st); })) return 2;
    scalars[0] = pl->pstar; scalars[1] = pl->error; scalars[2] = pl->num_tiles; scalars[3] = pl->out_count;
    scalars[4] = pl->digit_lo; scalars[5] = pl->digit_hi; scalars[6] = status[0]; scalars[7] = status[1];
    for (int d = 0; d < kRadix; d++)
    {
        seg[d * 3 + 0] = pl->seg[d].first_tile; seg[d * 3 + 1] = pl->seg[d].len; seg[d * 3 + 2] = pl->seg[d].out_start;
        xfer[0 * kRadix + d] = x->src_off[d]; xfer[1 * kRadix + d] = x->len[d]; xfer[2 * kRadix + d] = x->dst_off[d];
        xfer[3 * kRadix + d] = x->owner[d]; xfer[4 * kRadix + d] = x->round_of[d];
        part_offsets[d] = c->hist[pl->pstar][d];
    }
    for (uint32_t k = 0; k < rounds; k++) std::memcpy(cum_pairs + k * kRadix, x->cum_pairs[k], sizeof(uint32_t) * kRadix);
    for (uint32_t k = 0; k <= rounds; k++) { round_digit[k] = pl->round_digit[k]; round_tile[k] = pl->round_tile[k]; }
    return 0;
}

#ifdef PLAN_EMULATION_MAIN
// stand-alone run for ThreadSanitizer builds (a sanitized shared library cannot be loaded into the Python process): a few plans
// on pseudo-random histograms; the outputs are compared elsewhere, here only the absence of data races matters
#include <random>
int main()
{
    std::mt19937 rng(7);
    for (uint32_t world : {1u, 3u, 8u})
    {
        std::vector<uint32_t> h((size_t) world * kPasses * kRadix);
        for (auto& v : h) v = rng() % 5000u;
        for (uint32_t rounds : {1u, 4u})
        {
            std::vector<uint32_t> scalars(8), seg(256 * 3), xfer(5 * 256), cum(rounds * 256), rd(rounds + 1), rt(rounds + 1), part(256);
            std::vector<uint16_t> tile_seg(4096);
            const uint32_t rank = world - 1;
            if (emu_plan(h.data(), world, rank, 4, 4096, rounds, 4000, 4000, scalars.data(), seg.data(), xfer.data(), cum.data(), rd.data(),
                         rt.data(), tile_seg.data(), part.data()) != 0)
                return 1;
            std::printf("world %u rounds %u: pstar %u error %u tiles %u out %u\n", world, rounds, scalars[0], scalars[1], scalars[2], scalars[3]);
        }
    }
    std::printf("DONE\n");
    return 0;
}
#endif
