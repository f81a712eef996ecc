// The chained single-pass scan (vren_b200/csrc/scan.cu::exclusive_scan_u32_kernel) executed on the host: the test extracts the
// kernel's body and the definitions it needs from scan.cu (the text between the [[chained-scan-*]] markers) into two .inc
// files, this file wraps them and cta_emulator.hpp runs them, one CTA after the other, in an ORDER OF THE TEST'S CHOOSING.
// With ticket tile ids (the safe mode) any order must give the exclusive scan: a CTA's tile id is its start order, so every
// tile it looks back at is finished.  With block-index tile ids only the in-order run is legal (that is the assumption the safe
// mode removes); it is run too, as a check of this harness against the kernel the GPU suite has verified.
//   scan_emulation <ticket 0|1> <order: forward|reverse|shuffle> <misalign elements> <n> [<n> ...]
#include <cuda_runtime.h>      // uint4 and the host types of common.cuh; nothing of the CUDA runtime is called

#include "cta_emulator.hpp"

#include "../../vren_b200/csrc/common.cuh"

#include <algorithm>
#include <cstring>
#include <random>

namespace vrenb200 {
namespace {

// host stand-ins of the device-only helpers of common.cuh the body uses
inline uint64_t ld_relaxed_u64(const uint64_t* p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
inline void st_relaxed_u64(uint64_t* p, uint64_t v) { __atomic_store_n(p, v, __ATOMIC_RELAXED); }
inline uint4 ldg_stream_u4(const uint4* p) { return *p; }

#include "scan_chained_defs.inc"

template <int THREADS, bool TICKET>
void chained_scan_body(const uint32_t* in, uint32_t* out, uint32_t n, scan_state* state, uint32_t base)
{
#include "scan_chained_body.inc"
}

} // namespace
} // namespace vrenb200

using namespace vrenb200;

template <bool TICKET>
static bool run_case(uint32_t n, const char* order_name, uint32_t misalign)
{
    constexpr int THREADS = 256;
    constexpr uint32_t tile = THREADS * kScanVecs * 4;
    const uint32_t tiles = (n + tile - 1) / tile;
    std::mt19937 rng(n * 7u + misalign);
    std::vector<uint32_t> in_store(n + 8), out_store(n + 8, 0xDEADBEEFu);
    uint32_t* in = in_store.data() + misalign;        // misalign != 0: the guarded scalar path instead of the 128-bit one
    uint32_t* out = out_store.data() + misalign;
    for (uint32_t i = 0; i < n; i++) in[i] = rng();
    const uint32_t base = 0x9E3779B9u;
    std::vector<uint32_t> want(n);
    uint32_t acc = base;
    for (uint32_t i = 0; i < n; i++) { want[i] = acc; acc += in[i]; }

    std::vector<uint64_t> state_store((offsetof(scan_state, status) + (size_t) tiles * 8 + 7) / 8 + 1, 0);   // zeroed, as the host's memset leaves it
    scan_state* state = reinterpret_cast<scan_state*>(state_store.data());
    std::vector<unsigned> order(tiles);
    for (uint32_t t = 0; t < tiles; t++) order[t] = t;
    if (!std::strcmp(order_name, "reverse")) std::reverse(order.begin(), order.end());
    if (!std::strcmp(order_name, "shuffle")) std::shuffle(order.begin(), order.end(), rng);
    for (unsigned block : order)
        if (!cta_emu::run_cta(THREADS, [=]() { chained_scan_body<THREADS, TICKET>(in, out, n, state, base); }, block)) std::exit(3);
    bool ok = std::memcmp(out, want.data(), (size_t) n * 4) == 0;
    for (uint32_t i = 0; i < misalign; i++) ok = ok && out_store[i] == 0xDEADBEEFu;
    for (uint32_t i = misalign + n; i < n + 8; i++) ok = ok && out_store[i] == 0xDEADBEEFu;
    if (TICKET) ok = ok && state->ticket == tiles;
    if (!ok) std::printf("MISMATCH ticket=%d order=%s misalign=%u n=%u\n", (int) TICKET, order_name, misalign, n);
    return ok;
}

int main(int argc, char** argv)
{
    if (argc < 5) { std::fprintf(stderr, "usage: %s <ticket> <forward|reverse|shuffle> <misalign> <n> [<n> ...]\n", argv[0]); return 2; }
    const bool ticket = std::atoi(argv[1]) != 0;
    const uint32_t misalign = (uint32_t) std::atoi(argv[3]);
    if (!ticket && std::strcmp(argv[2], "forward")) { std::fprintf(stderr, "block-index tile ids are only defined for in-order execution\n"); return 2; }
    bool all = true;
    int cases = 0;
    for (int a = 4; a < argc; a++, cases++)
    {
        const uint32_t n = (uint32_t) std::strtoul(argv[a], nullptr, 10);
        all = (ticket ? run_case<true>(n, argv[2], misalign) : run_case<false>(n, argv[2], misalign)) && all;
    }
    std::printf("%s %d cases\n", all ? "ALL PASS" : "FAILED", cases);
    return all ? 0 : 1;
}
