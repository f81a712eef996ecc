// The body of the single-CTA radix sort (vren_b200/csrc/small_sort_body.cuh) executed on the host by cta_emulator.hpp and
// compared with std::stable_sort by key (the reference test's check, vren_test radix_sort.cpp:88, pairs extension).
//   small_sort_emulation <with_values 0|1> <n> [<n> ...]      every n with five key patterns
#include "cta_emulator.hpp"

#include "../../vren_b200/csrc/small_sort_body.cuh"

#include <algorithm>
#include <cstring>
#include <random>

using namespace vrenb200;

template <bool HV>
static bool run_case(uint32_t n, int pattern, uint32_t seed)
{
    std::mt19937 rng(seed);
    std::vector<uint32_t> keys(n), vals(n);
    for (uint32_t i = 0; i < n; i++)
    {
        const uint32_t r = rng();
        switch (pattern)
        {
        case 0: keys[i] = r; break;                                // uniform
        case 1: keys[i] = n - 1 - i; break;                        // reversed iota (TEST(radix_sort, main))
        case 2: keys[i] = (r % 7u) * 0x01010101u; break;           // few values: lane collisions, stability visible in the values
        case 3: keys[i] = 0xFFFFFFFFu; break;                      // equal to the padding key
        default: keys[i] = (r & 1u) ? 0xFFFFFFFFu : (r & 0xFF00FF00u); break;
        }
        vals[i] = i;
    }
    std::vector<uint32_t> order(n);
    for (uint32_t i = 0; i < n; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
    std::vector<uint32_t> want_k(n), want_v(n);
    for (uint32_t i = 0; i < n; i++) { want_k[i] = keys[order[i]]; want_v[i] = vals[order[i]]; }

    std::vector<uint64_t> smem((sizeof(small_sort_smem<HV>) + 7) / 8);      // 8-byte aligned
    // guard words around the arrays the kernel writes in place: it must not touch anything beyond n
    std::vector<uint32_t> gk(n + 2, 0xDEADBEEFu), gv(n + 2, 0xDEADBEEFu);
    std::memcpy(gk.data() + 1, keys.data(), n * 4);
    std::memcpy(gv.data() + 1, vals.data(), n * 4);
    unsigned char* raw = reinterpret_cast<unsigned char*>(smem.data());
    uint32_t* pk = gk.data() + 1;
    uint32_t* pv = HV ? gv.data() + 1 : nullptr;
    if (!cta_emu::run_cta(kSmallSortThreads, [=]() { single_cta_sort_body<HV>(raw, pk, pv, n, 0, 4); })) std::exit(3);
    bool ok = gk[0] == 0xDEADBEEFu && gk[n + 1] == 0xDEADBEEFu && gv[0] == 0xDEADBEEFu && gv[n + 1] == 0xDEADBEEFu;
    ok = ok && std::memcmp(pk, want_k.data(), n * 4) == 0;
    if (HV) ok = ok && std::memcmp(gv.data() + 1, want_v.data(), n * 4) == 0;
    if (!ok) std::printf("MISMATCH with_values=%d n=%u pattern=%d\n", (int) HV, n, pattern);
    return ok;
}

int main(int argc, char** argv)
{
    if (argc < 3) { std::fprintf(stderr, "usage: %s <with_values> <n> [<n> ...]\n", argv[0]); return 2; }
    const bool hv = std::atoi(argv[1]) != 0;
    bool all = true;
    int cases = 0;
    for (int a = 2; a < argc; a++)
    {
        const uint32_t n = (uint32_t) std::strtoul(argv[a], nullptr, 10);
        if (n == 0 || n > kSmallSortMax) { std::fprintf(stderr, "n out of range: %u\n", n); return 2; }
        for (int pattern = 0; pattern < 5; pattern++, cases++)
            all = (hv ? run_case<true>(n, pattern, 1000u + n + pattern) : run_case<false>(n, pattern, 1000u + n + pattern)) && all;
    }
    std::printf("%s %d cases\n", all ? "ALL PASS" : "FAILED", cases);
    return all ? 0 : 1;
}
