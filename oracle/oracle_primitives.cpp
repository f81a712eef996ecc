// TEST INFRASTRUCTURE — CPU restatement ("oracle") of vren's parallel primitives.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
// library; the product path (vren_b200/) never does.  Every function cites the reference file:line it
// restates (paths relative to the reference checkout).  The reference has no CPU implementation of these
// passes: where its unit tests carry a CPU check (std::sort, std::exclusive_scan, run_cpu_reduce, linear AABB
// scan) the oracle restates that check too, and tests/test_oracle.py verifies the shader restatement against it.
//
// Parity pinning: exact KATs exist only for calc_bvh_* (vren_test/vren_test/primitives/build_bvh.cpp:225-253);
// they are asserted in tests/test_oracle.py.  oracle/_ref (built by oracle/ref_extract.py from the reference
// sources where they lie) supplies the reference's own run_cpu_reduce, base.hpp helpers and calc_bvh_* for
// cross-checks.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <thread>
#include <vector>

namespace {

// base/base.hpp:37-47
uint32_t next_pow2(uint32_t v)
{
    v--;
    v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16;
    v++;
    return v;
}

struct vec4 { float x, y, z, w; };

// VRen.cmake:70-75 — VREN_OPERATION(b, a) with GLSL min/max semantics (min(x,y) = y<x ? y : x)
enum { OP_ADD = 0, OP_MIN = 1, OP_MAX = 2 };
enum { DT_U32 = 0, DT_VEC4 = 1, DT_F32 = 2 };

template <typename T> T op_apply(int op, T b, T a);
template <> uint32_t op_apply(int op, uint32_t b, uint32_t a)
{
    return op == OP_ADD ? b + a : op == OP_MIN ? (a < b ? a : b) : (b < a ? a : b);
}
template <> float op_apply(int op, float b, float a)
{
    return op == OP_ADD ? b + a : op == OP_MIN ? (a < b ? a : b) : (b < a ? a : b);
}
template <> vec4 op_apply(int op, vec4 b, vec4 a)
{
    return vec4{op_apply(op, b.x, a.x), op_apply(op, b.y, a.y), op_apply(op, b.z, a.z), op_apply(op, b.w, a.w)};
}
template <typename T> T op_identity(int op);
template <> uint32_t op_identity(int op) { return op == OP_MIN ? ~0u : 0u; }
template <> float op_identity(int op) { return op == OP_ADD ? 0.0f : op == OP_MIN ? 1e35f : -1e35f; }
template <> vec4 op_identity(int op) { float i = op_identity<float>(op); return vec4{i, i, i, i}; }

// reduce.comp:50-87 driven by reduce.cpp:80-113: levels in groups of 10, every touched slot written back.
// Flattened: tmp = in padded with identity; for l, for j: b = 2^l-1 + j*2^(l+1) + 2^l; tmp[b] = op(tmp[b], tmp[a]).
template <typename T>
void reduce_tree(int op, const T* in, uint32_t n, T* out, uint32_t blocks)
{
    const uint32_t P = next_pow2(n);
    std::vector<T> tmp(P);
    for (uint32_t y = 0; y < blocks; y++)
    {
        const T* row_in = in + (size_t) y * n;   // reduce.comp:52
        T* row_out = out + (size_t) y * P;       // reduce.comp:53
        for (uint32_t i = 0; i < P; i++) tmp[i] = i < n ? row_in[i] : op_identity<T>(op); // reduce.comp:61
        for (uint32_t s = 1; s < P; s <<= 1)
            for (uint32_t b = 2 * s - 1; b < P; b += 2 * s)
                tmp[b] = op_apply<T>(op, tmp[b], tmp[b - s]); // reduce.comp:74-77
        std::memcpy(row_out, tmp.data(), (size_t) P * sizeof(T)); // reduce.comp:83-86
    }
}

} // namespace

extern "C" {

// ---- a10: base/base.hpp:32-79 restated WITH the reference's double-precision log/pow formulation -----------
int oracle_is_power_of_2(uint32_t v) { return v > 0 && (v & (v - 1)) == 0; }
uint32_t oracle_round_to_next_power_of_2(uint32_t v) { return next_pow2(v); }
uint64_t oracle_round_to_next_multiple_of(uint64_t v, uint64_t m) { uint64_t r = v % m; return r == 0 ? v : v + m - r; }
uint32_t oracle_divide_and_ceil(uint32_t v, uint32_t d) { return (uint32_t) std::ceil(double(v) / double(d)); } // base.hpp:76-79
int oracle_is_power_of(uint32_t n, uint32_t base)
{
    double e = std::log((double) n) / std::log((double) base); // base.hpp:63-67
    return (e - std::floor(e)) < 2.220446049250313e-16;
}
uint32_t oracle_round_to_next_power_of(uint32_t n, uint32_t base)
{
    double e = std::log((double) n) / std::log((double) base); // base.hpp:70-74
    return (uint32_t) std::pow((double) base, std::ceil(e));
}

// ---- a5 sizing: build_bvh.cpp:99-136 ------------------------------------------------------------------------
uint32_t oracle_calc_bvh_padded_leaf_count(uint32_t leaf_count)
{
    return leaf_count <= 1 ? 32u : oracle_round_to_next_power_of(leaf_count, 32u);
}
uint32_t oracle_calc_bvh_buffer_length(uint32_t leaf_count)
{
    uint32_t padded = oracle_calc_bvh_padded_leaf_count(leaf_count);
    uint64_t length = 0;
    while (padded != 0) { length += padded; padded >>= 5; }
    return (uint32_t) length;
}
uint64_t oracle_calc_bvh_buffer_size(uint32_t leaf_count) { return (uint64_t) oracle_calc_bvh_buffer_length(leaf_count) * 32u; }
uint32_t oracle_calc_bvh_root_index(uint32_t leaf_count) { return oracle_calc_bvh_buffer_length(leaf_count) - 1; }
uint32_t oracle_calc_bvh_level_count(uint32_t leaf_count)
{
    uint32_t padded = oracle_calc_bvh_padded_leaf_count(leaf_count);
    return (uint32_t) (std::log2((double) padded) / 5); // glm::log2(padded) / 5
}

// ---- a1: reduce ----------------------------------------------------------------------------------------------------
// out must hold blocks * next_pow2(n) elements
void oracle_reduce(int dtype, int op, const void* in, uint32_t n, void* out, uint32_t blocks)
{
    if (dtype == DT_U32) reduce_tree<uint32_t>(op, (const uint32_t*) in, n, (uint32_t*) out, blocks);
    else if (dtype == DT_F32) reduce_tree<float>(op, (const float*) in, n, (float*) out, blocks);
    else reduce_tree<vec4>(op, (const vec4*) in, n, (vec4*) out, blocks);
}

// the reference TEST's CPU check, restated: vren_test/vren_test/primitives/reduce.cpp:72-98 (operation(a, b),
// glm::min/glm::max argument order) over an already padded pow2 buffer, in place
void oracle_test_cpu_reduce_u32(int op, uint32_t* data, uint32_t length)
{
    for (uint32_t i = 0; (1u << i) < length; i++)
        for (uint32_t j = 0; j < (length >> (i + 1)); j++)
        {
            uint32_t a = (1u << i) - 1 + (j << (i + 1));
            uint32_t b = a + (1u << i);
            uint32_t x = data[a], y = data[b];
            data[b] = op == OP_ADD ? x + y : op == OP_MIN ? (y < x ? y : x) : (x < y ? y : x);
        }
}
void oracle_test_cpu_reduce_f32(int op, float* data, uint32_t length, uint32_t comps)
{
    for (uint32_t i = 0; (1u << i) < length; i++)
        for (uint32_t j = 0; j < (length >> (i + 1)); j++)
        {
            uint32_t a = (1u << i) - 1 + (j << (i + 1));
            uint32_t b = a + (1u << i);
            for (uint32_t c = 0; c < comps; c++)
            {
                float x = data[a * comps + c], y = data[b * comps + c];
                data[b * comps + c] = op == OP_ADD ? x + y : op == OP_MIN ? (y < x ? y : x) : (x < y ? y : x);
            }
        }
}

// ---- a2: scan ------------------------------------------------------------------------------------------------------
// std::exclusive_scan — the reference test's check (vren_test/.../blelloch_scan.cpp:135)
void oracle_exclusive_scan_u32(const uint32_t* in, uint32_t* out, uint32_t n)
{
    uint32_t acc = 0;
    for (uint32_t i = 0; i < n; i++) { uint32_t v = in[i]; out[i] = acc; acc += v; }
}

// blelloch_scan::downsweep — blelloch_scan.cpp:57-139 + blelloch_scan_downsweep.comp:34-125, in place.
// Global levels log2(n)-1 .. 10 (clear_last only on the first), then the 1024-wide workgroup pass whose tile is
// zero-filled above block_length - clear_last.
void oracle_downsweep_u32(uint32_t* buf, uint32_t n, uint32_t blocks, int clear_last)
{
    int log2n = 0;
    while ((1u << log2n) < n) log2n++;
    for (uint32_t y = 0; y < blocks; y++)
    {
        uint32_t* d = buf + (size_t) y * n;
        int cl = clear_last;
        for (int level = log2n - 1; level > 9; level--) // blelloch_scan.cpp:90
        {
            const uint32_t stride = 1u << level, offset = stride - 1;
            for (uint32_t g = 0;; g++)
            {
                const uint64_t a = offset + (uint64_t) g * stride * 2, b = a + stride;
                if (b >= n) break;
                const uint32_t bv = (cl && b == n - 1) ? 0u : d[b]; // comp:48
                d[b] = d[a] + bv;
                d[a] = bv;
            }
            cl = 0; // blelloch_scan.cpp:118-121
        }
        // workgroup pass, one 1024-wide tile per workgroup (comp:59-125)
        for (uint32_t base = 0; base < n; base += 1024)
        {
            uint32_t s[1024];
            for (uint32_t t = 0; t < 1024; t++)
            {
                const uint64_t idx = (uint64_t) base + t;
                s[t] = idx < (uint64_t) n - (cl ? 1 : 0) ? d[idx] : 0u; // comp:68-75
            }
            for (int level = 9; level >= 0; level--)
            {
                const uint32_t m = (1u << (level + 1)) - 1;
                for (uint32_t t = 0; t < 1024; t++)
                    if ((t & m) == m)
                    {
                        const uint32_t a = t - (1u << level);
                        const uint32_t tmp = s[t];
                        s[t] = s[a] + tmp;
                        s[a] = tmp;
                    }
            }
            for (uint32_t t = 0; t < 1024 && (uint64_t) base + t < n; t++) d[base + t] = s[t];
        }
    }
}

// blelloch_scan::operator() — blelloch_scan.cpp:141-166: reduce<uint,add> in place with blocks_num hard-coded
// to 1 (:151), then downsweep(clear_last=true) over all `blocks` rows. n must be a power of two (:67).
void oracle_blelloch_scan_u32(uint32_t* buf, uint32_t n, uint32_t blocks)
{
    std::vector<uint32_t> tree(n);
    reduce_tree<uint32_t>(OP_ADD, buf, n, tree.data(), 1);
    std::memcpy(buf, tree.data(), (size_t) n * 4);
    oracle_downsweep_u32(buf, n, blocks, 1);
}

// ---- a3: radix sort -----------------------------------------------------------------------------------------------
// the reference test's check: std::sort (vren_test/.../radix_sort.cpp:88)
void oracle_sort_keys(uint32_t* keys, uint32_t n) { std::sort(keys, keys + n); }

// literal restatement of the reference algorithm (radix_sort.cpp:171-337; radix_sort_local_count.comp:49-70,
// radix_sort_global_offset.comp:38-56, radix_sort_reorder.comp:55-113): 8 stable passes of 4 bits, per-workgroup (1024 keys) digit
// counts stored digit-major, exclusive scan across workgroups, 16-wide global offsets, stable scatter.
void oracle_radix_sort_lsd4(uint32_t* keys, uint32_t n)
{
    std::vector<uint32_t> tmp(n);
    uint32_t* src = keys;
    uint32_t* dst = tmp.data();
    const uint32_t wgs = (n + 1023) / 1024;
    std::vector<uint32_t> local(16 * (size_t) wgs);
    for (int pass = 0; pass < 8; pass++)
    {
        const int shift = pass * 4;
        std::fill(local.begin(), local.end(), 0u);
        for (uint32_t i = 0; i < n; i++) local[(size_t) ((src[i] >> shift) & 15) * wgs + i / 1024]++;
        uint32_t global_offset[16], total[16];
        for (int d = 0; d < 16; d++)
        {
            uint32_t acc = 0;
            for (uint32_t w = 0; w < wgs; w++) { uint32_t c = local[(size_t) d * wgs + w]; local[(size_t) d * wgs + w] = acc; acc += c; }
            total[d] = acc;
        }
        uint32_t acc = 0;
        for (int d = 0; d < 16; d++) { global_offset[d] = acc; acc += total[d]; }
        std::vector<uint32_t> seen(16 * (size_t) wgs, 0u);
        for (uint32_t i = 0; i < n; i++)
        {
            const uint32_t d = (src[i] >> shift) & 15, w = i / 1024;
            dst[global_offset[d] + local[(size_t) d * wgs + w] + seen[(size_t) d * wgs + w]++] = src[i];
        }
        std::swap(src, dst);
    }
    // 8 passes: result is back in `keys` (radix_sort.cpp:173-174)
}

// KV extension: stable by key
struct kv_t { uint32_t k, v; };
void oracle_sort_pairs(uint32_t* keys, uint32_t* vals, uint32_t n)
{
    std::vector<kv_t> p(n);
    for (uint32_t i = 0; i < n; i++) p[i] = kv_t{keys[i], vals[i]};
    std::stable_sort(p.begin(), p.end(), [](const kv_t& a, const kv_t& b) { return a.k < b.k; });
    for (uint32_t i = 0; i < n; i++) { keys[i] = p[i].k; vals[i] = p[i].v; }
}

// timed baseline helpers (bench.py): interleaved pairs sorted in place, single- and multi-threaded
void oracle_sort_pairs_interleaved(uint64_t* pairs_kv_lo_key, uint32_t n)
{
    // element = key | value<<32 ; compare on the low 32 bits only -> stable_sort keeps ties in input order
    std::stable_sort(pairs_kv_lo_key, pairs_kv_lo_key + n,
                     [](uint64_t a, uint64_t b) { return (uint32_t) a < (uint32_t) b; });
}
void oracle_sort_pairs_interleaved_mt(uint64_t* pairs, uint32_t n, uint32_t threads)
{
    if (threads <= 1 || n < 1u << 16) { oracle_sort_pairs_interleaved(pairs, n); return; }
    auto cmp = [](uint64_t a, uint64_t b) { return (uint32_t) a < (uint32_t) b; };
    std::vector<size_t> cut(threads + 1);
    for (uint32_t t = 0; t <= threads; t++) cut[t] = (size_t) n * t / threads;
    std::vector<std::thread> pool;
    for (uint32_t t = 0; t < threads; t++)
        pool.emplace_back([&, t] { std::stable_sort(pairs + cut[t], pairs + cut[t + 1], cmp); });
    for (auto& th : pool) th.join();
    // pairwise parallel merges (stable: left run first)
    for (uint32_t width = 1; width < threads; width *= 2)
    {
        pool.clear();
        for (uint32_t t = 0; t + width < threads; t += 2 * width)
        {
            const size_t lo = cut[t], mid = cut[t + width], hi = cut[std::min(t + 2 * width, threads)];
            pool.emplace_back([=] { std::inplace_merge(pairs + lo, pairs + mid, pairs + hi, cmp); });
        }
        for (auto& th : pool) th.join();
    }
}
void oracle_sort_keys_mt(uint32_t* keys, uint32_t n, uint32_t threads)
{
    if (threads <= 1 || n < 1u << 16) { std::sort(keys, keys + n); return; }
    std::vector<size_t> cut(threads + 1);
    for (uint32_t t = 0; t <= threads; t++) cut[t] = (size_t) n * t / threads;
    std::vector<std::thread> pool;
    for (uint32_t t = 0; t < threads; t++) pool.emplace_back([&, t] { std::sort(keys + cut[t], keys + cut[t + 1]); });
    for (auto& th : pool) th.join();
    for (uint32_t width = 1; width < threads; width *= 2)
    {
        pool.clear();
        for (uint32_t t = 0; t + width < threads; t += 2 * width)
        {
            const size_t lo = cut[t], mid = cut[t + width], hi = cut[std::min(t + 2 * width, threads)];
            pool.emplace_back([=] { std::inplace_merge(keys + lo, keys + mid, keys + hi); });
        }
        for (auto& th : pool) th.join();
    }
}

// ---- a4: bucket sort ----------------------------------------------------------------------------------------------
// bucket_sort_count.comp:27-34, blelloch scan of the 65536 counters (bucket_sort.cpp:133-141),
// bucket_sort_write.comp:27-35 with the canonical tie-break "input order" (one legal outcome of the atomics).
// counters end as bucket END offsets because the write pass increments the scanned offsets.
void oracle_bucket_sort(const uint32_t* in_pairs, uint32_t n, uint32_t* out_pairs, uint32_t* counters)
{
    std::vector<uint32_t> cnt(65536, 0u);
    for (uint32_t i = 0; i < n; i++) cnt[in_pairs[2 * (size_t) i] & 0xFFFFu]++;
    uint32_t acc = 0;
    for (uint32_t k = 0; k < 65536; k++) { uint32_t c = cnt[k]; cnt[k] = acc; acc += c; }
    for (uint32_t i = 0; i < n; i++)
    {
        const uint32_t x = in_pairs[2 * (size_t) i], y = in_pairs[2 * (size_t) i + 1];
        const uint32_t o = cnt[x & 0xFFFFu]++;
        out_pairs[2 * (size_t) o] = x;
        out_pairs[2 * (size_t) o + 1] = y;
    }
    std::memcpy(counters, cnt.data(), 65536 * 4);
}

// ---- a5: BVH build --------------------------------------------------------------------------------------------------
struct bvh_node { float mn[3]; uint32_t next; float mx[3]; uint32_t pad; };
static const uint32_t LEAF = 0xFFFFFFFFu, INVALID = 0xFFFFFFFEu;

static uint32_t ford(float f)
{
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}

// build_bvh.cpp:69-94 + build_bvh.comp:32-55. Canonical choices where the reference is undefined
// (SURVEY 8c-iii): an all-invalid parent gets min=+1e35, max=-1e35 (never overlaps), _pad = 0.
void oracle_build_bvh(bvh_node* nodes, uint32_t padded_leaf_count)
{
    uint32_t src = 0, count = padded_leaf_count;
    while (count > 1)
    {
        const uint32_t dst = src + count;
        for (uint32_t j = 0; j < count / 32; j++)
        {
            bvh_node p;
            for (int c = 0; c < 3; c++) { p.mn[c] = 1e35f; p.mx[c] = -1e35f; }
            p.pad = 0;
            bool any = false;
            for (uint32_t t = 0; t < 32; t++)
            {
                const bvh_node& ch = nodes[src + 32 * j + t];
                if (ch.next == INVALID) continue;
                any = true;
                for (int c = 0; c < 3; c++)
                {
                    // total order with -0 < +0 so that the result does not depend on the visiting order
                    p.mn[c] = ford(ch.mn[c]) < ford(p.mn[c]) ? ch.mn[c] : p.mn[c];
                    p.mx[c] = ford(p.mx[c]) < ford(ch.mx[c]) ? ch.mx[c] : p.mx[c];
                }
            }
            p.next = any ? src + 32 * j : INVALID;
            nodes[dst + j] = p;
        }
        src = dst;
        count >>= 5;
    }
}

// the reference test's property check (vren_test/.../build_bvh.cpp:25-78): recursive traversal vs linear scan
static bool point_in(const bvh_node& n, const float* p)
{
    return p[0] >= n.mn[0] && p[1] >= n.mn[1] && p[2] >= n.mn[2] && p[0] <= n.mx[0] && p[1] <= n.mx[1] && p[2] <= n.mx[2];
}
static void traverse_r(const bvh_node* bvh, uint32_t offset, const float* p, std::vector<uint32_t>& out)
{
    for (uint32_t i = 0; i < 32; i++)
    {
        const bvh_node& n = bvh[offset + i];
        if (n.next == INVALID) continue;
        if (point_in(n, p))
        {
            if (n.next == LEAF) out.push_back(offset + i);
            else traverse_r(bvh, n.next, p, out);
        }
    }
}
// returns the hit count; writes up to max_hits sorted leaf indices
uint32_t oracle_bvh_traverse_point(const bvh_node* bvh, uint32_t root, const float* p, uint32_t* hits, uint32_t max_hits)
{
    std::vector<uint32_t> out;
    const bvh_node& r = bvh[root];
    if (r.next != INVALID && point_in(r, p)) traverse_r(bvh, r.next, p, out);
    std::sort(out.begin(), out.end());
    for (uint32_t i = 0; i < out.size() && i < max_hits; i++) hits[i] = out[i];
    return (uint32_t) out.size();
}
uint32_t oracle_bvh_linear_point(const bvh_node* leaves, uint32_t leaf_count, const float* p, uint32_t* hits, uint32_t max_hits)
{
    uint32_t c = 0;
    for (uint32_t i = 0; i < leaf_count; i++)
        if (leaves[i].next != INVALID && point_in(leaves[i], p))
        {
            if (c < max_hits) hits[c] = i;
            c++;
        }
    return c;
}

} // extern "C"
