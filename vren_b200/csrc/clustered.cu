// a7 — clustered_shading::find_unique_cluster_list (reference: vren/vren/pipeline/clustered_shading.cpp:363-448,
//      shaders/clustered_shading/find_unique_clusters.comp:46-121, clustered_shading.glsl:7-43)
// a8 — clustered_shading::assign_lights (reference: clustered_shading.cpp:473-690, assign_lights.comp:77-241,
//      clustered_shading.glsl:58-110)
//
// a7, reference: one 32x32 workgroup per tile, 1024-wide bitonic sort (55 barrier stages) + workgroup scan + one
// global atomicAdd per tile -> list order depends on atomic arrival order.
// a7, here: one CTA per tile; the 32x32 depth (and RGBA16F normal) tile is staged with a 2-D TMA tensor copy
// (cp.async.bulk.tensor.2d + mbarrier).  All keys of a tile share (i, j), so the 16 remaining key bits index a
// 65536-bit shared-memory bitmap: set bit = key present, rank = popcount prefix -> sorted unique list without a sort.
// Tile bases come from a decoupled look-back over tiles in tile-major order, which makes the list order
// deterministic (the canonical order of the oracle).  One launch, 4 B/px read (+8 B/px normals) + 4 B/px written.
//
// a8, reference: count traversal -> copy -> blelloch_scan(2^17) -> write traversal (the BVH is walked twice).
// a8, here: ONE traversal per cluster.  A warp takes a cluster from a ticket, walks the 32-ary light BVH with a
// per-level overlap bitmask (lane t tests child t, lowest set bit first; leaf level = sphere vs box on a compact
// 16-byte leaf sphere), remembers the (leaf group, hit mask) pairs and writes the list - in the order of
// assign_lights.comp:229-232, descending lane inside a leaf group - into a bump-allocated arena; the single-pass
// scan of the counts gives the reference's offsets, and a copy kernel moves every list to its final place.
//
// fp32 contract: every op is an explicitly rounded _rn intrinsic in GLSL source order; tan/pow/log are evaluated
// on the host once per camera (tables below) with the formulas stated in oracle/oracle_clustered.cpp.
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <deque>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace vrenb200 {

namespace {

constexpr uint32_t kInvalid = VRENB200_BVH_INVALID_NODE;
constexpr int kMaxBvhLevels = 6;

// ---- host-side camera constants (camera.cpp:40-50) -----------------------------------------------------------
struct proj_consts
{
    float i00, i11, iB, nAB;   // closed-form inverse(projection) entries
    float tan_half;
};

proj_consts make_proj(const vrenb200_camera& c)
{
    proj_consts p;
    p.tan_half = tanf(c.fov_y / 2.0f);
    const float m00 = 1.0f / (p.tan_half * c.aspect_ratio);
    const float m11 = 1.0f / p.tan_half;
    const float m22 = c.far_plane / (c.far_plane - c.near_plane);
    const float m32 = -(c.far_plane * c.near_plane) / (c.far_plane - c.near_plane);
    p.i00 = 1.0f / m00;
    p.i11 = 1.0f / m11;
    p.iB = 1.0f / m32;
    p.nAB = (-m22) / m32;
    return p;
}

// slice index of a view-space depth: uint(floor(log(z / near) / log(a))) with every operation in fp32, as the shader types
// it (find_unique_clusters.comp:65); log = the host's logf.  Monotone in z for every camera tried (all floats of
// [near/2, 2 far] checked for 3..135 tile rows), which is what the threshold table below relies on.
uint32_t slice_of(float z, float near_plane, float a)
{
    const float k = std::floor(logf(z / near_plane) / logf(a));
    if (!(k >= 0.0f)) return 0u;
    if (k > 4294967040.0f) return 0xFFFFFFFFu;
    return (uint32_t) k;
}

// thresholds[k] = smallest positive float z whose slice is >= k (k >= 1); the device finds the slice by comparing
// against this table, so it agrees with the host formula above for every float z by construction.
constexpr uint32_t kMaxSlices = 16384;

struct slice_table
{
    float near_plane = 0, a = 0;
    std::vector<float> thresholds; // [0] = 0 (unused), [1..count)
};

// Device copies of the small per-camera tables (slice thresholds, near_k), keyed by (device, kind, two floats): the
// tables only change with the camera, so a frame re-uses them instead of paying a host->device copy (and, for near_k,
// 1024 powf calls) per call.  A new camera is filled in with a synchronous copy (the one host synchronisation of the
// clustered path, once per camera and device); entries live as long as the library.
struct device_table
{
    int device, kind;
    float p0, p1;
    float* data;
    uint32_t count;
};

std::mutex g_table_mutex;
constexpr size_t kMaxDeviceTables = 512;
std::deque<device_table> g_tables;    // deque: growing it never moves the entries other threads hold pointers to

const device_table* find_device_table(int kind, float p0, float p1)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    for (const device_table& t : g_tables)
        if (t.device == dev && t.kind == kind && t.p0 == p0 && t.p1 == p1) return &t;
    return nullptr;
}

const device_table* add_device_table(int kind, float p0, float p1, const float* host, uint32_t count)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    // entries are never freed while the library is loaded: their pointers live in CUDA graphs recorded by callers and in
    // kernels already enqueued.  A table is <= 64 KB; 512 distinct (device, camera) pairs is the (generous) limit.
    if (g_tables.size() >= kMaxDeviceTables) return nullptr;
    device_table t{dev, kind, p0, p1, nullptr, count};
    if (cudaMalloc(&t.data, (size_t) count * sizeof(float)) != cudaSuccess) return nullptr;
    if (cudaMemcpy(t.data, host, (size_t) count * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess)
    {
        cudaFree(t.data);
        return nullptr;
    }
    g_tables.push_back(t);
    return &g_tables.back();
}

const slice_table& get_slice_table(float near_plane, float a)
{
    static thread_local slice_table tab;
    if (tab.near_plane == near_plane && tab.a == a && !tab.thresholds.empty()) return tab;
    tab.near_plane = near_plane;
    tab.a = a;
    tab.thresholds.clear();
    tab.thresholds.push_back(0.0f);
    // zfin = the largest z whose slice is a number: beyond it z / near overflows to +inf in fp32 and uint(floor(inf)) is taken
    // as 0xFFFFFFFF (key bits 0x3FF).  Bisection over the float bit patterns (slice_of is monotone).
    uint32_t lo = 0x00800000u, hi = 0x7F800000u;   // slice_of(bits lo) finite, slice_of(+inf) = 0xFFFFFFFF
    auto as_float = [](uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; };
    while (hi - lo > 1)
    {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (slice_of(as_float(mid), near_plane, a) == 0xFFFFFFFFu) hi = mid; else lo = mid;
    }
    const float zfin = as_float(lo);
    const uint32_t kfin = slice_of(zfin, near_plane, a);
    const uint32_t kmax = std::min<uint32_t>(kfin, kMaxSlices - 2);
    for (uint32_t k = 1; k <= kmax; k++)
    {
        float c = std::min((float) ((double) near_plane * std::pow((double) a, (double) k)), zfin);
        while (slice_of(c, near_plane, a) < k) c = std::nextafterf(c, INFINITY);
        while (slice_of(std::nextafterf(c, 0.0f), near_plane, a) >= k) c = std::nextafterf(c, 0.0f);
        tab.thresholds.push_back(c);
    }
    // the overflow region maps to the next index whose low 10 bits are all ones (what 0xFFFFFFFF & 0x3FF leaves in the key)
    if (kfin == kmax && kmax + 1024 < kMaxSlices)
    {
        const float over = as_float(hi);
        do tab.thresholds.push_back(over); while (((tab.thresholds.size() - 1) & 0x3FFu) != 0x3FFu);
    }
    return tab;
}

// ---- mbarrier / TMA wrappers ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
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
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* map, int x, int y, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst_smem)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}

// ---- a7 kernel ----------------------------------------------------------------------------------------------------
struct cluster_key_params
{
    uint32_t width, height, tiles_x, tiles_y;
    float iB, nAB;              // w' = d*iB + nAB ; z = 1/w'
    float inv_near, inv_log2a;  // first guess of the slice
    uint32_t table_len;
    uint32_t max_keys;
    int use_tma, has_normals;
};

// clustered_shading.glsl:7-43
__device__ __forceinline__ uint32_t discretize_normal(float nx, float ny, float nz)
{
    if (nx == 0.0f && ny == 0.0f && nz == 0.0f) return 0xFFFFFFFFu;
    const float n[3] = {nx, ny, nz};
    float min_t = 1e35f;
    uint32_t axis = 0, face_idx = 0;
#pragma unroll
    for (uint32_t i = 0; i < 3; i++)
    {
        const float sg = n[i] > 0.0f ? 1.0f : (n[i] < 0.0f ? -1.0f : 0.0f);
        const float t = __fdiv_rn(sg, n[i]);
        if (t < min_t)
        {
            min_t = t;
            axis = i;
            face_idx = (n[i] > 0.0f ? 1u : 0u) * 3u + i;
        }
    }
    const float p0 = __fmul_rn(n[0], min_t), p1 = __fmul_rn(n[1], min_t), p2 = __fmul_rn(n[2], min_t);
    const float uvx = axis == 0 ? p1 : (axis == 1 ? p2 : p0);
    const float uvy = axis == 0 ? p2 : (axis == 1 ? p0 : p1);
    const float fx = floorf(__fmul_rn(__fdiv_rn(__fadd_rn(uvx, 1.0f), 2.0f), 3.0f));
    const float fy = floorf(__fmul_rn(__fdiv_rn(__fadd_rn(uvy, 1.0f), 2.0f), 3.0f));
    const uint32_t dx = fx >= 0.0f ? (uint32_t) fx : 0u, dy = fy >= 0.0f ? (uint32_t) fy : 0u;
    return (face_idx * 9u + dx * 3u + dy) & 0x3Fu;
}

constexpr int kKeyThreads = 256;   // one CTA per 32x32 tile, 4 pixels (rows warp, warp+8, +16, +24) per thread

__global__ void __launch_bounds__(kKeyThreads)
find_unique_clusters_kernel(const __grid_constant__ CUtensorMap depth_map, const __grid_constant__ CUtensorMap normal_map,
                            const float* __restrict__ depth, const uint2* __restrict__ normals,
                            const float* __restrict__ thresholds, cluster_key_params prm,
                            uint32_t* tile_count, uint32_t* cluster_ref, uint32_t* tile_keys)
{
    __shared__ alignas(128) float s_depth[32 * 32];
    __shared__ alignas(128) uint2 s_normal[32 * 32];
    __shared__ alignas(16) uint32_t s_bitmap[2048];
    __shared__ alignas(16) uint32_t s_prefix[2048];
    __shared__ uint32_t s_warp[kKeyThreads / 32];
    __shared__ alignas(8) uint64_t s_bar;

    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // Tiles are independent here: keys, TILE-LOCAL cluster numbers and the number of unique keys of the tile.  The global
    // numbering (tile-major order) needs an exclusive prefix over the tiles; with a decoupled look-back inline, 46 % of the
    // warp samples sat at the barrier behind it (profiles/r1d_ncu_clustered.md), and a deferred in-kernel finalize stage
    // was slower still, so the prefix is a separate scan of the per-tile counts followed by finalize_clusters_kernel.
    const uint32_t tile = blockIdx.x;
    const uint32_t ti = tile % prm.tiles_x, tj = tile / prm.tiles_x;   // canonical tile-major order
    const bool interior = (ti << 5) + 32 <= prm.width && (tj << 5) + 32 <= prm.height;
    const bool tma = prm.use_tma && interior;
    if (tid == 0 && tma)
    {
        mbar_init(&s_bar, 1);
        mbar_fence_init();
        mbar_arrive_expect_tx(&s_bar, 32 * 32 * 4 + (prm.has_normals ? 32 * 32 * 8 : 0));
        tma_load_2d(s_depth, &depth_map, (int) (ti << 5), (int) (tj << 5), &s_bar);
        if (prm.has_normals) tma_load_2d(s_normal, &normal_map, (int) (ti << 5), (int) (tj << 5), &s_bar);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) s_bitmap[tid + i * kKeyThreads] = 0;
    __syncthreads();
    if (tma) mbar_wait(&s_bar, 0);

    uint32_t sub[4];
#pragma unroll
    for (int r = 0; r < 4; r++)
    {
        const uint32_t row = warp + 8 * r;
        const uint32_t x = (ti << 5) + lane, y = (tj << 5) + row;      // local invocation (lane, row)
        float d;
        uint2 nraw = make_uint2(0u, 0u);
        if (tma)
        {
            d = s_depth[row * 32 + lane];
            if (prm.has_normals) nraw = s_normal[row * 32 + lane];
        }
        else
        {
            // partial tiles wrap like the reference's REPEAT sampler (gbuffer.cpp:28-34)
            const uint32_t sx = x % prm.width, sy = y % prm.height;
            d = depth[(size_t) sy * prm.width + sx];
            if (prm.has_normals) nraw = normals[(size_t) sy * prm.width + sx];
        }
        // view-space depth and slice (find_unique_clusters.comp:48-65)
        const float w = __fadd_rn(__fmul_rn(d, prm.iB), __fmul_rn(1.0f, prm.nAB));
        const float z = __fdiv_rn(1.0f, w);
        uint32_t k = 0;
        if (z > 0.0f)
        {
            const float zc = z;   // +inf included: it lands in the table's overflow run
            int g = (int) fminf(__log2f(zc * prm.inv_near) * prm.inv_log2a, 65536.0f);
            g = max(0, min(g, (int) prm.table_len - 1));
            while (g > 0 && zc < __ldg(&thresholds[g])) g--;
            while (g + 1 < (int) prm.table_len && zc >= __ldg(&thresholds[g + 1])) g++;
            k = (uint32_t) g;
        }
        uint32_t nb = 0xFFFFFFFFu;
        if (prm.has_normals)
        {
            const __half2 h01 = *reinterpret_cast<const __half2*>(&nraw.x);
            const __half2 h23 = *reinterpret_cast<const __half2*>(&nraw.y);
            nb = discretize_normal(__low2float(h01), __high2float(h01), __low2float(h23));
        }
        sub[r] = ((nb & 0x3Fu) << 10) | (k & 0x3FFu);   // key bits 16..31
        atomicOr(&s_bitmap[sub[r] >> 5], 1u << (sub[r] & 31));
    }
    __syncthreads();

    // popcount prefix over the 2048 bitmap words: thread t owns words 8t .. 8t+7
    // (two 128-bit accesses per thread: eight 32-bit ones at a stride of 8 words are 8-way bank conflicts)
    uint32_t words[8], local[8], mine = 0;
    {
        const uint4 wa = reinterpret_cast<const uint4*>(s_bitmap)[2 * tid], wb = reinterpret_cast<const uint4*>(s_bitmap)[2 * tid + 1];
        words[0] = wa.x; words[1] = wa.y; words[2] = wa.z; words[3] = wa.w;
        words[4] = wb.x; words[5] = wb.y; words[6] = wb.z; words[7] = wb.w;
    }
#pragma unroll
    for (int i = 0; i < 8; i++)
    {
        local[i] = mine;
        mine += __popc(words[i]);
    }
    uint32_t inc = mine;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1)
    {
        const uint32_t t = __shfl_up_sync(kFullMask, inc, s);
        if (lane >= (unsigned) s) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t warp_prefix = 0, unique = 0;
#pragma unroll
    for (int w = 0; w < kKeyThreads / 32; w++)
    {
        const uint32_t t = s_warp[w];
        if (w < (int) warp) warp_prefix += t;
        unique += t;
    }
    const uint32_t excl = warp_prefix + inc - mine;
    reinterpret_cast<uint4*>(s_prefix)[2 * tid] = make_uint4(excl + local[0], excl + local[1], excl + local[2], excl + local[3]);
    reinterpret_cast<uint4*>(s_prefix)[2 * tid + 1] = make_uint4(excl + local[4], excl + local[5], excl + local[6], excl + local[7]);
    __syncthreads();   // s_prefix complete

    // per-pixel cluster reference (imageStore, find_unique_clusters.comp:113-120), tile-local for now: the finalizer
    // of this tile adds the tile's base.  Out-of-image pixels are dropped.
#pragma unroll
    for (int r = 0; r < 4; r++)
    {
        const uint32_t x = (ti << 5) + lane, y = (tj << 5) + warp + 8 * r;
        const uint32_t word = s_bitmap[sub[r] >> 5];
        const uint32_t ref = s_prefix[sub[r] >> 5] + __popc(word & ((1u << (sub[r] & 31)) - 1u));
        if (x < prm.width && y < prm.height) cluster_ref[(size_t) y * prm.width + x] = ref;
    }

    // unique keys, ascending inside the tile, parked in the tile's slot of the staging area
    const uint32_t tile_bits = (ti & 0xFFu) | ((tj & 0xFFu) << 8);
    uint32_t* parked = tile_keys + (size_t) tile * 1024;
    uint32_t out = excl;
#pragma unroll
    for (int i = 0; i < 8; i++)
    {
        uint32_t bits = words[i];
        while (bits)
        {
            const uint32_t b = __ffs(bits) - 1;
            bits &= bits - 1;
            parked[out++] = tile_bits | (((8 * tid + i) * 32 + b) << 16);
        }
    }
    if (tid == 0) tile_count[tile] = unique;
}

// second half of a7: global cluster numbers and the key list, from the exclusive scan of the per-tile counts.
// One CTA per tile, pure streaming (the tile's 4 KB of cluster numbers are read-modified-written, its parked keys copied).
__global__ void __launch_bounds__(kKeyThreads)
finalize_clusters_kernel(cluster_key_params prm, const uint32_t* __restrict__ tile_count, const uint32_t* __restrict__ tile_base,
                         const uint32_t* __restrict__ tile_keys, uint32_t* keys_out, uint32_t* dispatch_params, uint32_t* cluster_ref)
{
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile = blockIdx.x;
    const uint32_t base = tile_base[tile], unique = tile_count[tile];
    if (tid == 0)
    {
        if (base + unique > prm.max_keys) atomicOr(&dispatch_params[3], 1u);   // overflow, detected not silent
        if (tile == prm.tiles_x * prm.tiles_y - 1)
        {
            // indirect-dispatch block of the reference: {count, 1, 1, _} (clustered_shading.cpp:386-394)
            dispatch_params[0] = min(base + unique, prm.max_keys);
            dispatch_params[1] = 1;
            dispatch_params[2] = 1;
        }
    }
    const uint32_t* parked = tile_keys + (size_t) tile * 1024;
    for (uint32_t i = tid; i < unique; i += kKeyThreads)
        if (base + i < prm.max_keys) keys_out[base + i] = parked[i];
    if (base != 0)
    {
        const uint32_t ti = tile % prm.tiles_x, tj = tile / prm.tiles_x;
        uint32_t v[4];
#pragma unroll
        for (int r = 0; r < 4; r++)
        {
            const uint32_t x = (ti << 5) + lane, y = (tj << 5) + warp + 8 * r;
            v[r] = (x < prm.width && y < prm.height) ? cluster_ref[(size_t) y * prm.width + x] : 0u;
        }
#pragma unroll
        for (int r = 0; r < 4; r++)
        {
            const uint32_t x = (ti << 5) + lane, y = (tj << 5) + warp + 8 * r;
            if (x < prm.width && y < prm.height) cluster_ref[(size_t) y * prm.width + x] = v[r] + base;
        }
    }
}

// ---- a8 kernel ----------------------------------------------------------------------------------------------------
struct assign_params
{
    uint32_t tiles_x, tiles_y;
    float i00, i11, nAB;
    float near_plane, a;
    uint32_t bvh_root, levels, light_count;
    uint32_t max_keys, max_assigned;
};

struct float3x { float v[3]; };

// clustered_shading.glsl:72-110, exactly as written (corners are NOT re-ordered component-wise)
__device__ __forceinline__ void cluster_aabb(uint32_t key, const assign_params& prm, const float* __restrict__ near_table,
                                             float3x& cmin, float3x& cmax)
{
    const uint32_t ci = key & 0xFFu, cj = (key >> 8) & 0xFFu, ck = (key >> 16) & 0x3FFu;
    const float tx = (float) prm.tiles_x, ty = (float) prm.tiles_y;
    float p0x = __fdiv_rn((float) ci, tx), p0y = __fdiv_rn((float) cj, ty);
    p0x = __fsub_rn(__fmul_rn(p0x, 2.0f), 1.0f);
    p0y = __fsub_rn(__fmul_rn(__fsub_rn(1.0f, p0y), 2.0f), 1.0f);
    float p1x = __fdiv_rn((float) (ci + 1), tx), p1y = __fdiv_rn((float) (cj + 1), ty);
    p1x = __fsub_rn(__fmul_rn(p1x, 2.0f), 1.0f);
    p1y = __fsub_rn(__fmul_rn(__fsub_rn(1.0f, p1y), 2.0f), 1.0f);
    const float w = prm.nAB; // (0*iB) + (1*nAB)
    const float v0x = __fdiv_rn(__fmul_rn(prm.i00, p0x), w), v0y = __fdiv_rn(__fmul_rn(prm.i11, p0y), w), vz = __fdiv_rn(1.0f, w);
    const float v1x = __fdiv_rn(__fmul_rn(prm.i00, p1x), w), v1y = __fdiv_rn(__fmul_rn(prm.i11, p1y), w);
    const float near_k = near_table[ck];
    const float far_k = __fmul_rn(near_k, prm.a);
    const float l0 = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(v0x, v0x), __fmul_rn(v0y, v0y)), __fmul_rn(vz, vz)));
    const float l1 = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(v1x, v1x), __fmul_rn(v1y, v1y)), __fmul_rn(vz, vz)));
    const float d0[3] = {__fdiv_rn(v0x, l0), __fdiv_rn(v0y, l0), __fdiv_rn(vz, l0)};
    const float d1[3] = {__fdiv_rn(v1x, l1), __fdiv_rn(v1y, l1), __fdiv_rn(vz, l1)};
    const float s0 = __fdiv_rn(near_k, d0[2]), s1 = __fdiv_rn(far_k, d1[2]);
#pragma unroll
    for (int c = 0; c < 3; c++)
    {
        cmin.v[c] = __fmul_rn(s0, d0[c]);
        cmax.v[c] = __fmul_rn(s1, d1[c]);
    }
}

// test_aabb_aabb(cluster_min, cluster_max, node._min, node._max), assign_lights.comp:84-94; INVALID nodes never overlap
__device__ __forceinline__ bool node_overlaps(const float* cmin, const float* cmax, const float4& lo, const float4& hi)
{
    return __float_as_uint(lo.w) != kInvalid && cmax[0] >= lo.x && cmin[0] <= hi.x && cmax[1] >= lo.y && cmin[1] <= hi.y &&
           cmax[2] >= lo.z && cmin[2] <= hi.z;
}

// test_sphere_aabb(light centre, radius, cluster_min, cluster_max), assign_lights.comp:97-102, on the leaf record
// {centre.xyz, T}: T = sq_threshold(radius), so "length(p - o) < radius" == "dot(p - o, p - o) < T" bit for bit
__device__ __forceinline__ bool leaf_hits(const float* cmin, const float* cmax, const float4& o)
{
    // min/max pick identical values as the GLSL ternaries (only the sign of a zero may differ, squared away)
    const float q0 = __fsub_rn(fmaxf(cmin[0], fminf(o.x, cmax[0])), o.x);
    const float q1 = __fsub_rn(fmaxf(cmin[1], fminf(o.y, cmax[1])), o.y);
    const float q2 = __fsub_rn(fmaxf(cmin[2], fminf(o.z, cmax[2])), o.z);
    const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(q0, q0), __fmul_rn(q1, q1)), __fmul_rn(q2, q2));
    return d2 < o.w;
}

constexpr int kAssignThreads = 256;
constexpr int kAssignWarps = kAssignThreads / 32;
constexpr int kHitCap = 96;    // leaf groups with hits remembered per cluster before falling back to a second walk

struct assign_state
{
    uint32_t ticket;        // next cluster to process (dynamic: balances the very uneven per-cluster work)
    uint32_t arena_top;     // bump allocator over the arena of provisional light lists
    uint32_t _pad[62];
};

// One walk of the light BVH for one cluster (assign_lights.comp:121-241), warp-uniform control flow.
// The reference's state machine (ENTER / ADVANCE / POP over per-level overlap bitmasks) is a depth-first walk that
// visits the set bits of every level lowest first; with the level count known at launch it unrolls into LEVELS nested
// loops whose masks live in registers.
//   WRITE == false: counts hits, remembers (leaf group, hit mask) in `hits` (up to kHitCap entries)
//   WRITE == true : writes the light indices at out_base in traversal order, descending lane inside a group
struct walk_ctx
{
    float cmin[3], cmax[3];
    const float4* __restrict__ bvh;
    const float4* __restrict__ spheres;
    const uint2* __restrict__ sorted_pairs;
    uint32_t level_base[kMaxBvhLevels];
    uint32_t light_count, max_assigned;
    uint2* hits;
    uint32_t* out;
    uint32_t out_base;
    uint32_t running, groups;
    uint32_t stat_nodes, stat_leaves;
};

template <int LEVEL, int LEVELS, bool WRITE>
struct walker
{
    static __device__ __forceinline__ void visit(walk_ctx& c, uint32_t idx)
    {
        const unsigned lane = threadIdx.x & 31;
        const uint32_t addr = c.level_base[LEVEL] + idx * 32 + lane;
        if constexpr (LEVEL + 1 < LEVELS)
        {
            // test_aabb_aabb(cluster_min, cluster_max, node.min, node.max), assign_lights.comp:84-94
            const float4 lo = c.bvh[2 * (size_t) addr], hi = c.bvh[2 * (size_t) addr + 1];
            const bool overlap = node_overlaps(c.cmin, c.cmax, lo, hi);
            unsigned mask = __ballot_sync(kFullMask, overlap);
            c.stat_nodes += 32;
            while (mask != 0)
            {
                const uint32_t child = __ffs(mask) - 1;
                mask &= mask - 1;
                walker<LEVEL + 1, LEVELS, WRITE>::visit(c, idx * 32 + child);
            }
        }
        else
        {
            // leaf group: sphere (light) vs cluster box, assign_lights.comp:97-102,209-214.  The leaf record
            // {view_pos.xyz, T} comes from the light-BVH build; T is the smallest float s with sqrt(s) >= radius, so
            // "length(p - o) < radius" == "dot(p - o, p - o) < T" bit for bit, without the square root.
            bool hit = false;
            if (addr < c.light_count)
            {
                hit = leaf_hits(c.cmin, c.cmax, c.spheres[addr]);
            }
            const unsigned mask = __ballot_sync(kFullMask, hit);
            c.stat_leaves += 32;
            if (mask != 0)
            {
                if (WRITE)
                {
                    if (hit)
                    {
                        const uint32_t slot = c.out_base + c.running + (__popc(mask & lanemask_ge()) - 1);   // descending lane order
                        if (slot < c.max_assigned) c.out[slot] = c.sorted_pairs[addr].y;
                    }
                }
                else
                {
                    if (c.groups < (uint32_t) kHitCap && lane == 0) c.hits[c.groups] = make_uint2(addr, mask);
                    c.groups++;
                }
                c.running += __popc(mask);
            }
        }
    }
};

__device__ __forceinline__ void init_walk_ctx(walk_ctx& c, const assign_params& prm, const float4* bvh, const float4* spheres,
                                              const uint2* sorted_pairs)
{
    c.bvh = bvh; c.spheres = spheres; c.sorted_pairs = sorted_pairs;
    c.light_count = prm.light_count; c.max_assigned = prm.max_assigned;
    // level_base[l] = address of the first node of traversal level l (0 = children of the root), assign_lights.comp:108-119
    uint32_t acc = 0, p = 1;
#pragma unroll
    for (int l = 0; l < kMaxBvhLevels; l++)
    {
        p *= 32;
        acc += p;
        c.level_base[l] = prm.bvh_root - acc;
    }
    c.stat_nodes = c.stat_leaves = 0;
}

// Phase 1 — ONE traversal per cluster: count the hits, remember (leaf group, hit mask), bump-allocate the list in the
// arena and write the light indices there in their final within-list order.
template <int LEVELS>
__global__ void __launch_bounds__(kAssignThreads)
assign_lights_walk_kernel(const uint32_t* __restrict__ cluster_keys, const uint32_t* __restrict__ dispatch_params,
                          const float4* __restrict__ bvh, const uint2* __restrict__ sorted_pairs, const float4* __restrict__ spheres,
                          const float* __restrict__ near_table, assign_params prm, assign_state* state,
                          uint32_t* counts, uint32_t* alloc, uint32_t* arena, uint32_t* status_out)
{
    __shared__ uint2 s_hits[kAssignWarps][kHitCap];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t count = min(dispatch_params[0], prm.max_keys);
    const unsigned ge = lanemask_ge();
    walk_ctx ctx;
    init_walk_ctx(ctx, prm, bvh, spheres, sorted_pairs);
    ctx.hits = s_hits[warp];

    while (true)
    {
        uint32_t c = 0;
        if (lane == 0) c = atomicAdd(&state->ticket, 1u);
        c = __shfl_sync(kFullMask, c, 0);
        if (c >= count) break;

        float3x cmin, cmax;
        cluster_aabb(cluster_keys[c], prm, near_table, cmin, cmax);
#pragma unroll
        for (int k = 0; k < 3; k++) { ctx.cmin[k] = cmin.v[k]; ctx.cmax[k] = cmax.v[k]; }
        ctx.running = ctx.groups = 0;
        walker<0, LEVELS, false>::visit(ctx, 0);
        const uint32_t total = ctx.running, groups = ctx.groups;
        uint32_t at = 0;
        if (lane == 0)
        {
            at = total ? atomicAdd(&state->arena_top, total) : 0u;
            counts[c] = total;
            alloc[c] = at;
        }
        at = __shfl_sync(kFullMask, at, 0);
        if (total == 0 || (uint64_t) at + total > prm.max_assigned) continue;   // arena full: phase 3 re-walks everything
        if (groups <= (uint32_t) kHitCap)
        {
            __syncwarp();
            uint32_t running = 0;
            for (uint32_t g = 0; g < groups; g++)
            {
                const uint2 h = ctx.hits[g];
                if (h.y & (1u << lane)) arena[at + running + (__popc(h.y & ge) - 1)] = sorted_pairs[h.x + lane].y;
                running += __popc(h.y);
            }
            __syncwarp();
        }
        else
        {
            ctx.out = arena; ctx.out_base = at; ctx.max_assigned = 0xFFFFFFFFu; // bounds were checked against the arena above
            ctx.running = ctx.groups = 0;
            walker<0, LEVELS, true>::visit(ctx, 0);
            ctx.max_assigned = prm.max_assigned;
        }
    }
    if (status_out != nullptr && lane == 0)
    {
        if (ctx.stat_nodes) atomicAdd(&status_out[2], ctx.stat_nodes);
        if (ctx.stat_leaves) atomicAdd(&status_out[3], ctx.stat_leaves);
    }
}

// Phase 3 — lists move from the arena to offsets[c] (exclusive scan of the counts in cluster-list order, phase 2).
// If the arena overflowed (more hits than max_assigned) every cluster is walked again and written in place, clamped.
template <int LEVELS>
__global__ void __launch_bounds__(kAssignThreads)
assign_lights_place_kernel(const uint32_t* __restrict__ cluster_keys, const uint32_t* __restrict__ dispatch_params,
                           const float4* __restrict__ bvh, const uint2* __restrict__ sorted_pairs, const float4* __restrict__ spheres,
                           const float* __restrict__ near_table, assign_params prm, const assign_state* state,
                           const uint32_t* __restrict__ counts, const uint32_t* __restrict__ alloc, const uint32_t* __restrict__ arena,
                           const uint32_t* __restrict__ offsets, uint32_t* indices, uint32_t* status_out)
{
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t count = min(dispatch_params[0], prm.max_keys);
    const uint32_t total = state->arena_top;
    const bool overflow = total > prm.max_assigned;
    if (blockIdx.x == 0 && threadIdx.x == 0 && status_out != nullptr)
    {
        status_out[0] = total;                 // reference overruns VREN_MAX_ASSIGNED_LIGHT_COUNT silently (config.hpp:24)
        status_out[1] = overflow ? 1u : 0u;
    }
    const uint32_t total_warps = gridDim.x * kAssignWarps;
    if (!overflow)
    {
        // four clusters per step: their (count, arena slot, offset) loads, then their first two 32-wide chunks, are
        // independent and in flight together (one cluster at a time is three dependent L2 round trips per ~50 indices)
        constexpr int U = 4;
        for (uint32_t c0 = blockIdx.x * kAssignWarps + warp; c0 < count; c0 += U * total_warps)
        {
            uint32_t n[U], src[U], dst[U], a[U], b[U];
#pragma unroll
            for (int u = 0; u < U; u++)
            {
                const uint32_t c = c0 + u * total_warps;
                n[u] = c < count ? counts[c] : 0u;
                src[u] = c < count ? alloc[c] : 0u;
                dst[u] = c < count ? offsets[c] : 0u;
            }
#pragma unroll
            for (int u = 0; u < U; u++)
            {
                a[u] = lane < n[u] ? arena[src[u] + lane] : 0u;
                b[u] = lane + 32 < n[u] ? arena[src[u] + lane + 32] : 0u;
            }
#pragma unroll
            for (int u = 0; u < U; u++)
            {
                if (lane < n[u]) indices[dst[u] + lane] = a[u];
                if (lane + 32 < n[u]) indices[dst[u] + lane + 32] = b[u];
                for (uint32_t i = lane + 64; i < n[u]; i += 32) indices[dst[u] + i] = arena[src[u] + i];
            }
        }
        return;
    }
    walk_ctx ctx;
    init_walk_ctx(ctx, prm, bvh, spheres, sorted_pairs);
    ctx.hits = nullptr;
    ctx.out = indices;
    for (uint32_t c = blockIdx.x * kAssignWarps + warp; c < count; c += total_warps)
    {
        float3x cmin, cmax;
        cluster_aabb(cluster_keys[c], prm, near_table, cmin, cmax);
#pragma unroll
        for (int k = 0; k < 3; k++) { ctx.cmin[k] = cmin.v[k]; ctx.cmax[k] = cmax.v[k]; }
        ctx.running = ctx.groups = 0;
        ctx.out_base = offsets[c];
        walker<0, LEVELS, true>::visit(ctx, 0);
    }
}

template <int LEVELS>
int launch_assign(cudaStream_t s, const uint32_t* cluster_keys, const uint32_t* dispatch_params, const float4* bvh, const uint2* pairs,
                  const float4* spheres, const float* near_table, const assign_params& prm, assign_state* state, uint32_t* counts,
                  uint32_t* alloc, uint32_t* arena, uint32_t* offsets, uint32_t* indices, uint32_t* status_out,
                  vrenb200_stream_t stream, void* scan_scratch, size_t scan_bytes)
{
    const int grid = kNumSMs * 8;
    assign_lights_walk_kernel<LEVELS><<<grid, kAssignThreads, 0, s>>>(cluster_keys, dispatch_params, bvh, pairs, spheres, near_table, prm,
                                                                     state, counts, alloc, arena, status_out);
    VRENB200_TRY(check_launch());
    // copy counts -> offsets + blelloch_scan over all max_keys slots (clustered_shading.cpp:586-619), one pass here
    VRENB200_TRY(vrenb200_exclusive_scan_u32(stream, counts, offsets, prm.max_keys, scan_scratch, scan_bytes));
    assign_lights_place_kernel<LEVELS><<<grid, kAssignThreads, 0, s>>>(cluster_keys, dispatch_params, bvh, pairs, spheres, near_table, prm,
                                                                      state, counts, alloc, arena, offsets, indices, status_out);
    return check_launch();
}

// Diagnostics: the device functions above evaluated one element per thread (vrenb200_cluster_tests)
__global__ void __launch_bounds__(256)
cluster_tests_kernel(const uint32_t* __restrict__ keys, uint32_t count, assign_params prm, const float* __restrict__ near_table,
                     float* out_min3, float* out_max3, const float* __restrict__ node_boxes6, const float* __restrict__ spheres4,
                     uint8_t* out_flags)
{
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i >= count) return;
    float3x cmin, cmax;
    cluster_aabb(keys[i], prm, near_table, cmin, cmax);
    if (out_min3 != nullptr)
    {
#pragma unroll
        for (int k = 0; k < 3; k++) { out_min3[3 * (size_t) i + k] = cmin.v[k]; out_max3[3 * (size_t) i + k] = cmax.v[k]; }
    }
    uint32_t flags = 0;
    if (node_boxes6 != nullptr)
    {
        const float* b = node_boxes6 + 6 * (size_t) i;
        if (node_overlaps(cmin.v, cmax.v, make_float4(b[0], b[1], b[2], __uint_as_float(0u)), make_float4(b[3], b[4], b[5], 0.0f))) flags |= 1u;
    }
    if (spheres4 != nullptr)
    {
        const float* sp = spheres4 + 4 * (size_t) i;
        if (leaf_hits(cmin.v, cmax.v, make_float4(sp[0], sp[1], sp[2], sq_threshold(sp[3])))) flags |= 2u;
    }
    if (out_flags != nullptr) out_flags[i] = (uint8_t) flags;
}

// ---- tensor maps ----------------------------------------------------------------------------------------------------
using encode_fn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

encode_fn get_encode_fn()
{
    static encode_fn fn = []() -> encode_fn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<encode_fn>(p);
    }();
    return fn;
}

bool make_tile_map(CUtensorMap* map, const void* base, uint32_t width, uint32_t height, uint32_t elem_bytes)
{
    encode_fn enc = get_encode_fn();
    if (!enc) return false;
    if ((reinterpret_cast<uintptr_t>(base) & 15) || ((uint64_t) width * elem_bytes) % 16 != 0) return false;
    const cuuint64_t dims[2] = {width, height};
    const cuuint64_t strides[1] = {(cuuint64_t) width * elem_bytes};
    const cuuint32_t box[2] = {32, 32};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt = elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT64;
    return enc(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

assign_params make_assign_params(uint32_t width, uint32_t height, const vrenb200_camera& camera)
{
    const proj_consts pc = make_proj(camera);
    assign_params prm{};
    prm.tiles_x = (width + 31) / 32; prm.tiles_y = (height + 31) / 32;
    prm.i00 = pc.i00; prm.i11 = pc.i11; prm.nAB = pc.nAB;
    prm.near_plane = camera.near_plane;
    prm.a = 2.0f * pc.tan_half / (float) prm.tiles_y + 1.0f;            // clustered_shading.glsl:96
    return prm;
}

// near_k = near * pow(a, k), k < 1024 (clustered_shading.glsl:97) evaluated on the host with powf; cached per device
const float* get_near_table(const vrenb200_camera& camera, float a)
{
    std::lock_guard<std::mutex> lock(g_table_mutex);
    const device_table* cached = find_device_table(1, camera.near_plane, a);
    if (cached == nullptr)
    {
        float near_host[1024];
        for (int k = 0; k < 1024; k++) near_host[k] = camera.near_plane * powf(a, (float) k);
        cached = add_device_table(1, camera.near_plane, a, near_host, 1024);
        if (cached == nullptr) return nullptr;
    }
    return cached->data;
}

} // namespace
} // namespace vrenb200

using namespace vrenb200;

extern "C" int vrenb200_cluster_tests(vrenb200_stream_t stream, uint32_t width, uint32_t height, const vrenb200_camera* camera,
                                      const uint32_t* cluster_keys, uint32_t count, float* out_min3, float* out_max3,
                                      const float* node_boxes6, const float* spheres4, uint8_t* out_flags)
{
    if (!camera || (count > 0 && !cluster_keys) || ((out_min3 == nullptr) != (out_max3 == nullptr))) return VRENB200_EINVAL_ARG;
    if (count == 0) return VRENB200_OK;
    const assign_params prm = make_assign_params(width, height, *camera);
    const float* near_table = get_near_table(*camera, prm.a);
    if (near_table == nullptr) return VRENB200_ECUDA;
    cluster_tests_kernel<<<(count + 255) / 256, 256, 0, as_stream(stream)>>>(cluster_keys, count, prm, near_table, out_min3, out_max3,
                                                                               node_boxes6, spheres4, out_flags);
    return check_launch();
}

extern "C" size_t vrenb200_find_unique_clusters_scratch_bytes(uint32_t width, uint32_t height)
{
    const size_t tiles = (size_t) ((width + 31) / 32) * ((height + 31) / 32);
    // slice thresholds | per-tile counts | per-tile bases | scan scratch | parked keys: 1024 slots per tile (a 32x32 tile
    // cannot hold more distinct keys)
    const size_t t = std::max<size_t>(tiles, 1);
    return align_up(kMaxSlices * sizeof(float), 256) + 2 * align_up(t * 4, 256) + align_up(vrenb200_scan_scratch_bytes((uint32_t) t), 256) +
           t * 1024 * sizeof(uint32_t);
}

extern "C" int vrenb200_find_unique_clusters(vrenb200_stream_t stream,
                                             const float* depth, const void* normals_rgba16f,
                                             uint32_t width, uint32_t height, const vrenb200_camera* camera,
                                             uint32_t* keys_out, uint32_t max_keys, uint32_t* dispatch_params,
                                             uint32_t* cluster_ref, void* scratch, size_t scratch_bytes)
{
    if (!depth || !camera || !keys_out || !dispatch_params || !cluster_ref) return VRENB200_EINVAL_ARG;
    if (width == 0 || height == 0) return VRENB200_EINVAL_LENGTH;
    const uint32_t tiles_x = (width + 31) / 32, tiles_y = (height + 31) / 32;
    if (tiles_x > 256 || tiles_y > 256) return VRENB200_ELIMIT;      // 8-bit tile fields of the key (find_unique_clusters.comp:72-76)
    if (scratch == nullptr || scratch_bytes < vrenb200_find_unique_clusters_scratch_bytes(width, height)) return VRENB200_ESCRATCH;
    if (reinterpret_cast<uintptr_t>(scratch) & 255) return VRENB200_EALIGN;
    cudaStream_t s = as_stream(stream);

    const proj_consts pc = make_proj(*camera);
    const float a = 1.0f + (2.0f * pc.tan_half) / (float) tiles_y;
    const float* thresholds = nullptr;
    uint32_t table_len = 0;
    {
        std::lock_guard<std::mutex> lock(g_table_mutex);
        const device_table* cached = find_device_table(0, camera->near_plane, a);
        if (cached == nullptr)
        {
            const slice_table& tab = get_slice_table(camera->near_plane, a);
            cached = add_device_table(0, camera->near_plane, a, tab.thresholds.data(), (uint32_t) tab.thresholds.size());
            if (cached == nullptr) return VRENB200_ECUDA;
        }
        thresholds = cached->data;
        table_len = cached->count;
    }

    const uint32_t num_tiles = tiles_x * tiles_y;
    char* cursor = static_cast<char*>(scratch);
    cursor += align_up(kMaxSlices * sizeof(float), 256);                // (formerly the per-call copy of the thresholds)
    uint32_t* tile_count = reinterpret_cast<uint32_t*>(cursor);         cursor += align_up((size_t) num_tiles * 4, 256);
    uint32_t* tile_base = reinterpret_cast<uint32_t*>(cursor);          cursor += align_up((size_t) num_tiles * 4, 256);
    void* scan_scratch = cursor;                                        cursor += align_up(vrenb200_scan_scratch_bytes(num_tiles), 256);
    uint32_t* tile_keys = reinterpret_cast<uint32_t*>(cursor);
    VRENB200_TRY(check_cuda(cudaMemsetAsync(dispatch_params, 0, 16, s)));                  // vkCmdUpdateBuffer {0,1,1}

    cluster_key_params prm{};
    prm.width = width; prm.height = height; prm.tiles_x = tiles_x; prm.tiles_y = tiles_y;
    prm.iB = pc.iB; prm.nAB = pc.nAB;
    prm.inv_near = 1.0f / camera->near_plane;
    prm.inv_log2a = (float) (1.0 / std::log2((double) a));
    prm.table_len = table_len;
    prm.max_keys = max_keys;
    prm.has_normals = normals_rgba16f != nullptr;
    CUtensorMap dmap, nmap;
    std::memset(&dmap, 0, sizeof(dmap));
    std::memset(&nmap, 0, sizeof(nmap));
    prm.use_tma = make_tile_map(&dmap, depth, width, height, 4) &&
                  (!prm.has_normals || make_tile_map(&nmap, normals_rgba16f, width, height, 8));
    find_unique_clusters_kernel<<<num_tiles, kKeyThreads, 0, s>>>(dmap, nmap, depth, static_cast<const uint2*>(normals_rgba16f), thresholds, prm,
                                                         tile_count, cluster_ref, tile_keys);
    VRENB200_TRY(check_launch());
    VRENB200_TRY(vrenb200_exclusive_scan_u32(stream, tile_count, tile_base, num_tiles, scan_scratch, vrenb200_scan_scratch_bytes(num_tiles)));
    finalize_clusters_kernel<<<num_tiles, kKeyThreads, 0, s>>>(prm, tile_count, tile_base, tile_keys, keys_out, dispatch_params, cluster_ref);
    return check_launch();
}

namespace vrenb200 { size_t light_leaf_sphere_offset(uint32_t light_count); }

extern "C" size_t vrenb200_assign_lights_scratch_bytes(uint32_t max_keys, uint32_t max_assigned)
{
    // near_k table | state | alloc[max_keys] | scan scratch | arena[max_assigned]
    return 1024 * sizeof(float) + 256 + align_up((size_t) max_keys * 4, 256) + vrenb200_scan_scratch_bytes(max_keys) +
           align_up((size_t) max_assigned * 4, 256);
}

extern "C" int vrenb200_assign_lights(vrenb200_stream_t stream,
                                      uint32_t width, uint32_t height, const vrenb200_camera* camera,
                                      const uint32_t* cluster_keys, const uint32_t* dispatch_params, uint32_t max_keys,
                                      const void* bvh_buffer, uint32_t bvh_root_index, uint32_t light_count,
                                      const void* light_index_buffer, const float* view_pos,
                                      uint32_t* indices_out, uint32_t max_assigned,
                                      uint32_t* counts_out, uint32_t* offsets_out, uint32_t* status_out,
                                      void* scratch, size_t scratch_bytes)
{
    if (!camera || !cluster_keys || !dispatch_params || !counts_out || !offsets_out || !indices_out) return VRENB200_EINVAL_ARG;
    if (max_keys == 0) return VRENB200_EINVAL_LENGTH;
    cudaStream_t s = as_stream(stream);
    // vkCmdFillBuffer(counts, 0) — the only effect when there are no lights (clustered_shading.cpp:492-496)
    VRENB200_TRY(check_cuda(cudaMemsetAsync(counts_out, 0, (size_t) max_keys * 4, s)));
    if (status_out) VRENB200_TRY(check_cuda(cudaMemsetAsync(status_out, 0, 16, s)));
    if (light_count == 0) return VRENB200_OK;
    if (!bvh_buffer || !light_index_buffer || !view_pos) return VRENB200_EINVAL_ARG;
    if (scratch == nullptr || scratch_bytes < vrenb200_assign_lights_scratch_bytes(max_keys, max_assigned)) return VRENB200_ESCRATCH;
    if (reinterpret_cast<uintptr_t>(scratch) & 255) return VRENB200_EALIGN;
    const uint32_t levels = vrenb200_calc_bvh_level_count(light_count);
    if (levels > (uint32_t) kMaxBvhLevels) return VRENB200_ELIMIT;

    assign_params prm = make_assign_params(width, height, *camera);
    prm.bvh_root = bvh_root_index; prm.levels = levels; prm.light_count = light_count;
    prm.max_keys = max_keys; prm.max_assigned = max_assigned;
    const float* near_table = get_near_table(*camera, prm.a);
    if (near_table == nullptr) return VRENB200_ECUDA;
    char* sp = static_cast<char*>(scratch);
    assign_state* state = reinterpret_cast<assign_state*>(sp + 1024 * sizeof(float));
    uint32_t* alloc = reinterpret_cast<uint32_t*>(sp + 1024 * sizeof(float) + 256);
    void* scan_scratch = sp + 1024 * sizeof(float) + 256 + align_up((size_t) max_keys * 4, 256);
    const size_t scan_bytes = vrenb200_scan_scratch_bytes(max_keys);
    uint32_t* arena = reinterpret_cast<uint32_t*>(static_cast<char*>(scan_scratch) + scan_bytes);
    VRENB200_TRY(check_cuda(cudaMemsetAsync(state, 0, sizeof(assign_state), s)));

    const float4* bvh = static_cast<const float4*>(bvh_buffer);
    const uint2* pairs = static_cast<const uint2*>(light_index_buffer);
    // compact leaf spheres written by vrenb200_construct_point_light_bvh behind the bucket-sort counters
    const float4* spheres = reinterpret_cast<const float4*>(static_cast<const char*>(light_index_buffer) + light_leaf_sphere_offset(light_count));
    (void) view_pos; // its xyz live in the leaf spheres; kept in the signature like the reference (assign_lights.comp:72-75)
#define VRENB200_ASSIGN(L) case L: return launch_assign<L>(s, cluster_keys, dispatch_params, bvh, pairs, spheres, near_table, prm, state, \
                                                       counts_out, alloc, arena, offsets_out, indices_out, status_out, stream,      \
                                                       scan_scratch, scan_bytes)
    switch (levels)
    {
        VRENB200_ASSIGN(1); VRENB200_ASSIGN(2); VRENB200_ASSIGN(3); VRENB200_ASSIGN(4); VRENB200_ASSIGN(5); VRENB200_ASSIGN(6);
    default: return VRENB200_ELIMIT;
    }
#undef VRENB200_ASSIGN
}

// ---- n1: the consumer side of the light lists ---------------------------------------------------------------
// Per pixel: cluster = cluster_reference(x, y); walk indices[offsets[cluster] .. + counts[cluster]) — the access pattern of
// shade.comp:101-105 — and emit {count, XOR of the light indices}, the integer core of the demo's list visualiser
// (vren_demo/resources/shaders/show_clusters.comp:97-118).  The reference walks the list once per PIXEL; every pixel of
// a cluster sees the same list, so here the XOR is folded once per CLUSTER (one warp each) and the per-pixel kernel is a
// pure gather: 4 B/px read + 8 B/px written instead of ~count x 4 B/px of L2 traffic.
namespace vrenb200 {
namespace {

__global__ void __launch_bounds__(256)
cluster_list_hash_kernel(const uint32_t* __restrict__ dispatch_params, uint32_t max_keys, const uint32_t* __restrict__ counts,
                         const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ indices, uint32_t max_assigned,
                         uint32_t* __restrict__ cluster_hash)
{
    const uint32_t count = min(dispatch_params[0], max_keys);
    const unsigned lane = threadIdx.x & 31;
    for (uint32_t c = (blockIdx.x * 256u + threadIdx.x) >> 5; c < count; c += (gridDim.x * 256u) >> 5)
    {
        const uint32_t n = counts[c], off = offsets[c];
        uint32_t h = 0;
        for (uint32_t i = lane; i < n; i += 32)
            if (off + i < max_assigned) h ^= indices[off + i];
        h = __reduce_xor_sync(kFullMask, h);
        if (lane == 0) cluster_hash[c] = h;
    }
}

__global__ void __launch_bounds__(256)
pixel_light_list_kernel(const uint32_t* __restrict__ cluster_ref, uint32_t pixels, uint32_t max_keys, const uint32_t* __restrict__ counts,
                        const uint32_t* __restrict__ cluster_hash, uint2* __restrict__ out)
{
    for (uint32_t p = blockIdx.x * 256u + threadIdx.x; p < pixels; p += gridDim.x * 256u)
    {
        const uint32_t c = cluster_ref[p];
        out[p] = c < max_keys ? make_uint2(counts[c], cluster_hash[c]) : make_uint2(0u, 0u);
    }
}

} // namespace
} // namespace vrenb200

extern "C" size_t vrenb200_light_list_hash_scratch_bytes(uint32_t max_keys) { return align_up((size_t) max_keys * 4, 256); }

extern "C" int vrenb200_light_list_hash(vrenb200_stream_t stream, uint32_t width, uint32_t height, const uint32_t* cluster_ref,
                                        const uint32_t* dispatch_params, uint32_t max_keys, const uint32_t* counts,
                                        const uint32_t* offsets, const uint32_t* indices, uint32_t max_assigned,
                                        uint32_t* out_count_hash, void* scratch, size_t scratch_bytes)
{
    if (!cluster_ref || !dispatch_params || !counts || !offsets || !indices || !out_count_hash) return VRENB200_EINVAL_ARG;
    if (width == 0 || height == 0 || max_keys == 0) return VRENB200_EINVAL_LENGTH;
    if (scratch == nullptr || scratch_bytes < vrenb200_light_list_hash_scratch_bytes(max_keys)) return VRENB200_ESCRATCH;
    if (reinterpret_cast<uintptr_t>(out_count_hash) & 7) return VRENB200_EALIGN;
    cudaStream_t s = as_stream(stream);
    uint32_t* cluster_hash = static_cast<uint32_t*>(scratch);
    cluster_list_hash_kernel<<<kNumSMs * 8, 256, 0, s>>>(dispatch_params, max_keys, counts, offsets, indices, max_assigned, cluster_hash);
    VRENB200_TRY(check_launch());
    const uint32_t pixels = width * height;
    pixel_light_list_kernel<<<kNumSMs * 16, 256, 0, s>>>(cluster_ref, pixels, max_keys, counts, cluster_hash, reinterpret_cast<uint2*>(out_count_hash));
    return check_launch();
}
