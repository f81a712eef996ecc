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

#include "common.cuh"

namespace vrenb200 {

namespace {

constexpr int kScanVecs = 8;                                        // uint4 per thread
constexpr uint32_t kScanMinTile = 256 * kScanVecs * 4;              // smallest tile of any variant: sizes the status array
constexpr uint32_t kWarpChunk = 32 * kScanVecs * 4;                 // 1024 contiguous elements per warp

constexpr uint64_t kFlagAggregate = 1ull << 32;
constexpr uint64_t kFlagInclusive = 2ull << 32;

struct scan_state
{
    uint32_t ticket;      // unused: tile id = block index (CTAs are dispatched in index order, so a tile only waits on
                          // tiles that already started; saves an L2 atomic round trip before the first load)
    uint32_t _pad[63];
    uint64_t status[1];   // [tiles]
};

// THREADS sets the tile (THREADS x 32 elements).  The look-back walk is as long as the number of older tiles still in
// flight, so for a fixed number of bytes in flight larger tiles mean proportionally fewer L2 round trips per tile.
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
exclusive_scan_u32_kernel(const uint32_t* in, uint32_t* out, uint32_t n, scan_state* state, uint32_t base)
{
    constexpr int kScanWarps = THREADS / 32;
    constexpr uint32_t kScanTile = THREADS * kScanVecs * 4;
    __shared__ uint32_t s_warp_total[kScanWarps];
    __shared__ uint32_t s_tile_prefix;

    const uint32_t tile = blockIdx.x;
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
int g_scan_variant = 0;
constexpr int kScanVariantThreads[] = { 256, 512, 1024 };
}

// tuning hook (bench.py / tests): CTA size of the scan kernel, 0: 256, 1: 512, 2: 1024 threads
extern "C" int vrenb200_scan_set_variant(int v)
{
    if (v < 0 || v >= (int) (sizeof(kScanVariantThreads) / sizeof(int))) return VRENB200_EINVAL_ARG;
    g_scan_variant = v;
    return VRENB200_OK;
}

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
    if (in == nullptr || out == nullptr) return VRENB200_EINVAL_ARG;
    if (n == 0) return VRENB200_EINVAL_LENGTH;
    const size_t need = vrenb200_scan_scratch_bytes(n);
    if (scratch == nullptr || scratch_bytes < need) return VRENB200_ESCRATCH;
    if ((reinterpret_cast<uintptr_t>(scratch) & 7) != 0) return VRENB200_EALIGN;
    cudaStream_t s = as_stream(stream);
    VRENB200_TRY(check_cuda(cudaMemsetAsync(scratch, 0, need, s)));
    const uint32_t threads = kScanVariantThreads[g_scan_variant];
    const uint32_t tile = threads * kScanVecs * 4;
    const uint32_t tiles = (uint32_t) (((size_t) n + tile - 1) / tile);
    scan_state* st = static_cast<scan_state*>(scratch);
    if (threads == 256) exclusive_scan_u32_kernel<256><<<tiles, 256, 0, s>>>(in, out, n, st, base);
    else if (threads == 512) exclusive_scan_u32_kernel<512><<<tiles, 512, 0, s>>>(in, out, n, st, base);
    else exclusive_scan_u32_kernel<1024><<<tiles, 1024, 0, s>>>(in, out, n, st, base);
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
