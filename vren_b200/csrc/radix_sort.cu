// a3 — vren::radix_sort (reference: vren/vren/primitives/radix_sort.{hpp,cpp}, shaders/radix_sort_*.comp).
//
// Reference: LSD, 4-bit digits, 8 passes, each pass = fill + local_count + reduce + global_offset + downsweep
// + reorder (~14 dispatches, 12n bytes of key traffic per pass, one key per thread, n pow2 >= 1024).
//
// Here (B200-first): "onesweep" — 8-bit digits, 4 passes.
//   1. ONE histogram kernel reads the keys once (128-bit loads) and builds all four 256-bin digit histograms.
//   2. One tiny kernel turns them into exclusive digit offsets.
//   3. Per pass ONE fused kernel.  A CTA takes a tile (dynamic ticket), stages keys (and values) into shared
//      memory with a single-thread TMA bulk copy (cp.async.bulk + mbarrier: no register staging, values land
//      while keys are being ranked), ranks keys stably with warp-ballot digit matching against warp-private
//      histograms, resolves the tile's global digit offsets with a decoupled look-back over a flag|count word
//      per (tile, digit), regroups keys/values by digit in shared memory and writes them out coalesced.
// HBM traffic: 4n (histogram) + 4 passes x 8n (keys) [+ 4 x 8n values] = 36 B/key, 68 B/pair.
// Stability: warp-striped order (warp, item, lane) == element order, so equal keys keep input order (needed
// by LSD passes, and it is the pairs contract).
#include "common.cuh"

namespace vrenb200 {

namespace {

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kPasses = 32 / kRadixBits;

constexpr uint32_t kLbFlagAggregate = 1u << 30;
constexpr uint32_t kLbFlagInclusive = 2u << 30;
constexpr uint32_t kLbValueMask = (1u << 30) - 1;

// ---- control block carved from scratch ----------------------------------------------------------------------
struct sort_control
{
    uint32_t tickets[kPasses];            // dynamic tile ids per pass
    uint32_t _pad[60];
    uint32_t hist[kPasses][kRadix];       // global digit counts, then exclusive offsets
    // followed by look-back words: [kPasses][tiles][kRadix]
};

// ---- mbarrier / bulk-copy (TMA 1D) wrappers --------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
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

// ---- 1. histogram of all four digits in one read of the keys ----------------------------------------------
constexpr int kHistThreads = 512;

__global__ void __launch_bounds__(kHistThreads)
radix_histogram_kernel(const uint32_t* __restrict__ keys, uint32_t n, sort_control* ctl)
{
    __shared__ uint32_t s_hist[kPasses][kRadix];
    for (int i = threadIdx.x; i < kPasses * kRadix; i += kHistThreads) (&s_hist[0][0])[i] = 0;
    __syncthreads();

    const uint32_t n4 = n / 4;
    const uint4* keys4 = reinterpret_cast<const uint4*>(keys);
    const uint32_t stride = gridDim.x * kHistThreads;
    auto count = [&](uint32_t k) {
        atomicAdd(&s_hist[0][k & 0xFF], 1u);
        atomicAdd(&s_hist[1][(k >> 8) & 0xFF], 1u);
        atomicAdd(&s_hist[2][(k >> 16) & 0xFF], 1u);
        atomicAdd(&s_hist[3][k >> 24], 1u);
    };
    uint32_t i = blockIdx.x * kHistThreads + threadIdx.x;
    // two independent 128-bit loads in flight per thread per iteration
    for (; i + stride < n4; i += 2 * stride)
    {
        const uint4 a = ldg_stream_u4(keys4 + i);
        const uint4 b = ldg_stream_u4(keys4 + i + stride);
        count(a.x); count(a.y); count(a.z); count(a.w);
        count(b.x); count(b.y); count(b.z); count(b.w);
    }
    for (; i < n4; i += stride)
    {
        const uint4 a = ldg_stream_u4(keys4 + i);
        count(a.x); count(a.y); count(a.z); count(a.w);
    }
    if (blockIdx.x == 0)
        for (uint32_t t = n4 * 4 + threadIdx.x; t < n; t += kHistThreads) count(keys[t]);
    __syncthreads();
    for (int t = threadIdx.x; t < kPasses * kRadix; t += kHistThreads)
    {
        const uint32_t c = (&s_hist[0][0])[t];
        if (c != 0) atomicAdd(&(&ctl->hist[0][0])[t], c);
    }
}

// ---- 2. exclusive scan of each 256-bin histogram (grid = kPasses, block = 256) ---------------------------
__global__ void __launch_bounds__(kRadix)
radix_scan_histograms_kernel(sort_control* ctl)
{
    __shared__ uint32_t s_warp[kRadix / 32];
    uint32_t* h = ctl->hist[blockIdx.x];
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

// ---- 3. the fused onesweep pass ------------------------------------------------------------------------------
template <int THREADS, int ITEMS, bool KV>
struct onesweep_smem
{
    static constexpr int WARPS = THREADS / 32;
    static constexpr int TILE = THREADS * ITEMS;
    alignas(128) uint32_t keys[TILE];
    alignas(128) uint32_t vals[KV ? TILE : 4];
    uint32_t warp_hist[WARPS][kRadix];
    uint32_t digit_base[kRadix];
    uint32_t scan_warp[kRadix / 32];
    alignas(8) uint64_t bar_keys;
    alignas(8) uint64_t bar_vals;
    uint32_t tile;
};

enum { MATCH_BALLOT = 0, MATCH_HW = 1 };

template <int MATCH>
__device__ __forceinline__ unsigned match_digit(uint32_t d)
{
    if (MATCH == MATCH_HW)
    {
        return __match_any_sync(kFullMask, d);
    }
    else
    {
        unsigned mask = kFullMask;
#pragma unroll
        for (int b = 0; b < kRadixBits; b++)
        {
            const bool bit = (d >> b) & 1u;
            const unsigned bal = __ballot_sync(kFullMask, bit);
            mask &= bit ? bal : ~bal;
        }
        return mask;
    }
}

template <int THREADS, int ITEMS, bool KV, int MATCH, int MIN_BLOCKS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
onesweep_pass_kernel(const uint32_t* __restrict__ keys_in, uint32_t* __restrict__ keys_out,
                     const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ vals_out,
                     uint32_t n, int pass, sort_control* ctl, uint32_t* lookback, uint32_t num_tiles)
{
    using smem_t = onesweep_smem<THREADS, ITEMS, KV>;
    constexpr int WARPS = smem_t::WARPS;
    constexpr int TILE = smem_t::TILE;
    static_assert(THREADS >= kRadix, "one thread per digit needed");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    smem_t& sm = *reinterpret_cast<smem_t*>(smem_raw);

    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int shift = pass * kRadixBits;

    if (tid == 0)
    {
        sm.tile = atomicAdd(&ctl->tickets[pass], 1u);
        mbar_init(&sm.bar_keys, 1);
        mbar_init(&sm.bar_vals, 1);
        mbar_fence_init();
    }
    // warp-private digit counters
#pragma unroll
    for (int i = lane; i < kRadix; i += 32) sm.warp_hist[warp][i] = 0;
    __syncthreads();

    const uint32_t tile = sm.tile;
    const uint64_t tile_base = (uint64_t) tile * TILE;
    const uint32_t valid = (n - tile_base) < (uint64_t) TILE ? (uint32_t) (n - tile_base) : (uint32_t) TILE;
    const bool full = valid == (uint32_t) TILE;

    if (full)
    {
        if (tid == 0)
        {
            mbar_arrive_expect_tx(&sm.bar_keys, TILE * 4);
            bulk_copy_g2s(sm.keys, keys_in + tile_base, TILE * 4, &sm.bar_keys);
            if (KV)
            {
                mbar_arrive_expect_tx(&sm.bar_vals, TILE * 4);
                bulk_copy_g2s(sm.vals, vals_in + tile_base, TILE * 4, &sm.bar_vals);
            }
        }
        mbar_wait(&sm.bar_keys, 0);
    }
    else
    {
        // ragged last tile: guarded loads, padding keys 0xFFFFFFFF sort behind every real key of the tile
        for (uint32_t i = tid; i < (uint32_t) TILE; i += THREADS)
        {
            sm.keys[i] = i < valid ? keys_in[tile_base + i] : 0xFFFFFFFFu;
            if (KV) sm.vals[i] = i < valid ? vals_in[tile_base + i] : 0u;
        }
        __syncthreads();
    }

    // warp-striped register tile: element (warp, j, lane)
    const uint32_t warp_off = warp * (ITEMS * 32) + lane;
    uint32_t key[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; j++) key[j] = sm.keys[warp_off + j * 32];

    // stable in-warp ranking: ballot match + warp-private running digit counters
    uint32_t rank[ITEMS];
    uint32_t* my_hist = sm.warp_hist[warp];
#pragma unroll
    for (int j = 0; j < ITEMS; j++)
    {
        const uint32_t d = (key[j] >> shift) & 0xFFu;
        const unsigned mask = match_digit<MATCH>(d);
        const unsigned leader = 31 - __clz(mask);
        uint32_t prior = 0;
        if (lane == leader)
        {
            prior = my_hist[d];
            my_hist[d] = prior + __popc(mask);
        }
        prior = __shfl_sync(kFullMask, prior, leader);
        rank[j] = prior + __popc(mask & lanemask_lt());
        __syncwarp();
    }
    __syncthreads();

    // per digit: tile count, publish aggregate, tile-local exclusive offsets
    uint32_t cnt = 0, inc = 0, real_cnt = 0;
    uint32_t* lb = lookback + ((size_t) pass * num_tiles + tile) * kRadix;
    if (tid < kRadix)
    {
#pragma unroll
        for (int w = 0; w < WARPS; w++) cnt += sm.warp_hist[w][tid];
        real_cnt = cnt - ((tid == kRadix - 1) ? (uint32_t) TILE - valid : 0u);
        st_relaxed_u32(&lb[tid], (tile == 0 ? kLbFlagInclusive : kLbFlagAggregate) | real_cnt);
        inc = cnt;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1)
        {
            const uint32_t t = __shfl_up_sync(kFullMask, inc, s);
            if (lane >= (unsigned) s) inc += t;
        }
        if (lane == 31) sm.scan_warp[warp] = inc;
    }
    __syncthreads();
    uint32_t tile_off = 0;
    if (tid < kRadix)
    {
        uint32_t wp = 0;
#pragma unroll
        for (int w = 0; w < kRadix / 32; w++)
            if (w < (int) warp) wp += sm.scan_warp[w];
        tile_off = wp + inc - cnt;
        uint32_t run = tile_off;
#pragma unroll
        for (int w = 0; w < WARPS; w++)
        {
            const uint32_t c = sm.warp_hist[w][tid];
            sm.warp_hist[w][tid] = run;
            run += c;
        }
    }
    __syncthreads();

    // in-tile destination of every item; fetch the staged values with the same striping
#pragma unroll
    for (int j = 0; j < ITEMS; j++) rank[j] += my_hist[(key[j] >> shift) & 0xFFu];
    uint32_t val[KV ? ITEMS : 1];
    if (KV)
    {
        if (full) mbar_wait(&sm.bar_vals, 0);
#pragma unroll
        for (int j = 0; j < ITEMS; j++) val[j] = sm.vals[warp_off + j * 32];
    }
    __syncthreads(); // every warp has consumed the staged inputs: the buffers become the regroup area

#pragma unroll
    for (int j = 0; j < ITEMS; j++)
    {
        sm.keys[rank[j]] = key[j];
        if (KV) sm.vals[rank[j]] = val[j];
    }

    // decoupled look-back, one thread per digit
    if (tid < kRadix)
    {
        uint32_t exclusive = 0;
        if (tile > 0)
        {
            const uint32_t* p = lb - kRadix + tid;
            for (int64_t t = (int64_t) tile - 1; t >= 0; t--, p -= kRadix)
            {
                uint32_t s;
                do { s = ld_relaxed_u32(p); } while ((s >> 30) == 0);
                exclusive += s & kLbValueMask;
                if ((s >> 30) == 2) break;
            }
            st_relaxed_u32(&lb[tid], kLbFlagInclusive | (exclusive + real_cnt));
        }
        sm.digit_base[tid] = ctl->hist[pass][tid] + exclusive - tile_off;
    }
    __syncthreads();

    // coalesced write-out: in-tile position p -> digit run -> global slot
#pragma unroll
    for (int j = 0; j < ITEMS; j++)
    {
        const uint32_t p = j * THREADS + tid;
        if (full || p < valid)
        {
            const uint32_t k = sm.keys[p];
            const uint32_t g = sm.digit_base[(k >> shift) & 0xFFu] + p;
            keys_out[g] = k;
            if (KV) vals_out[g] = sm.vals[p];
        }
    }
}

// ---- host side ------------------------------------------------------------------------------------------------
struct sort_variant
{
    const char* name;
    uint32_t tile;
    int (*launch)(cudaStream_t, const uint32_t*, uint32_t*, const uint32_t*, uint32_t*, uint32_t, int,
                  sort_control*, uint32_t*, uint32_t, bool);
};

template <int THREADS, int ITEMS, int MATCH, int MIN_BLOCKS>
int launch_onesweep(cudaStream_t s, const uint32_t* kin, uint32_t* kout, const uint32_t* vin, uint32_t* vout,
                    uint32_t n, int pass, sort_control* ctl, uint32_t* lookback, uint32_t tiles, bool kv)
{
    if (kv)
    {
        auto kern = onesweep_pass_kernel<THREADS, ITEMS, true, MATCH, MIN_BLOCKS>;
        constexpr size_t smem = sizeof(onesweep_smem<THREADS, ITEMS, true>);
        VRENB200_TRY(check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)));
        kern<<<tiles, THREADS, smem, s>>>(kin, kout, vin, vout, n, pass, ctl, lookback, tiles);
    }
    else
    {
        auto kern = onesweep_pass_kernel<THREADS, ITEMS, false, MATCH, MIN_BLOCKS>;
        constexpr size_t smem = sizeof(onesweep_smem<THREADS, ITEMS, false>);
        VRENB200_TRY(check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)));
        kern<<<tiles, THREADS, smem, s>>>(kin, kout, nullptr, nullptr, n, pass, ctl, lookback, tiles);
    }
    return check_launch();
}

#define VARIANT(T, I, M, B) { #T "x" #I "/" #M "/occ" #B, (T) * (I), launch_onesweep<T, I, M, B> }
const sort_variant g_variants[] = {
    VARIANT(384, 16, MATCH_BALLOT, 2),   // 0: default
    VARIANT(256, 16, MATCH_BALLOT, 3),
    VARIANT(512, 16, MATCH_BALLOT, 2),
    VARIANT(512, 12, MATCH_BALLOT, 2),
    VARIANT(384, 16, MATCH_HW, 2),
    VARIANT(256, 16, MATCH_HW, 3),
    VARIANT(512, 16, MATCH_HW, 2),
    VARIANT(256, 24, MATCH_BALLOT, 2),
};
constexpr int kNumVariants = sizeof(g_variants) / sizeof(g_variants[0]);
// the scratch layout must not depend on the variant: size the look-back for the smallest tile
constexpr uint32_t kMinTile = 256 * 16;

int g_variant = 0;

size_t lookback_words(uint32_t n)
{
    const size_t tiles = ((size_t) n + kMinTile - 1) / kMinTile;
    return (size_t) kPasses * (tiles > 0 ? tiles : 1) * kRadix;
}

size_t control_bytes(uint32_t n)
{
    return align_up(sizeof(sort_control) + lookback_words(n) * sizeof(uint32_t), 256);
}

} // namespace
} // namespace vrenb200

// per-kernel device timing for bench.py's roofline: events recorded on the launching stream between launches
struct vrenb200_sort_profile
{
    cudaEvent_t ev[vrenb200::kPasses + 3]; // start | after hist | after scan | after pass 0..3
};

namespace vrenb200 {
namespace {

// control + look-back live in `ctl_mem`; alt buffers given explicitly
int radix_sort_impl(cudaStream_t s, uint32_t* keys, uint32_t* vals, uint32_t n, uint32_t* alt_keys, uint32_t* alt_vals,
                    void* ctl_mem, vrenb200_sort_profile* prof = nullptr)
{
    if (n == 0) return VRENB200_OK;
    if (n >= (1u << 30)) return VRENB200_ELIMIT;
    if ((reinterpret_cast<uintptr_t>(keys) | reinterpret_cast<uintptr_t>(alt_keys) |
         reinterpret_cast<uintptr_t>(vals) | reinterpret_cast<uintptr_t>(alt_vals) |
         reinterpret_cast<uintptr_t>(ctl_mem)) & 15)
        return VRENB200_EALIGN;
    const bool kv = vals != nullptr;
    const sort_variant& var = g_variants[g_variant];
    const uint32_t tiles = (uint32_t) (((size_t) n + var.tile - 1) / var.tile);
    sort_control* ctl = static_cast<sort_control*>(ctl_mem);
    uint32_t* lookback = reinterpret_cast<uint32_t*>(ctl + 1);
    // only the part of the look-back this variant touches needs clearing
    const size_t clear = sizeof(sort_control) + (size_t) kPasses * tiles * kRadix * sizeof(uint32_t);
    VRENB200_TRY(check_cuda(cudaMemsetAsync(ctl_mem, 0, clear, s)));
    if (prof) cudaEventRecord(prof->ev[0], s);
    radix_histogram_kernel<<<kNumSMs * 2, kHistThreads, 0, s>>>(keys, n, ctl);
    VRENB200_TRY(check_launch());
    if (prof) cudaEventRecord(prof->ev[1], s);
    radix_scan_histograms_kernel<<<kPasses, kRadix, 0, s>>>(ctl);
    VRENB200_TRY(check_launch());
    if (prof) cudaEventRecord(prof->ev[2], s);
    for (int pass = 0; pass < kPasses; pass++)
    {
        const bool even = (pass & 1) == 0;
        VRENB200_TRY(var.launch(s, even ? keys : alt_keys, even ? alt_keys : keys, even ? vals : alt_vals,
                                even ? alt_vals : vals, n, pass, ctl, lookback, tiles, kv));
        if (prof) cudaEventRecord(prof->ev[3 + pass], s);
    }
    return VRENB200_OK;
}

} // namespace
} // namespace vrenb200

using namespace vrenb200;

// tuning hook (not part of the reference surface): select the kernel configuration used by subsequent calls
extern "C" int vrenb200_radix_sort_set_variant(int v)
{
    if (v < 0 || v >= kNumVariants) return VRENB200_EINVAL_ARG;
    g_variant = v;
    return VRENB200_OK;
}
extern "C" int vrenb200_radix_sort_num_variants(void) { return kNumVariants; }
extern "C" const char* vrenb200_radix_sort_variant_name(int v)
{
    return (v < 0 || v >= kNumVariants) ? "" : g_variants[v].name;
}

extern "C" size_t vrenb200_radix_sort_scratch_bytes(uint32_t n, int with_values)
{
    const size_t alt = align_up((size_t) n * 4, 256);
    return alt * (with_values ? 2 : 1) + control_bytes(n);
}

extern "C" int vrenb200_radix_sort_keys(vrenb200_stream_t stream, uint32_t* keys, uint32_t n,
                                        void* scratch, size_t scratch_bytes)
{
    if (n == 0) return VRENB200_OK;
    if (keys == nullptr) return VRENB200_EINVAL_ARG;
    if (scratch == nullptr || scratch_bytes < vrenb200_radix_sort_scratch_bytes(n, 0)) return VRENB200_ESCRATCH;
    char* p = static_cast<char*>(scratch);
    const size_t alt = align_up((size_t) n * 4, 256);
    return radix_sort_impl(as_stream(stream), keys, nullptr, n, reinterpret_cast<uint32_t*>(p), nullptr, p + alt);
}

extern "C" int vrenb200_radix_sort_pairs(vrenb200_stream_t stream, uint32_t* keys, uint32_t* values, uint32_t n,
                                         void* scratch, size_t scratch_bytes)
{
    if (n == 0) return VRENB200_OK;
    if (keys == nullptr || values == nullptr) return VRENB200_EINVAL_ARG;
    if (scratch == nullptr || scratch_bytes < vrenb200_radix_sort_scratch_bytes(n, 1)) return VRENB200_ESCRATCH;
    char* p = static_cast<char*>(scratch);
    const size_t alt = align_up((size_t) n * 4, 256);
    return radix_sort_impl(as_stream(stream), keys, values, n, reinterpret_cast<uint32_t*>(p),
                           reinterpret_cast<uint32_t*>(p + alt), p + 2 * alt);
}

extern "C" vrenb200_sort_profile* vrenb200_sort_profile_create(void)
{
    vrenb200_sort_profile* p = new vrenb200_sort_profile();
    for (auto& e : p->ev)
        if (cudaEventCreate(&e) != cudaSuccess) { delete p; return nullptr; }
    return p;
}
extern "C" void vrenb200_sort_profile_destroy(vrenb200_sort_profile* p)
{
    if (!p) return;
    for (auto& e : p->ev) cudaEventDestroy(e);
    delete p;
}
// ms_out[6] = {histogram, histogram scan, pass0, pass1, pass2, pass3}; call after the stream has been synchronised
extern "C" int vrenb200_sort_profile_read(vrenb200_sort_profile* p, float* ms_out)
{
    if (!p || !ms_out) return VRENB200_EINVAL_ARG;
    for (int i = 0; i < kPasses + 2; i++)
        VRENB200_TRY(check_cuda(cudaEventElapsedTime(&ms_out[i], p->ev[i], p->ev[i + 1])));
    return VRENB200_OK;
}
extern "C" int vrenb200_radix_sort_pairs_profiled(vrenb200_stream_t stream, uint32_t* keys, uint32_t* values, uint32_t n,
                                                  void* scratch, size_t scratch_bytes, vrenb200_sort_profile* prof)
{
    if (n == 0) return VRENB200_OK;
    if (keys == nullptr) return VRENB200_EINVAL_ARG;
    const int kv = values != nullptr;
    if (scratch == nullptr || scratch_bytes < vrenb200_radix_sort_scratch_bytes(n, kv)) return VRENB200_ESCRATCH;
    char* p = static_cast<char*>(scratch);
    const size_t alt = align_up((size_t) n * 4, 256);
    return radix_sort_impl(as_stream(stream), keys, values, n, reinterpret_cast<uint32_t*>(p),
                           kv ? reinterpret_cast<uint32_t*>(p + alt) : nullptr, p + (kv ? 2 : 1) * alt, prof);
}

extern "C" size_t vrenb200_radix_sort_scratch_buffer_1_bytes(uint32_t n) { return control_bytes(n); }
extern "C" size_t vrenb200_radix_sort_scratch_buffer_2_bytes(uint32_t n) { return align_up((size_t) n * 4, 256); }

extern "C" int vrenb200_radix_sort_compat(vrenb200_stream_t stream, uint32_t* keys, uint32_t n,
                                          void* scratch_1, size_t scratch_1_bytes,
                                          void* scratch_2, size_t scratch_2_bytes)
{
    // radix_sort.cpp:158-161: "Length must be higher than 1024 and a power of 2"
    if (!(n >= 1024 && (n & (n - 1)) == 0)) return VRENB200_EINVAL_LENGTH;
    if (keys == nullptr) return VRENB200_EINVAL_ARG;
    if (scratch_1 == nullptr || scratch_1_bytes < control_bytes(n)) return VRENB200_ESCRATCH;
    if (scratch_2 == nullptr || scratch_2_bytes < (size_t) n * 4) return VRENB200_ESCRATCH;
    return radix_sort_impl(as_stream(stream), keys, nullptr, n, static_cast<uint32_t*>(scratch_2), nullptr, scratch_1);
}

extern "C" size_t vrenb200_radix_sort_host_work_bytes(uint32_t n, int with_values)
{
    const size_t buf = align_up((size_t) n * 4, 256);
    return buf * (with_values ? 2 : 1) + vrenb200_radix_sort_scratch_bytes(n, with_values);
}

extern "C" int vrenb200_radix_sort_pairs_host(vrenb200_stream_t stream, uint32_t* keys_host, uint32_t* values_host,
                                              uint32_t n, void* dev_work, size_t dev_work_bytes)
{
    if (n == 0) return VRENB200_OK;
    if (keys_host == nullptr || dev_work == nullptr) return VRENB200_EINVAL_ARG;
    const int kv = values_host != nullptr;
    if (dev_work_bytes < vrenb200_radix_sort_host_work_bytes(n, kv)) return VRENB200_ESCRATCH;
    cudaStream_t s = as_stream(stream);
    char* p = static_cast<char*>(dev_work);
    const size_t buf = align_up((size_t) n * 4, 256);
    uint32_t* dk = reinterpret_cast<uint32_t*>(p);
    uint32_t* dv = kv ? reinterpret_cast<uint32_t*>(p + buf) : nullptr;
    char* scratch = p + buf * (kv ? 2 : 1);
    const size_t scratch_bytes = dev_work_bytes - buf * (kv ? 2 : 1);
    VRENB200_TRY(check_cuda(cudaMemcpyAsync(dk, keys_host, (size_t) n * 4, cudaMemcpyHostToDevice, s)));
    if (kv) VRENB200_TRY(check_cuda(cudaMemcpyAsync(dv, values_host, (size_t) n * 4, cudaMemcpyHostToDevice, s)));
    int st = kv ? vrenb200_radix_sort_pairs(stream, dk, dv, n, scratch, scratch_bytes)
                : vrenb200_radix_sort_keys(stream, dk, n, scratch, scratch_bytes);
    if (st != VRENB200_OK) return st;
    VRENB200_TRY(check_cuda(cudaMemcpyAsync(keys_host, dk, (size_t) n * 4, cudaMemcpyDeviceToHost, s)));
    if (kv) VRENB200_TRY(check_cuda(cudaMemcpyAsync(values_host, dv, (size_t) n * 4, cudaMemcpyDeviceToHost, s)));
    return check_cuda(cudaStreamSynchronize(s));
}
