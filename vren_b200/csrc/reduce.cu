// a1 — vren::reduce<T,op> (reference: vren/vren/primitives/reduce.{hpp,cpp}, shaders/reduce.comp:50-87).
//
// Semantics kept: Blelloch up-sweep tree over P = next_pow2(n) slots, slots >= n read as the identity,
// slot (j+1)*2^l-1 ends holding op over its aligned 2^l block, result at out[P-1]; `blocks` independent rows
// (in + y*n, out + y*P).  fp32 adds follow the reference's fixed pairwise order (b = op(b, a)), so results are
// bit-identical to the tree the reference (and its test oracle run_cpu_reduce) produces.
//
// B200 design: one pass over the input with 128-bit loads; a CTA owns a tile of 256 threads x 4 vectors and
// resolves 10 (vec4) / 12 (scalar) tree levels in registers + warp shuffles + a 32-slot shared-memory stage.
// The remaining log2(P/tile) levels are a second, tiny launch over the tile tops.  In FINAL mode nothing but
// the tile tops is written (4 B/elt of HBM traffic instead of the reference's 8 B/elt + one dispatch per
// 10 levels).
#include "common.cuh"

namespace vrenb200 {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kVecPerThread = 4;

template <typename T> struct elem_traits;
template <> struct elem_traits<uint32_t> { static constexpr int epv = 4; };
template <> struct elem_traits<float>    { static constexpr int epv = 4; };
template <> struct elem_traits<float4>   { static constexpr int epv = 1; };

template <typename T> __host__ __device__ constexpr uint32_t tile_elems()
{
    return kThreads * kVecPerThread * elem_traits<T>::epv;
}

// ---- operations: identity values from VRen.cmake:70-75, operand order from reduce.comp:77 -------------------
template <typename T, int OP> struct op_t;

template <> struct op_t<uint32_t, VRENB200_ADD> {
    static constexpr bool order_free = true;
    __device__ static uint32_t identity() { return 0u; }
    __device__ static uint32_t apply(uint32_t b, uint32_t a) { return b + a; }
};
template <> struct op_t<uint32_t, VRENB200_MIN> {
    static constexpr bool order_free = true;
    __device__ static uint32_t identity() { return ~0u; }
    __device__ static uint32_t apply(uint32_t b, uint32_t a) { return a < b ? a : b; }
};
template <> struct op_t<uint32_t, VRENB200_MAX> {
    static constexpr bool order_free = true;
    __device__ static uint32_t identity() { return 0u; }
    __device__ static uint32_t apply(uint32_t b, uint32_t a) { return b < a ? a : b; }
};
template <> struct op_t<float, VRENB200_ADD> {
    static constexpr bool order_free = false; // fp32 add is not associative: keep the reference's tree
    __device__ static float identity() { return 0.0f; }
    __device__ static float apply(float b, float a) { return __fadd_rn(b, a); }
};
template <> struct op_t<float, VRENB200_MIN> {
    static constexpr bool order_free = true;
    __device__ static float identity() { return 1e35f; }
    __device__ static float apply(float b, float a) { return a < b ? a : b; } // GLSL min(b, a)
};
template <> struct op_t<float, VRENB200_MAX> {
    static constexpr bool order_free = true;
    __device__ static float identity() { return -1e35f; }
    __device__ static float apply(float b, float a) { return b < a ? a : b; } // GLSL max(b, a)
};
template <int OP> struct op_t<float4, OP> {
    using s = op_t<float, OP>;
    static constexpr bool order_free = s::order_free;
    __device__ static float4 identity() { float i = s::identity(); return make_float4(i, i, i, i); }
    __device__ static float4 apply(float4 b, float4 a)
    {
        return make_float4(s::apply(b.x, a.x), s::apply(b.y, a.y), s::apply(b.z, a.z), s::apply(b.w, a.w));
    }
};

template <typename T> __device__ __forceinline__ T shfl_up_t(T v, int d) { return __shfl_up_sync(kFullMask, v, d); }
template <> __device__ __forceinline__ float4 shfl_up_t<float4>(float4 v, int d)
{
    return make_float4(__shfl_up_sync(kFullMask, v.x, d), __shfl_up_sync(kFullMask, v.y, d),
                       __shfl_up_sync(kFullMask, v.z, d), __shfl_up_sync(kFullMask, v.w, d));
}
template <typename T> __device__ __forceinline__ T shfl_xor_t(T v, int d) { return __shfl_xor_sync(kFullMask, v, d); }
template <> __device__ __forceinline__ float4 shfl_xor_t<float4>(float4 v, int d)
{
    return make_float4(__shfl_xor_sync(kFullMask, v.x, d), __shfl_xor_sync(kFullMask, v.y, d),
                       __shfl_xor_sync(kFullMask, v.z, d), __shfl_xor_sync(kFullMask, v.w, d));
}

// one 16-byte vector = EPV consecutive logical elements
template <typename T> struct vec_t { T e[elem_traits<T>::epv]; };

template <typename T, typename Op>
__device__ __forceinline__ vec_t<T> load_vec(const T* row, uint64_t idx, uint32_t n, bool vec_ok)
{
    constexpr int EPV = elem_traits<T>::epv;
    vec_t<T> r;
    if (vec_ok && idx + EPV <= n)
    {
        const uint4 raw = ldg_stream_u4(reinterpret_cast<const uint4*>(row + idx));
        r = *reinterpret_cast<const vec_t<T>*>(&raw);
    }
    else
    {
#pragma unroll
        for (int k = 0; k < EPV; k++)
            r.e[k] = (idx + k < n) ? row[idx + k] : Op::identity();
    }
    return r;
}

template <typename T>
__device__ __forceinline__ void store_vec(T* row, uint64_t idx, uint32_t limit, bool vec_ok, const vec_t<T>& v)
{
    constexpr int EPV = elem_traits<T>::epv;
    if (vec_ok && idx + EPV <= limit)
    {
        *reinterpret_cast<uint4*>(row + idx) = *reinterpret_cast<const uint4*>(&v);
    }
    else
    {
#pragma unroll
        for (int k = 0; k < EPV; k++)
            if (idx + k < limit) row[idx + k] = v.e[k];
    }
}

// ---- level-0 kernel: tile of 256 x 4 vectors; exact tree order with per-slot values ---------------------
// MODE TREE : every slot of the tile's sub-tree is written to out (in-tile tree), tile top included.
// MODE FINAL: only the tile's result slot is written, to dst[tile] (dst = partials or the final slot).
template <typename T, int OP, int MODE>
__global__ void __launch_bounds__(kThreads)
reduce_tile_exact_kernel(const T* in, T* out, uint32_t n, uint32_t P,
                         uint64_t in_row_stride, uint64_t out_row_stride, uint32_t final_slot_stride)
{
    using Op = op_t<T, OP>;
    constexpr int EPV = elem_traits<T>::epv;
    constexpr uint32_t TILE = tile_elems<T>();
    constexpr uint32_t SEG = kThreads * EPV; // logical elements covered by one vector index v

    __shared__ T s_top[kVecPerThread][kWarps];

    const T* row_in = in + blockIdx.y * in_row_stride;
    T* row_out = out + blockIdx.y * out_row_stride;
    const uint64_t tile_base = (uint64_t) blockIdx.x * TILE;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool in_vec_ok = (reinterpret_cast<uintptr_t>(row_in) & 15) == 0;

    vec_t<T> x[kVecPerThread];
#pragma unroll
    for (int v = 0; v < kVecPerThread; v++)
        x[v] = load_vec<T, Op>(row_in, tile_base + (uint64_t) v * SEG + threadIdx.x * EPV, n, in_vec_ok);

#pragma unroll
    for (int v = 0; v < kVecPerThread; v++)
    {
        // levels inside the 16-byte vector
        if constexpr (EPV == 4)
        {
            x[v].e[1] = Op::apply(x[v].e[1], x[v].e[0]);
            x[v].e[3] = Op::apply(x[v].e[3], x[v].e[2]);
            x[v].e[3] = Op::apply(x[v].e[3], x[v].e[1]);
        }
        // 5 levels across the warp: lane whose low (l+1) bits are all ones absorbs lane - 2^l
        T top = x[v].e[EPV - 1];
#pragma unroll
        for (int s = 1; s < 32; s <<= 1)
        {
            T other = shfl_up_t(top, s);
            if ((lane & (2 * s - 1)) == (unsigned) (2 * s - 1)) top = Op::apply(top, other);
        }
        x[v].e[EPV - 1] = top;
        if (lane == 31) s_top[v][warp] = top;
    }
    __syncthreads();
    // 3 levels across the 8 warps (one thread per v), then 2 levels across v
    if (threadIdx.x < kVecPerThread)
    {
        T* t = s_top[threadIdx.x];
        t[1] = Op::apply(t[1], t[0]); t[3] = Op::apply(t[3], t[2]);
        t[5] = Op::apply(t[5], t[4]); t[7] = Op::apply(t[7], t[6]);
        t[3] = Op::apply(t[3], t[1]); t[7] = Op::apply(t[7], t[5]);
        t[7] = Op::apply(t[7], t[3]);
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        s_top[1][7] = Op::apply(s_top[1][7], s_top[0][7]);
        s_top[3][7] = Op::apply(s_top[3][7], s_top[2][7]);
        s_top[3][7] = Op::apply(s_top[3][7], s_top[1][7]);
    }
    __syncthreads();

    if constexpr (MODE == VRENB200_REDUCE_TREE)
    {
        const bool out_vec_ok = (reinterpret_cast<uintptr_t>(row_out) & 15) == 0;
#pragma unroll
        for (int v = 0; v < kVecPerThread; v++)
        {
            if (lane == 31) x[v].e[EPV - 1] = s_top[v][warp];
            store_vec<T>(row_out, tile_base + (uint64_t) v * SEG + threadIdx.x * EPV, P, out_vec_ok, x[v]);
        }
    }
    else
    {
        // result slot of this tile: TILE-1, or P-1 when the whole problem is smaller than one tile
        const uint32_t rs = (P < TILE ? P : TILE) - 1;
        const uint32_t rv = rs / SEG, rt = (rs % SEG) / EPV, rk = rs % EPV;
        if (threadIdx.x == rt)
        {
            T r = x[0].e[0];
#pragma unroll
            for (int v = 0; v < kVecPerThread; v++)
#pragma unroll
                for (int k = 0; k < EPV; k++)
                    if (v == (int) rv && k == (int) rk) r = x[v].e[k];
            if ((rt & 31) == 31 && rk == EPV - 1) r = s_top[rv][rt >> 5];
            row_out[(uint64_t) blockIdx.x * final_slot_stride + (final_slot_stride - 1)] = r;
        }
    }
}

// ---- level-0 kernel, order-free ops in FINAL mode: plain accumulate, butterfly, one value per tile -------
template <typename T, int OP>
__global__ void __launch_bounds__(kThreads)
reduce_tile_fast_kernel(const T* __restrict__ in, T* out, uint32_t n,
                        uint64_t in_row_stride, uint64_t out_row_stride, uint32_t final_slot_stride)
{
    using Op = op_t<T, OP>;
    constexpr int EPV = elem_traits<T>::epv;
    constexpr uint32_t TILE = tile_elems<T>();
    constexpr uint32_t SEG = kThreads * EPV;

    __shared__ T s_top[kWarps];

    const T* row_in = in + blockIdx.y * in_row_stride;
    T* row_out = out + blockIdx.y * out_row_stride;
    const uint64_t tile_base = (uint64_t) blockIdx.x * TILE;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool in_vec_ok = (reinterpret_cast<uintptr_t>(row_in) & 15) == 0;

    vec_t<T> x[kVecPerThread];
#pragma unroll
    for (int v = 0; v < kVecPerThread; v++)
        x[v] = load_vec<T, Op>(row_in, tile_base + (uint64_t) v * SEG + threadIdx.x * EPV, n, in_vec_ok);

    T acc = Op::identity();
#pragma unroll
    for (int v = 0; v < kVecPerThread; v++)
#pragma unroll
        for (int k = 0; k < EPV; k++)
            acc = Op::apply(acc, x[v].e[k]);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc = Op::apply(acc, shfl_xor_t(acc, s));
    if (lane == 0) s_top[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0)
    {
#pragma unroll
        for (int w = 1; w < kWarps; w++) acc = Op::apply(acc, s_top[w]);
        row_out[(uint64_t) blockIdx.x * final_slot_stride + (final_slot_stride - 1)] = acc;
    }
}

// ---- upper levels: logical element i lives at buf[(i+1)*stride-1]; 1024 per CTA, 10 levels per launch -----
// (same shape as the reference's reduce.comp with base_level>0, reduce.comp:55-86)
template <typename T, int OP, int MODE>
__global__ void __launch_bounds__(1024)
reduce_strided_kernel(T* buf, uint32_t count, uint64_t stride, uint64_t row_stride)
{
    using Op = op_t<T, OP>;
    __shared__ T s_data[1024];
    T* row = buf + blockIdx.y * row_stride;
    const uint32_t gi = blockIdx.x * 1024u + threadIdx.x;
    const uint64_t idx = ((uint64_t) gi + 1) * stride - 1;
    s_data[threadIdx.x] = gi < count ? row[idx] : Op::identity();
    __syncthreads();
    for (uint32_t level = 0; level < 10; level++)
    {
        const uint32_t m = (1u << (level + 1)) - 1;
        if ((threadIdx.x & m) == m)
            s_data[threadIdx.x] = Op::apply(s_data[threadIdx.x], s_data[threadIdx.x - (1u << level)]);
        __syncthreads();
    }
    if (gi < count)
    {
        if (MODE == VRENB200_REDUCE_TREE)
            row[idx] = s_data[threadIdx.x];
        else
        {
            // FINAL: only block tops matter; the last valid slot of a short block is its top as well
            const uint32_t top = (count < 1024u ? count : 1024u) - 1;
            if (threadIdx.x == top) row[idx] = s_data[threadIdx.x];
        }
    }
}

template <typename T>
__global__ void gather_row_tops_kernel(const T* partials, uint32_t tiles, T* out, uint32_t P, uint32_t blocks)
{
    const uint32_t y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y < blocks) out[(uint64_t) y * P + (P - 1)] = partials[(uint64_t) y * tiles + (tiles - 1)];
}

template <typename T, int OP>
int launch_reduce(cudaStream_t stream, int mode, const T* in, uint32_t n, T* out, uint32_t blocks,
                  void* scratch, size_t scratch_bytes)
{
    using Op = op_t<T, OP>;
    constexpr uint32_t TILE = tile_elems<T>();
    const uint32_t P = next_pow2_u32(n);
    const uint32_t tiles = P <= TILE ? 1u : P / TILE;
    const dim3 grid0(tiles, blocks);

    if (mode == VRENB200_REDUCE_TREE)
    {
        reduce_tile_exact_kernel<T, OP, VRENB200_REDUCE_TREE><<<grid0, kThreads, 0, stream>>>(in, out, n, P, n, P, 0);
        VRENB200_TRY(check_launch());
        uint64_t stride = TILE;
        uint32_t count = tiles;
        while (count > 1)
        {
            const dim3 grid((count + 1023) / 1024, blocks);
            reduce_strided_kernel<T, OP, VRENB200_REDUCE_TREE><<<grid, 1024, 0, stream>>>(out, count, stride, P);
            VRENB200_TRY(check_launch());
            stride *= 1024;
            count = (count + 1023) / 1024;
        }
        return VRENB200_OK;
    }

    // FINAL: tile results go to the slot they would occupy in the tree when a single tile covers the row,
    // otherwise into a compact partials array in scratch that is then reduced by the strided kernel
    if (tiles == 1)
    {
        // final_slot_stride = P places the single result at out[y*P + P-1]
        if (Op::order_free)
            reduce_tile_fast_kernel<T, OP><<<grid0, kThreads, 0, stream>>>(in, out, n, n, P, P);
        else
            reduce_tile_exact_kernel<T, OP, VRENB200_REDUCE_FINAL><<<grid0, kThreads, 0, stream>>>(in, out, n, P, n, P, P);
        return check_launch();
    }
    const size_t need = (size_t) blocks * tiles * sizeof(T);
    if (scratch == nullptr || scratch_bytes < need) return VRENB200_ESCRATCH;
    T* partials = static_cast<T*>(scratch);
    if (Op::order_free)
        reduce_tile_fast_kernel<T, OP><<<grid0, kThreads, 0, stream>>>(in, partials, n, n, tiles, 1);
    else
        reduce_tile_exact_kernel<T, OP, VRENB200_REDUCE_FINAL><<<grid0, kThreads, 0, stream>>>(in, partials, n, P, n, tiles, 1);
    VRENB200_TRY(check_launch());
    uint64_t stride = 1;
    uint32_t count = tiles;
    while (count > 1)
    {
        const dim3 grid((count + 1023) / 1024, blocks);
        reduce_strided_kernel<T, OP, VRENB200_REDUCE_FINAL><<<grid, 1024, 0, stream>>>(partials, count, stride, tiles);
        VRENB200_TRY(check_launch());
        stride *= 1024;
        count = (count + 1023) / 1024;
    }
    // partials[y*tiles + tiles-1] -> out[y*P + P-1]
    gather_row_tops_kernel<T><<<(blocks + 255) / 256, 256, 0, stream>>>(partials, tiles, out, P, blocks);
    return check_launch();
}

template <typename T>
int dispatch_op(cudaStream_t stream, int op, int mode, const void* in, uint32_t n, void* out, uint32_t blocks,
                void* scratch, size_t scratch_bytes)
{
    const T* i = static_cast<const T*>(in);
    T* o = static_cast<T*>(out);
    switch (op)
    {
    case VRENB200_ADD: return launch_reduce<T, VRENB200_ADD>(stream, mode, i, n, o, blocks, scratch, scratch_bytes);
    case VRENB200_MIN: return launch_reduce<T, VRENB200_MIN>(stream, mode, i, n, o, blocks, scratch, scratch_bytes);
    case VRENB200_MAX: return launch_reduce<T, VRENB200_MAX>(stream, mode, i, n, o, blocks, scratch, scratch_bytes);
    default: return VRENB200_EINVAL_ARG;
    }
}

size_t dtype_size(int dtype)
{
    return dtype == VRENB200_VEC4 ? 16 : 4;
}

uint32_t dtype_tile(int dtype)
{
    return dtype == VRENB200_VEC4 ? tile_elems<float4>() : tile_elems<uint32_t>();
}

} // namespace

} // namespace vrenb200

using namespace vrenb200;

extern "C" uint32_t vrenb200_calc_reduce_output_buffer_length(uint32_t count)
{
    return next_pow2_u32(count);
}

extern "C" size_t vrenb200_reduce_scratch_bytes(int dtype, int mode, uint32_t n, uint32_t blocks)
{
    if (mode == VRENB200_REDUCE_TREE) return 0;
    const uint32_t P = next_pow2_u32(n);
    const uint32_t tile = dtype_tile(dtype);
    const uint32_t tiles = P <= tile ? 1u : P / tile;
    return tiles == 1 ? 0 : align_up((size_t) blocks * tiles * dtype_size(dtype), 256);
}

extern "C" int vrenb200_reduce(vrenb200_stream_t stream, int dtype, int op, int mode,
                               const void* in, uint32_t n, void* out, uint32_t blocks,
                               void* scratch, size_t scratch_bytes)
{
    if (in == nullptr || out == nullptr) return VRENB200_EINVAL_ARG;
    if (n == 0 || blocks == 0 || blocks > 65535) return VRENB200_EINVAL_LENGTH;
    if (n > (1u << 31)) return VRENB200_ELIMIT; // next_pow2 must fit 32 bits like the reference
    if (mode != VRENB200_REDUCE_TREE && mode != VRENB200_REDUCE_FINAL) return VRENB200_EINVAL_ARG;
    const size_t esz = dtype_size(dtype);
    if ((reinterpret_cast<uintptr_t>(in) % esz) != 0 || (reinterpret_cast<uintptr_t>(out) % esz) != 0)
        return VRENB200_EALIGN;
    cudaStream_t s = as_stream(stream);
    switch (dtype)
    {
    case VRENB200_U32:  return dispatch_op<uint32_t>(s, op, mode, in, n, out, blocks, scratch, scratch_bytes);
    case VRENB200_F32:  return dispatch_op<float>(s, op, mode, in, n, out, blocks, scratch, scratch_bytes);
    case VRENB200_VEC4: return dispatch_op<float4>(s, op, mode, in, n, out, blocks, scratch, scratch_bytes);
    default: return VRENB200_EINVAL_ARG;
    }
}
