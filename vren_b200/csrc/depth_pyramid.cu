// n2 — depth-buffer pyramid: vren::depth_buffer_pyramid / vren::depth_buffer_reductor::copy_and_reduce
// (reference: vren/vren/pipeline/depth_buffer_pyramid.{hpp,cpp}:13-305, shaders/depth_buffer_copy.comp:8-18,
// shaders/depth_buffer_reduce.comp:10-35).
//
// Semantics kept: level_count = floor(log2(max(W,H))) + 1; level l is max(W>>l,1) x max(H>>l,1); level 0 is a copy
// of the depth buffer; level l+1(x,y) = max over the existing texels of the 2x2 block (2x..2x+1, 2y..2y+1) of level
// l, accumulated from 0.0 with GLSL max (depth_buffer_reduce.comp:20-31) — texels of an odd last row/column that no
// parent covers are dropped, exactly like the reference's floor-sized mips.
// Storage: the Vulkan mip chain becomes one flat float buffer, levels back to back (offset = sum of earlier sizes).
//
// Reference: one copy dispatch + one reduce dispatch (and barrier) per level — 12 launches at 4K.
// Here: ONE kernel does level 0..6.  A CTA stages a 64x64 base tile with a 2-D TMA tensor copy (out-of-image texels
// arrive as 0.0, the neutral element of this max), every thread owns a 4x4 block: levels 1 and 2 stay in registers,
// levels 3..6 go through a 256-float shared stage; a second one-CTA kernel finishes the few remaining levels.
// HBM traffic: 4 B/px read + 4/3 x 4 B/px written.
#include <cuda.h>

#include <cstring>

#include "common.cuh"

namespace vrenb200 {
namespace {

__host__ __device__ inline uint32_t level_dim(uint32_t base, uint32_t level)
{
    const uint32_t d = level >= 32 ? 0u : base >> level;
    return d > 0 ? d : 1u;
}

inline uint32_t level_count_of(uint32_t w, uint32_t h)
{
    uint32_t m = w > h ? w : h, l = 0;
    while (m >>= 1) l++;
    return l + 1; // glm::log2(glm::max(width, height)) + 1, depth_buffer_pyramid.cpp:18
}

// GLSL max(x, y) = x < y ? y : x
__device__ __forceinline__ float gl_max(float x, float y) { return x < y ? y : x; }

struct pyramid_params
{
    uint32_t width, height, level_count;
    unsigned long long level_offset[16]; // element offset of each level inside the pyramid buffer (first 7 used here)
    int use_tma;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256)
depth_pyramid_tile_kernel(const __grid_constant__ CUtensorMap depth_map, const float* __restrict__ depth, float* __restrict__ pyramid,
                          pyramid_params prm)
{
    __shared__ alignas(128) float s_tile[64 * 64];
    __shared__ float s_l2[16 * 16];
    __shared__ float s_l3[8 * 8 * 4];
    __shared__ alignas(8) uint64_t s_bar;

    const unsigned tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const uint32_t x0 = blockIdx.x * 64, y0 = blockIdx.y * 64;
    const uint32_t W = prm.width, H = prm.height;

    if (prm.use_tma)
    {
        if (tid == 0)
        {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&s_bar)), "r"(64 * 64 * 4) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(s_tile)), "l"(&depth_map), "r"((int) x0), "r"((int) y0), "r"(smem_u32(&s_bar)) : "memory");
        }
        __syncthreads();
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "W_%=:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
            "@p bra D_%=;\n"
            "bra W_%=;\n"
            "D_%=:\n"
            "}\n" ::"r"(smem_u32(&s_bar)) : "memory");
    }
    else
    {
        for (uint32_t i = tid; i < 64 * 64; i += 256)
        {
            const uint32_t x = x0 + (i & 63), y = y0 + (i >> 6);
            s_tile[i] = (x < W && y < H) ? depth[(size_t) y * W + x] : 0.0f;
        }
        __syncthreads();
    }

    // level 0 (depth_buffer_copy.comp) + levels 1, 2 in registers: thread (tx, ty) owns base texels 4x4 at (4tx, 4ty)
    float v[4][4];
#pragma unroll
    for (int r = 0; r < 4; r++)
    {
        const float4 row = *reinterpret_cast<const float4*>(&s_tile[(4 * ty + r) * 64 + 4 * tx]);
        v[r][0] = row.x; v[r][1] = row.y; v[r][2] = row.z; v[r][3] = row.w;
        const uint32_t y = y0 + 4 * ty + r, x = x0 + 4 * tx;
        if (y < H)
        {
            float* dst = pyramid + (size_t) y * W + x;
            if (x + 4 <= W && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0))
                *reinterpret_cast<float4*>(dst) = row;
            else
            {
#pragma unroll
                for (int c = 0; c < 4; c++)
                    if (x + c < W) dst[c] = v[r][c];
            }
        }
    }
    // a texel that does not exist at its level contributes the neutral 0.0 to its parent (it has no parent texel in
    // the reference: floor-sized mips)
    const uint32_t w1 = level_dim(W, 1), h1 = level_dim(H, 1), w2 = level_dim(W, 2), h2 = level_dim(H, 2);
    float l1[2][2];
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
        for (int c = 0; c < 2; c++)
        {
            // accumulation order of depth_buffer_reduce.comp:21-31: x outer, y inner
            float m = 0.0f;
            m = gl_max(m, v[2 * r][2 * c]);
            m = gl_max(m, v[2 * r + 1][2 * c]);
            m = gl_max(m, v[2 * r][2 * c + 1]);
            m = gl_max(m, v[2 * r + 1][2 * c + 1]);
            const uint32_t x = (x0 >> 1) + 2 * tx + c, y = (y0 >> 1) + 2 * ty + r;
            const bool exists = prm.level_count > 1 && x < w1 && y < h1;
            l1[r][c] = exists ? m : 0.0f;
            if (exists) pyramid[prm.level_offset[1] + (size_t) y * w1 + x] = m;
        }
    float l2 = 0.0f;
    {
        float m = 0.0f;
        m = gl_max(m, l1[0][0]); m = gl_max(m, l1[1][0]); m = gl_max(m, l1[0][1]); m = gl_max(m, l1[1][1]);
        const uint32_t x = (x0 >> 2) + tx, y = (y0 >> 2) + ty;
        const bool exists = prm.level_count > 2 && x < w2 && y < h2;
        l2 = exists ? m : 0.0f;
        if (exists) pyramid[prm.level_offset[2] + (size_t) y * w2 + x] = m;
    }
    s_l2[ty * 16 + tx] = l2;
    __syncthreads();

    // levels 3..6 of the tile: 8x8, 4x4, 2x2, 1x1, ping-ponging between two shared arrays
    float* src = s_l2;
    float* dst = s_l3;
    uint32_t side = 8;
#pragma unroll
    for (uint32_t level = 3; level <= 6; level++, side >>= 1)
    {
        if (tid < side * side)
        {
            const uint32_t cx = tid % side, cy = tid / side, src_side = side * 2;
            float m = 0.0f;
            m = gl_max(m, src[(2 * cy) * src_side + 2 * cx]);
            m = gl_max(m, src[(2 * cy + 1) * src_side + 2 * cx]);
            m = gl_max(m, src[(2 * cy) * src_side + 2 * cx + 1]);
            m = gl_max(m, src[(2 * cy + 1) * src_side + 2 * cx + 1]);
            const uint32_t wl = level_dim(W, level), hl = level_dim(H, level);
            const uint32_t x = (x0 >> level) + cx, y = (y0 >> level) + cy;
            const bool exists = level < prm.level_count && x < wl && y < hl;
            dst[cy * side + cx] = exists ? m : 0.0f;
            if (exists) pyramid[prm.level_offset[level] + (size_t) y * wl + x] = m;
        }
        __syncthreads();
        float* t = src; src = dst; dst = t;
    }
}

// remaining levels (7 and up): tiny, one CTA walks them in order
__global__ void __launch_bounds__(1024)
depth_pyramid_top_kernel(float* pyramid, pyramid_params prm, uint32_t first_level)
{
    for (uint32_t level = first_level; level < prm.level_count; level++)
    {
        const uint32_t wl = level_dim(prm.width, level), hl = level_dim(prm.height, level);
        const uint32_t ws = level_dim(prm.width, level - 1), hs = level_dim(prm.height, level - 1);
        const float* src = pyramid + prm.level_offset[level - 1];
        float* dst = pyramid + prm.level_offset[level];
        for (uint32_t i = threadIdx.x; i < wl * hl; i += 1024)
        {
            const uint32_t x = i % wl, y = i / wl;
            float m = 0.0f;
            for (uint32_t dx = 0; dx < 2; dx++)
                for (uint32_t dy = 0; dy < 2; dy++)
                {
                    const uint32_t sx = 2 * x + dx, sy = 2 * y + dy;
                    if (sx < ws && sy < hs) m = gl_max(m, src[(size_t) sy * ws + sx]);
                }
            dst[i] = m;
        }
        __threadfence_block();
        __syncthreads();
    }
}

using encode_fn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

encode_fn get_encode_fn()
{
    static encode_fn fn = []() -> encode_fn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<encode_fn>(p);
    }();
    return fn;
}

void fill_params(pyramid_params& prm, uint32_t w, uint32_t h)
{
    prm.width = w; prm.height = h; prm.level_count = level_count_of(w, h);
    unsigned long long off = 0;
    for (uint32_t l = 0; l < 16; l++)
    {
        prm.level_offset[l] = off;
        if (l < prm.level_count) off += (unsigned long long) level_dim(w, l) * level_dim(h, l);
    }
}

} // namespace
} // namespace vrenb200

using namespace vrenb200;

extern "C" uint32_t vrenb200_depth_pyramid_level_count(uint32_t width, uint32_t height)
{
    return (width == 0 || height == 0) ? 0u : level_count_of(width, height);
}

extern "C" uint32_t vrenb200_depth_pyramid_level_width(uint32_t width, uint32_t level) { return level_dim(width, level); }   // depth_buffer_pyramid.hpp:43-46
extern "C" uint32_t vrenb200_depth_pyramid_level_height(uint32_t height, uint32_t level) { return level_dim(height, level); } // :48-51

extern "C" size_t vrenb200_depth_pyramid_level_offset(uint32_t width, uint32_t height, uint32_t level)
{
    size_t off = 0;
    for (uint32_t l = 0; l < level; l++) off += (size_t) level_dim(width, l) * level_dim(height, l);
    return off;
}

extern "C" size_t vrenb200_depth_pyramid_bytes(uint32_t width, uint32_t height)
{
    return vrenb200_depth_pyramid_level_offset(width, height, vrenb200_depth_pyramid_level_count(width, height)) * sizeof(float);
}

extern "C" int vrenb200_depth_pyramid_build(vrenb200_stream_t stream, const float* depth, uint32_t width, uint32_t height, float* pyramid)
{
    if (!depth || !pyramid) return VRENB200_EINVAL_ARG;
    if (width == 0 || height == 0) return VRENB200_EINVAL_LENGTH;
    if (level_count_of(width, height) > 16) return VRENB200_ELIMIT;   // k_max_depth_buffer_pyramid_level_count, depth_buffer_pyramid.hpp:24
    cudaStream_t s = as_stream(stream);
    pyramid_params prm{};
    fill_params(prm, width, height);
    CUtensorMap map;
    std::memset(&map, 0, sizeof(map));
    prm.use_tma = 0;
    if (encode_fn enc = get_encode_fn())
    {
        if ((reinterpret_cast<uintptr_t>(depth) & 15) == 0 && ((uint64_t) width * 4) % 16 == 0)
        {
            const cuuint64_t dims[2] = {width, height};
            const cuuint64_t strides[1] = {(cuuint64_t) width * 4};
            const cuuint32_t box[2] = {64, 64};
            const cuuint32_t estr[2] = {1, 1};
            prm.use_tma = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(depth), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        }
    }
    const dim3 grid((width + 63) / 64, (height + 63) / 64);
    depth_pyramid_tile_kernel<<<grid, 256, 0, s>>>(map, depth, pyramid, prm);
    VRENB200_TRY(check_launch());
    if (prm.level_count > 7)
    {
        depth_pyramid_top_kernel<<<1, 1024, 0, s>>>(pyramid, prm, 7);
        VRENB200_TRY(check_launch());
    }
    return VRENB200_OK;
}
