// n3 — vren_demo::point_light_bouncer::bounce (reference: vren_demo/vren_demo/point_light_bouncer.{hpp,cpp},
// vren_demo/resources/shaders/bounce_point_lights.comp:33-73): the producer of the per-frame point-light positions that
// construct_point_light_bvh consumes.  Every light moves speed*dt along its direction inside an AABB and reflects off
// the faces it hits (at most 32 reflections per call).  One thread per light, 32 B in + 32 B out per light.
//
// fp32 contract as everywhere: the shader's op order with round-to-nearest IEEE ops (__fdiv_rn,
// __fmul_rn, __fadd_rn; the file is compiled with -fmad=false), min/max with fminf/fmaxf semantics (a NaN operand
// loses), so the oracle's plain C++ restatement is bit-identical.
#include "common.cuh"

namespace vrenb200 {
namespace {

constexpr float kBounceInf = 1e35f;
constexpr float kBounceEps = 1e-5f;
constexpr int kMaxBouncingIter = 32;

__device__ __forceinline__ float face_time(float face, float p, float d)
{
    const float t = __fdiv_rn(__fsub_rn(face, p), d);
    return t <= 0.0f ? kBounceInf : t;
}

__global__ void __launch_bounds__(256)
bounce_point_lights_kernel(float4* __restrict__ positions, float4* __restrict__ directions, uint32_t count,
                           float3 lo, float3 hi, float speed, float dt)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float4 p4 = positions[i], d4 = directions[i];
    float px = fminf(fmaxf(p4.x, lo.x), hi.x), py = fminf(fmaxf(p4.y, lo.y), hi.y), pz = fminf(fmaxf(p4.z, lo.z), hi.z);
    float dx = d4.x, dy = d4.y, dz = d4.z;
    float rem_t = __fmul_rn(speed, dt);
    for (int j = 0; j < kMaxBouncingIter && rem_t > 0.0f; j++)
    {
        const float tx1 = face_time(lo.x, px, dx), tx2 = face_time(hi.x, px, dx);
        const float ty1 = face_time(lo.y, py, dy), ty2 = face_time(hi.y, py, dy);
        const float tz1 = face_time(lo.z, pz, dz), tz2 = face_time(hi.z, pz, dz);
        const float min_t = fminf(tx1, fminf(tx2, fminf(ty1, fminf(ty2, fminf(tz1, tz2)))));
        const float step = fminf(__fsub_rn(min_t, kBounceEps), rem_t);
        px = __fadd_rn(px, __fmul_rn(step, dx));
        py = __fadd_rn(py, __fmul_rn(step, dy));
        pz = __fadd_rn(pz, __fmul_rn(step, dz));
        if (min_t < rem_t)
        {
            if (tx1 == min_t || tx2 == min_t) dx = -dx;
            if (ty1 == min_t || ty2 == min_t) dy = -dy;
            if (tz1 == min_t || tz2 == min_t) dz = -dz;
        }
        rem_t = __fsub_rn(rem_t, step);
    }
    positions[i] = make_float4(px, py, pz, 1.0f);
    directions[i] = make_float4(dx, dy, dz, 0.0f);
}

} // namespace
} // namespace vrenb200

using namespace vrenb200;

extern "C" int vrenb200_bounce_point_lights(vrenb200_stream_t stream, float* positions, float* directions, uint32_t count,
                                            const float aabb_min[3], const float aabb_max[3], float speed, float dt)
{
    if (positions == nullptr || directions == nullptr || aabb_min == nullptr || aabb_max == nullptr) return VRENB200_EINVAL_ARG;
    if (count == 0) return VRENB200_OK;   // the shader's bounds check makes an empty dispatch a no-op
    if (((reinterpret_cast<uintptr_t>(positions) | reinterpret_cast<uintptr_t>(directions)) & 15) != 0) return VRENB200_EALIGN;
    const float3 lo = make_float3(aabb_min[0], aabb_min[1], aabb_min[2]);
    const float3 hi = make_float3(aabb_max[0], aabb_max[1], aabb_max[2]);
    bounce_point_lights_kernel<<<(count + 255) / 256, 256, 0, as_stream(stream)>>>(
        reinterpret_cast<float4*>(positions), reinterpret_cast<float4*>(directions), count, lo, hi, speed, dt);
    return check_launch();
}
