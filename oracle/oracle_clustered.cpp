// TEST INFRASTRUCTURE — CPU restatement ("oracle") of vren's light-clustering pass (a6, a7, a8).
//
// PARITY UNPINNED UPSTREAM: the reference has no test, golden vector or CPU implementation for clustered shading
// (SURVEY 4, 8c), and it cannot be built here (no Vulkan/glslc).  This file restates the shaders line by line and
// fixes the choices the reference leaves open (SURVEY 8c i-vii); each is marked CANONICAL below.
//
// fp32 contract (CANONICAL vii): every operation is a separately rounded IEEE fp32 op in GLSL source order (this
// file is compiled with -ffp-contract=off); mat4*vec4 = ((m0*x + m1*y) + (m2*z + m3*w)); normalize(v) = v / sqrt(dot);
// dot(v,v) = (x*x + y*y) + z*z.  Transcendentals (tan, pow, log) are evaluated once on the host:
//   a        = 1.0f + (2.0f * tanf(fov_y / 2.0f)) / (float) Ty                (find_unique_clusters.comp:65)
//   slice k  = floor(logf(z / near) / logf(a)) in fp32, as the shader types it (r2: was evaluated in double, which disagrees
//              with the fp32 form for 0.0005-0.003 % of all depths, always by one slice at a boundary); negative -> 0
//   near_k   = near * powf(a, (float) k)                                       (clustered_shading.glsl:97-98)
//   inverse(projection) in closed form: i00 = 1/m00, i11 = 1/m11, row3 = (0, 0, 1/m32, -m22/m32), row2 = (0,0,0,1)
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

extern "C" {
void oracle_reduce(int dtype, int op, const void* in, uint32_t n, void* out, uint32_t blocks);
void oracle_bucket_sort(const uint32_t* in_pairs, uint32_t n, uint32_t* out_pairs, uint32_t* counters);
uint32_t oracle_calc_bvh_padded_leaf_count(uint32_t leaf_count);
uint32_t oracle_calc_bvh_buffer_length(uint32_t leaf_count);
uint32_t oracle_calc_bvh_level_count(uint32_t leaf_count);
uint32_t oracle_round_to_next_power_of_2(uint32_t v);
void oracle_build_bvh(void* nodes, uint32_t padded_leaf_count);
}

namespace {

struct bvh_node { float mn[3]; uint32_t next; float mx[3]; uint32_t pad; };
const uint32_t LEAF = 0xFFFFFFFFu, INVALID = 0xFFFFFFFEu;

struct camera_t { float fov_y, aspect, near_plane, far_plane; };

// camera.cpp:40-50 (glm column-major: m[col][row])
struct proj_t
{
    float m00, m11, m22, m32;      // m23 = 1
    float i00, i11, iB, nAB;       // closed-form inverse entries (CANONICAL vii)
    float tan_half;
};

proj_t make_projection(const camera_t& c)
{
    proj_t p;
    p.tan_half = tanf(c.fov_y / 2.0f);
    p.m00 = 1.0f / (p.tan_half * c.aspect);
    p.m11 = 1.0f / p.tan_half;
    p.m22 = c.far_plane / (c.far_plane - c.near_plane);
    p.m32 = -(c.far_plane * c.near_plane) / (c.far_plane - c.near_plane);
    p.i00 = 1.0f / p.m00;
    p.i11 = 1.0f / p.m11;
    p.iB = 1.0f / p.m32;
    p.nAB = (-p.m22) / p.m32;
    return p;
}

// find_unique_clusters.comp:52-57: frag_pos = inverse(P) * (., ., d, 1); frag_pos /= frag_pos.w  ->  z = 1 / w'
float view_z(float d, const proj_t& pr)
{
    const float w = d * pr.iB + 1.0f * pr.nAB;
    return 1.0f / w;
}

float slice_base(const proj_t& p, uint32_t tiles_y)
{
    return 1.0f + (2.0f * p.tan_half) / (float) tiles_y;
}

uint32_t slice_of(float z, float near_plane, float a)
{
    // find_unique_clusters.comp:65 as GLSL types it: every operand and operation is fp32 (libm logf; division rounded once)
    const float k = std::floor(logf(z / near_plane) / logf(a));
    if (!(k >= 0.0f)) return 0u;                // CANONICAL: uint(negative or NaN) -> 0
    if (k > 4294967040.0f) return 0xFFFFFFFFu;
    return (uint32_t) k;
}

float half_to_float(uint16_t h)
{
    const uint32_t s = (h >> 15) & 1u, e = (h >> 10) & 31u, m = h & 1023u;
    uint32_t u;
    if (e == 0)
    {
        if (m == 0) u = s << 31;
        else
        {
            int ee = -1;
            uint32_t mm = m;
            do { ee++; mm <<= 1; } while ((mm & 1024u) == 0);
            u = (s << 31) | ((uint32_t) (127 - 15 - ee) << 23) | ((mm & 1023u) << 13);
        }
    }
    else if (e == 31) u = (s << 31) | 0x7F800000u | (m << 13);
    else u = (s << 31) | ((e + 112u) << 23) | (m << 13);
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

// clustered_shading.glsl:7-43
uint32_t discretize_normal(float nx, float ny, float nz)
{
    if (nx == 0.0f && ny == 0.0f && nz == 0.0f) return 0xFFFFFFFFu;
    const float n[3] = {nx, ny, nz};
    float min_t = 1e35f;
    uint32_t axis = 0, face_idx = 0; // CANONICAL: defined start values (the shader leaves them uninitialised)
    for (uint32_t i = 0; i < 3; i++)
    {
        const float sg = n[i] > 0.0f ? 1.0f : (n[i] < 0.0f ? -1.0f : 0.0f);
        const float t = sg / n[i];
        if (t < min_t)
        {
            min_t = t;
            axis = i;
            face_idx = (n[i] > 0.0f ? 1u : 0u) * 3u + i;
        }
    }
    const float p[3] = {n[0] * min_t, n[1] * min_t, n[2] * min_t};
    const float uvx = p[(axis + 1) % 3], uvy = p[(axis + 2) % 3];
    const float fx = std::floor((uvx + 1.0f) / 2.0f * 3.0f), fy = std::floor((uvy + 1.0f) / 2.0f * 3.0f);
    const uint32_t dx = fx >= 0.0f ? (uint32_t) fx : 0u, dy = fy >= 0.0f ? (uint32_t) fy : 0u;
    return (face_idx * 9u + dx * 3u + dy) & 0x3Fu;
}

// point_light_position_to_view_space.comp:30 — mat4 * vec4(p.xyz, 1) as ((m0*x + m1*y) + (m2*z + m3*1))
void to_view_space(const float* view, const float* p, float* out)
{
    for (int c = 0; c < 4; c++)
    {
        const float a = view[0 + c] * p[0] + view[4 + c] * p[1];
        const float b = view[8 + c] * p[2] + view[12 + c] * 1.0f;
        out[c] = a + b;
    }
}

// discretize_point_light_positions.comp:31-39
uint32_t morton_code(const float* vp, const float* mn, const float* mx)
{
    uint32_t q[3];
    for (int c = 0; c < 3; c++)
    {
        const float t = (vp[c] - mn[c]) / (mx[c] - mn[c]) * 32.0f;
        const float f = std::floor(t);
        q[c] = f >= 0.0f ? (uint32_t) f : 0u; // CANONICAL v: NaN (max == min) -> bin 0
    }
    uint32_t code = 0;
    for (uint32_t b = 0; b < 5; b++)
    {
        code |= ((q[0] >> b) & 1u) << (b * 3);
        code |= ((q[1] >> b) & 1u) << (b * 3 + 1);
        code |= ((q[2] >> b) & 1u) << (b * 3 + 2);
    }
    return code;
}

// init_light_array_bvh.comp:52-53
void light_leaf_box(const float* vp, float intensity, float* mn, float* mx)
{
    for (int c = 0; c < 3; c++) { mn[c] = vp[c] - intensity; mx[c] = vp[c] + intensity; }
}

// assign_lights.comp:84-94, (min_1, max_1) = cluster corners as written, (min_2, max_2) = node box
bool test_aabb_aabb(const float* min1, const float* max1, const float* min2, const float* max2)
{
    return max1[0] >= min2[0] && min1[0] <= max2[0] && max1[1] >= min2[1] && min1[1] <= max2[1] && max1[2] >= min2[2] && min1[2] <= max2[2];
}

// assign_lights.comp:97-102
bool test_sphere_aabb(const float* o, float r, const float* amin, const float* amax)
{
    float q[3];
    for (int k = 0; k < 3; k++)
    {
        const float m = amax[k] < o[k] ? amax[k] : o[k];      // min(sphere_o, aabb_max)
        q[k] = amin[k] < m ? m : amin[k];                     // max(aabb_min, .)
        q[k] = q[k] - o[k];
    }
    const float d = std::sqrt((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]);
    return d < r;
}

// get_node_address, assign_lights.comp:108-119 (exp2 of small integers is exact: integer powers of 32)
int64_t node_address(uint32_t bvh_root_index, uint32_t level, const uint32_t* overlaps)
{
    int64_t sum = 0, p32 = 1;
    for (uint32_t e = 0; e <= level + 1; e++) { sum += p32; p32 *= 32; }
    int64_t addr = (int64_t) bvh_root_index - (sum - 1);
    for (uint32_t i = 0; i < level; i++)
    {
        int64_t w = 1;
        for (uint32_t e = 0; e < level - i; e++) w *= 32;
        addr += w * (int64_t) __builtin_ctz(overlaps[i]);
    }
    return addr;
}

} // namespace

extern "C" {

// ---- a6 ------------------------------------------------------------------------------------------------------------
// positions vec4[L], lights {vec3 color; float intensity}[L], view = column-major mat4.
// Outputs: view_pos vec4[L], nodes[calc_bvh_buffer_length(L)], sorted uvec2[L].
void oracle_construct_point_light_bvh(const float* positions, const float* lights, uint32_t L, const float* view,
                                      float* view_pos, bvh_node* nodes, uint32_t* sorted_pairs)
{
    // K9: point_light_position_to_view_space.comp:30
    for (uint32_t i = 0; i < L; i++) to_view_space(view, positions + 4 * (size_t) i, view_pos + 4 * (size_t) i);
    // clustered_shading.cpp:141-178: reduce<vec4,max>, reduce<vec4,min> over next_pow2(L) slots, result = last slot
    const uint32_t P = oracle_round_to_next_power_of_2(L);
    std::vector<float> tree((size_t) P * 4);
    float mx[4], mn[4];
    oracle_reduce(1, 2, view_pos, L, tree.data(), 1);
    std::memcpy(mx, &tree[(size_t) (P - 1) * 4], 16);
    oracle_reduce(1, 1, view_pos, L, tree.data(), 1);
    std::memcpy(mn, &tree[(size_t) (P - 1) * 4], 16);
    // K10: discretize_point_light_positions.comp:26-43
    std::vector<uint32_t> pairs((size_t) L * 2);
    for (uint32_t i = 0; i < L; i++)
    {
        pairs[2 * (size_t) i] = morton_code(view_pos + 4 * (size_t) i, mn, mx);
        pairs[2 * (size_t) i + 1] = i;
    }
    // bucket sort (CANONICAL i: ties keep input order)
    std::vector<uint32_t> counters(65536);
    oracle_bucket_sort(pairs.data(), L, sorted_pairs, counters.data());
    // K11: init_light_array_bvh.comp:42-65 (CANONICAL iii: invalid leaves get the empty box, _pad = 0)
    const uint32_t padded = oracle_calc_bvh_padded_leaf_count(L);
    for (uint32_t i = 0; i < padded; i++)
    {
        bvh_node& nd = nodes[i];
        nd.pad = 0;
        if (i < L)
        {
            const uint32_t l = sorted_pairs[2 * (size_t) i + 1];
            light_leaf_box(view_pos + 4 * (size_t) l, lights[4 * (size_t) l + 3], nd.mn, nd.mx);
            nd.next = LEAF;
        }
        else
        {
            for (int c = 0; c < 3; c++) { nd.mn[c] = 1e35f; nd.mx[c] = -1e35f; }
            nd.next = INVALID;
        }
    }
    oracle_build_bvh(nodes, padded);
}

// ---- a7 ------------------------------------------------------------------------------------------------------------
// find_unique_clusters.comp:46-121.  depth float[W*H]; normals = RGBA16F as uint16[W*H*4] or NULL (all-zero).
// keys_out must hold Tx*Ty*1024 entries at worst; cluster_ref uint[W*H].
// CANONICAL ii: tiles visited in tile-major order (j*Tx + i), ascending key inside a tile.
// CANONICAL vi: pixels of partial tiles wrap (sampler REPEAT), nearest texel.
// Returns the number of unique cluster keys (dispatch_params.x).
uint32_t oracle_find_unique_clusters(const float* depth, const uint16_t* normals, uint32_t W, uint32_t H,
                                     const camera_t* cam, uint32_t* keys_out, uint32_t* cluster_ref)
{
    const proj_t pr = make_projection(*cam);
    const uint32_t Tx = (W + 31) / 32, Ty = (H + 31) / 32;
    const float a = slice_base(pr, Ty);
    uint32_t count = 0;
    std::vector<uint32_t> keys(1024);
    for (uint32_t j = 0; j < Ty; j++)
        for (uint32_t i = 0; i < Tx; i++)
        {
            for (uint32_t t = 0; t < 1024; t++)
            {
                const uint32_t x = ((i << 5) + (t & 31)) % W, y = ((j << 5) + (t >> 5)) % H;
                const float d = depth[(size_t) y * W + x];
                const uint32_t k = slice_of(view_z(d, pr), cam->near_plane, a);
                uint32_t nb = 0xFFFFFFFFu;
                if (normals)
                {
                    const uint16_t* h = normals + ((size_t) y * W + x) * 4;
                    nb = discretize_normal(half_to_float(h[0]), half_to_float(h[1]), half_to_float(h[2]));
                }
                keys[t] = (i & 0xFFu) | ((j & 0xFFu) << 8) | ((k & 0x3FFu) << 16) | (nb << 26);
            }
            std::vector<uint32_t> uniq(keys);
            std::sort(uniq.begin(), uniq.end());
            uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
            for (uint32_t t = 0; t < 1024; t++)
            {
                const uint32_t x = (i << 5) + (t & 31), y = (j << 5) + (t >> 5);
                if (x < W && y < H)
                    cluster_ref[(size_t) y * W + x] =
                        count + (uint32_t) (std::lower_bound(uniq.begin(), uniq.end(), keys[t]) - uniq.begin());
            }
            for (uint32_t u : uniq) keys_out[count++] = u;
        }
    return count;
}

// ---- a8 ------------------------------------------------------------------------------------------------------------
// clustered_shading.glsl:72-110 — cluster "min/max" corners exactly as written (NOT component-wise ordered)
namespace {
void cluster_aabb(uint32_t ci, uint32_t cj, uint32_t ck, uint32_t Tx, uint32_t Ty, const camera_t& cam,
                         const proj_t& pr, float a, float* cmin, float* cmax)
{
    const float tx = (float) Tx, ty = (float) Ty;
    float p0x = (float) ci / tx, p0y = (float) cj / ty;
    p0x = p0x * 2.0f - 1.0f;
    p0y = (1.0f - p0y) * 2.0f - 1.0f;
    float p1x = (float) (ci + 1) / tx, p1y = (float) (cj + 1) / ty;
    p1x = p1x * 2.0f - 1.0f;
    p1y = (1.0f - p1y) * 2.0f - 1.0f;
    // inverse(proj) * (px, py, 0, 1): x' = i00*px, y' = i11*py, z' = 1, w' = (iB*0 + nAB*1)
    const float w = 0.0f * pr.iB + 1.0f * pr.nAB;
    const float v0[3] = {(pr.i00 * p0x) / w, (pr.i11 * p0y) / w, 1.0f / w};
    const float v1[3] = {(pr.i00 * p1x) / w, (pr.i11 * p1y) / w, 1.0f / w};
    const float near_k = cam.near_plane * powf(a, (float) ck);
    const float far_k = near_k * a;
    const float l0 = std::sqrt((v0[0] * v0[0] + v0[1] * v0[1]) + v0[2] * v0[2]);
    const float l1 = std::sqrt((v1[0] * v1[0] + v1[1] * v1[1]) + v1[2] * v1[2]);
    const float d0[3] = {v0[0] / l0, v0[1] / l0, v0[2] / l0};
    const float d1[3] = {v1[0] / l1, v1[1] / l1, v1[2] / l1};
    const float s0 = near_k / d0[2], s1 = far_k / d1[2];
    for (int c = 0; c < 3; c++) { cmin[c] = s0 * d0[c]; cmax[c] = s1 * d1[c]; }
}
} // namespace

// assign_lights.comp:121-241 + clustered_shading.cpp:473-690.
// counts/offsets: uint[max_keys] (offsets = exclusive scan of counts over all max_keys slots);
// indices: hits in traversal order, within a leaf group in DESCENDING lane order (:229-232).
// CANONICAL iii/iv: INVALID nodes and padding leaves never overlap.
// Returns the total number of assigned lights (indices beyond max_assigned are counted but not written).
uint64_t oracle_assign_lights(uint32_t W, uint32_t H, const camera_t* cam, const uint32_t* cluster_keys, uint32_t key_count,
                              uint32_t max_keys, const bvh_node* bvh, uint32_t bvh_root_index, uint32_t light_count,
                              const uint32_t* sorted_pairs, const float* view_pos,
                              uint32_t* indices, uint64_t max_assigned, uint32_t* counts, uint32_t* offsets)
{
    std::fill(counts, counts + max_keys, 0u);                     // vkCmdFillBuffer, clustered_shading.cpp:492
    if (light_count == 0) return 0;                                // :494-496 nothing else is touched
    const proj_t pr = make_projection(*cam);
    const uint32_t Tx = (W + 31) / 32, Ty = (H + 31) / 32;
    const float a = 2.0f * pr.tan_half / (float) Ty + 1.0f;        // clustered_shading.glsl:96
    const uint32_t levels = oracle_calc_bvh_level_count(light_count);
    std::vector<std::vector<uint32_t>> lists(key_count);
    for (uint32_t c = 0; c < key_count; c++)
    {
        const uint32_t key = cluster_keys[c];
        float cmin[3], cmax[3];
        cluster_aabb(key & 0xFFu, (key >> 8) & 0xFFu, (key >> 16) & 0x3FFu, Tx, Ty, *cam, pr, a, cmin, cmax);
        // the state machine of assign_lights.comp:133-239 is a depth-first walk, lowest set bit first
        uint32_t overlaps[8] = {0};
        uint32_t level = 0;
        int state = 2;
        std::vector<uint32_t>& out = lists[c];
        while (true)
        {
            if (state == 0)
            {
                overlaps[level] &= ~(1u << __builtin_ctz(overlaps[level]));
                if (overlaps[level] == 0) state = 1;
                else { level++; state = 2; }
            }
            else if (state == 1)
            {
                if (level > 0) { level--; state = 0; }
                else break;
            }
            else
            {
                const uint32_t addr = (uint32_t) node_address(bvh_root_index, level, overlaps);
                if (level < levels - 1)
                {
                    uint32_t mask = 0;
                    for (uint32_t t = 0; t < 32; t++)
                    {
                        const bvh_node& n = bvh[addr + t];
                        if (n.next == INVALID) continue;
                        if (test_aabb_aabb(cmin, cmax, n.mn, n.mx)) mask |= 1u << t;
                    }
                    overlaps[level] = mask;
                    if (mask == 0) state = 1;
                    else level++;
                }
                else
                {
                    uint32_t mask = 0;
                    for (uint32_t t = 0; t < 32; t++)
                    {
                        if (addr + t >= light_count) continue;
                        const bvh_node& n = bvh[addr + t];
                        const uint32_t l = sorted_pairs[2 * (size_t) (addr + t) + 1];
                        const float* o = view_pos + 4 * (size_t) l;
                        if (test_sphere_aabb(o, (n.mx[0] - n.mn[0]) / 2.0f, cmin, cmax)) mask |= 1u << t;
                    }
                    for (int t = 31; t >= 0; t--)
                        if (mask & (1u << t)) out.push_back(sorted_pairs[2 * (size_t) (addr + t) + 1]);
                    state = 1;
                }
            }
        }
        counts[c] = (uint32_t) out.size();
    }
    uint32_t acc = 0;
    for (uint32_t c = 0; c < max_keys; c++) { offsets[c] = acc; acc += counts[c]; }  // blelloch_scan over max_keys
    uint64_t total = 0;
    for (uint32_t c = 0; c < key_count; c++)
    {
        for (size_t r = 0; r < lists[c].size(); r++)
            if ((uint64_t) offsets[c] + r < max_assigned) indices[offsets[c] + r] = lists[c][r];
        total += lists[c].size();
    }
    return total;
}

// ---- the pure functions above, one call per array: what tests/test_clustered_reference.py compares with the reference's own
// GLSL compiled through oracle/glsl_shim.hpp (oracle/_ref/libvrenref_glsl.so) and with tests/golden/clustered_reference_outputs.npz
void oracle_projection(const camera_t* cam, float* proj16, float* inverse16)
{
    const proj_t p = make_projection(*cam);
    for (int i = 0; i < 16; i++) proj16[i] = inverse16[i] = 0.0f;
    proj16[0] = p.m00; proj16[5] = p.m11; proj16[10] = p.m22; proj16[11] = 1.0f; proj16[14] = p.m32;          // camera.cpp:40-50, m[col][row]
    inverse16[0] = p.i00; inverse16[5] = p.i11; inverse16[11] = p.iB; inverse16[14] = 1.0f; inverse16[15] = p.nAB;   // closed form
}
void oracle_discretize_normal(const float* n3, uint32_t count, uint32_t* out)
{
    for (uint32_t i = 0; i < count; i++) out[i] = discretize_normal(n3[3 * i], n3[3 * i + 1], n3[3 * i + 2]);
}
// key of tile (0, 0): slice << 16 | normal bin << 26; normals3 may be NULL (bin 63)
void oracle_cluster_key(const float* depth, const float* normals3, uint32_t count, uint32_t tiles_y, const camera_t* cam, uint32_t* out_key, float* out_view_z)
{
    const proj_t pr = make_projection(*cam);
    const float a = slice_base(pr, tiles_y);
    for (uint32_t i = 0; i < count; i++)
    {
        const float z = view_z(depth[i], pr);
        const uint32_t nb = normals3 ? discretize_normal(normals3[3 * i], normals3[3 * i + 1], normals3[3 * i + 2]) : 0xFFFFFFFFu;
        out_key[i] = ((slice_of(z, cam->near_plane, a) & 0x3FFu) << 16) | (nb << 26);
        out_view_z[i] = z;
    }
}
void oracle_cluster_aabb(const uint32_t* ijk3, uint32_t count, uint32_t tiles_x, uint32_t tiles_y, const camera_t* cam, float* out_min3, float* out_max3)
{
    const proj_t pr = make_projection(*cam);
    const float a = 2.0f * pr.tan_half / (float) tiles_y + 1.0f;
    for (uint32_t i = 0; i < count; i++)
        cluster_aabb(ijk3[3 * i], ijk3[3 * i + 1], ijk3[3 * i + 2], tiles_x, tiles_y, *cam, pr, a, out_min3 + 3 * (size_t) i, out_max3 + 3 * (size_t) i);
}
void oracle_test_aabb_aabb(const float* b12, uint32_t count, uint8_t* out)
{
    for (uint32_t i = 0; i < count; i++) out[i] = test_aabb_aabb(b12 + 12 * (size_t) i, b12 + 12 * (size_t) i + 3, b12 + 12 * (size_t) i + 6, b12 + 12 * (size_t) i + 9);
}
void oracle_test_sphere_aabb(const float* s10, uint32_t count, uint8_t* out)
{
    for (uint32_t i = 0; i < count; i++) out[i] = test_sphere_aabb(s10 + 10 * (size_t) i, s10[10 * (size_t) i + 3], s10 + 10 * (size_t) i + 4, s10 + 10 * (size_t) i + 7);
}
int32_t oracle_get_node_address(uint32_t bvh_root_index, uint32_t level, const uint32_t* level_overlaps4)
{
    return (int32_t) node_address(bvh_root_index, level, level_overlaps4);
}
void oracle_morton_code(const float* pos3, uint32_t count, const float* mn, const float* mx, uint32_t* out)
{
    for (uint32_t i = 0; i < count; i++) out[i] = morton_code(pos3 + 3 * (size_t) i, mn, mx);
}
void oracle_position_to_view_space(const float* view16, const float* pos4, uint32_t count, float* out4)
{
    for (uint32_t i = 0; i < count; i++) to_view_space(view16, pos4 + 4 * (size_t) i, out4 + 4 * (size_t) i);
}
void oracle_light_leaf_box(const float* view_pos4, const float* intensity, uint32_t count, float* out_min3, float* out_max3)
{
    for (uint32_t i = 0; i < count; i++) light_leaf_box(view_pos4 + 4 * (size_t) i, intensity[i], out_min3 + 3 * (size_t) i, out_max3 + 3 * (size_t) i);
}

// ---- n2: depth-buffer pyramid -------------------------------------------------------------------------------------------
// depth_buffer_copy.comp:8-18 (level 0 = copy) + depth_buffer_reduce.comp:10-35 per level (depth_buffer_pyramid.cpp:177-305).
// pyramid: levels back to back, level l = max(W>>l,1) x max(H>>l,1), level_count = floor(log2(max(W,H))) + 1 (:18).
uint32_t oracle_depth_pyramid(const float* depth, uint32_t W, uint32_t H, float* pyramid)
{
    uint32_t levels = 1;
    for (uint32_t m = W > H ? W : H; m >>= 1;) levels++;
    std::memcpy(pyramid, depth, (size_t) W * H * sizeof(float));
    const float* from = pyramid;
    uint32_t fw = W, fh = H;
    float* to = pyramid + (size_t) W * H;
    for (uint32_t l = 1; l < levels; l++)
    {
        const uint32_t tw = (W >> l) ? (W >> l) : 1u, th = (H >> l) ? (H >> l) : 1u;
        for (uint32_t y = 0; y < th; y++)
            for (uint32_t x = 0; x < tw; x++)
            {
                float max_depth = 0.0f;
                for (uint32_t dx = 0; dx < 2; dx++)
                    for (uint32_t dy = 0; dy < 2; dy++)
                    {
                        const uint32_t sx = 2 * x + dx, sy = 2 * y + dy;
                        if (sx < fw && sy < fh)
                        {
                            const float d = from[(size_t) sy * fw + sx];
                            max_depth = max_depth < d ? d : max_depth;   // GLSL max(max_depth, depth)
                        }
                    }
                to[(size_t) y * tw + x] = max_depth;
            }
        from = to;
        to += (size_t) tw * th;
        fw = tw; fh = th;
    }
    return levels;
}

// ---- n1: per-pixel walk of the light lists ------------------------------------------------------------------------------
// shade.comp:101-105 access pattern with the integer payload of show_clusters.comp:97-118 (count, XOR of the indices)
void oracle_light_list_hash(uint32_t W, uint32_t H, const uint32_t* cluster_ref, const uint32_t* counts, const uint32_t* offsets,
                            const uint32_t* indices, uint32_t* out_count_hash)
{
    for (size_t p = 0; p < (size_t) W * H; p++)
    {
        const uint32_t c = cluster_ref[p];
        uint32_t h = 0;
        for (uint32_t i = 0; i < counts[c]; i++) h ^= indices[offsets[c] + i];
        out_count_hash[2 * p] = counts[c];
        out_count_hash[2 * p + 1] = h;
    }
}

} // extern "C"

// ---- n3: light animation producer ---------------------------------------------------------------------------------------
// bounce_point_lights.comp:33-73, one light after the other.  GLSL min/max restated with fmin/fmax semantics (a NaN
// operand loses; the shader leaves that case undefined), divisions and the p += step * d update as separate IEEE fp32
// operations (this file is built with -ffp-contract=off).
extern "C" void oracle_bounce_point_lights(float* positions, float* directions, uint32_t count, const float* lo, const float* hi,
                                           float speed, float dt)
{
    const float INF = 1e35f, EPS = 1e-5f;
    for (uint32_t i = 0; i < count; i++)
    {
        float p[3], d[3];
        for (int a = 0; a < 3; a++)
        {
            p[a] = std::fmin(std::fmax(positions[i * 4 + a], lo[a]), hi[a]);
            d[a] = directions[i * 4 + a];
        }
        float rem_t = speed * dt;
        for (uint32_t j = 0; j < 32 && rem_t > 0; j++)
        {
            float t1[3], t2[3];
            for (int a = 0; a < 3; a++)
            {
                t1[a] = (lo[a] - p[a]) / d[a]; t1[a] = t1[a] <= 0 ? INF : t1[a];
                t2[a] = (hi[a] - p[a]) / d[a]; t2[a] = t2[a] <= 0 ? INF : t2[a];
            }
            const float min_t = std::fmin(t1[0], std::fmin(t2[0], std::fmin(t1[1], std::fmin(t2[1], std::fmin(t1[2], t2[2])))));
            const float step = std::fmin(min_t - EPS, rem_t);
            for (int a = 0; a < 3; a++)
            {
                const float sd = step * d[a];
                p[a] = p[a] + sd;
            }
            if (min_t < rem_t)
                for (int a = 0; a < 3; a++)
                    if (t1[a] == min_t || t2[a] == min_t) d[a] = -d[a];
            rem_t = rem_t - step;
        }
        for (int a = 0; a < 3; a++) { positions[i * 4 + a] = p[a]; directions[i * 4 + a] = d[a]; }
        positions[i * 4 + 3] = 1.0f;
        directions[i * 4 + 3] = 0.0f;
    }
}

// ---- n4: BVH debug consumer ----------------------------------------------------------------------------------------------
// visualize_bvh.cpp:59-94 (one dispatch per level, leaves first) + show_bvh.comp:62-78 (12 lines per node, written with the
// VREN_WRITE_DEBUG_DRAW_BUFFER_AABB macro, show_bvh.comp:46-60).  nodes: {min[3], next, max[3], pad} x N; vertices: {pos[3], color}.
extern "C" void oracle_visualize_bvh(const uint32_t* nodes, uint32_t level_count, uint32_t* vertices)
{
    static const uint32_t colors[7] = {0xff0000, 0xffff00, 0x00ff00, 0x0000ff, 0x00ffff, 0xff00ff, 0xffffff};
    uint32_t offset = 0;
    for (int32_t level = (int32_t) level_count; level >= 0; level--)
    {
        const uint32_t node_count = 1u << (5 * level);
        const uint32_t color = colors[level_count - level + 2];
        for (uint32_t i = 0; i < node_count; i++)
        {
            const uint32_t node_idx = i + offset;
            const uint32_t* n = nodes + (size_t) node_idx * 8;
            uint32_t m[3] = {n[0], n[1], n[2]}, M[3] = {n[4], n[5], n[6]}, c = color;
            if (n[3] == 0xFFFFFFFEu) { m[0] = m[1] = m[2] = M[0] = M[1] = M[2] = 0u; c = 0u; }   // float 0.0 bits
            uint32_t* out = vertices + (size_t) node_idx * 24 * 4;
            auto line = [&](uint32_t ax, uint32_t ay, uint32_t az, uint32_t bx, uint32_t by, uint32_t bz) {
                out[0] = ax; out[1] = ay; out[2] = az; out[3] = c;
                out[4] = bx; out[5] = by; out[6] = bz; out[7] = c;
                out += 8;
            };
            line(m[0], m[1], m[2], M[0], m[1], m[2]);
            line(M[0], m[1], m[2], M[0], m[1], M[2]);
            line(M[0], m[1], M[2], m[0], m[1], M[2]);
            line(m[0], m[1], M[2], m[0], m[1], m[2]);
            line(m[0], M[1], m[2], M[0], M[1], m[2]);
            line(M[0], M[1], m[2], M[0], M[1], M[2]);
            line(M[0], M[1], M[2], m[0], M[1], M[2]);
            line(m[0], M[1], M[2], m[0], M[1], m[2]);
            line(m[0], m[1], m[2], m[0], M[1], m[2]);
            line(M[0], m[1], m[2], M[0], M[1], m[2]);
            line(M[0], m[1], M[2], M[0], M[1], M[2]);
            line(m[0], m[1], M[2], m[0], M[1], M[2]);
        }
        offset += node_count;
    }
}
