// Exercises the C++ facade (include/vren/) the way the reference's own tests exercise vren
// (vren_test/vren_test/primitives/*.cpp): fill a buffer, compute the check on the CPU (std::sort, std::exclusive_scan,
// tree reduce, linear AABB scan), run the primitive through immediate_graphics_queue_submit, compare element-wise.
// Prints one line per test and exits non-zero on the first mismatch.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <random>
#include <vector>

#include "vren/base/kd_tree.hpp"
#include "vren/context.hpp"
#include "vren/pipeline/clustered_shading.hpp"
#include "vren/pipeline/depth_buffer_pyramid.hpp"
#include "vren_demo/point_light_bouncer.hpp"
#include "vren_demo/visualize_bvh.hpp"

static int g_failures = 0;
#define EXPECT(cond, ...)                                            \
    do                                                               \
    {                                                                \
        if (!(cond))                                                 \
        {                                                            \
            std::printf("FAIL %s:%d: ", __FILE__, __LINE__);         \
            std::printf(__VA_ARGS__);                                \
            std::printf("\n");                                       \
            g_failures++;                                            \
            return;                                                  \
        }                                                            \
    } while (0)

template <typename T> static void upload(vren::vk_utils::buffer const& b, std::vector<T> const& v, size_t byte_offset = 0)
{
    cudaMemcpy(b.ptr<char>(byte_offset), v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
}
template <typename T> static std::vector<T> download(vren::vk_utils::buffer const& b, size_t count, size_t byte_offset = 0)
{
    std::vector<T> v(count);
    cudaMemcpy(v.data(), b.ptr<char>(byte_offset), count * sizeof(T), cudaMemcpyDeviceToHost);
    return v;
}

// TEST(reduce, var_length) + TEST(reduce, type_uint): whole padded tree equals run_cpu_reduce (reduce.cpp:72-98,245-291);
// TEST(reduce, type_vec4): component-wise vec4 max, in place
static void test_reduce(vren::context& ctx)
{
    for (uint32_t length : {1u, 10u, 100u, 1000u, 10000u, 100000u})
    {
        const uint32_t P = vren::round_to_next_power_of_2(length);
        std::vector<uint32_t> cpu(P, 0u);
        for (uint32_t i = 0; i < length; i++) cpu[i] = 1 + (i % 7);
        auto in = vren::vk_utils::alloc_device_only_buffer(ctx, P * 4), out = vren::vk_utils::alloc_device_only_buffer(ctx, P * 4);
        upload(in, cpu);
        vren::vk_utils::immediate_graphics_queue_submit(ctx, [&](VkCommandBuffer cmd, vren::resource_container& res) {
            ctx.m_toolbox->m_reduce_uint_add(cmd, res, in, length, 0, out, 0, 1);
        });
        for (uint32_t i = 0; (1u << i) < P; i++)
            for (uint32_t j = 0; j < (P >> (i + 1)); j++)
            {
                const uint32_t a = (1u << i) - 1 + (j << (i + 1)), b = a + (1u << i);
                cpu[b] = cpu[a] + cpu[b];
            }
        auto gpu = download<uint32_t>(out, P);
        EXPECT(gpu == cpu, "reduce uint add length %u", length);
    }
    // vec4 max, in place
    const uint32_t n = 10000, P = vren::round_to_next_power_of_2(n);
    std::vector<float> v((size_t) P * 4, -1e35f);
    std::mt19937 rng(1);
    for (uint32_t i = 0; i < n * 4; i++) v[i] = (float) (rng() % 100);
    auto buf = vren::vk_utils::alloc_device_only_buffer(ctx, (size_t) P * 16);
    upload(buf, v);
    vren::vk_utils::immediate_graphics_queue_submit(ctx, [&](VkCommandBuffer cmd, vren::resource_container& res) {
        ctx.m_toolbox->m_reduce_vec4_max(cmd, res, buf, n, 0, 1);
    });
    auto got = download<float>(buf, (size_t) P * 4);
    for (int c = 0; c < 4; c++)
    {
        float m = -1e35f;
        for (uint32_t i = 0; i < n; i++) m = std::max(m, v[i * 4 + c]);
        EXPECT(got[(size_t) (P - 1) * 4 + c] == m, "reduce vec4 max comp %d", c);
    }
    std::printf("ok reduce\n");
}

// TEST(blelloch_scan, main): all ones, 2^0..2^19 vs std::exclusive_scan (blelloch_scan.cpp:122-176)
static void test_scan(vren::context& ctx)
{
    for (uint32_t log2n = 0; log2n < 20; log2n++)
    {
        const uint32_t n = 1u << log2n;
        std::vector<uint32_t> cpu(n, 1u), want(n);
        std::exclusive_scan(cpu.begin(), cpu.end(), want.begin(), 0u);
        auto buf = vren::vk_utils::alloc_device_only_buffer(ctx, (size_t) n * 4);
        upload(buf, cpu);
        vren::vk_utils::immediate_graphics_queue_submit(ctx, [&](VkCommandBuffer cmd, vren::resource_container& res) {
            ctx.m_toolbox->m_blelloch_scan(cmd, res, buf, n, 0, 1);
        });
        EXPECT(download<uint32_t>(buf, n) == want, "scan length %u", n);
    }
    bool threw = false;
    try
    {
        auto buf = vren::vk_utils::alloc_device_only_buffer(ctx, 4000);
        vren::vk_utils::immediate_graphics_queue_submit(ctx, [&](VkCommandBuffer cmd, vren::resource_container& res) {
            ctx.m_toolbox->m_blelloch_scan(cmd, res, buf, 1000, 0, 1);
        });
    }
    catch (std::invalid_argument const&) { threw = true; }
    EXPECT(threw, "scan must reject non power-of-2 lengths");
    std::printf("ok blelloch_scan\n");
}

// TEST(radix_sort, main): reversed iota 2^10 vs std::sort (+ the commented-out loop to 2^20 with random keys)
static void test_radix_sort(vren::context& ctx)
{
    vren::radix_sort& radix_sort = ctx.m_toolbox->m_radix_sort;
    std::mt19937 rng(2);
    for (uint32_t n = 1u << 10; n <= 1u << 20; n <<= 2)
    {
        std::vector<uint32_t> cpu(n);
        for (uint32_t i = 0; i < n; i++) cpu[i] = n == 1024 ? n - i - 1 : (uint32_t) rng();
        auto buf = vren::vk_utils::alloc_device_only_buffer(ctx, (size_t) n * 4);
        auto s1 = radix_sort.create_scratch_buffer_1(n), s2 = radix_sort.create_scratch_buffer_2(n);
        upload(buf, cpu);
        std::sort(cpu.begin(), cpu.end());
        vren::vk_utils::immediate_graphics_queue_submit(ctx, [&](VkCommandBuffer cmd, vren::resource_container& res) {
            radix_sort(cmd, res, buf, n, s1, s2);
        });
        EXPECT(download<uint32_t>(buf, n) == cpu, "radix sort length %u", n);
    }
    bool threw = false;
    try
    {
        auto buf = vren::vk_utils::alloc_device_only_buffer(ctx, 4096);
        auto s1 = radix_sort.create_scratch_buffer_1(1000), s2 = radix_sort.create_scratch_buffer_2(1000);
        vren::vk_utils::immediate_graphics_queue_submit(ctx, [&](VkCommandBuffer cmd, vren::resource_container& res) {
            radix_sort(cmd, res, buf, 1000, s1, s2);
        });
    }
    catch (std::invalid_argument const&) { threw = true; }
    EXPECT(threw, "radix_sort must throw std::invalid_argument (radix_sort.cpp:158-161)");
    std::printf("ok radix_sort\n");
}

// TEST(bucket_sort, main): keys equal std::sort on the masked key at every index, values a permutation
static void test_bucket_sort(vren::context& ctx)
{
    std::mt19937 rng(3);
    for (uint32_t n : {1u, 120u, 14400u, 1728000u})
    {
        std::vector<uint32_t> pairs((size_t) n * 2);
        for (uint32_t i = 0; i < n; i++) { pairs[2 * i] = rng() % 65536; pairs[2 * i + 1] = i; }
        auto in = vren::vk_utils::alloc_device_only_buffer(ctx, (size_t) n * 8);
        auto out = vren::vk_utils::alloc_device_only_buffer(ctx, vren::bucket_sort::get_required_output_buffer_size(n));
        upload(in, pairs);
        vren::vk_utils::immediate_graphics_queue_submit(ctx, [&](VkCommandBuffer cmd, vren::resource_container& res) {
            ctx.m_toolbox->m_bucket_sort(cmd, res, in, n, 0, out, 0);
        });
        auto got = download<uint32_t>(out, (size_t) n * 2);
        std::vector<uint32_t> keys(n), vals(n);
        for (uint32_t i = 0; i < n; i++) { keys[i] = pairs[2 * i] & 0xFFFF; }
        std::sort(keys.begin(), keys.end());
        for (uint32_t i = 0; i < n; i++)
        {
            EXPECT((got[2 * i] & 0xFFFF) == keys[i], "bucket sort key at %u (n=%u)", i, n);
            vals[i] = got[2 * i + 1];
        }
        std::sort(vals.begin(), vals.end());
        for (uint32_t i = 0; i < n; i++) EXPECT(vals[i] == i, "bucket sort values must be a permutation (n=%u)", n);
    }
    std::printf("ok bucket_sort\n");
}

static bool point_in(vren::bvh_node const& n, float const* p)
{
    return p[0] >= n.m_min[0] && p[1] >= n.m_min[1] && p[2] >= n.m_min[2] && p[0] <= n.m_max[0] && p[1] <= n.m_max[1] && p[2] <= n.m_max[2];
}
static void traverse_r(vren::bvh_node const* bvh, uint32_t offset, float const* p, std::vector<uint32_t>& out)
{
    for (uint32_t i = 0; i < 32; i++)
    {
        vren::bvh_node const& n = bvh[offset + i];
        if (n.is_invalid()) continue;
        if (point_in(n, p))
        {
            if (n.is_leaf()) out.push_back(offset + i);
            else traverse_r(bvh, n.m_next, p, out);
        }
    }
}

// TEST(build_bvh, main) + TEST(build_bvh, utils)
static void test_build_bvh(vren::context& ctx)
{
    EXPECT(vren::calc_bvh_padded_leaf_count(129) == 1024u && vren::calc_bvh_padded_leaf_count(2193819) == 33554432u, "padded KAT");
    EXPECT(vren::calc_bvh_buffer_length(1024) == 1057u && vren::calc_bvh_root_index(582) == 1056u, "length/root KAT");
    EXPECT(vren::calc_bvh_level_count(0) == 1u && vren::calc_bvh_level_count(582) == 2u && vren::calc_bvh_level_count(2193819) == 5u, "levels KAT");
    std::mt19937 rng(4);
    std::uniform_real_distribution<float> dist(0, 100);
    for (uint32_t leaf_count : {1u, 10u, 100u, 1000u, 10000u})
    {
        const uint32_t padded = vren::calc_bvh_padded_leaf_count(leaf_count), length = vren::calc_bvh_buffer_length(leaf_count);
        std::vector<vren::bvh_node> cpu(length);
        for (uint32_t i = 0; i < padded; i++)
        {
            if (i < leaf_count)
            {
                for (int c = 0; c < 3; c++)
                {
                    const float pos = dist(rng), ext = dist(rng);
                    cpu[i].m_min[c] = pos - ext;
                    cpu[i].m_max[c] = pos + ext;
                }
                cpu[i].m_next = vren::bvh_node::k_leaf_node;
            }
            else cpu[i].m_next = vren::bvh_node::k_invalid_node;
        }
        auto buf = vren::vk_utils::alloc_device_only_buffer(ctx, vren::build_bvh::get_required_buffer_size(leaf_count));
        upload(buf, cpu);
        vren::vk_utils::immediate_graphics_queue_submit(ctx, [&](VkCommandBuffer cmd, vren::resource_container& res) {
            ctx.m_toolbox->m_build_bvh(cmd, res, buf, padded);
        });
        auto gpu = download<vren::bvh_node>(buf, length);
        for (int q = 0; q < 8; q++)
        {
            const float p[3] = {dist(rng), dist(rng), dist(rng)};
            std::vector<uint32_t> lin, trav;
            for (uint32_t i = 0; i < padded; i++)
                if (!cpu[i].is_invalid() && point_in(cpu[i], p)) lin.push_back(i);
            vren::bvh_node const& root = gpu[length - 1];
            if (point_in(root, p)) traverse_r(gpu.data(), root.m_next, p, trav);
            std::sort(trav.begin(), trav.end());
            EXPECT(lin == trav, "bvh traversal leaves=%u query=%d (%zu vs %zu hits)", leaf_count, q, lin.size(), trav.size());
        }
    }
    std::printf("ok build_bvh\n");
}

// cluster_and_shade steps 1-3: structural checks on the outputs (the reference has no test for this pass)
static void test_cluster_and_shade(vren::context& ctx)
{
    const uint32_t W = 640, H = 360, L = 4096;
    vren::cluster_and_shade::limits lim;
    lim.max_screen_width = W; lim.max_screen_height = H; lim.max_point_light_count = L;
    vren::cluster_and_shade pass(ctx, lim);
    std::mt19937 rng(5);
    std::uniform_real_distribution<float> u(0, 1);
    std::vector<float> depth((size_t) W * H), pos((size_t) L * 4), lights((size_t) L * 4);
    for (auto& d : depth) d = u(rng) < 0.1f ? 1.0f : 0.9990f + 0.0009f * u(rng);
    for (uint32_t i = 0; i < L; i++)
    {
        const float z = 2 + 60 * u(rng);
        pos[4 * i] = (2 * u(rng) - 1) * z * 0.7f; pos[4 * i + 1] = (2 * u(rng) - 1) * z * 0.4f; pos[4 * i + 2] = z; pos[4 * i + 3] = 1;
        lights[4 * i + 3] = 0.5f + 2 * u(rng);
    }
    vren::light_array la;
    la.m_point_light_position_buffer = vren::vk_utils::alloc_device_only_buffer(ctx, pos.size() * 4);
    la.m_point_light_buffer = vren::vk_utils::alloc_device_only_buffer(ctx, lights.size() * 4);
    la.m_point_light_count = L;
    upload(la.m_point_light_position_buffer, pos);
    upload(la.m_point_light_buffer, lights);
    vren::vk_utils::depth_buffer_t db{vren::vk_utils::alloc_device_only_buffer(ctx, depth.size() * 4)};
    upload(db.m_image, depth);
    vren::gbuffer gb;
    gb.m_width = W; gb.m_height = H;
    vren::camera cam;
    cam.m_aspect_ratio = (float) W / H;
    vren::vk_utils::immediate_graphics_queue_submit(ctx, [&](VkCommandBuffer cmd, vren::resource_container& res) {
        pass(cmd, res, vren::uvec2{W, H}, cam, gb, db, la);
    });
    auto disp = download<uint32_t>(pass.m_cluster_key_dispatch_params_buffer, 4);
    auto status = download<uint32_t>(pass.m_status_buffer, 4);
    EXPECT(disp[0] > 0 && disp[1] == 1 && disp[2] == 1 && disp[3] == 0, "dispatch params {%u,%u,%u,%u}", disp[0], disp[1], disp[2], disp[3]);
    auto keys = download<uint32_t>(pass.m_cluster_key_buffer, disp[0]);
    auto ref = download<uint32_t>(pass.m_cluster_reference_buffer.m_image, (size_t) W * H);
    for (uint32_t y = 0; y < H; y += 7)
        for (uint32_t x = 0; x < W; x += 5)
        {
            const uint32_t k = keys[ref[(size_t) y * W + x]];
            EXPECT((k & 0xFF) == (x >> 5) && ((k >> 8) & 0xFF) == (y >> 5), "cluster reference of pixel (%u,%u)", x, y);
        }
    auto counts = download<uint32_t>(pass.m_assigned_light_counts_buffer, disp[0]);
    auto offsets = download<uint32_t>(pass.m_assigned_light_offsets_buffer, disp[0]);
    uint64_t total = 0;
    for (uint32_t c = 0; c < disp[0]; c++) { EXPECT(offsets[c] == total, "offsets are the exclusive scan of counts at %u", c); total += counts[c]; }
    EXPECT(total == status[0] && status[1] == 0 && total > 0, "assigned total %llu vs status %u", (unsigned long long) total, status[0]);
    auto indices = download<uint32_t>(pass.m_assigned_light_indices_buffer, total);
    for (uint32_t v : indices) EXPECT(v < L, "light index out of range");
    std::printf("ok cluster_and_shade (%u clusters, %llu assignments)\n", disp[0], (unsigned long long) total);
}

// depth pyramid: every level is the 2x2 max of the level below (depth_buffer_reduce.comp:20-31)
static void test_depth_pyramid(vren::context& ctx)
{
    const uint32_t W = 333, H = 200;
    std::mt19937 rng(6);
    std::uniform_real_distribution<float> u(0, 1);
    std::vector<float> depth((size_t) W * H);
    for (auto& d : depth) d = u(rng);
    vren::vk_utils::depth_buffer_t db{vren::vk_utils::alloc_device_only_buffer(ctx, depth.size() * 4)};
    upload(db.m_image, depth);
    vren::depth_buffer_pyramid pyramid(ctx, W, H);
    vren::depth_buffer_reductor reductor(ctx);
    vren::vk_utils::immediate_graphics_queue_submit(ctx, [&](VkCommandBuffer cmd, vren::resource_container&) {
        reductor.copy_and_reduce(cmd, db, pyramid);
    });
    EXPECT(pyramid.get_level_count() == 9, "level count %u", pyramid.get_level_count());
    auto all = download<float>(pyramid.m_image, pyramid.m_image.m_size / 4);
    std::vector<float> prev = depth;
    uint32_t pw = W, ph = H;
    size_t off = (size_t) W * H;
    EXPECT(std::equal(depth.begin(), depth.end(), all.begin()), "level 0 must be a copy");
    for (uint32_t l = 1; l < pyramid.get_level_count(); l++)
    {
        const uint32_t w = pyramid.get_image_width(l), h = pyramid.get_image_height(l);
        std::vector<float> cur((size_t) w * h);
        for (uint32_t y = 0; y < h; y++)
            for (uint32_t x = 0; x < w; x++)
            {
                float m = 0.0f;
                for (uint32_t dx = 0; dx < 2; dx++)
                    for (uint32_t dy = 0; dy < 2; dy++)
                        if (2 * x + dx < pw && 2 * y + dy < ph) m = std::max(m, prev[(size_t) (2 * y + dy) * pw + 2 * x + dx]);
                cur[(size_t) y * w + x] = m;
                EXPECT(all[off + (size_t) y * w + x] == m, "pyramid level %u texel (%u,%u)", l, x, y);
            }
        off += cur.size();
        prev.swap(cur);
        pw = w; ph = h;
    }
    std::printf("ok depth_buffer_pyramid\n");
}

static void test_point_light_bouncer(vren::context& ctx)
{
    // a light flying along +x in the unit box: 0.25 to the face, reflected, 0.25 back (minus the shader's EPS)
    std::vector<float> pos = { 0.75f, 0.5f, 0.5f, 7.0f, 0.1f, 0.2f, 0.3f, 7.0f }, dir = { 1, 0, 0, 7.0f, 0, 0, 0, 7.0f };
    auto pb = vren::vk_utils::alloc_device_only_buffer(ctx, pos.size() * 4);
    auto dbuf = vren::vk_utils::alloc_device_only_buffer(ctx, dir.size() * 4);
    upload(pb, pos);
    upload(dbuf, dir);
    vren_demo::point_light_bouncer bouncer(ctx);
    vren::vk_utils::immediate_graphics_queue_submit(ctx, [&](VkCommandBuffer cmd, vren::resource_container& rc) {
        bouncer.bounce(0, cmd, rc, pb, dbuf, 2, vren_demo::vec3{0, 0, 0}, vren_demo::vec3{1, 1, 1}, 1.0f, 0.5f);
    });
    auto p = download<float>(pb, 8), d = download<float>(dbuf, 8);
    EXPECT(d[0] == -1.0f && d[3] == 0.0f && p[3] == 1.0f, "reflection: d.x %f d.w %f p.w %f", d[0], d[3], p[3]);
    EXPECT(std::fabs(p[0] - 0.75f) < 1e-3f && p[1] == 0.5f && p[2] == 0.5f, "position after the bounce: %f %f %f", p[0], p[1], p[2]);
    EXPECT(p[4] == 0.1f && p[5] == 0.2f && p[6] == 0.3f, "a light without direction stays where it is");
    std::printf("ok point_light_bouncer\n");
}

static void test_visualize_bvh(vren::context& ctx)
{
    // 32 leaves (one level above the root): boxes [i, i+1]^3; the root box must span [0, 32]^3 and carry the second colour
    const uint32_t leaves = 32, level_count = vren::calc_bvh_level_count(leaves);
    std::vector<vren::bvh_node> nodes(vren::calc_bvh_buffer_length(leaves));
    for (uint32_t i = 0; i < leaves; i++)
    {
        for (int c = 0; c < 3; c++)
        {
            nodes[i].m_min[c] = (float) i;
            nodes[i].m_max[c] = (float) i + 1;
        }
        nodes[i].m_next = vren::bvh_node::k_leaf_node;
    }
    auto bvh = vren::vk_utils::alloc_device_only_buffer(ctx, nodes.size() * sizeof(vren::bvh_node));
    upload(bvh, nodes);
    vren_demo::visualize_bvh vis(ctx);
    auto draw = vren::vk_utils::alloc_device_only_buffer(ctx, vren_demo::visualize_bvh::get_required_vertex_buffer_size(level_count));
    vren::vk_utils::immediate_graphics_queue_submit(ctx, [&](VkCommandBuffer cmd, vren::resource_container& rc) {
        ctx.m_toolbox->m_build_bvh(cmd, rc, bvh, leaves);
        vis.write(cmd, bvh, level_count, draw);
    });
    auto v = download<vren_demo::debug_draw_vertex>(draw, 33 * 24);
    EXPECT(level_count == 1, "level count %u", level_count);
    EXPECT(v[0].color == 0x00ff00u && v[32 * 24].color == 0x0000ffu, "level colours %06x %06x", v[0].color, v[32 * 24].color);
    float lo = 1e30f, hi = -1e30f;
    for (int k = 0; k < 24; k++)
        for (int a = 0; a < 3; a++)
        {
            lo = std::min(lo, v[32 * 24 + k].position[a]);
            hi = std::max(hi, v[32 * 24 + k].position[a]);
        }
    EXPECT(lo == 0.0f && hi == 32.0f, "root box %f %f", lo, hi);
    EXPECT(v[5 * 24].position[0] == 5.0f && v[5 * 24 + 1].position[0] == 6.0f, "first edge of leaf 5: %f -> %f", v[5 * 24].position[0], v[5 * 24 + 1].position[0]);
    std::printf("ok visualize_bvh\n");
}

// TEST(KDTree, NearestNeigborSearch), vren_test/vren_test/kd_tree.cpp:29: tree search == linear search
static void test_kd_tree()
{
    std::mt19937 rng(11);
    std::uniform_real_distribution<float> u(-50, 50);
    const size_t n = 100000;
    std::vector<float> pts(n * 3);
    for (auto& v : pts) v = u(rng);
    std::vector<uint32_t> indices(n);
    std::iota(indices.begin(), indices.end(), 0u);
    std::vector<vren::kd_tree_node> tree(2 * n);
    const size_t nodes = vren::kd_tree_build(pts.data(), 3, indices.data(), n, tree.data(), 0, 32);
    EXPECT(nodes > 0 && nodes <= 2 * n, "node count %zu", nodes);
    for (int q = 0; q < 100; q++)
    {
        const float s[3] = {u(rng), u(rng), u(rng)};
        uint32_t best = ~0u, lin = ~0u;
        float best_d2 = INFINITY, lin_d2 = INFINITY;
        vren::kd_tree_search(pts.data(), 3, indices.data(), n, tree.data(), 0, s, vren::k_kd_tree_default_search_filter, best, best_d2);
        for (uint32_t i = 0; i < n; i++)
        {
            const float dx = s[0] - pts[3 * i], dy = s[1] - pts[3 * i + 1], dz = s[2] - pts[3 * i + 2];
            const float d2 = dx * dx + dy * dy + dz * dz;
            if (d2 < lin_d2) { lin_d2 = d2; lin = i; }
        }
        EXPECT(best == lin, "kd-tree query %d: %u vs %u", q, best, lin);
    }
    std::printf("ok kd_tree\n");
}

int main()
{
    try
    {
        test_kd_tree();
        vren::context ctx(0);
        test_reduce(ctx);
        test_scan(ctx);
        test_radix_sort(ctx);
        test_bucket_sort(ctx);
        test_build_bvh(ctx);
        test_cluster_and_shade(ctx);
        test_depth_pyramid(ctx);
        test_point_light_bouncer(ctx);
        test_visualize_bvh(ctx);
    }
    catch (std::exception const& e)
    {
        std::printf("FAIL exception: %s\n", e.what());
        return 2;
    }
    std::printf(g_failures == 0 ? "ALL PASS\n" : "FAILURES: %d\n", g_failures);
    return g_failures == 0 ? 0 : 1;
}
