/*
 * vrenb200.h — C ABI of the B200-native replacement for vren's data-parallel compute core.
 *
 * Every entry point is what a binding from the reference (loryruta/vren, C++/Vulkan) would call in
 * place of recording Vulkan dispatches.  All `*_run`-style calls are:
 *   - asynchronous on the given CUDA stream (the analogue of "record into a VkCommandBuffer"),
 *   - allocation-free and host-sync-free (caller owns inputs, outputs and scratch),
 *   - re-entrant: no global mutable state; how a call runs is an argument of the call (vrenb200_sort_config).
 * Pointers are raw device pointers unless the name ends in `_host`.  Lengths are ELEMENTS,
 * sizes are BYTES.  Return value: 0 on success, one of VRENB200_E* otherwise.
 *
 * Reference interface each group replaces (paths relative to the reference checkout):
 *   helpers   vren/vren/base/base.hpp:32-79
 *   reduce    vren/vren/primitives/reduce.hpp:28-46, reduce.cpp:33-128, shaders/reduce.comp:50-87
 *   scan      vren/vren/primitives/blelloch_scan.hpp:31-48, blelloch_scan.cpp:57-166
 *   radix     vren/vren/primitives/radix_sort.hpp:42-52, radix_sort.cpp:124-337
 *   bucket    vren/vren/primitives/bucket_sort.hpp:33-44, bucket_sort.cpp:62-161
 *   bvh       vren/vren/primitives/build_bvh.hpp:8-56, build_bvh.cpp:38-136
 *   lights    vren/vren/pipeline/clustered_shading.hpp:19-108, clustered_shading.cpp:29-690
 */
#ifndef VRENB200_H_
#define VRENB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cudaStream_t without dragging cuda_runtime.h into C callers */
typedef struct CUstream_st* vrenb200_stream_t;

enum {
    VRENB200_OK = 0,
    VRENB200_EINVAL_LENGTH = 1,   /* length precondition violated (e.g. radix_sort compat: n>=1024 && pow2) */
    VRENB200_EALIGN = 2,          /* pointer / offset alignment violated */
    VRENB200_ESCRATCH = 3,        /* scratch buffer missing or too small */
    VRENB200_ECUDA = 4,           /* a CUDA runtime call failed; see vrenb200_last_cuda_error() */
    VRENB200_EINVAL_ARG = 5,      /* null pointer / bad enum */
    VRENB200_ELIMIT = 6           /* exceeds an implementation limit (documented per call) */
};

/* element types of vren::reduce<T,op> (reduce.cpp:130-135) + scalar f32 as an extension */
enum { VRENB200_U32 = 0, VRENB200_VEC4 = 1, VRENB200_F32 = 2 };
/* vren::reduce_operation (reduce.hpp:8-13) */
enum { VRENB200_ADD = 0, VRENB200_MIN = 1, VRENB200_MAX = 2 };
/* output layout of reduce: TREE = full Blelloch up-sweep tree in out[0..P) (what the reference writes,
 * reduce.comp:83-86); FINAL = only out[P-1] is written (what every in-repo caller consumes) */
enum { VRENB200_REDUCE_TREE = 0, VRENB200_REDUCE_FINAL = 1 };

const char* vrenb200_version(void);
const char* vrenb200_status_string(int status);
/* last cudaError_t seen by this thread inside the library (0 if none) */
int vrenb200_last_cuda_error(void);

/* ---- a10: integer helpers (base/base.hpp:32-79), restated in pure integer arithmetic ---------- */
int      vrenb200_is_power_of_2(uint32_t v);
uint32_t vrenb200_round_to_next_power_of_2(uint32_t v);
uint64_t vrenb200_round_to_next_multiple_of(uint64_t v, uint64_t multiple);
uint32_t vrenb200_divide_and_ceil(uint32_t v, uint32_t d);
int      vrenb200_is_power_of(uint32_t n, uint32_t base);
uint32_t vrenb200_round_to_next_power_of(uint32_t n, uint32_t base);

/* ---- a1: reduce ------------------------------------------------------------------------------- */
/* calc_reduce_output_buffer_length (reduce.cpp:137-140) */
uint32_t vrenb200_calc_reduce_output_buffer_length(uint32_t count);
/* scratch needed by FINAL mode (TREE mode needs none: partials live in the tree itself) */
size_t vrenb200_reduce_scratch_bytes(int dtype, int mode, uint32_t n, uint32_t blocks);
/* rows: in + y*n, out + y*P  (reduce.comp:52-53); `out` may alias `in` (reduce.cpp:117-128) */
int vrenb200_reduce(vrenb200_stream_t stream, int dtype, int op, int mode,
                    const void* in, uint32_t n, void* out, uint32_t blocks,
                    void* scratch, size_t scratch_bytes);

/* ---- a2: exclusive add-scan of uint32 ---------------------------------------------------------- */
size_t vrenb200_scan_scratch_bytes(uint32_t n);
/* in-place when out==in; any n>=1 (reference requires pow2: blelloch_scan.cpp:67) */
int vrenb200_exclusive_scan_u32(vrenb200_stream_t stream, const uint32_t* in, uint32_t* out, uint32_t n,
                                void* scratch, size_t scratch_bytes);
/* out[i] = base + sum_{k<i} in[k]: local step of the sharded scan */
int vrenb200_exclusive_scan_u32_base(vrenb200_stream_t stream, const uint32_t* in, uint32_t* out, uint32_t n,
                                     uint32_t base, void* scratch, size_t scratch_bytes);
/* the same with flags.  VRENB200_SCAN_TILE_IDS_TICKET ("safe mode"): tile ids come from an atomic ticket taken when a CTA starts
 * instead of the block index, so the look-back of the chained scan never waits for a tile whose CTA has not started — no
 * assumption about the order in which the hardware dispatches CTAs (MPS, time slicing, debuggers).  It uses the chained
 * register-tile kernel at every size (the faster run-ahead kernel for n >= 2^24 ties its roles to the block index).
 * VRENB200_SCAN_TILE_IDS=ticket in the environment makes it the default of the process.  Opt-in: written after the round's
 * GPU time was spent; the kernel text runs on the host CTA emulator (tests/test_scan_emulation.py), first hardware run =
 * tests/test_scan_safe_mode.py. */
#define VRENB200_SCAN_TILE_IDS_TICKET 1u
int vrenb200_exclusive_scan_u32_ex(vrenb200_stream_t stream, const uint32_t* in, uint32_t* out, uint32_t n,
                                   uint32_t base, void* scratch, size_t scratch_bytes, uint32_t flags);
/* blelloch_scan::downsweep (blelloch_scan.cpp:57-139): turns `blocks` up-sweep trees of pow2 length n
 * into exclusive scans (+root when clear_last==0), in place */
int vrenb200_blelloch_downsweep_u32(vrenb200_stream_t stream, uint32_t* buf, uint32_t n, uint32_t blocks,
                                    int clear_last);

/* ---- a3: radix sort ---------------------------------------------------------------------------- */
/* How a sort call runs (all fields 0 = the library's defaults; pass NULL for the same).  Part of the call, not of the process:
 * there is no global selection state.  The defaults can be overridden through the environment, read once:
 * VRENB200_SORT_RANKING=match|verified|sampled|atomic, VRENB200_SORT_TILE_IDS=block|ticket. */
enum {
    VRENB200_RANKING_AUTO = 0,               /* = ATOMIC_SAMPLED */
    VRENB200_RANKING_MATCH = 1,              /* warp-ballot digit match: order by construction */
    VRENB200_RANKING_ATOMIC_VERIFIED = 2,    /* one returning shared atomic per pair; EVERY row is checked against the ballot match
                                                and a pass that fails the check is repeated by the match kernel before the next
                                                pass starts: correct whatever order the hardware serves same-address lanes in (PTX
                                                does not specify it); slower than MATCH on B200, kept for validation */
    VRENB200_RANKING_ATOMIC_SAMPLED = 3,     /* same with one row in eight checked (the checks run in production, on every sort:
                                                ~10^5 checked rows with lane collisions per 2^28-pair pass): 1.5 % slower than
                                                unchecked; a device that does not serve the lanes in order is caught in the
                                                first tile and every such pass is redone by construction */
    VRENB200_RANKING_ATOMIC_UNVERIFIED = 4,  /* no check: relies on ascending lane order of same-address shared atomics */
    VRENB200_RANKING_SELFTEST_REDO = 5       /* tests: the verified kernels report a failed check and write nothing, so the
                                                result is what the repeat pass alone produces */
};
enum {
    VRENB200_TILE_IDS_AUTO = 0,              /* = TICKET (1.3 % slower than BLOCK_INDEX at 2^28 pairs) */
    VRENB200_TILE_IDS_BLOCK_INDEX = 1,       /* tile = block index: assumes CTAs of a 1-D grid start in index order (as CUB's
                                                decoupled-look-back scan does) */
    VRENB200_TILE_IDS_TICKET = 2             /* tile = value of an atomic counter taken when the CTA starts: forward progress of
                                                the look-back without any assumption about the dispatch order */
};
typedef struct vrenb200_sort_config {
    int ranking;
    int tile_ids;
    int variant;     /* 0 = automatic; > 0: 1-based index into the kernel table (vrenb200_radix_sort_variant_name);
                        VRENB200_SORT_VARIANT_SINGLE_CTA: see below */
} vrenb200_sort_config;
/* vrenb200_radix_sort_ex only, n <= 8192: the whole sort (all four passes) in ONE launch of ONE CTA, in shared memory — the
 * reference's own test size (1024 keys, vren_test radix_sort.cpp:130-143) without the 7 launches of the tiled path.  Opt-in:
 * written after the round's GPU time was spent, its body has run on the host CTA emulator only (tests/cpp/cta_emulator.hpp),
 * first hardware run = tests/test_small_sort.py.  Larger n, or a profiled call, silently takes the tiled path. */
#define VRENB200_SORT_VARIANT_SINGLE_CTA (-1)

/* one scratch blob holds the ping-pong buffers, digit histograms and look-back state. n < 2^30 */
size_t vrenb200_radix_sort_scratch_bytes(uint32_t n, int with_values);
/* ascending sort of uint32 keys, result in `keys` (radix_sort.cpp:149-337 semantics, any n>=0) */
int vrenb200_radix_sort_keys(vrenb200_stream_t stream, uint32_t* keys, uint32_t n,
                             void* scratch, size_t scratch_bytes);
/* key-value extension: stable ascending by key, result in keys/values */
int vrenb200_radix_sort_pairs(vrenb200_stream_t stream, uint32_t* keys, uint32_t* values, uint32_t n,
                              void* scratch, size_t scratch_bytes);
typedef struct vrenb200_sort_profile vrenb200_sort_profile;   /* CUDA events around every kernel of one sort */
/* the general form: values may be NULL (keys only), cfg may be NULL (defaults), prof may be NULL */
int vrenb200_radix_sort_ex(vrenb200_stream_t stream, uint32_t* keys, uint32_t* values, uint32_t n,
                           void* scratch, size_t scratch_bytes, const vrenb200_sort_config* cfg, vrenb200_sort_profile* prof);
/* device word inside `scratch`, != 0 after a sort in which the ranking check failed (and the repeat pass ran); read it after
 * the stream has been synchronised */
const uint32_t* vrenb200_radix_sort_violation_word(void* scratch, uint32_t n, int with_values);
/* sizes of the two scratch buffers of the reference-shaped call (radix_sort.cpp:124-147 are SMALLER;
 * the facade's create_scratch_buffer_1/2 use these) */
size_t vrenb200_radix_sort_scratch_buffer_1_bytes(uint32_t n);
size_t vrenb200_radix_sort_scratch_buffer_2_bytes(uint32_t n);
/* reference-shaped call: two separate scratch buffers, enforces n>=1024 && pow2 (radix_sort.cpp:158-161) */
int vrenb200_radix_sort_compat(vrenb200_stream_t stream, uint32_t* keys, uint32_t n,
                               void* scratch_1, size_t scratch_1_bytes,
                               void* scratch_2, size_t scratch_2_bytes);
/* end-to-end variant with HOST buffers: H2D, sort, D2H on `stream`, then stream sync.
 * `dev_work` must hold 8n (keys-only: 4n) bytes + scratch; used by bench.py's e2e leg */
size_t vrenb200_radix_sort_host_work_bytes(uint32_t n, int with_values);
int vrenb200_radix_sort_pairs_host(vrenb200_stream_t stream, uint32_t* keys_host, uint32_t* values_host,
                                   uint32_t n, void* dev_work, size_t dev_work_bytes);
/* enqueue-only form: separate source / destination host buffers (equal pointers = in place), no synchronisation.  Calls on
 * different streams, each with its own dev_work, overlap one call's upload with another's download (pinned memory). */
int vrenb200_radix_sort_pairs_host_async(vrenb200_stream_t stream, const uint32_t* keys_in_host, const uint32_t* values_in_host,
                                         uint32_t* keys_out_host, uint32_t* values_out_host, uint32_t n,
                                         void* dev_work, size_t dev_work_bytes);

/* building blocks of the NCCL form of the multi-GPU sort: all four 256-bin digit histograms of the keys
 * (hist_out: device uint32[4][256]) and a stable sort restricted to the digits [first_pass, first_pass+num_passes) */
int vrenb200_radix_digit_histograms(vrenb200_stream_t stream, const uint32_t* keys, uint32_t n, uint32_t* hist_out);
size_t vrenb200_radix_sort_range_scratch_bytes(uint32_t n);
int vrenb200_radix_sort_pairs_range(vrenb200_stream_t stream, uint32_t* keys, uint32_t* values, uint32_t* alt_keys,
                                    uint32_t* alt_values, uint32_t n, int first_pass, int num_passes,
                                    void* scratch, size_t scratch_bytes, int* result_in_alt);

/* reporting: the kernel configuration a sort of n elements would use, and the table behind vrenb200_sort_config::variant */
const char* vrenb200_radix_sort_selected_variant_name(uint32_t n, int with_values, const vrenb200_sort_config* cfg);
int vrenb200_radix_sort_num_variants(void);
const char* vrenb200_radix_sort_variant_name(int variant);
vrenb200_sort_profile* vrenb200_sort_profile_create(void);
void vrenb200_sort_profile_destroy(vrenb200_sort_profile* p);
/* ms_out[6] = {histogram, histogram-scan, pass0..pass3}, read after stream sync */
int vrenb200_sort_profile_read(vrenb200_sort_profile* p, float* ms_out);

/* ---- e1: the multi-GPU form of the key-value radix sort (single-device in the reference: radix_sort.cpp:149-337) ----
 * One context per rank.  Every rank owns a SYMMETRIC region (vrenb200_sharded_sort_symmetric_bytes) that all ranks can address:
 * peer-mapped device memory (torch symmetric memory, cudaIpc handles, or cudaDeviceEnablePeerAccess inside one process).  The
 * call is collective — every rank calls it once per sort with its shard — and, like every other entry point, only enqueues:
 * no NCCL, no host synchronisation, no allocation.  Concatenating the ranks' outputs in rank order gives the stable sort by key
 * of the concatenation of their inputs.  See vren_b200/csrc/sharded_sort.cu for the pipeline.
 *   capacity    pairs a rank can receive (its share after the exchange, padded to whole tiles: about n_total / world * 1.25
 *               + 256 tiles for balanced keys); a plan that does not fit sets status[0] and leaves the output untouched
 *   rounds      1..8: how many pieces the exchange is cut into; the transfer of a piece overlaps the sorting of the one before
 *   peer_regions[r]  base of rank r's symmetric region as addressable from THIS rank (256-byte aligned), r = 0..world-1
 *   local       device scratch of this rank, vrenb200_sharded_sort_local_bytes(max_n, capacity, cfg) bytes, 256-byte aligned
 *               (the same cfg as create() gets: the tile of the passes depends on it)
 * create() zeroes this rank's flag words with a synchronous memset: call it on every rank, then synchronise the ranks once
 * (any barrier) before the first sort.  key_bits (32, 24, 16 or 8): how many low bits of the key take part (16 = the
 * bucket-sort key of bucket_sort.hpp:15-16; the whole 32-bit word is carried).
 * status (device uint32[8], valid once the stream has finished the call): {error, pairs in the output, first and one-past-last
 * value of the partition digit this rank owns, index of the partition digit}; error bit 0: a rank's share exceeds the capacity,
 * bit 1: a round exceeds its launch bound (retry with rounds = 1 or a larger capacity). */
typedef struct vrenb200_sharded_sort vrenb200_sharded_sort;
size_t vrenb200_sharded_sort_symmetric_bytes(uint32_t capacity);
size_t vrenb200_sharded_sort_local_bytes(uint32_t max_n, uint32_t capacity, const vrenb200_sort_config* cfg);
int vrenb200_sharded_sort_create(vrenb200_sharded_sort** out, uint32_t rank, uint32_t world, uint32_t max_n, uint32_t capacity,
                                 uint32_t rounds, void* const* peer_regions, void* local, size_t local_bytes,
                                 const vrenb200_sort_config* cfg);
void vrenb200_sharded_sort_destroy(vrenb200_sharded_sort* ctx);
int vrenb200_sharded_sort_pairs(vrenb200_sharded_sort* ctx, vrenb200_stream_t stream, const uint32_t* keys, const uint32_t* values,
                                uint32_t n, int key_bits);
const uint32_t* vrenb200_sharded_sort_out_keys(const vrenb200_sharded_sort* ctx);
const uint32_t* vrenb200_sharded_sort_out_values(const vrenb200_sharded_sort* ctx);
const uint32_t* vrenb200_sharded_sort_status(const vrenb200_sharded_sort* ctx);
/* phases of the last call from CUDA events (after the stream has been synchronised): ms_out[3] = {histograms + plan + local
 * partition, the transfers, the segmented passes}; the last two start together and overlap */
int vrenb200_sharded_sort_phases(const vrenb200_sharded_sort* ctx, float* ms_out);

#ifdef VRENB200_TUNING
/* tuning builds only (VRENB200_TUNING=1 python -m vren_b200.build): process-global selection of experimental scan kernels.
 * The release library exports nothing that mutates process-wide state. */
int vrenb200_scan_set_variant(int variant);
int vrenb200_scan_set_runahead(int finalize_lag_tiles, int scan_lag_tiles);
#endif

/* ---- a4: bucket sort (16-bit key counting sort of uvec2) --------------------------------------- */
/* bucket_sort::get_required_output_buffer_size (bucket_sort.cpp:67-70) */
size_t vrenb200_bucket_sort_output_bytes(uint32_t n);
size_t vrenb200_bucket_sort_scratch_bytes(uint32_t n);
/* out: sorted uvec2[n], then at round_up(8n,256) 65536 counters holding bucket END offsets
 * (bucket_sort.cpp:86, bucket_sort_write.comp:32). Ties keep input order (canonical choice). */
int vrenb200_bucket_sort(vrenb200_stream_t stream, const void* in_pairs, uint32_t n, void* out,
                         void* scratch, size_t scratch_bytes);
/* general form.  end_offsets: -1 automatic (from 2^20 pairs on the END offsets are found by a search in the sorted output instead
 * of per-key global atomics), 0 atomics, 1 search; results are identical either way.  cfg as for the radix sort (may be NULL). */
int vrenb200_bucket_sort_ex(vrenb200_stream_t stream, const void* in_pairs, uint32_t n, void* out,
                            void* scratch, size_t scratch_bytes, const vrenb200_sort_config* cfg, int end_offsets);

/* ---- a5: 32-ary implicit BVH ------------------------------------------------------------------- */
typedef struct vrenb200_bvh_node {   /* == vren::bvh_node (build_bvh.hpp:8-18), 32 bytes */
    float min[3]; uint32_t next;
    float max[3]; uint32_t _pad;
} vrenb200_bvh_node;
#define VRENB200_BVH_LEAF_NODE    0xFFFFFFFFu
#define VRENB200_BVH_INVALID_NODE 0xFFFFFFFEu

uint32_t vrenb200_calc_bvh_padded_leaf_count(uint32_t leaf_count);  /* build_bvh.cpp:99-104 */
uint32_t vrenb200_calc_bvh_buffer_length(uint32_t leaf_count);      /* build_bvh.cpp:106-118 */
size_t   vrenb200_calc_bvh_buffer_size(uint32_t leaf_count);        /* build_bvh.cpp:120-124 */
uint32_t vrenb200_calc_bvh_root_index(uint32_t leaf_count);         /* build_bvh.cpp:126-130 */
uint32_t vrenb200_calc_bvh_level_count(uint32_t leaf_count);        /* build_bvh.cpp:132-136 */
/* leaves pre-filled in nodes[0..padded); builds every upper level (build_bvh.cpp:38-97) */
int vrenb200_build_bvh(vrenb200_stream_t stream, vrenb200_bvh_node* nodes, uint32_t padded_leaf_count);

/* ---- a6: construct_point_light_bvh ------------------------------------------------------------- */
size_t vrenb200_light_bvh_buffer_bytes(uint32_t light_count);        /* clustered_shading.cpp:34-47 */
size_t vrenb200_light_index_buffer_bytes(uint32_t light_count);      /* clustered_shading.cpp:54-58 (+ bucket-sort counters) */
size_t vrenb200_light_bvh_scratch_bytes(uint32_t light_count);
/* positions: vec4[L] world space; lights: point_light{vec3 color; float intensity}[L];
 * view: column-major mat4 (glm layout). Outputs: view_pos vec4[L]; bvh nodes at offset 0 of bvh_buffer;
 * sorted uvec2{morton, light_idx}[L] at offset 0 of index_buffer (clustered_shading.cpp:60-346) */
int vrenb200_construct_point_light_bvh(vrenb200_stream_t stream,
                                       const float* positions, const float* lights, uint32_t light_count,
                                       const float* view_col_major,
                                       float* view_pos, void* bvh_buffer, void* index_buffer,
                                       void* scratch, size_t scratch_bytes);

/* ---- a7/a8: cluster keys and light assignment --------------------------------------------------- */
typedef struct vrenb200_camera {    /* vren::camera projection inputs (camera.hpp:14-23, camera.cpp:40-50) */
    float fov_y, aspect_ratio, near_plane, far_plane;
} vrenb200_camera;

typedef struct vrenb200_cluster_limits {   /* config.hpp:23-24 made runtime */
    uint32_t max_unique_cluster_keys;      /* default 1<<17 */
    uint32_t max_assigned_lights;          /* default 1<<23 */
} vrenb200_cluster_limits;

size_t vrenb200_find_unique_clusters_scratch_bytes(uint32_t width, uint32_t height);
/* depth: float[W*H] (D32 aspect); normals: half4[W*H] (RGBA16F) or NULL (=all-zero normals).
 * keys_out uint[max_keys]; dispatch_params uvec4 {count,1,1,overflow_flag}; cluster_ref uint[W*H]
 * (find_unique_clusters.comp:46-121). Emits the canonical tile-major order. */
int vrenb200_find_unique_clusters(vrenb200_stream_t stream,
                                  const float* depth, const void* normals_rgba16f,
                                  uint32_t width, uint32_t height, const vrenb200_camera* camera,
                                  uint32_t* keys_out, uint32_t max_keys, uint32_t* dispatch_params,
                                  uint32_t* cluster_ref,
                                  void* scratch, size_t scratch_bytes);

size_t vrenb200_assign_lights_scratch_bytes(uint32_t max_keys, uint32_t max_assigned);
/* counts/offsets: uint[max_keys]; indices: uint[max_assigned] (assign_lights.comp:121-241,
 * clustered_shading.cpp:473-690). status_out (device uint[4], may be NULL):
 * {total_assigned, overflow_flag, node_visits, leaf_tests} */
int vrenb200_assign_lights(vrenb200_stream_t stream,
                           uint32_t width, uint32_t height, const vrenb200_camera* camera,
                           const uint32_t* cluster_keys, const uint32_t* dispatch_params, uint32_t max_keys,
                           const void* bvh_buffer, uint32_t bvh_root_index, uint32_t light_count,
                           const void* light_index_buffer, const float* view_pos,
                           uint32_t* indices_out, uint32_t max_assigned,
                           uint32_t* counts_out, uint32_t* offsets_out, uint32_t* status_out,
                           void* scratch, size_t scratch_bytes);

/* Diagnostics of a8, one element per thread: the device functions the light-assignment walk is made of, for the parity tests
 * pinned to the reference's own GLSL (tests/test_clustered_reference.py).  Per cluster key i: out_min3/out_max3[i] = the
 * cluster corners of clustered_shading.glsl:72-110 (both NULL to skip); out_flags[i] bit 0 = test_aabb_aabb against
 * node_boxes6[i] = {min.xyz, max.xyz}, bit 1 = test_sphere_aabb against spheres4[i] = {centre.xyz, radius}
 * (assign_lights.comp:84-102); either input and out_flags may be NULL. */
int vrenb200_cluster_tests(vrenb200_stream_t stream, uint32_t width, uint32_t height, const vrenb200_camera* camera,
                           const uint32_t* cluster_keys, uint32_t count, float* out_min3, float* out_max3,
                           const float* node_boxes6, const float* spheres4, uint8_t* out_flags);

/* ---- n1: consumer side of the light lists (shade.comp:101-105, vren_demo show_clusters.comp:97-118) -------- */
/* out_count_hash: uvec2[W*H] = {assigned_light_counts[cluster_reference(x,y)], XOR of that cluster's light indices} */
size_t vrenb200_light_list_hash_scratch_bytes(uint32_t max_keys);
int vrenb200_light_list_hash(vrenb200_stream_t stream, uint32_t width, uint32_t height, const uint32_t* cluster_ref,
                             const uint32_t* dispatch_params, uint32_t max_keys, const uint32_t* counts,
                             const uint32_t* offsets, const uint32_t* indices, uint32_t max_assigned,
                             uint32_t* out_count_hash, void* scratch, size_t scratch_bytes);

/* ---- n2: depth-buffer pyramid (pipeline/depth_buffer_pyramid.{hpp,cpp}, depth_buffer_{copy,reduce}.comp) ----- */
/* level_count = floor(log2(max(W,H))) + 1 (depth_buffer_pyramid.cpp:18); level l is max(W>>l,1) x max(H>>l,1) */
uint32_t vrenb200_depth_pyramid_level_count(uint32_t width, uint32_t height);
uint32_t vrenb200_depth_pyramid_level_width(uint32_t width, uint32_t level);
uint32_t vrenb200_depth_pyramid_level_height(uint32_t height, uint32_t level);
/* the mip chain is one flat float buffer, levels back to back: element offset of a level, total size in bytes */
size_t vrenb200_depth_pyramid_level_offset(uint32_t width, uint32_t height, uint32_t level);
size_t vrenb200_depth_pyramid_bytes(uint32_t width, uint32_t height);
/* depth_buffer_reductor::copy_and_reduce (depth_buffer_pyramid.cpp:177-305): level 0 = copy, level l+1 = 2x2 max */
int vrenb200_depth_pyramid_build(vrenb200_stream_t stream, const float* depth, uint32_t width, uint32_t height, float* pyramid);

/* ---- n3: light animation producer ----------------------------------------------------------------------------------
 * vren_demo::point_light_bouncer::bounce (vren_demo/vren_demo/point_light_bouncer.hpp:22-33,
 * vren_demo/resources/shaders/bounce_point_lights.comp:33-73).  positions / directions: vec4 per light (16-byte aligned),
 * updated in place: the light advances speed*dt along its direction inside [aabb_min, aabb_max] and reflects off the
 * faces it reaches (<= 32 reflections per call); position.w := 1, direction.w := 0.  aabb_* are HOST pointers
 * (push constants in the reference). */
int vrenb200_bounce_point_lights(vrenb200_stream_t stream, float* positions, float* directions, uint32_t count,
                                 const float aabb_min[3], const float aabb_max[3], float speed, float dt);

/* ---- n4: BVH debug consumer ---------------------------------------------------------------------------------------
 * vren_demo::visualize_bvh::write (vren_demo/vren_demo/visualize_bvh.cpp:59-94, show_bvh.comp:62-78): the 12 box edges of
 * every node as 24 vertices {float position[3]; uint32 color} at vertices[node * 24], colour by level
 * (colors[level_count - level + 2]), INVALID nodes as degenerate black lines.  level_count = log32(padded leaf count)
 * (vrenb200_calc_bvh_level_count), at most 4.  vertices: vrenb200_visualize_bvh_vertex_count(level_count) x 16 bytes. */
uint64_t vrenb200_visualize_bvh_vertex_count(uint32_t level_count);
int vrenb200_visualize_bvh(vrenb200_stream_t stream, const void* bvh_nodes, uint32_t level_count, void* vertices);

/* ---- n4: kd-tree (vren/vren/base/kd_tree.hpp:12-45) ------------------------------------------------------------------
 * Pre-order node array: an inner node {split, axis 0..2, right_child_distance} is followed by its left subtree; a leaf is a
 * run of nodes {point index, axis 3, run length in the first node / 0x3FFFFFFF in the others}.  axis = axis_and_link & 3,
 * link = axis_and_link >> 2.  The tree needs at most 2 * count nodes. */
typedef struct vrenb200_kd_tree_node
{
    union { float split; uint32_t index; };
    uint32_t axis_and_link;
} vrenb200_kd_tree_node;
/* host: vren::kd_tree_build (kd_tree.cpp:5-75).  indices[count] (a permutation of the point indices to organise) is
 * reordered in place; returns the number of nodes written (0 on bad arguments). */
size_t vrenb200_kd_tree_build(const float* points, size_t point_stride, uint32_t* indices, size_t count,
                              vrenb200_kd_tree_node* nodes, size_t max_leaf_point_count);
/* host: vren::kd_tree_search (kd_tree.cpp:77-129): best_point / best_distance_squared are in-out (start with +inf);
 * filter may be NULL (k_kd_tree_default_search_filter) */
void vrenb200_kd_tree_search(const float* points, size_t point_stride, const vrenb200_kd_tree_node* nodes, const float sample[3],
                             int (*filter)(uint32_t point, void* user), void* user, uint32_t* best_point, float* best_distance_squared);
/* device: the same search for sample_count queries at once (points, nodes, samples[sample_count][3] and the outputs are
 * device pointers); no reference counterpart */
int vrenb200_kd_tree_search_batch(vrenb200_stream_t stream, const float* points, uint32_t point_stride,
                                  const vrenb200_kd_tree_node* nodes, const float* samples, uint32_t sample_count,
                                  uint32_t* best_point, float* best_distance_squared);

#ifdef __cplusplus
}
#endif
#endif /* VRENB200_H_ */
