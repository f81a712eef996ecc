"""Seeded inputs of the reference-pinned clustered-shading parity tests (tests/test_clustered_reference.py) and of the
generator of their golden file (tests/golden/make_clustered_golden.py).  Only OUTPUTS of the reference's compiled GLSL
are stored in the fixture; the inputs are rebuilt from these functions on every machine (integer arithmetic on a
splitmix64 stream and correctly rounded IEEE operations; the GPU box runs the same image as the build container)."""
import math

import numpy as np

from conftest import splitmix64

N = 1 << 17                      # elements per pure-function case (>= 10^5)
FOV_Y, NEAR, FAR = np.float32(math.radians(45.0)), np.float32(0.01), np.float32(1000.0)   # camera.hpp:20-23
IMAGES = {"4k_rows": (64, 2160), "1080p_rows": (64, 1080)}     # W x H: same tile rows (68 / 34) and slice base as 4K / 1080p


def unit(seed, n):
    """U[0,1) on a 2^-24 grid as float64 (exact in float32 too)"""
    return (splitmix64(seed, n) >> np.uint64(40)).astype(np.float64) / float(1 << 24)


def to_half_exact(x):
    """float32 values that are exactly representable in binary16 (what an RGBA16F g-buffer can hold)"""
    return np.asarray(x, np.float32).astype(np.float16).astype(np.float32)


def normals(seed, n):
    """[n,3] float32, half-representable: random directions (not normalised, like the interpolated g-buffer normals),
    axis-aligned, zero components, exact zeros and equal-magnitude ties"""
    v = to_half_exact(unit(seed, 3 * n).reshape(n, 3) * 2.0 - 1.0)
    kind = splitmix64(seed + 1, n) % np.uint64(20)
    axis = (splitmix64(seed + 2, n) % np.uint64(3)).astype(np.int64)
    sign = np.where(splitmix64(seed + 3, n) & np.uint64(1), 1.0, -1.0).astype(np.float32)
    rows = np.arange(n)
    aligned = np.zeros((n, 3), np.float32)
    aligned[rows, axis] = sign
    v = np.where((kind == 0)[:, None], aligned, v)                       # axis-aligned
    zc = v.copy(); zc[rows, axis] = 0.0
    v = np.where((kind == 1)[:, None], zc, v)                            # one zero component
    v = np.where((kind == 2)[:, None], np.float32(0.0), v)               # null normal -> UINT32_MAX
    tie = v.copy(); tie[rows, (axis + 1) % 3] = tie[rows, axis] * sign   # two components of equal magnitude
    v = np.where((kind == 3)[:, None], tie, v)
    return np.ascontiguousarray(v, np.float32)


def depth_values(seed, n):
    """float32 depth-buffer values: view-space z = (1 + u) 2^e, e uniform in [-6, 9] (roughly log-uniform over
    [1.5 near, far], exact arithmetic only) through d = f/(f-n) - f n/((f-n) z) (camera.cpp:40-50), 5 % exactly 1.0
    (cleared far plane)"""
    f, nr = float(FAR), float(NEAR)
    e = (splitmix64(seed + 2, n) % np.uint64(16)).astype(np.int64) - 6
    z = np.clip(np.ldexp(1.0 + unit(seed, n), e), nr * 1.5, f)
    d = (f / (f - nr) - f * nr / ((f - nr) * z)).astype(np.float32)
    far_px = splitmix64(seed + 1, n) % np.uint64(20) == 0
    return np.where(far_px, np.float32(1.0), d).astype(np.float32)


def image(name):
    """(depth [H,W] f32, normals [H,W,3] f32 half-exact, uv [H,W,2] f32 = frag_coord of find_unique_clusters.comp:48)"""
    w, h = IMAGES[name]
    seed = 100 + sorted(IMAGES).index(name) * 10
    depth = depth_values(seed, w * h).reshape(h, w)
    nrm = normals(seed + 5, w * h).reshape(h, w, 3)
    xs, ys = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    uv = np.stack([xs / np.float32(w), ys / np.float32(h)], axis=-1).astype(np.float32)
    return np.ascontiguousarray(depth), np.ascontiguousarray(nrm), np.ascontiguousarray(uv)


def rgba16f(nrm):
    """[H,W,3] float32 (half-exact) -> [H,W,4] uint16 bit patterns, alpha = 0"""
    h, w, _ = nrm.shape
    out = np.zeros((h, w, 4), np.uint16)
    out[..., :3] = nrm.astype(np.float16).view(np.uint16)
    return out


def cluster_cells():
    """every tile of a 4K frame (120 x 68) at 16 slices: uvec3 ijk [N', 3] with N' = 130 560"""
    ks = np.array([0, 1, 2, 5, 17, 63, 100, 200, 333, 478, 500, 640, 777, 900, 950, 1023], np.uint32)
    i, j, k = np.meshgrid(np.arange(120, dtype=np.uint32), np.arange(68, dtype=np.uint32), ks, indexing="ij")
    return np.ascontiguousarray(np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1), np.uint32)


def keys_of(ijk):
    return (ijk[:, 0] | (ijk[:, 1] << np.uint32(8)) | (ijk[:, 2] << np.uint32(16)) | np.uint32(63 << 26)).astype(np.uint32)


def _nudge(x, steps):
    """move float32 values by a few ulps (int32 steps on the bit pattern; x > 0)"""
    return (np.asarray(x, np.float32).view(np.int32) + steps.astype(np.int32)).view(np.float32)


def node_boxes(cmin, cmax, seed):
    """[n,6] node boxes {min, max} around the given cluster corners: overlapping, disjoint and touching (shared bounds)"""
    n = cmin.shape[0]
    lo, hi = np.minimum(cmin, cmax), np.maximum(cmin, cmax)
    ext = np.maximum(hi - lo, np.float32(1e-3))
    # (the corners are not ordered component-wise — y is flipped —, so on such an axis only a node that spans the whole
    # cluster interval passes the test as written: boxes up to three cluster extents wide, centred near the cluster)
    c = lo + (unit(seed, 3 * n).reshape(n, 3) * 2.0 - 0.5).astype(np.float32) * ext
    half = (unit(seed + 1, 3 * n).reshape(n, 3) * 3.0).astype(np.float32) * ext
    bmin, bmax = (c - half).astype(np.float32), (c + half).astype(np.float32)
    touch = splitmix64(seed + 2, n) % np.uint64(4)
    ax = (splitmix64(seed + 3, n) % np.uint64(3)).astype(np.int64)
    rows = np.arange(n)
    # touching cases: node.min == cluster max (>= holds) or node.max == cluster min (<= holds) on one axis, exactly
    bmin[rows[touch == 0], ax[touch == 0]] = cmax[rows[touch == 0], ax[touch == 0]]
    bmax[rows[touch == 0], ax[touch == 0]] = np.maximum(bmax[rows[touch == 0], ax[touch == 0]], bmin[rows[touch == 0], ax[touch == 0]])
    bmax[rows[touch == 1], ax[touch == 1]] = cmin[rows[touch == 1], ax[touch == 1]]
    bmin[rows[touch == 1], ax[touch == 1]] = np.minimum(bmin[rows[touch == 1], ax[touch == 1]], bmax[rows[touch == 1], ax[touch == 1]])
    return np.ascontiguousarray(np.concatenate([bmin, bmax], axis=1), np.float32)


def spheres(cmin, cmax, seed):
    """[n,4] {centre, radius}: centres around the cluster box; half of the radii sit within +-2 ulp of the exact
    clamp distance (the `d < r` boundary), the rest are random"""
    n = cmin.shape[0]
    lo, hi = np.minimum(cmin, cmax), np.maximum(cmin, cmax)
    ext = np.maximum(hi - lo, np.float32(1e-3))
    o = (lo + (unit(seed, 3 * n).reshape(n, 3) * 3.0 - 1.0).astype(np.float32) * ext).astype(np.float32)
    # the shader's clamp max(aabb_min, min(o, aabb_max)) with the corners as written (not re-ordered)
    p = np.maximum(cmin, np.minimum(o, cmax)).astype(np.float32)
    q = (p - o).astype(np.float32)
    d = np.sqrt(((q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1]).astype(np.float32) + q[:, 2] * q[:, 2]).astype(np.float32)).astype(np.float32)
    steps = (splitmix64(seed + 1, n) % np.uint64(5)).astype(np.int64) - 2
    near_tie = _nudge(np.maximum(d, np.float32(1e-30)), steps)
    rnd = (unit(seed + 2, n) * 2.0).astype(np.float32) * (d + np.float32(1e-3))
    r = np.where(splitmix64(seed + 3, n) & np.uint64(1), near_tie, rnd).astype(np.float32)
    return np.ascontiguousarray(np.concatenate([o, r[:, None]], axis=1), np.float32)


def light_chain(n=N):
    """a6 inputs: positions [n,4] (w = 1), intensities [n], view matrix (16 floats, column-major): a rotation about y and x
    plus a translation, entries rounded to float32"""
    pos = np.ones((n, 4), np.float32)
    pos[:, :3] = (unit(7, 3 * n).reshape(n, 3) * 40.0 - 20.0).astype(np.float32)
    intensity = (0.25 + unit(8, n) * 3.0).astype(np.float32)
    yaw, pitch = 0.7, -0.3
    cy, sy, cp, sp = math.cos(yaw), math.sin(yaw), math.cos(pitch), math.sin(pitch)
    ry = np.array([[cy, 0, sy, 0], [0, 1, 0, 0], [-sy, 0, cy, 0], [0, 0, 0, 1]], np.float64)
    rx = np.array([[1, 0, 0, 0], [0, cp, -sp, 0], [0, sp, cp, 0], [0, 0, 0, 1]], np.float64)
    tr = np.eye(4); tr[:3, 3] = [1.5, -2.25, 30.0]
    view = (tr @ rx @ ry).astype(np.float32)            # row-major math; stored column-major below
    return pos, intensity, np.ascontiguousarray(view.T.reshape(-1), np.float32)


def node_address_cases(n=4096):
    """(root, level, overlaps[4]) for 3- and 4-level trees (32^3 and 32^4 leaves), non-zero masks"""
    roots = np.where(splitmix64(21, n) & np.uint64(1), np.uint32(1082400), np.uint32(33824)).astype(np.uint32)
    depth = np.where(roots == 1082400, 4, 3)
    level = (splitmix64(22, n) % depth.astype(np.uint64)).astype(np.uint32)
    masks = (splitmix64(23, 4 * n) & np.uint64(0xFFFFFFFF)).astype(np.uint32).reshape(n, 4)
    masks |= np.uint32(1) << (splitmix64(24, 4 * n) % np.uint64(32)).astype(np.uint32).reshape(n, 4)
    return roots, level, np.ascontiguousarray(masks)


def depth_image_n2():
    """odd-sized depth image for one reduction level (depth_buffer_reduce.comp): 257 x 131 -> 128 x 65"""
    w, h = 257, 131
    return unit(31, w * h).astype(np.float32).reshape(h, w)


def bounce_case(n=4096):
    lo, hi = np.array([-10, -5, -2.5], np.float32), np.array([10, 5, 7.5], np.float32)
    pos = np.ones((n, 4), np.float32)
    pos[:, :3] = (lo + unit(41, 3 * n).reshape(n, 3).astype(np.float32) * (hi - lo)).astype(np.float32)
    d = (unit(42, 3 * n).reshape(n, 3) * 2.0 - 1.0)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    dirs = np.zeros((n, 4), np.float32)
    dirs[:, :3] = d.astype(np.float32)
    return pos, dirs, lo, hi, np.float32(7.5), np.float32(0.25)      # a step of 1.875 units, three calls in the tests


def light_lists(n=4096):
    counts = (splitmix64(51, n) % np.uint64(40)).astype(np.uint32)
    offsets = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.uint32)
    indices = (splitmix64(52, int(counts.sum())) % np.uint64(65536)).astype(np.uint32)
    return counts, offsets, indices


# ---- evaluation of every case through one backend --------------------------------------------------------------------------
HEAD = 1 << 14      # scalars of every output kept verbatim in the golden file (all of it is covered by a SHA-256)


def camera_matrices(aspect):
    """(projection, closed-form inverse) as 16 floats column-major, from the oracle's restatement of camera.cpp:40-50"""
    import ctypes as C

    import oracle

    cam = oracle.Camera(FOV_Y, np.float32(aspect), NEAR, FAR)
    proj, inv = np.zeros(16, np.float32), np.zeros(16, np.float32)
    oracle.load().oracle_projection(C.byref(cam), proj, inv)
    return cam, proj, inv


def evaluate(backend, generic_inverse=False):
    """every case through `backend`: "ref" = the reference's GLSL compiled by g++ (oracle/_ref/libvrenref_glsl.so),
    "oracle" = oracle/oracle_clustered.cpp.  generic_inverse (ref only): use the shim's cofactor inverse() instead of the
    closed-form inverse(projection) the oracle and the product use.  Returns {name: ndarray}."""
    import ctypes as C

    import oracle

    ref = backend == "ref"
    lib = oracle.load_ref_glsl() if ref else oracle.load()
    if lib is None:
        raise RuntimeError("oracle/_ref/libvrenref_glsl.so is not built (needs /root/reference)")
    half_fov = np.float32(FOV_Y / np.float32(2.0))          # clustered_shading.cpp:419
    out = {}

    def inverse_mode(inv):
        if ref:
            lib.refglsl_set_inverse_override(None if generic_inverse else inv.ctypes.data_as(C.c_void_p))

    # a7: normal bins
    nrm = normals(61, N)
    bins = np.zeros(N, np.uint32)
    (lib.refglsl_discretize_normal if ref else lib.oracle_discretize_normal)(nrm.reshape(-1), N, bins)
    out["normal_bins"] = bins
    # a7: depth -> view z -> slice -> key, per image
    for name, (w, h) in IMAGES.items():
        depth, nimg, uv = image(name)
        cam, proj, inv = camera_matrices(16.0 / 9.0)
        tiles_y = (h + 31) // 32
        keys, vz = np.zeros(w * h, np.uint32), np.zeros(w * h, np.float32)
        if ref:
            inverse_mode(inv)
            lib.refglsl_cluster_key(uv.reshape(-1), depth.reshape(-1), nimg.reshape(-1), w * h, 0, 0, tiles_y, NEAR, half_fov, proj, keys, vz)
        else:
            lib.oracle_cluster_key(depth.reshape(-1), nimg.reshape(-1), w * h, tiles_y, C.byref(cam), keys, vz)
        out[f"key_hi_{name}"] = (keys >> np.uint32(16)).astype(np.uint16).reshape(h, w)
        out[f"view_z_{name}"] = vz.reshape(h, w)
    # a8: cluster corners, then the two overlap predicates on boxes / spheres placed around them
    cells = cluster_cells()
    n = cells.shape[0]
    cam, proj, inv = camera_matrices(3840.0 / 2160.0)
    cmin, cmax = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
    if ref:
        inverse_mode(inv)
        mn4, mx4 = np.zeros((n, 4), np.float32), np.zeros((n, 4), np.float32)
        lib.refglsl_calc_cluster_aabb(cells.reshape(-1), n, 120, 68, NEAR, half_fov, proj, mn4.reshape(-1), mx4.reshape(-1))
        cmin, cmax = np.ascontiguousarray(mn4[:, :3]), np.ascontiguousarray(mx4[:, :3])
        dec = np.zeros((n, 4), np.uint32)
        lib.refglsl_decode_cluster_key(keys_of(cells), n, dec.reshape(-1))
        assert np.array_equal(dec[:, :3], cells) and bool((dec[:, 3] == 63).all())      # clustered_shading.glsl:58-70
    else:
        lib.oracle_cluster_aabb(cells.reshape(-1), n, 120, 68, C.byref(cam), cmin.reshape(-1), cmax.reshape(-1))
    out["aabb_min"], out["aabb_max"] = cmin, cmax
    if not generic_inverse:
        boxes, sph = node_boxes(cmin, cmax, 71), spheres(cmin, cmax, 81)
        b12 = np.ascontiguousarray(np.concatenate([cmin, cmax, boxes], axis=1), np.float32)
        s10 = np.ascontiguousarray(np.concatenate([sph, cmin, cmax], axis=1), np.float32)
        f1, f2 = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        (lib.refglsl_test_aabb_aabb if ref else lib.oracle_test_aabb_aabb)(b12.reshape(-1), n, f1)
        (lib.refglsl_test_sphere_aabb if ref else lib.oracle_test_sphere_aabb)(s10.reshape(-1), n, f2)
        out["aabb_aabb"], out["sphere_aabb"] = f1, f2
    # a8: node addresses
    roots, level, masks = node_address_cases()
    fn = lib.refglsl_get_node_address if ref else lib.oracle_get_node_address
    out["node_address"] = np.array([fn(int(roots[i]), int(level[i]), masks[i]) for i in range(roots.size)], np.int32)
    # a6: view transform -> Morton code (over the exact min / max of the view positions) -> leaf boxes
    pos, intensity, view = light_chain()
    vp = np.zeros((N, 4), np.float32)
    (lib.refglsl_position_to_view_space if ref else lib.oracle_position_to_view_space)(view, pos.reshape(-1), N, vp.reshape(-1))
    out["view_pos"] = vp
    mn, mx = vp.min(axis=0).astype(np.float32), vp.max(axis=0).astype(np.float32)
    p3 = np.ascontiguousarray(vp[:, :3])
    codes = np.zeros(N, np.uint32)
    (lib.refglsl_morton_code if ref else lib.oracle_morton_code)(p3.reshape(-1), N, mn, mx, codes)
    out["morton"] = codes
    lmin, lmax = np.zeros((N, 3), np.float32), np.zeros((N, 3), np.float32)
    (lib.refglsl_light_leaf_box if ref else lib.oracle_light_leaf_box)(vp.reshape(-1), intensity, N, lmin.reshape(-1), lmax.reshape(-1))
    out["leaf_min"], out["leaf_max"] = lmin, lmax
    # n2: one pyramid level
    dimg = depth_image_n2()
    h, w = dimg.shape
    if ref:
        to = np.zeros((h >> 1, w >> 1), np.float32)
        lib.refglsl_depth_reduce(dimg.reshape(-1), w, h, to.reshape(-1), w >> 1, h >> 1)
    else:
        pyr, _ = oracle.depth_pyramid(dimg)
        to = pyr[w * h:w * h + (w >> 1) * (h >> 1)].reshape(h >> 1, w >> 1).copy()
    out["depth_reduce"] = to
    # n3: three bounce steps
    bp, bd, lo, hi, speed, dt = bounce_case()
    bp, bd = bp.copy(), bd.copy()
    for _ in range(3):
        if ref:
            lib.refglsl_bounce_point_lights(bp.reshape(-1), bd.reshape(-1), bp.shape[0], lo, hi, speed, dt)
        else:
            bp, bd = oracle.bounce_point_lights(bp, bd, lo, hi, speed, dt)
    out["bounce_pos"], out["bounce_dir"] = bp, bd
    # n1: XOR of every cluster's light list
    counts, offsets, indices = light_lists()
    if ref:
        x = np.zeros(counts.size, np.uint32)
        lib.refglsl_light_list_xor(counts, offsets, indices, indices.size, counts.size, x)
    else:
        cref = np.arange(counts.size, dtype=np.uint32).reshape(1, -1)
        x = oracle.light_list_hash(cref, counts, offsets, indices)[0, :, 1].copy()
    out["list_xor"] = x
    return out


def digest(arr):
    import hashlib

    a = np.ascontiguousarray(arr)
    return hashlib.sha256(a.view(np.uint8).reshape(-1).tobytes()).hexdigest()


def head(arr):
    a = np.ascontiguousarray(arr)
    return a.reshape(-1)[:HEAD].copy()
