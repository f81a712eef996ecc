"""SURVEY 8b, checked mechanically against the reference's headers where they lie (/root/reference; skipped on the GPU box): the
C++ facade (include/vren/**) declares every public entry point of the hot path with the reference's method name and the
reference's parameter TYPES in the reference's ORDER, and carries the same constants — so a call site of the reference compiles
against the facade unchanged.  Deliberate differences are listed, not ignored."""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference/vren/vren")
FACADE = ROOT / "include" / "vren"

pytestmark = pytest.mark.skipif(not REF.exists(), reason="/root/reference is not present here")


def strip_comments(text):
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.sub(r"//[^\n]*", "", text)


def signatures(text, name):
    """parameter type lists of every declaration / definition of `name(` in the text"""
    text = strip_comments(text)
    out = []
    for m in re.finditer(re.escape(name) + r"\s*\(", text):
        if name != "operator()" and not re.search(r"[\w&*>]\s+$", text[max(0, m.start() - 40):m.start()]):
            continue                                   # a call, not a declaration
        i, depth = m.end(), 1
        while depth and i < len(text):
            depth += {"(": 1, ")": -1}.get(text[i], 0)
            i += 1
        params, cur, depth = [], "", 0
        for ch in text[m.end():i - 1]:
            if ch == "," and depth == 0:
                params.append(cur)
                cur = ""
            else:
                depth += {"(": 1, "<": 1, ")": -1, ">": -1}.get(ch, 0)
                cur += ch
        if cur.strip():
            params.append(cur)
        types = []
        for p in params:
            defaulted = "=" in p
            p = re.sub(r"=.*$", "", p)                                  # default argument
            p = re.sub(r"\s+", " ", p).strip()
            if not re.search(r"[&*]$", p) and " " in p:                 # a trailing identifier is the parameter's name
                p = p.rsplit(" ", 1)[0]
            p = re.sub(r"\s*([&*])\s*", r"\1", p).replace("vren::", "")
            p = re.sub(r"\bglm(_compat)?::", "", p)                     # the facade carries its own vec / mat types (no glm dependency)
            types.append(p + (" =default" if defaulted else ""))
        out.append(tuple(types))
    return out


def accepts(facade_sig, ref_sig):
    """a call written for the reference's signature compiles against the facade's: same types in the same order; the facade may
    append parameters that have defaults"""
    plain = tuple(t.replace(" =default", "") for t in facade_sig)
    ref_plain = tuple(t.replace(" =default", "") for t in ref_sig)
    return plain[:len(ref_plain)] == ref_plain and all(t.endswith(" =default") for t in facade_sig[len(ref_plain):])


HOT_PATH = [
    # (reference header, facade header, method names whose every reference signature the facade must declare)
    ("primitives/reduce.hpp", "primitives/reduce.hpp", ["operator()"]),
    ("primitives/blelloch_scan.hpp", "primitives/blelloch_scan.hpp", ["operator()", "downsweep"]),
    ("primitives/radix_sort.hpp", "primitives/radix_sort.hpp", ["operator()", "create_scratch_buffer_1", "create_scratch_buffer_2"]),
    ("primitives/bucket_sort.hpp", "primitives/bucket_sort.hpp", ["operator()", "get_required_output_buffer_size"]),
    ("primitives/build_bvh.hpp", "primitives/build_bvh.hpp", ["operator()", "get_required_buffer_size", "calc_bvh_padded_leaf_count", "calc_bvh_buffer_length",
                                                              "calc_bvh_root_index", "calc_bvh_level_count"]),
    ("pipeline/clustered_shading.hpp", "pipeline/clustered_shading.hpp", ["operator()"]),
]


@pytest.mark.parametrize("ref_header,facade_header,names", HOT_PATH, ids=[h[0] for h in HOT_PATH])
def test_facade_declares_the_reference_signatures(ref_header, facade_header, names):
    ref, fac = (REF / ref_header).read_text(), (FACADE / facade_header).read_text()
    for name in names:
        want = signatures(ref, name)
        have = signatures(fac, name)
        assert want, f"{name} not found in the reference's {ref_header}"
        for sig in want:
            if ref_header.endswith("clustered_shading.hpp") and "material_buffer const&" in sig:
                # the two entry points that involve what SURVEY 8 puts out of scope — the shade step (shade.comp) and the render graph:
                if "render_graph_allocator&" not in sig:
                    continue                            # clustered_shading::shade::operator(): not built
                # cluster_and_shade::operator() builds a render-graph node in the reference; the facade's runs steps 1-3 directly: the
                # reference's parameters without the allocator and the two the shade step alone reads, behind the command buffer
                # and resource container every stage functor takes
                keep = tuple(t for t in sig if t not in ("render_graph_allocator&", "material_buffer const&", "vk_utils::combined_image_view const&"))
                assert ("VkCommandBuffer", "resource_container&") + keep in have
                continue
            assert any(accepts(h, sig) for h in have), f"{ref_header}: {name}{sig} is not declared by the facade (it has {have})"


def constants(text, pattern):
    return {m.group(1): re.sub(r"\s+", " ", m.group(2)).strip() for m in re.finditer(pattern, strip_comments(text))}


def test_facade_constants_equal_the_reference():
    define = r"#define\s+(VREN_\w+)\s+(.+)"
    ref = constants((REF / "config.hpp").read_text(), define)
    fac = {}
    for h in ("pipeline/clustered_shading.hpp", "vk_helpers/buffer.hpp"):
        fac.update(constants((FACADE / h).read_text(), define))
    for name in ("VREN_MAX_SCREEN_WIDTH", "VREN_MAX_SCREEN_HEIGHT", "VREN_MAX_POINT_LIGHT_COUNT", "VREN_MAX_UNIQUE_CLUSTER_KEY_COUNT",
                 "VREN_MAX_ASSIGNED_LIGHT_COUNT", "VREN_MIN_STORAGE_BUFFER_OFFSET_ALIGNMENT"):
        assert fac[name] == ref[name], name
    member = r"inline static const(?:expr)? uint32_t (k_\w+)\s*=\s*([^;]+);"
    deliberate = {("primitives/radix_sort.hpp", "k_radix_bits"): ("4", "8")}   # 4-bit x 8 passes there, 8-bit x 4 passes here
    for header in ("primitives/blelloch_scan.hpp", "primitives/bucket_sort.hpp", "primitives/build_bvh.hpp", "primitives/radix_sort.hpp",
                   "primitives/reduce.hpp"):
        r, f = constants((REF / header).read_text(), member), constants((FACADE / header).read_text(), member)
        assert r, header
        for name, value in r.items():
            assert name in f, f"{header}: {name} missing in the facade"
            if (header, name) in deliberate:
                assert (value, f[name]) == deliberate[(header, name)]
            else:
                assert f[name] == value, f"{header}: {name} = {f[name]}, reference {value}"


def test_bvh_node_layout_equals_the_reference():
    """build_bvh.hpp:8-18: min, next, max, pad — 32 bytes; the C ABI's vrenb200_bvh_node and the oracle's numpy dtype follow it"""
    import oracle

    text = strip_comments((REF / "primitives/build_bvh.hpp").read_text())
    body = re.search(r"struct bvh_node\s*\{(.*?)\};", text, flags=re.S).group(1)
    fields = re.findall(r"(glm::vec3|uint32_t)\s+(\w+)\s*;", body)
    assert [t for t, _ in fields] == ["glm::vec3", "uint32_t", "glm::vec3", "uint32_t"]
    assert oracle.BVH_NODE.itemsize == 32 and [oracle.BVH_NODE.fields[n][1] for n in ("min", "next", "max", "pad")] == [0, 12, 16, 28]
    abi = strip_comments((ROOT / "include" / "vrenb200.h").read_text())
    node = re.search(r"typedef struct vrenb200_bvh_node\s*\{(.*?)\}", abi, flags=re.S).group(1)
    assert re.findall(r"(float|uint32_t)\s+\w+(\[3\])?\s*;", node) == [("float", "[3]"), ("uint32_t", ""), ("float", "[3]"), ("uint32_t", "")]


def test_every_reference_test_has_a_facade_counterpart():
    """vren_test's own test list (TEST(suite, name) in vren_test/vren_test/**) against tests/cpp/facade_test.cpp, which drives the
    facade the way vren_test drives vren: each of them is named there by the function that restates it"""
    names = set()
    for f in (Path("/root/reference") / "vren_test" / "vren_test").rglob("*.cpp"):
        names.update(re.findall(r"^\s*TEST\((\w+),\s*(\w+)\)", f.read_text(errors="replace"), flags=re.M))
    assert len(names) >= 9
    facade = (ROOT / "tests" / "cpp" / "facade_test.cpp").read_text()
    for suite, name in sorted(names):
        assert f"TEST({suite}, {name})" in facade, f"TEST({suite}, {name}) has no counterpart in tests/cpp/facade_test.cpp"
