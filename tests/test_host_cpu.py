"""CPU-only tests of the product's host side: the C-ABI library loads and exports every symbol
include/b200pt.h declares, the host SAH builder is structurally identical to the oracle's
restatement of BVHAccel::new, and compute entry points fail loudly without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest


def test_cabi_exports_every_declared_symbol(pkg):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "b200pt.h")).read()
    names = set(re.findall(r"\b(b200pt_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 20
    L = C.CDLL(pkg.LIB_PATH)
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, "symbols declared in include/b200pt.h but not exported: %s" % missing


def test_integration_md_binds_every_declared_symbol(pkg):
    """INTEGRATION.md's b200pt-sys extern block is generated from the header (tools/gen_rust_ffi.py --update): it must name
    every entry point, and its struct mirror must carry the fields the ctypes mirror has."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "b200pt.h")).read()
    names = set(re.findall(r"\b(b200pt_[a-z0-9_]+)\s*\(", hdr))
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    missing = [n for n in sorted(names) if "pub fn %s(" % n not in doc]
    assert not missing, "entry points missing from INTEGRATION.md (run python tools/gen_rust_ffi.py --update): %s" % missing
    for field, _ in pkg.SceneDesc._fields_:
        assert "pub %s:" % field in doc, field


def test_struct_sizes_match_header(pkg):
    assert pkg.RAY_DTYPE.itemsize == 32 and pkg.HIT_DTYPE.itemsize == 16 and pkg.NODE_DTYPE.itemsize == 32
    assert C.sizeof(pkg.Material) == 4 + 15 * 4 + 3 * 4 + 4
    assert C.sizeof(pkg.Light) == 168 + 8 + 4 + 4 + 8 + 8  # 164 bytes of scalars, padded to 168 for the map pointer; map size; the spot light's two cosines; the projection light's fov + padding
    assert C.sizeof(pkg.Film) == 8 + 16 + 8 + 1024 + 8
    assert C.sizeof(pkg.FloatTexture) == 4 + 16 + 8 + 12 + 8  # 40 bytes of scalars, then the texel pointer


@pytest.mark.parametrize("max_prims", [1, 4, 8, 255])
@pytest.mark.parametrize("mesh", ["sphere", "soup", "tiny", "coincident"])
def test_host_sah_builder_matches_oracle(pkg, oracle, mesh, max_prims):
    from pbrt_v3_rs_b200 import workloads as wl
    if mesh == "sphere":
        tv = wl.displaced_sphere(60, 30)
    elif mesh == "soup":
        tv = wl.triangle_soup(5000)
    elif mesh == "tiny":
        tv = wl.ground_quad()
    else:  # many primitives with identical centroids -> zero-extent centroid bounds -> fat leaf (sah.rs:61)
        tv = np.tile(wl.ground_quad()[:1], (37, 1))
    pb = pkg.triangle_bounds(tv)
    assert np.array_equal(pb, oracle.triangle_bounds(tv))
    n1, o1 = pkg.build_bvh_sah(pb, max_prims)
    n2, o2 = oracle.build_bvh_sah(pb, max_prims)
    assert len(n1) == len(n2)
    assert n1.tobytes() == n2.tobytes(), "LinearBVHNode arrays differ"
    assert np.array_equal(o1, o2), "ordered_prims differ"


def test_host_sah_builder_keeps_the_reference_sign_of_zero(pkg, oracle):
    """min(a, b) = a < b ? a : b (core/src/pbrt/common.rs:83-108) keeps the LAST of two equal operands, so whether a
    box coordinate is +0 or -0 depends on the order the reference unites in: leaves fold their primitives, interior
    nodes unite child0 with child1 (accelerators/src/bvh/common.rs:150-159)."""
    rng = np.random.default_rng(11)
    n = 400
    lo = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    hi = lo + rng.uniform(0.01, 0.2, (n, 3)).astype(np.float32)
    zero = rng.integers(0, 2, (n, 3)).astype(bool)
    lo[zero] = np.where(rng.integers(0, 2, zero.sum()) == 1, np.float32(0.0), np.float32(-0.0))
    hi = np.maximum(hi, lo + np.float32(0.01))
    neg = rng.integers(0, 4, (n, 3)) == 0
    hi[neg & zero] = np.where(rng.integers(0, 2, (neg & zero).sum()) == 1, np.float32(0.0), np.float32(-0.0))
    lo[neg & zero] = hi[neg & zero] - np.float32(0.05)
    pb = np.concatenate([lo, hi], axis=1).astype(np.float32)
    for max_prims in (1, 4):
        n1, o1 = pkg.build_bvh_sah(pb, max_prims)
        n2, o2 = oracle.build_bvh_sah(pb, max_prims)
        assert n1.tobytes() == n2.tobytes() and np.array_equal(o1, o2)
    z = n1["bounds"][n1["bounds"] == 0]
    assert np.signbit(z).any() and (~np.signbit(z)).any()  # the case is exercised


def test_empty_and_single_primitive(pkg, oracle):
    n, o = pkg.build_bvh_sah(np.zeros((0, 6), np.float32))
    assert len(n) == 0  # BVHAccel::new with no primitives has no nodes (mod.rs:47-53)
    from pbrt_v3_rs_b200 import workloads as wl
    tv = wl.ground_quad()[:1]
    n, o = pkg.build_bvh_sah(pkg.triangle_bounds(tv))
    assert len(n) == 1 and n[0]["n_primitives"] == 1 and list(o) == [0]


def test_scene_setup_math_matches_oracle(pkg, oracle):
    """look_at / perspective matrices of the host package vs the oracle's restatement of transform.rs."""
    from pbrt_v3_rs_b200 import scene as sc
    eye, look, up = np.array([0.3, 1.2, -4.0], np.float32), np.array([0, -0.1, 0], np.float32), np.array([0, 1, 0], np.float32)
    c2w = np.zeros((4, 4), np.float32)
    r2c = np.zeros((4, 4), np.float32)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    oracle.lib().orc_camera_matrices(P(eye), P(look), P(up), 40.0, 400, 300, None, P(c2w), P(r2c))
    assert np.allclose(sc.look_at_camera_to_world(eye, look, up), c2w, atol=1e-6)
    assert np.allclose(sc.perspective_raster_to_camera(40.0, 400, 300), r2c, rtol=1e-5, atol=1e-7)


def test_compute_calls_fail_loudly_without_device(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.B200PTError):
        pkg.init(0)
    from pbrt_v3_rs_b200 import workloads as wl
    with pytest.raises(pkg.B200PTError):
        pkg.BVHAccel.from_params({}, wl.ground_quad())


def test_workload_generators_are_deterministic(pkg):
    from pbrt_v3_rs_b200 import workloads as wl
    a, b = wl.c2_mesh(wl.C2_SMALL), wl.c2_mesh(wl.C2_SMALL)
    assert a.tobytes() == b.tobytes() and a.shape == (10000, 9)
    r = wl.primary_rays(64, 32)
    assert r.shape == (2048,) and np.allclose(np.linalg.norm(r["d"], axis=1), 1, atol=1e-6)


# ---- SplitMethod::HLBVH (accelerators/src/bvh/hlbvh.rs, morton.rs) --------------------------------------------

def _ref_morton_py(pb):
    """Independent numpy restatement of compute_morton_primitives + encode_morton_3 (hlbvh.rs:104-141, morton.rs:43-49,
    102-120): note float_to_bits -> the code interleaves bits of the float's bit pattern."""
    lo, hi = pb[:, :3].min(0), pb[:, 3:].max(0)
    cen = np.float32(0.5) * (pb[:, :3] + pb[:, 3:])
    off = cen - lo
    ext = hi - lo
    off = np.where(hi > lo, off / np.where(hi > lo, ext, 1).astype(np.float32), off).astype(np.float32)
    bits = (off * np.float32(1024.0)).astype(np.float32).view(np.uint32)

    def spread(x):
        x = np.where(x == 1024, x - 1, x).astype(np.uint64)
        x = (x | (x << 16)) & 0x030000FF
        x = (x | (x << 8)) & 0x0300F00F
        x = (x | (x << 4)) & 0x030C30C3
        x = (x | (x << 2)) & 0x09249249
        return x
    return ((spread(bits[:, 2]) << 2) | (spread(bits[:, 1]) << 1) | spread(bits[:, 0])).astype(np.uint32)


def test_hlbvh_morton_codes_and_sort(pkg, oracle):
    from pbrt_v3_rs_b200 import workloads as wl
    pb = pkg.triangle_bounds(wl.triangle_soup(3000))
    codes = pkg.hlbvh_morton_codes(pb)
    assert np.array_equal(codes, _ref_morton_py(pb))
    assert codes.max() < (1 << 30)
    _, ordered, sorted_codes = oracle.build_bvh_hlbvh(pb, 4, with_codes=True)
    order = np.argsort(codes, kind="stable")  # the reference's LSD radix sort is a stable sort of the 30-bit code
    assert np.array_equal(sorted_codes, codes[order])
    assert np.array_equal(ordered, order.astype(np.uint32))  # leaves take their primitives in sorted order


@pytest.mark.parametrize("max_prims", [1, 4, 8, 255])
@pytest.mark.parametrize("mesh", ["sphere", "soup", "tiny", "coincident+soup"])
def test_host_hlbvh_builder_matches_oracle(pkg, oracle, mesh, max_prims):
    from pbrt_v3_rs_b200 import workloads as wl
    if mesh == "sphere":
        tv = wl.displaced_sphere(60, 30)
    elif mesh == "soup":
        tv = wl.triangle_soup(5000)
    elif mesh == "tiny":
        tv = wl.ground_quad()
    else:
        tv = np.concatenate([np.tile(wl.ground_quad()[:1], (37, 1)), wl.triangle_soup(200)])
    pb = pkg.triangle_bounds(tv)
    n1, o1 = pkg.build_bvh_hlbvh(pb, max_prims)
    n2, o2 = oracle.build_bvh_hlbvh(pb, max_prims)
    assert len(n1) == len(n2)
    assert np.array_equal(o1, o2), "ordered_prims differ"
    assert n1.tobytes() == n2.tobytes(), "LinearBVHNode arrays differ"
    # structural invariants: every primitive in exactly one leaf, children inside parents
    leaves = n1[n1["n_primitives"] > 0]
    assert leaves["n_primitives"].sum() == len(tv) and sorted(o1.tolist()) == list(range(len(tv)))
    for i in np.flatnonzero(n1["n_primitives"] == 0)[:200]:
        for c in (i + 1, n1["offset"][i]):
            assert np.all(n1["bounds"][c][:3] >= n1["bounds"][i][:3]) and np.all(n1["bounds"][c][3:] <= n1["bounds"][i][3:])


def test_hlbvh_tree_traces_like_brute_force(pkg, oracle):
    """The HLBVH tree (whatever its quality) must return the same closest hits as the SAH tree: same primitive ids and t."""
    from pbrt_v3_rs_b200 import workloads as wl
    tv = wl.displaced_sphere(40, 20)
    pb = pkg.triangle_bounds(tv)
    rays = wl.primary_rays(48, 24)
    n1, o1 = pkg.build_bvh_hlbvh(pb, 4)
    n2, o2 = pkg.build_bvh_sah(pb, 4, where="host")
    h1 = oracle.OracleAccel(n1, o1, tv).intersect(rays, counters=False, diag=False)[0]
    h2 = oracle.OracleAccel(n2, o2, tv).intersect(rays, counters=False, diag=False)[0]
    assert (h1["prim"] != pkg.MISS).sum() > 100
    assert np.array_equal(h1["t"].view(np.uint32), h2["t"].view(np.uint32))
    same = h1["prim"] == h2["prim"]
    assert same.mean() > 0.995  # equal-t ties on shared edges may go to the other triangle (different visiting order)


def test_hlbvh_single_and_empty(pkg, oracle):
    from pbrt_v3_rs_b200 import workloads as wl
    n, o = pkg.build_bvh_hlbvh(np.zeros((0, 6), np.float32))
    assert len(n) == 0
    pb = pkg.triangle_bounds(wl.ground_quad()[:1])
    n1, o1 = pkg.build_bvh_hlbvh(pb, 4)
    n2, o2 = oracle.build_bvh_hlbvh(pb, 4)
    assert len(n1) == 1 and n1.tobytes() == n2.tobytes() and list(o1) == [0]
