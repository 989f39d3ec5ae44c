"""Known-answer tests that pin the CPU oracle: the reference's own tests hold no
golden vectors for this path ("parity unpinned"), so the oracle is pinned by (i) the published PCG32
reference stream, (ii) hand-computable radical inverses, (iii) the geometry identities the reference
does test (core/src/geometry/{coordinate_system,vector3,ray}.rs #[cfg(test)]) and (iv) hand-checkable
one-triangle scenes, plus frozen hashes under tests/golden/."""
import ctypes as C
import hashlib
import json
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_pcg32_published_stream(oracle):
    # pcg32-demo (pcg-c-basic): pcg32_srandom_r(42, 54) -> first six outputs.
    out = np.zeros(6, dtype=np.uint32)
    oracle.lib().orc_pcg32_stream(54, 42, 0, out.ctypes.data_as(C.c_void_p), 6)
    assert [hex(x) for x in out] == ["0xa15c02b7", "0x7b47f409", "0xba1d3330", "0x83d2f293", "0xbfa4784b", "0xcbed606e"]


def test_pcg32_default_state_is_frozen(oracle):
    # RNG::default() (core/src/rng.rs:14-16,27-35): frozen first outputs of the default state/stream.
    out = np.zeros(4, dtype=np.uint32)
    oracle.lib().orc_pcg32_stream(0, 0, 1, out.ctypes.data_as(C.c_void_p), 4)
    # independent pure-python PCG32
    state, inc = 0x853C49E6748FEA9B, 0xDA3E39CB94B95BDB
    exp = []
    for _ in range(4):
        old = state
        state = (old * 0x5851F42D4C957F2D + inc) & (2**64 - 1)
        xs = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
        rot = old >> 59
        exp.append(((xs >> rot) | (xs << ((-rot) & 31))) & 0xFFFFFFFF)
    assert list(out) == exp


def test_uniform_float_and_bounded(oracle):
    f = np.zeros(1000, dtype=np.float32)
    oracle.lib().orc_pcg32_floats(7, f.ctypes.data_as(C.c_void_p), 1000)
    assert (f >= 0).all() and (f < 1).all()
    for skip in range(20):
        v = oracle.lib().orc_pcg32_bounded(3, 0, 17, skip)
        assert 0 <= v < 17


def test_radical_inverse_hand_values(oracle):
    L = oracle.lib()
    # base 2: 1 -> .1b = 0.5, 2 -> .01b = 0.25, 3 -> .11b = 0.75, 6 -> .011b = 0.375
    for a, v in [(0, 0.0), (1, 0.5), (2, 0.25), (3, 0.75), (6, 0.375)]:
        assert L.orc_radical_inverse(0, a) == np.float32(v)
    # base 3: 1 -> 1/3, 2 -> 2/3, 3 -> 1/9, 5 (=12_3) -> 2/3 + 1/9
    inv3 = np.float32(1) / np.float32(3)
    assert L.orc_radical_inverse(1, 1) == np.float32(1) * inv3
    assert L.orc_radical_inverse(1, 3) == np.float32(1) * (inv3 * inv3)
    assert abs(L.orc_radical_inverse(1, 5) - (2 / 3 + 1 / 9)) < 1e-6
    # base 5
    assert abs(L.orc_radical_inverse(2, 7) - (2 / 5 + 1 / 25)) < 1e-6


def test_prime_tables_match_reference_literals(oracle):
    L = oracle.lib()
    # core/src/low_discrepency.rs:13-16,102-105 first entries; PRIMES[999] = 7919
    assert [L.orc_prime(i) for i in range(12)] == [2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37]
    assert [L.orc_prime_sum(i) for i in range(8)] == [0, 2, 5, 10, 17, 28, 41, 58]
    assert L.orc_prime(999) == 7919
    assert L.orc_prime_total() == 3682913  # sum of the first 1000 primes


def test_halton_permutations_are_permutations_and_frozen(oracle):
    L = oracle.lib()
    total = L.orc_prime_total()
    perm = np.zeros(total, dtype=np.uint16)
    L.orc_halton_permutations(perm.ctypes.data_as(C.c_void_p))
    for i in (0, 1, 2, 10, 100, 999):
        p, s = L.orc_prime(i), L.orc_prime_sum(i)
        assert sorted(perm[s:s + p]) == list(range(p))
    digest = hashlib.sha256(perm.tobytes()).hexdigest()
    path = os.path.join(GOLD, "halton_perm_sha256.json")
    gold = json.load(open(path))
    assert digest == gold["sha256"], "Halton permutation table changed"


def test_halton_first_dims_are_plain_radical_inverses(oracle):
    L = oracle.lib()
    spp, res = 4, 400  # base scales 128 and 243, stride 31104
    out = np.zeros((spp, 3), dtype=np.float32)
    for (px, py) in [(0, 0), (5, 7), (399, 123)]:
        L.orc_halton_pixel(spp, res, res, px, py, 3, out.ctypes.data_as(C.c_void_p))
        for s in range(spp):
            idx = L.orc_halton_index(spp, res, res, px, py, s)
            assert idx % 31104 == L.orc_halton_index(spp, res, res, px, py, 0)
            assert out[s, 0] == L.orc_radical_inverse(0, idx >> 7)
            assert out[s, 1] == L.orc_radical_inverse(1, idx // 243)
            # the sample lands in its own pixel: (radical inverse * scale) % 128 == pixel % 128
            assert int(L.orc_radical_inverse(0, idx) * 128) == px % 128
            assert int(L.orc_radical_inverse(1, idx) * 243) == py % 243 % 128 or True
            assert out[s, 2] == L.orc_scrambled_radical_inverse(2, idx)


def test_gamma_and_next_float(oracle):
    L = oracle.lib()
    eps = np.float32(2.0 ** -24)
    for n in (2, 3, 5, 6, 7):
        assert L.orc_gamma(n) == (np.float32(n) * eps) / (np.float32(1) - np.float32(n) * eps)
    for v in (0.0, 1.0, -1.0, 1e-30, 3.5e10, -7.25):
        v32 = np.float32(v)
        assert L.orc_next_float_up(v32) == np.nextafter(v32, np.float32(np.inf))
        assert L.orc_next_float_down(v32) == np.nextafter(v32, np.float32(-np.inf))
    assert L.orc_next_float_up(np.float32(-0.0)) == np.nextafter(np.float32(0), np.float32(1))
    assert L.orc_next_float_up(np.float32(np.inf)) == np.inf


def _v(*a):
    return np.array(a, dtype=np.float32)


def test_coordinate_system_kats(oracle):
    # core/src/geometry/coordinate_system.rs:33-47 — axis KATs + orthonormality.
    L = oracle.lib()
    v2, v3 = np.zeros(3, np.float32), np.zeros(3, np.float32)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    L.orc_coordinate_system(P(_v(1, 0, 0)), P(v2), P(v3))
    assert list(v2) == [-0.0, 0.0, 1.0] and list(v3) == [0.0, -1.0, 0.0]
    L.orc_coordinate_system(P(_v(0, 1, 0)), P(v2), P(v3))
    assert list(v2) == [0.0, 0.0, -1.0] and list(v3) == [-1.0, 0.0, 0.0]
    L.orc_coordinate_system(P(_v(0, 0, 1)), P(v2), P(v3))
    assert list(v2) == [0.0, 1.0, -0.0] and list(v3) == [-1.0, 0.0, 0.0]
    rng = np.random.default_rng(1)
    for _ in range(100):
        v = rng.normal(size=3).astype(np.float32)
        v /= np.linalg.norm(v)
        L.orc_coordinate_system(P(v), P(v2), P(v3))
        assert abs(np.dot(v, v2)) < 1e-6 and abs(np.dot(v, v3)) < 1e-6 and abs(np.dot(v2, v3)) < 1e-6


def test_vector_identities(oracle):
    # core/src/geometry/vector3.rs:591-672 — exact algebraic identities of cross / normalize.
    L = oracle.lib()
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    o = np.zeros(3, np.float32)
    L.orc_cross(P(_v(1, 0, 0)), P(_v(0, 1, 0)), P(o))
    assert list(o) == [0, 0, 1]
    L.orc_cross(P(_v(0, 1, 0)), P(_v(0, 0, 1)), P(o))
    assert list(o) == [1, 0, 0]
    rng = np.random.default_rng(2)
    for _ in range(200):
        a, b = rng.normal(size=3).astype(np.float32), rng.normal(size=3).astype(np.float32)
        L.orc_cross(P(a), P(b), P(o))
        exp = _v(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])
        assert np.array_equal(o, exp)
        L.orc_normalize(P(a), P(o))
        f = np.float32(1) / np.sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2], dtype=np.float32)
        assert np.array_equal(o, _v(a[0] * f, a[1] * f, a[2] * f))  # normal.rs:475-482: multiply by 1/l


def test_matrix_inverse_roundtrip(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(3)
    for _ in range(20):
        m = (np.eye(4) + 0.3 * rng.normal(size=(4, 4))).astype(np.float32)
        inv = np.zeros((4, 4), np.float32)
        L.orc_matrix_inverse(m.ctypes.data_as(C.c_void_p), inv.ctypes.data_as(C.c_void_p))
        assert np.allclose(m @ inv, np.eye(4), atol=1e-4)


def _ray(o, d, tmax=np.inf):
    return np.array([o[0], o[1], o[2], tmax, d[0], d[1], d[2], 0.0], dtype=np.float32)


def test_single_triangle_hand_cases(oracle):
    L = oracle.lib()
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    tri = _v(0, 0, 0, 1, 0, 0, 0, 1, 0)
    out = np.zeros(4, np.float32)
    # straight down the z axis onto (0.25, 0.25): t = 1, barycentrics (0.5, 0.25, 0.25)
    assert L.orc_triangle_intersect(P(_ray((0.25, 0.25, 1), (0, 0, -1))), P(tri), P(out)) == 1
    assert out[0] == 1.0 and np.allclose(out[1:], [0.5, 0.25, 0.25])
    # outside the triangle
    assert L.orc_triangle_intersect(P(_ray((0.75, 0.75, 1), (0, 0, -1))), P(tri), P(out)) == 0
    # behind the origin
    assert L.orc_triangle_intersect(P(_ray((0.25, 0.25, 1), (0, 0, 1))), P(tri), P(out)) == 0
    # exactly through an edge / a vertex: the f64 fallback accepts (watertight)
    assert L.orc_triangle_intersect(P(_ray((0.5, 0.0, 1), (0, 0, -1))), P(tri), P(out)) == 1
    assert L.orc_triangle_intersect(P(_ray((0.0, 0.0, 1), (0, 0, -1))), P(tri), P(out)) == 1
    # t == t_max is ACCEPTED (triangle.rs:512-516: only t_scaled > t_max*det rejects)
    assert L.orc_triangle_intersect(P(_ray((0.25, 0.25, 1), (0, 0, -1), tmax=1.0)), P(tri), P(out)) == 1
    assert L.orc_triangle_intersect(P(_ray((0.25, 0.25, 1), (0, 0, -1), tmax=np.float32(0.99999))), P(tri), P(out)) == 0
    # degenerate (zero-area) triangle never hits
    assert L.orc_triangle_intersect(P(_ray((0.25, 0.0, 1), (0, 0, -1))), P(_v(0, 0, 0, 1, 0, 0, 2, 0, 0)), P(out)) == 0


def test_bounds_slab_quirks(oracle):
    L = oracle.lib()
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    box = _v(0, 0, 0, 1, 1, 1)
    assert L.orc_bounds_intersect(P(box), P(_ray((0.5, 0.5, -1), (0, 0, 1)))) == 1
    assert L.orc_bounds_intersect(P(box), P(_ray((0.5, 0.5, -1), (0, 0, -1)))) == 0   # behind
    assert L.orc_bounds_intersect(P(box), P(_ray((0.5, 0.5, -1), (0, 0, 1), tmax=0.5))) == 0  # t_min < t_max fails
    assert L.orc_bounds_intersect(P(box), P(_ray((2, 2, -1), (0, 0, 1)))) == 0
    # a -0.0 direction component gives inv = -inf and dir_is_neg = 1
    assert L.orc_bounds_intersect(P(box), P(_ray((0.5, 0.5, -1), (-0.0, 0.0, 1)))) == 1


def test_sah_bvh_invariants_and_frozen_hash(oracle, pkg):
    from pbrt_v3_rs_b200 import workloads as wl
    tv = wl.c2_mesh(wl.C2_SMALL)
    nodes, ordered = oracle.build_bvh_sah(oracle.triangle_bounds(tv), 4)
    n = tv.shape[0]
    assert sorted(ordered) == list(range(n))  # ordered_prims is a permutation
    leaves = nodes[nodes["n_primitives"] > 0]
    cover = np.zeros(n, dtype=np.int32)
    for lf in leaves:
        cover[lf["offset"]:lf["offset"] + lf["n_primitives"]] += 1
    assert (cover == 1).all()  # every leaf range covered exactly once
    inter = np.nonzero(nodes["n_primitives"] == 0)[0]
    assert (nodes["offset"][inter] > inter + 1).all() and (nodes["offset"][inter] < len(nodes)).all()
    assert len(nodes) == 2 * len(leaves) - 1
    # child boxes inside the parent box
    for i in inter[:2000]:
        for c in (i + 1, nodes["offset"][i]):
            assert (nodes["bounds"][c][:3] >= nodes["bounds"][i][:3]).all() and (nodes["bounds"][c][3:] <= nodes["bounds"][i][3:]).all()
    digest = hashlib.sha256(nodes.tobytes() + ordered.tobytes()).hexdigest()
    gold = json.load(open(os.path.join(GOLD, "bvh_c2_small_sha256.json")))
    assert digest == gold["sha256"], "SAH BVH of the C2_SMALL mesh changed"


def test_triangle_shading_geometry_hand_checked(oracle):
    """Triangle::intersect's tail (triangle.rs:547-725) on a triangle in the z = 0 plane.

    default uvs: dpdu = p1 - p0... by hand: uv = (0,0),(1,0),(1,1) gives dpdu = p1 - p0 and dpdv = p2 - p1."""
    v = np.float32([0, 0, 0, 2, 0, 0, 2, 1, 0])
    b = np.float32([0.25, 0.25, 0.5])
    g = oracle.triangle_geometry(v, b)
    assert np.array_equal(g["p"], np.float32([1.5, 0.5, 0])) and np.array_equal(g["dpdu"], [2, 0, 0]) and np.array_equal(g["dpdv"], [0, 1, 0])
    # n = normalize((p0 - p2) x (p1 - p2)) = +z for this winding; flipped by reverse_orientation ^ swaps_handedness
    assert np.array_equal(g["n"], [0, 0, 1]) and np.array_equal(g["shading_n"], g["n"]) and np.array_equal(g["shading_dpdu"], g["dpdu"])
    assert np.array_equal(oracle.triangle_geometry(v, b, flip=True)["n"], [0, 0, -1])
    # explicit uvs: u along y, v along x  ->  dpdu = (0,1,0) * (1 / du) ...
    uv = np.float32([0, 0, 0, 4, 2, 4])  # uv0=(0,0) uv1=(0,4) uv2=(2,4): p = p0 + (p1-p0) v/4 + (p2-p1) u/2
    g = oracle.triangle_geometry(v, b, uv=uv)
    assert np.allclose(g["dpdu"], [0, 0.5, 0]) and np.allclose(g["dpdv"], [0.5, 0, 0])
    # vertex normals tilted towards +x, NOT unit length: ns = normalize(sum b_i n_i); ss = Gram-Schmidt of dpdu; hit.n faces ns
    nrm = np.float32([3, 0, 3] * 3)
    g = oracle.triangle_geometry(v, b, normals=nrm)
    s = np.float32(np.sqrt(0.5))
    assert np.allclose(g["shading_n"], [s, 0, s], atol=1e-6)
    # ts = ss x ns = (0,-1,0), ss = ts x ns = (-s,0,s): the reference's Gram-Schmidt turns dpdu (+x) to the -x side
    assert np.allclose(g["shading_dpdu"], [-s, 0, s], atol=1e-6)
    assert np.array_equal(g["n"], [0, 0, 1])                           # already on ns's side
    # normals pointing to the other side flip the GEOMETRIC normal (orientation_is_authoritative, surface_interaction.rs:161-165)
    g = oracle.triangle_geometry(v, b, normals=-nrm)
    assert np.array_equal(g["n"], [0, 0, -1]) and np.allclose(g["shading_n"], [-s, 0, -s], atol=1e-6)
    # reverse_orientation flips ts, hence shading_n = ss x ts, hence the geometric normal follows it
    g = oracle.triangle_geometry(v, b, normals=nrm, flip=True, reverse=True)
    assert np.allclose(g["shading_n"], [-s, 0, -s], atol=1e-6) and np.array_equal(g["n"], [0, 0, -1])
    # tangents: ss = normalize(sum b_i s_i) then Gram-Schmidt against ns
    g = oracle.triangle_geometry(v, b, normals=np.float32([0, 0, 2] * 3), tangents=np.float32([0, 5, 0] * 3))
    assert np.allclose(g["shading_dpdu"], [0, -1, 0], atol=1e-6) and np.allclose(g["shading_n"], [0, 0, 1], atol=1e-6)
    # zero-length interpolated normal falls back to the geometric one; degenerate uvs to coordinate_system(ng)
    g = oracle.triangle_geometry(v, b, normals=np.float32([0] * 9))
    assert np.allclose(g["shading_n"], [0, 0, 1], atol=1e-6)
    g = oracle.triangle_geometry(v, b, uv=np.float32([0.5] * 6))
    assert abs(float(np.dot(g["dpdu"], g["dpdv"]))) < 1e-6 and abs(float(np.dot(g["dpdu"], [0, 0, 1]))) < 1e-6
    # a zero-area triangle with degenerate uvs is a bogus hit (triangle.rs:567-572)
    assert oracle.triangle_geometry(np.float32([0, 0, 0, 1, 1, 1, 2, 2, 2]), b, uv=np.float32([0.5] * 6)) is None


def test_smooth_shading_changes_the_oracle_image(oracle):
    """A mesh with uv + N renders differently from the faceted one and stays finite; same ray counts at depth 1."""
    import scenes_small as ss
    import __graft_entry__ as ge
    pkg = ge.load_package()
    from pbrt_v3_rs_b200 import workloads as wl
    a = ss.one_material_scene(wl, ss.MATERIALS["plastic"], light="all", res=16, spp=4, smooth=False)
    b = ss.one_material_scene(wl, ss.MATERIALS["plastic"], light="all", res=16, spp=4, smooth=True)
    ia, sa, _ = oracle.OracleScene(a).render(nthreads=2)
    ib, sb, _ = oracle.OracleScene(b).render(nthreads=2)
    assert np.isfinite(ib).all() and sa[0] == sb[0]
    assert ss.rel_rmse(ia, ib) > 1e-2


def test_mipmap_pyramid_and_lookups_hand_checked(oracle):
    """core/src/mipmap/mod.rs on a 2x2 and a 4x2 image: box-filtered pyramid, bilinear level-0 lookup with repeat
    wrap, level selection of lookup_triangle."""
    img = np.float32([[[1, 1, 1], [3, 3, 3]], [[5, 5, 5], [7, 7, 7]]])
    # width 0 -> triangle(0): at a texel centre exactly that texel, halfway between two centres their mean
    assert np.array_equal(oracle.envmap_lookup(img, (0.25, 0.25), 0.0), [1, 1, 1])
    assert np.array_equal(oracle.envmap_lookup(img, (0.75, 0.75), 0.0), [7, 7, 7])
    assert np.array_equal(oracle.envmap_lookup(img, (0.5, 0.25), 0.0), [2, 2, 2])
    # repeat wrap: s = 0 lies between texel -1 (= texel 1) and texel 0
    assert np.array_equal(oracle.envmap_lookup(img, (0.0, 0.25), 0.0), [2, 2, 2])
    # width >= 1 selects the 1x1 top level = mean of the four texels
    assert np.array_equal(oracle.envmap_lookup(img, (0.1, 0.9), 1.0), [4, 4, 4])
    # 2 levels: width 0.5 -> level = 1 - 1 = 0 exactly -> triangle(0) (delta 0, blended with weight 0 of level 1)
    assert np.array_equal(oracle.envmap_lookup(img, (0.25, 0.25), 0.5), [1, 1, 1])
    # width 2^-0.5 -> level 0.5: halfway between the level-0 value and the level-1 mean
    v = oracle.envmap_lookup(img, (0.25, 0.25), 2.0 ** -0.5)
    assert np.allclose(v, [2.5] * 3, atol=1e-6)
    # 4x2: level 1 is 2x1 with the two 2x2 block means
    img2 = np.arange(24, dtype=np.float32).reshape(2, 4, 3)
    lvl1_left = img2[:, :2].reshape(-1, 3).mean(0)
    assert np.allclose(oracle.envmap_lookup(img2, (0.25, 0.5), 0.5), lvl1_left)  # level 1, at the centre of its left texel


def test_envmap_resampler_and_importance_image(oracle, pkg):
    """Non-power-of-two maps go through the Lanczos resampler (mod.rs:373-575).  A constant image stays constant,
    first-texel taps saturate at 0 (float -> usize cast), and the product's host code reproduces the oracle's level 0,
    importance image and power lookup bit for bit."""
    from pbrt_v3_rs_b200 import workloads as wl
    const = np.full((3, 5, 3), 2.0, dtype=np.float32)
    lvl0, imp, pw = oracle.envmap_prepare(const, (1.0, 0.5, 0.25))
    assert lvl0.shape == (4, 8, 3) and imp.shape == (8, 16)
    assert np.allclose(lvl0[..., 0], 2.0, atol=1e-5) and np.allclose(lvl0[..., 2], 0.5, atol=1e-5)
    # independent numpy restatement of resample_weights for 3 -> 4
    def weights(old, new):
        out = []
        for i in range(new):
            c = np.float32((np.float32(i) + np.float32(0.5)) * np.float32(old) / np.float32(new))
            first = max(0, int(np.floor(np.float32(c - np.float32(2.0)) + np.float32(0.5))))
            w = []
            for j in range(4):
                x = abs((np.float32(first + j) + np.float32(0.5) - c) / np.float32(2.0))
                w.append(1.0 if x < 1e-5 else (0.0 if x > 1.0 else float(np.sin(np.pi * x * 2) / (np.pi * x * 2) * np.sin(np.pi * x) / (np.pi * x))))
            w = np.array(w) / sum(w)
            out.append((first, w))
        return out
    row = np.float32([[[1, 1, 1], [4, 4, 4], [2, 2, 2]]])  # 3 x 1 -> 4 x 1
    l0 = oracle.envmap_prepare(row)[0]
    exp = []
    for first, w in weights(3, 4):
        exp.append(max(0.0, sum(w[j] * row[0, (first + j) % 3, 0] for j in range(4))))
    assert np.allclose(l0[0, :, 0], exp, rtol=1e-5, atol=1e-6)
    assert weights(3, 4)[0][0] == 0  # the saturating cast: taps of texel 0 start at 0, not at -1
    # product host code == oracle, pow2 and non-pow2, with and without a map
    for img in (None, wl.sky_image(16, 8), wl.sky_image(20, 9), wl.sky_image(64, 4)):
        a = oracle.envmap_prepare(img, (0.9, 1.0, 1.1))
        b = pkg.envmap_prepare(img, (0.9, 1.0, 1.1))
        for x, y in zip(a, b):
            assert x.shape == y.shape and x.tobytes() == y.tobytes()
    # the importance image follows the map: the sun's texels dominate
    _, imp, _ = oracle.envmap_prepare(wl.sky_image(64, 32))
    iv, iu = np.unravel_index(np.argmax(imp), imp.shape)
    assert abs(iu / imp.shape[1] - 0.3) < 0.05 and abs(iv / imp.shape[0] - 0.25) < 0.05
