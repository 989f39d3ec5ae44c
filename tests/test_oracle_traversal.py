"""Oracle traversal against its frozen fixture (tests/golden/c2_small_hits.npz) and against
size-independent properties of the domain."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _setup(pkg, oracle):
    from pbrt_v3_rs_b200 import workloads as wl
    cfg = wl.C2_SMALL
    tv = wl.c2_mesh(cfg)
    nodes, ordered = oracle.build_bvh_sah(oracle.triangle_bounds(tv), 4)
    return wl, cfg, tv, oracle.OracleAccel(nodes, ordered, tv)


def test_oracle_matches_frozen_fixture(pkg, oracle):
    wl, cfg, tv, acc = _setup(pkg, oracle)
    g = np.load(os.path.join(GOLD, "c2_small_hits.npz"))
    rays = wl.primary_rays(cfg["width"], cfg["height"])
    hits, _, ct = acc.intersect(rays)
    assert np.array_equal(hits["prim"], g["primary_prim"])
    assert np.array_equal(hits["t"].view(np.uint32), g["primary_t"].view(np.uint32))
    br = wl.bounce_rays(tv, rays, hits, rays.shape[0])
    bh, _, bct = acc.intersect(br)
    assert np.array_equal(bh["prim"], g["bounce_prim"])
    occ, _ = acc.occluded(wl.shadow_rays(br))
    assert np.array_equal(np.packbits(occ), g["occluded"])
    assert np.array_equal(np.stack([ct.sum(0), bct.sum(0)]), g["counters"])


def test_closest_hit_agrees_with_brute_force(pkg, oracle):
    """BVH traversal returns the same closest primitive as testing every triangle (no BVH)."""
    wl, cfg, tv, acc = _setup(pkg, oracle)
    rays = wl.primary_rays(32, 16)
    hits, diag, _ = acc.intersect(rays)
    # brute force = a one-leaf "BVH" holding every triangle
    n = tv.shape[0]
    pb = oracle.triangle_bounds(tv)
    root = np.zeros(1, dtype=oracle.NODE_DTYPE)
    root["bounds"][0, :3] = pb[:, :3].min(0)
    root["bounds"][0, 3:] = pb[:, 3:].max(0)
    root["n_primitives"] = 0  # u16 cannot hold n; emulate with chunks of 60000
    best_t = np.full(rays.shape[0], np.inf, dtype=np.float32)
    best_p = np.full(rays.shape[0], 0xFFFFFFFF, dtype=np.uint32)
    for s in range(0, n, 60000):
        m = min(60000, n - s)
        leaf = root.copy()
        leaf["offset"], leaf["n_primitives"] = 0, m
        a = oracle.OracleAccel(leaf, np.arange(m, dtype=np.uint32), tv[s:s + m])
        h, _, _ = a.intersect(rays)
        better = h["t"] < best_t
        best_t[better] = h["t"][better]
        best_p[better] = h["prim"][better] + s
    ok = (best_p == hits["prim"]) | oracle.exempt_mask(hits, diag)
    assert ok.all()
    assert np.array_equal(best_t[best_p == hits["prim"]], hits["t"][best_p == hits["prim"]])


def test_any_hit_consistent_with_closest_hit(pkg, oracle):
    """intersect_p(ray) is true exactly when intersect(ray) finds a hit (same t range)."""
    wl, cfg, tv, acc = _setup(pkg, oracle)
    rays = wl.primary_rays(64, 32)
    hits, _, _ = acc.intersect(rays)
    br = wl.bounce_rays(tv, rays, hits, 4096)
    sr = wl.shadow_rays(br)
    occ, _ = acc.occluded(sr)
    ch, _, _ = acc.intersect(sr)
    assert np.array_equal(occ.astype(bool), ch["prim"] != 0xFFFFFFFF)


def test_ray_scaling_property(pkg, oracle):
    """Scaling d by 2 halves t for power-of-two scales (exact in binary floating point)."""
    wl, cfg, tv, acc = _setup(pkg, oracle)
    rays = wl.primary_rays(32, 32)
    h1, _, _ = acc.intersect(rays)
    r2 = rays.copy()
    r2["d"] *= np.float32(2)
    h2, _, _ = acc.intersect(r2)
    assert np.array_equal(h1["prim"], h2["prim"])
    m = h1["prim"] != 0xFFFFFFFF
    assert np.array_equal(h1["t"][m], h2["t"][m] * np.float32(2))
