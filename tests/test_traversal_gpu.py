"""GPU parity tests (call through the C ABI): closest-hit primitive ids and t bit-identical to the
oracle, any-hit booleans identical, on the fixed ray sets of the C2 microbench; plus edge cases."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
MISS = 0xFFFFFFFF


def _raysets(wl, cfg, tv, accel):
    rays = wl.primary_rays(cfg["width"], cfg["height"])
    hits = accel.intersect_batch(rays)
    br = wl.bounce_rays(tv, rays, hits, rays.shape[0])
    return rays, hits, br, wl.shadow_rays(br)


def _check_closest(gpu_hits, o_hits, diag, oracle):
    same_prim = gpu_hits["prim"] == o_hits["prim"]
    same_t = gpu_hits["t"].view(np.uint32) == o_hits["t"].view(np.uint32)
    exempt = oracle.exempt_mask(o_hits, diag)
    bad = ~(same_prim & same_t) & ~exempt
    assert not bad.any(), "%d non-exempt rays differ (first %s)" % (bad.sum(), np.nonzero(bad)[0][:5])
    # report how many needed the exemption band at all
    return int((~(same_prim & same_t)).sum())


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5])
def test_closest_and_any_hit_bit_exact_small(gpu, oracle, variant):
    import torch
    from pbrt_v3_rs_b200 import workloads as wl
    cfg = wl.C2_SMALL
    tv = wl.c2_mesh(cfg)
    accel = gpu.BVHAccel.from_params({"maxnodeprims": 4}, tv)
    oacc = oracle.OracleAccel(accel.nodes, accel.ordered_prims, tv)
    for rays in _raysets(wl, cfg, tv, accel)[0:3:2]:
        d_r = torch.from_numpy(rays.view(np.float32).reshape(-1, 8)).cuda()
        d_h = torch.empty((rays.shape[0], 4), dtype=torch.float32, device="cuda")
        accel.intersect_batch_device(d_r.data_ptr(), rays.shape[0], d_h.data_ptr(), torch.cuda.current_stream().cuda_stream, variant)
        torch.cuda.synchronize()
        g = d_h.cpu().numpy().view(gpu.HIT_DTYPE).reshape(-1)
        o, diag, _ = oacc.intersect(rays)
        assert _check_closest(g, o, diag, oracle) == 0  # on this set no ray needs the exemption
        assert np.array_equal(g["b0"].view(np.uint32), o["b0"].view(np.uint32))
        assert np.array_equal(g["b1"].view(np.uint32), o["b1"].view(np.uint32))
    sr = _raysets(wl, cfg, tv, accel)[3]
    d_r = torch.from_numpy(sr.view(np.float32).reshape(-1, 8)).cuda()
    d_o = torch.empty(sr.shape[0], dtype=torch.uint8, device="cuda")
    accel.occluded_batch_device(d_r.data_ptr(), sr.shape[0], d_o.data_ptr(), torch.cuda.current_stream().cuda_stream, variant)
    torch.cuda.synchronize()
    oo, _ = oacc.occluded(sr)
    assert np.array_equal(d_o.cpu().numpy(), oo)


def test_matches_frozen_fixture(gpu):
    from pbrt_v3_rs_b200 import workloads as wl
    cfg = wl.C2_SMALL
    tv = wl.c2_mesh(cfg)
    accel = gpu.BVHAccel.from_params({"maxnodeprims": 4}, tv)
    g = np.load(os.path.join(GOLD, "c2_small_hits.npz"))
    rays, hits, br, sr = _raysets(wl, cfg, tv, accel)
    assert np.array_equal(hits["prim"], g["primary_prim"])
    assert np.array_equal(hits["t"].view(np.uint32), g["primary_t"].view(np.uint32))
    bh = accel.intersect_batch(br)
    assert np.array_equal(bh["prim"], g["bounce_prim"])
    assert np.array_equal(np.packbits(accel.occluded_batch(sr)), g["occluded"])


@pytest.mark.parametrize("mesh,max_prims", [("sphere", 1), ("sphere", 8), ("soup", 4), ("soup", 255)])
def test_other_meshes_and_leaf_sizes(gpu, oracle, mesh, max_prims):
    from pbrt_v3_rs_b200 import workloads as wl
    tv = wl.displaced_sphere(200, 100) if mesh == "sphere" else wl.triangle_soup(30000, edge=0.05)
    accel = gpu.BVHAccel.from_params({"maxnodeprims": max_prims}, tv)
    oacc = oracle.OracleAccel(accel.nodes, accel.ordered_prims, tv)
    rays = wl.primary_rays(256, 128)
    g = accel.intersect_batch(rays)
    o, diag, _ = oacc.intersect(rays)
    _check_closest(g, o, diag, oracle)
    if (g["prim"] != MISS).any():
        br = wl.bounce_rays(tv, rays, g, rays.shape[0])
        gb = accel.intersect_batch(br)
        ob, dg, _ = oacc.intersect(br)
        _check_closest(gb, ob, dg, oracle)
        sr = wl.shadow_rays(br)
        assert np.array_equal(accel.occluded_batch(sr), oacc.occluded(sr)[0])


def test_edge_cases(gpu, oracle):
    """Empty batch, empty accelerator, single triangle, exact-edge rays (f64 fallback), t == t_max ties,
    zero direction components (inv = +-inf), alpha flags, axis-aligned rays in a shared-vertex fan."""
    from pbrt_v3_rs_b200 import workloads as wl
    tv = np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0], [1, 0, 0, 1, 1, 0, 0, 1, 0]], dtype=np.float32)
    accel = gpu.BVHAccel.from_params({}, tv)
    oacc = oracle.OracleAccel(accel.nodes, accel.ordered_prims, tv)
    assert accel.intersect_batch(np.zeros(0, dtype=gpu.RAY_DTYPE)).shape == (0,)
    rays = np.zeros(9, dtype=gpu.RAY_DTYPE)
    rays["tmax"] = np.inf
    rays["d"] = (0, 0, -1)
    pts = [(0.25, 0.25), (0.5, 0.5), (0.0, 0.0), (1.0, 0.0), (0.5, 0.0), (0.75, 0.75), (2.0, 2.0), (1.0, 1.0), (0.3, 0.7)]
    rays["o"] = [(x, y, 1.0) for x, y in pts]
    rays["tmax"][5] = 1.0          # t == t_max: the triangle would accept, but the flat leaf box fails t_min < t_max
    g = accel.intersect_batch(rays)
    o, diag, _ = oacc.intersect(rays)
    assert np.array_equal(g["prim"], o["prim"]) and np.array_equal(g["t"].view(np.uint32), o["t"].view(np.uint32))
    assert g["prim"][6] == MISS and g["prim"][0] != MISS and g["prim"][5] == MISS
    # duplicated triangles: the second candidate has t == t_max and is ACCEPTED (triangle.rs:512-516),
    # so the last tested primitive wins — traversal order is observable and must match
    dup = np.concatenate([tv, tv, tv[:1] + np.float32([0, 0, 0.5] * 3)])
    for mp in (1, 4):
        ad = gpu.BVHAccel.from_params({"maxnodeprims": mp}, dup)
        od = oracle.OracleAccel(ad.nodes, ad.ordered_prims, dup)
        gd, odh = ad.intersect_batch(rays), od.intersect(rays)[0]
        assert np.array_equal(gd["prim"], odh["prim"]) and np.array_equal(gd["t"].view(np.uint32), odh["t"].view(np.uint32))
        assert (gd["prim"] != MISS).sum() >= 3
    # shared diagonal edge (0.5,0.5): both triangles accept with equal t -> the LAST tested wins in both
    assert g["prim"][1] == o["prim"][1]
    # single-ray entry points: Primitive::intersect lowers tmax
    r1 = rays[:1].copy()
    h = accel.intersect(r1)
    assert h is not None and h["t"] == 1.0
    assert accel.intersect_p(rays[0]) and not accel.intersect_p(rays[6])
    # world_bound = root bounds
    assert np.array_equal(accel.world_bound(), accel.nodes[0]["bounds"])
    # alpha == 0 primitives are invisible to intersect; shadowalpha == 0 only to intersect_p
    flags = np.array([gpu.PRIM_ALPHA_ZERO, gpu.PRIM_SHADOW_ALPHA_ZERO], dtype=np.uint32)
    a2 = gpu.BVHAccel(tv, accel.nodes, accel.ordered_prims, flags)
    o2 = oracle.OracleAccel(accel.nodes, accel.ordered_prims, tv, flags)
    g2 = a2.intersect_batch(rays)
    assert np.array_equal(g2["prim"], o2.intersect(rays)[0]["prim"])
    sr = rays.copy()
    sr["tmax"] = 5.0
    assert np.array_equal(a2.occluded_batch(sr), o2.occluded(sr)[0])
    assert not a2.occluded_batch(sr).any()
    # empty accelerator
    e = gpu.BVHAccel(np.zeros((0, 9), np.float32), np.zeros(0, gpu.NODE_DTYPE), np.zeros(0, np.uint32))
    assert (e.intersect_batch(rays)["prim"] == MISS).all() and not e.occluded_batch(rays).any()


@pytest.mark.parametrize("variant", [0, 2, 4, 5])
def test_axis_parallel_rays_mixed_with_general_rays(gpu, oracle, variant):
    """Warps holding rays with zero / denormal direction components (inv = +-inf: the 0 * inf NaN
    cases of the slab test) next to general rays: the min/max box test must hand such warps to the literal one."""
    import torch
    from pbrt_v3_rs_b200 import workloads as wl
    rng = np.random.Generator(np.random.PCG64(21))
    # axis-aligned quads on integer planes (box faces coincide with ray origins) + a displaced sphere
    quads = []
    for k in range(-2, 3):
        for ax in range(3):
            c = np.zeros((4, 3), dtype=np.float32)
            u, v = (ax + 1) % 3, (ax + 2) % 3
            c[:, ax] = k
            c[:, u] = [-2, 2, 2, -2]
            c[:, v] = [-2, -2, 2, 2]
            quads += [np.concatenate([c[0], c[1], c[2]]), np.concatenate([c[0], c[2], c[3]])]
    tv = np.concatenate([np.stack(quads).astype(np.float32), wl.displaced_sphere(40, 20, radius=1.5)])
    accel = gpu.BVHAccel.from_params({"maxnodeprims": 4}, tv)
    oacc = oracle.OracleAccel(accel.nodes, accel.ordered_prims, tv)
    n = 1 << 14
    rays = np.zeros(n, dtype=gpu.RAY_DTYPE)
    rays["o"] = rng.uniform(-3, 3, size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    rays["tmax"] = np.inf
    special = np.arange(n) % 5 == 0
    k = np.arange(n)
    d[special & (k % 3 == 0), 0] = 0.0
    d[special & (k % 3 == 1), 1] = -0.0
    d[special & (k % 7 == 0), 2] = 0.0
    # denormal: 1/d overflows to inf (kept off the rays whose other components were zeroed: an all-denormal direction
    # yields NaN hit distances, whose payload bits differ between x86 and sm_100)
    d[special & (k % 11 == 0) & (k % 3 == 2) & (k % 7 != 0), 0] = 1e-42
    rays["d"] = d
    snap = special & (k % 2 == 0)
    rays["o"][snap] = np.round(rays["o"][snap])     # origins on the integer planes of the quads
    d_r = torch.from_numpy(rays.view(np.float32).reshape(-1, 8)).cuda()
    d_h = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    accel.intersect_batch_device(d_r.data_ptr(), n, d_h.data_ptr(), torch.cuda.current_stream().cuda_stream, variant)
    torch.cuda.synchronize()
    g = d_h.cpu().numpy().view(gpu.HIT_DTYPE).reshape(-1)
    o, diag, _ = oacc.intersect(rays)
    assert np.array_equal(g["prim"], o["prim"]), np.nonzero(g["prim"] != o["prim"])[0][:8]
    assert np.array_equal(g["t"].view(np.uint32), o["t"].view(np.uint32)), np.nonzero(g["t"].view(np.uint32) != o["t"].view(np.uint32))[0][:8]
    assert (g["prim"][special] != MISS).sum() > 100
    sr = rays.copy()
    sr["tmax"] = 2.0
    d_o = torch.empty(n, dtype=torch.uint8, device="cuda")
    accel.occluded_batch_device(d_r.data_ptr(), n, d_o.data_ptr(), torch.cuda.current_stream().cuda_stream, variant)
    d_r2 = torch.from_numpy(sr.view(np.float32).reshape(-1, 8)).cuda()
    accel.occluded_batch_device(d_r2.data_ptr(), n, d_o.data_ptr(), torch.cuda.current_stream().cuda_stream, variant)
    torch.cuda.synchronize()
    assert np.array_equal(d_o.cpu().numpy(), oacc.occluded(sr)[0])


def test_full_size_microbench_parity(gpu, oracle):
    """BASELINE config C2 at full size (1M triangles): a 2^18-ray sample of each set is compared with
    the oracle bit-for-bit; the full 2^24 set is checked through size-independent properties."""
    import torch
    from pbrt_v3_rs_b200 import workloads as wl
    cfg = wl.C2_FULL
    tv = wl.c2_mesh(cfg)
    accel = gpu.BVHAccel.from_params({"maxnodeprims": 4}, tv)
    rays = wl.primary_rays(cfg["width"], cfg["height"])
    hits = accel.intersect_batch(rays)
    br = wl.bounce_rays(tv, rays, hits, rays.shape[0])
    bh = accel.intersect_batch(br)
    sr = wl.shadow_rays(br)
    occ = accel.occluded_batch(sr)
    # property: any-hit(ray) == closest-hit(ray) finds something, over the full set
    sh = accel.intersect_batch(sr)
    assert np.array_equal(occ.astype(bool), sh["prim"] != MISS)
    # property: hit points lie on the reported triangle (barycentric reconstruction equals o + t d)
    m = bh["prim"] != MISS
    tri = tv.reshape(-1, 3, 3)[bh["prim"][m].astype(np.int64)]
    b0, b1 = bh["b0"][m][:, None], bh["b1"][m][:, None]
    p_bary = b0 * tri[:, 0] + b1 * tri[:, 1] + (1 - b0 - b1) * tri[:, 2]
    p_ray = br["o"][m] + bh["t"][m][:, None] * br["d"][m]
    assert np.abs(p_bary - p_ray).max() < 1e-4
    # sampled bit-exact comparison
    oacc = oracle.OracleAccel(accel.nodes, accel.ordered_prims, tv)
    sel = np.random.default_rng(7).choice(rays.shape[0], 1 << 18, replace=False)
    for rs, gh in ((rays, hits), (br, bh)):
        o, diag, _ = oacc.intersect(rs[sel])
        _check_closest(gh[sel], o, diag, oracle)
    assert np.array_equal(occ[sel], oacc.occluded(sr[sel])[0])


def test_uvs_reach_the_degenerate_hit_rejection(gpu, oracle):
    """triangle.rs:551-572: with a mesh's own uvs a hit is rejected only when the uv-derived dpdu x dpdv (or the uv
    determinant) AND the geometric normal degenerate.  Random uvs incl. degenerate ones must not change any hit, and
    the device walk must agree with the oracle walk that uses the same uvs."""
    from pbrt_v3_rs_b200 import workloads as wl
    cfg = wl.C2_SMALL
    tv = wl.c2_mesh(cfg)
    rng = np.random.Generator(np.random.PCG64(8))
    uv = rng.uniform(0, 1, size=(tv.shape[0], 6)).astype(np.float32)
    uv[::7] = 0.25                      # degenerate parameterisation -> coordinate_system(ng) branch
    accel = gpu.BVHAccel.from_params({"maxnodeprims": 4}, tv, tri_uvs=uv)
    plain = gpu.BVHAccel.from_params({"maxnodeprims": 4}, tv)
    oacc = oracle.OracleAccel(accel.nodes, accel.ordered_prims, tv, tri_uvs=uv)
    rays = wl.primary_rays(cfg["width"], cfg["height"])
    g, o = accel.intersect_batch(rays), oacc.intersect(rays)[0]
    assert np.array_equal(g["prim"], o["prim"]) and np.array_equal(g["t"].view(np.uint32), o["t"].view(np.uint32))
    assert np.array_equal(g["prim"], plain.intersect_batch(rays)["prim"])
    sr = wl.shadow_rays(wl.bounce_rays(tv, rays, g, rays.shape[0]))
    assert np.array_equal(accel.occluded_batch(sr), oacc.occluded(sr)[0])


@pytest.mark.parametrize("variant", [0, 1])
def test_hlbvh_accelerator_bit_exact(gpu, oracle, variant):
    """splitmethod "hlbvh" (accelerators/src/bvh/hlbvh.rs): the device walk over the reference's HLBVH tree returns the
    oracle's hits over the same tree bit for bit (the tree is poor — see oracle/oracle_hlbvh.h — but it is the reference's)."""
    import torch
    from pbrt_v3_rs_b200 import workloads as wl
    tv = wl.displaced_sphere(120, 60)
    accel = gpu.BVHAccel.from_params({"splitmethod": "hlbvh", "maxnodeprims": 4}, tv)
    n2, o2 = oracle.build_bvh_hlbvh(gpu.triangle_bounds(tv), 4)
    assert accel.nodes.tobytes() == n2.tobytes() and np.array_equal(accel.ordered_prims, o2)
    oacc = oracle.OracleAccel(accel.nodes, accel.ordered_prims, tv)
    rays = wl.primary_rays(256, 128)
    hits = accel.intersect_batch(rays)
    br = wl.bounce_rays(tv, rays, hits, rays.shape[0])
    for rs in (rays, br):
        d_r = torch.from_numpy(rs.view(np.float32).reshape(-1, 8)).cuda()
        d_h = torch.empty((rs.shape[0], 4), dtype=torch.float32, device="cuda")
        accel.intersect_batch_device(d_r.data_ptr(), rs.shape[0], d_h.data_ptr(), torch.cuda.current_stream().cuda_stream, variant)
        torch.cuda.synchronize()
        g = d_h.cpu().numpy().view(gpu.HIT_DTYPE).reshape(-1)
        o, diag, _ = oacc.intersect(rs)
        _check_closest(g, o, diag, oracle)
    sr = wl.shadow_rays(br)
    oo, _ = oacc.occluded(sr)
    assert np.array_equal(accel.occluded_batch(sr), oo)


def test_count_work_matches_oracle_counters(gpu, oracle):
    """The roofline numerator (bench.py: 32 B x N_node + 36 B x N_tri per ray) comes from b200pt_count_work_device; its
    totals and per-ray counts must be the oracle's counters of the reference walk (BVHAccel::intersect / intersect_p,
    accelerators/src/bvh/mod.rs:173-283) on the same rays: 2^16 primary, 2^16 incoherent bounce and 2^16 shadow rays
    of the full-size C2 mesh."""
    import torch
    from pbrt_v3_rs_b200 import workloads as wl
    cfg = wl.C2_FULL
    tv = wl.c2_mesh(cfg)
    accel = gpu.BVHAccel.from_params({"maxnodeprims": 4}, tv)
    oacc = oracle.OracleAccel(accel.nodes, accel.ordered_prims, tv)
    rays = wl.primary_rays(cfg["width"], cfg["height"])
    sel = np.random.default_rng(3).choice(rays.shape[0], 1 << 16, replace=False)
    prim = np.ascontiguousarray(rays[sel])
    hits = accel.intersect_batch(prim)
    bounce = wl.bounce_rays(tv, prim, hits, prim.shape[0])
    shadow = wl.shadow_rays(bounce)
    for rs, any_hit in ((prim, False), (bounce, False), (shadow, True)):
        n = rs.shape[0]
        d_r = torch.from_numpy(rs.view(np.float32).reshape(-1, 8)).cuda()
        d_c = torch.zeros((n, 2), dtype=torch.int32, device="cuda")
        tot = gpu.count_work_device(accel, d_r.data_ptr(), n, any_hit=any_hit, d_per_ray_ptr=d_c.data_ptr())
        per_ray = d_c.cpu().numpy().view(np.uint32)
        oct_ = oacc.occluded(rs)[1] if any_hit else oacc.intersect(rs)[2]
        assert np.array_equal(per_ray, oct_), "per-ray (N_node, N_tri) differ from the oracle's counters"
        assert tot == (int(oct_[:, 0].astype(np.int64).sum()), int(oct_[:, 1].astype(np.int64).sum()))
        assert tot[0] > 10 * n  # a real walk, not a root miss
