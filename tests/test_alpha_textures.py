"""Alpha-masked triangles: float textures as a mesh's "alpha" / "shadowalpha"
(shapes/src/triangle.rs:278-312, 587-607, 840-899; textures/src/{checkerboard_2d,dots,imagemap}.rs).

CPU part: the oracle's texture restatement against an independent numpy restatement and hand-checkable cases, the
scene-file loader against the Python mirror.  GPU part: closest-hit / any-hit parity and renders on the geometry of the
reference's scenes/shapes/triangles-alpha-mask.pbrt (a unit cube with "st" coordinates over a ground quad)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
import scenes_small as ss  # noqa: E402

F32 = np.float32


# ---- independent numpy restatement of Perlin's improved noise as core/src/texture/common.rs evaluates it (f32) ----------
def np_noise(perm, x, y, z):
    x, y, z = F32(x), F32(y), F32(z)
    P = lambda i: int(perm[int(i) & 255])
    ix, iy, iz = int(np.floor(x)), int(np.floor(y)), int(np.floor(z))
    dx, dy, dz = F32(x - F32(ix)), F32(y - F32(iy)), F32(z - F32(iz))
    ix, iy, iz = ix & 255, iy & 255, iz & 255

    def grad(a, b, c, gx, gy, gz):
        h = P(P(P(a) + b) + c) & 15
        u = gx if (h < 8 or h in (12, 13)) else gy
        v = gy if (h < 4 or h in (12, 13)) else gz
        return F32((-u if h & 1 else u) + (-v if h & 2 else v))

    def weight(t):
        t3 = F32(F32(t * t) * t)
        t4 = F32(t3 * t)
        return F32(F32(F32(F32(F32(6.0) * t4) * t) - F32(F32(15.0) * t4)) + F32(F32(10.0) * t3))

    def lerp(t, a, b):
        return F32(F32(F32(F32(1.0) - t) * a) + F32(t * b))

    one = F32(1.0)
    w = [[[grad(ix + i, iy + j, iz + k, F32(dx - one) if i else dx, F32(dy - one) if j else dy, F32(dz - one) if k else dz) for k in (0, 1)] for j in (0, 1)] for i in (0, 1)]
    wx, wy, wz = weight(dx), weight(dy), weight(dz)
    x00, x10 = lerp(wx, w[0][0][0], w[1][0][0]), lerp(wx, w[0][1][0], w[1][1][0])
    x01, x11 = lerp(wx, w[0][0][1], w[1][0][1]), lerp(wx, w[0][1][1], w[1][1][1])
    return lerp(wz, lerp(wy, x00, x10), lerp(wy, x01, x11))


def np_dots(perm, t, u, v):
    s, tt = F32(F32(F32(t["uscale"]) * F32(u)) + F32(t.get("udelta", 0.0))), F32(F32(F32(t["vscale"]) * F32(v)) + F32(t.get("vdelta", 0.0)))
    sc, tc = np.floor(F32(s + F32(0.5))), np.floor(F32(tt + F32(0.5)))
    outside_dot, inside_dot = F32(t.get("inside", 1.0)), F32(t.get("outside", 0.0))  # the reference's swapped constructor call
    if np_noise(perm, F32(sc + F32(0.5)), F32(tc + F32(0.5)), 0.5) > 0:
        radius = F32(0.35)
        shift = F32(F32(0.5) - radius)
        cs = F32(sc + F32(shift * np_noise(perm, F32(sc + F32(1.5)), F32(tc + F32(2.8)), 0.5)))
        ct = F32(tc + F32(shift * np_noise(perm, F32(sc + F32(4.5)), F32(tc + F32(9.8)), 0.5)))
        ds, dt = F32(s - cs), F32(tt - ct)
        if F32(F32(ds * ds) + F32(dt * dt)) < F32(radius * radius):
            return inside_dot
    return outside_dot


def _tex_struct(pkg, t):
    keep = []
    arr = pkg.float_texture_array([t], keep)
    return arr, keep


def _pkg():
    import __graft_entry__ as ge
    return ge.load_package()


def test_noise_perm_is_perlins_permutation():
    perm = _pkg().noise_perm()
    assert perm.shape == (256,) and sorted(perm.tolist()) == list(range(256))
    assert perm[:8].tolist() == [151, 160, 137, 91, 90, 15, 131, 13]  # Perlin, "Improving Noise" (2002), reference implementation


def test_oracle_noise_matches_numpy_restatement_and_lattice_zeros():
    import oracle_lib as ol
    perm = _pkg().noise_perm()
    rng = np.random.Generator(np.random.PCG64(3))
    for x, y, z in rng.uniform(-300, 300, size=(400, 3)).astype(F32):
        got = ol.lib().orc_noise_3d(ol._p(perm), float(x), float(y), float(z))
        assert F32(got).view(np.uint32) == np_noise(perm, x, y, z).view(np.uint32), (x, y, z)
    for p in [(0, 0, 0), (3, -7, 11), (255, 256, -256)]:  # gradient noise vanishes on the integer lattice
        assert ol.lib().orc_noise_3d(ol._p(perm), *map(float, p)) == 0.0


def test_oracle_checkerboard_dots_and_imagemap_closed_forms():
    import oracle_lib as ol
    pkg = _pkg()
    perm = pkg.noise_perm()
    ev = lambda arr, u, v: ol.lib().orc_float_texture_evaluate(C.cast(arr, C.c_void_p), ol._p(perm), float(u), float(v))
    chk = dict(type="checkerboard", uscale=4.0, vscale=2.0, tex1=1.0, tex2=0.0)
    arr, keep = _tex_struct(pkg, chk)
    for u, v in [(0.1, 0.1), (0.3, 0.1), (0.3, 0.6), (0.99, 0.99), (-0.1, 0.1), (-0.1, -0.1)]:
        cell = int(np.floor(F32(4.0) * F32(u))) + int(np.floor(F32(2.0) * F32(v)))
        want = 1.0 if int(np.fmod(cell, 2)) == 0 else 0.0  # truncating remainder: -1 % 2 == -1 -> tex2
        assert ev(arr, u, v) == want, (u, v)
    dots = dict(type="dots", uscale=10.0, vscale=10.0, inside=1.0, outside=0.0)
    arr, keep = _tex_struct(pkg, dots)
    rng = np.random.Generator(np.random.PCG64(5))
    vals = []
    for u, v in rng.uniform(0, 1, size=(600, 2)).astype(F32):
        got = ev(arr, u, v)
        assert got == np_dots(perm, dots, u, v), (u, v)
        vals.append(got)
    assert 0.05 < np.mean(np.array(vals) == 0.0) < 0.6  # "inside 1 outside 0" cuts the DOTS out (dots.rs:86 swaps them)
    img = np.array([[1.0, 2.0], [3.0, 5.0]], dtype=F32)  # rows t = 0, 1
    for wrap in ("repeat", "black", "clamp"):
        arr, keep = _tex_struct(pkg, dict(type="imagemap", texels=img, wrap=wrap))
        assert ev(arr, 0.25, 0.25) == 1.0 and ev(arr, 0.75, 0.25) == 2.0 and ev(arr, 0.25, 0.75) == 3.0 and ev(arr, 0.75, 0.75) == 5.0  # texel centres
        assert ev(arr, 0.5, 0.5) == 2.75  # bilinear blend of the four
    arr, keep = _tex_struct(pkg, dict(type="imagemap", texels=img, wrap="repeat"))
    assert ev(arr, 0.0, 0.25) == 1.5   # halfway between texel 0 and the wrapped texel 1 of row 0
    arr, keep = _tex_struct(pkg, dict(type="imagemap", texels=img, wrap="black"))
    assert ev(arr, 0.0, 0.25) == 0.5   # halfway between texel 0 and black
    arr, keep = _tex_struct(pkg, dict(type="imagemap", texels=img, wrap="clamp"))
    assert ev(arr, 0.0, 0.25) == 1.0


# ---- the geometry of scenes/shapes/triangles-alpha-mask.pbrt -----------------------------------------------------------
CUBE_P = np.array([[-1, -1, -1], [-1, 1, -1], [1, 1, -1], [1, -1, -1], [-1, -1, 1], [-1, 1, 1], [1, 1, 1], [1, -1, 1]], dtype=F32)
CUBE_ST = np.array([[0, 0], [0, 1], [1, 1], [1, 0], [1, 0], [1, 1], [0, 1], [0, 0]], dtype=F32)
CUBE_IDX = np.array([0, 1, 2, 3, 0, 2, 1, 5, 6, 2, 1, 6, 4, 5, 1, 0, 4, 1, 3, 2, 6, 7, 3, 6, 6, 5, 4, 6, 4, 7, 4, 0, 3, 7, 4, 3]).reshape(-1, 3)


def alpha_mask_scene(texture, shadow_texture=None, res=48, spp=4, integrator="path", light="point"):
    """Cube (rotated 135 degrees about z as in the scene file) with an alpha texture over a ground quad at z = -1."""
    from pbrt_v3_rs_b200 import workloads as wl
    from pbrt_v3_rs_b200.scene import SceneDescription
    sd = SceneDescription()
    red = sd.add_material(type="matte", Kd=(0.2, 0.01, 0.01))
    grey = sd.add_material(type="matte", Kd=(0.55, 0.55, 0.55))
    a = np.deg2rad(135.0)
    R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]], dtype=F32)
    P = (CUBE_P @ R.T).astype(F32)
    tv = P[CUBE_IDX].reshape(-1, 9)
    uv = CUBE_ST[CUBE_IDX].reshape(-1, 6)
    ta = sd.add_float_texture(**texture)
    kw = dict(alpha=("texture", ta))
    if shadow_texture is not None:
        kw["shadowalpha"] = ("texture", sd.add_float_texture(**shadow_texture))
    sd.add_mesh(tv, red, uv=uv, **kw)
    g = np.array([[-20, -20, -1], [20, -20, -1], [20, 20, -1], [-20, 20, -1]], dtype=F32)
    sd.add_mesh(g[[0, 1, 2, 0, 2, 3]].reshape(-1, 9), grey, uv=np.array([[0, 0, 1, 0, 1, 1], [0, 0, 1, 1, 0, 1]], dtype=F32))
    if light in ("point", "all"):
        sd.add_point_light((-5, 0, 5), (80, 90, 100))
    if light in ("area", "all"):
        q = np.array([[-1, -1, 6], [1, -1, 6], [1, 1, 6], [-1, 1, 6]], dtype=F32)
        sd.add_mesh(q[[0, 2, 1, 0, 3, 2]].reshape(-1, 9), grey, area_light=dict(L=(30, 30, 30)))
    sd.camera.update(eye=(0.0, 5.0, 3.0), look=(0.0, 0.0, 0.0), up=(0, 0, 1), fov=60.0)
    sd.film.update(xresolution=res, yresolution=res)
    sd.sampler.update(pixelsamples=spp)
    sd.integrator.update(name=integrator, maxdepth=4, lightsamplestrategy="uniform")
    return sd


TEXTURES = {
    "dots": dict(type="dots", uscale=10.0, vscale=10.0, inside=1.0, outside=0.0),  # as in the reference's scene file
    "checkerboard": dict(type="checkerboard", uscale=6.0, vscale=6.0, tex1=1.0, tex2=0.0),
    "imagemap": dict(type="imagemap", wrap="repeat", uscale=2.0, vscale=2.0,
                     texels=(np.random.Generator(np.random.PCG64(9)).uniform(0, 1, size=(8, 8)) > 0.45).astype(F32)),
}


def _rays_at_cube(n, seed=2):
    rng = np.random.Generator(np.random.PCG64(seed))
    from pbrt_v3_rs_b200 import RAY_DTYPE
    r = np.zeros(n, dtype=RAY_DTYPE)
    o = rng.normal(size=(n, 3)).astype(F32)
    o = (o / np.linalg.norm(o, axis=1, keepdims=True) * F32(6.0)).astype(F32)
    t = rng.uniform(-1.2, 1.2, size=(n, 3)).astype(F32)
    d = (t - o).astype(F32)
    r["o"], r["d"], r["tmax"] = o, d, np.inf
    return r


def test_oracle_alpha_masks_cut_holes():
    """Oracle alone: with the dots mask some rays that hit the opaque cube pass through it and reach a farther face or
    the ground; any-hit uses alpha and shadowalpha, closest-hit alpha only."""
    import oracle_lib as ol
    sd = alpha_mask_scene(TEXTURES["dots"], shadow_texture=TEXTURES["checkerboard"])
    sd.build_accel(ol.build_bvh_sah)
    rays = _rays_at_cube(4000)
    plain = ol.OracleAccel(sd.nodes, sd.ordered_prims, sd.tri_verts, sd.prim_flags & ~np.uint32(128), sd.tri_uvs)
    masked = ol.OracleAccel(sd.nodes, sd.ordered_prims, sd.tri_verts, sd.prim_flags & ~np.uint32(128), sd.tri_uvs)
    masked.set_alpha_textures(sd.float_textures, sd.prim_alpha_tex)
    h0, _, _ = plain.intersect(rays)
    h1, _, _ = masked.intersect(rays)
    changed = (h0["prim"] != h1["prim"]).mean()
    assert 0.05 < changed < 0.6, changed
    assert (h1["t"][h0["prim"] != h1["prim"]] >= h0["t"][h0["prim"] != h1["prim"]]).all()  # holes only let rays go farther (the cube's bottom face is coplanar with the ground)
    sh = rays.copy()
    sh["tmax"] = 0.999
    o0, _ = plain.occluded(sh)
    o1, _ = masked.occluded(sh)
    assert (o1 <= o0).all() and (o1 < o0).any()
    only_alpha = ol.OracleAccel(sd.nodes, sd.ordered_prims, sd.tri_verts, sd.prim_flags & ~np.uint32(128), sd.tri_uvs)
    pat = sd.prim_alpha_tex.copy()
    pat[:, 1] = -1
    only_alpha.set_alpha_textures(sd.float_textures, pat)
    o2, _ = only_alpha.occluded(sh)
    assert (o1 <= o2).all() and (o1 < o2).any()  # the shadowalpha checkerboard removes more occluders


def test_loader_float_textures_match_the_mirror(tmp_path):
    """Texture "name" "float" ... + "texture alpha" in a scene file give the same description as the Python mirror
    (identical flags, texture table, per-primitive indices), incl. the Attribute scoping of named textures."""
    pkg = _pkg()
    import oracle_lib as ol
    scene = tmp_path / "alpha.pbrt"
    scene.write_text('''
LookAt 0 5 3  0 0 0  0 0 1
Camera "perspective" "float fov" 60
Sampler "halton" "integer pixelsamples" 4
Integrator "path" "integer maxdepth" 4 "string lightsamplestrategy" "uniform"
Film "image" "string filename" "alpha.pfm" "integer xresolution" [48] "integer yresolution" [48]
WorldBegin
  LightSource "point" "rgb I" [80 90 100] "point from" [-5 0 5]
  AttributeBegin
    Texture "alpha" "float" "dots" "float inside" 1 "float outside" 0 "float uscale" 10 "float vscale" 10
    Texture "sh" "float" "checkerboard" "float uscale" 6 "float vscale" 6
    Rotate 135 0 0 1
    Material "matte" "rgb Kd" [.2 .01 .01]
    Shape "trianglemesh"
      "point P" [ -1 -1 -1   -1  1 -1   1  1 -1   1 -1 -1  -1 -1  1   -1  1  1   1  1  1   1 -1  1 ]
      "float st" [ 0 0   0 1   1 1   1 0  1 0   1 1   0 1   0 0 ]
      "integer indices" [ 0 1 2   3 0 2   1 5 6   2 1 6  4 5 1   0 4 1   3 2 6   7 3 6  6 5 4   6 4 7   4 0 3   7 4 3 ]
      "texture alpha" "alpha" "texture shadowalpha" "sh"
  AttributeEnd
  AttributeBegin
    Material "matte" "rgb Kd" [.55 .55 .55]
    Shape "trianglemesh" "point P" [ -20 -20 -1   20 -20 -1   20 20 -1   -20 20 -1 ] "float st" [ 0 0   1 0   1 1   0 1 ]
          "integer indices" [ 0 1 2   0 2 3 ] "texture alpha" "alpha"
  AttributeEnd
WorldEnd
''')
    loaded = pkg.load_pbrt(str(scene))
    d = loaded.to_desc()
    assert d.n_float_textures == 2
    ft = C.cast(d.float_textures, C.POINTER(pkg.FloatTexture))
    assert ft[0].type == pkg.TEX_DOTS and (ft[0].su, ft[0].sv) == (10.0, 10.0) and tuple(ft[0].value) == (1.0, 0.0)
    assert ft[1].type == pkg.TEX_CHECKERBOARD and tuple(ft[1].value) == (1.0, 0.0)
    n = d.n_prims
    flags = np.ctypeslib.as_array(C.cast(d.prim_flags, C.POINTER(C.c_uint32)), (n,))
    pat = np.ctypeslib.as_array(C.cast(d.prim_alpha_tex, C.POINTER(C.c_int32)), (n, 2))
    assert n == 14 and (flags[:12] & 128).all() and not (flags[12:] & 128).any()  # the ground's "alpha" is out of scope after AttributeEnd
    assert (pat[:12] == [0, 1]).all() and (pat[12:] == -1).all()
    mirror = alpha_mask_scene(TEXTURES["dots"], shadow_texture=TEXTURES["checkerboard"])
    md = mirror.to_desc()
    mflags = np.ctypeslib.as_array(C.cast(md.prim_flags, C.POINTER(C.c_uint32)), (n,))
    assert np.array_equal(flags, mflags)
    mtv = np.ctypeslib.as_array(C.cast(md.tri_verts, C.POINTER(C.c_float)), (n, 9))
    ltv = np.ctypeslib.as_array(C.cast(d.tri_verts, C.POINTER(C.c_float)), (n, 9))
    assert np.allclose(ltv, mtv, atol=1e-6)


# ---- GPU ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["dots", "checkerboard", "imagemap"])
def test_alpha_masked_traversal_matches_oracle(gpu, oracle, name):
    """Closest-hit ids / t / barycentrics and any-hit booleans bit-identical to the oracle with an alpha (and a
    shadowalpha) texture evaluated in the accept path, for the default kernel and the A/B variants."""
    sd = alpha_mask_scene(TEXTURES[name], shadow_texture=TEXTURES["checkerboard" if name != "checkerboard" else "dots"])
    sd.build_accel(None)
    base_flags = sd.prim_flags & ~np.uint32(128)
    acc = gpu.BVHAccel(sd.tri_verts, sd.nodes, sd.ordered_prims, base_flags, sd.tri_uvs)
    acc.set_alpha_textures(sd.float_textures, sd.prim_alpha_tex)
    orc = oracle.OracleAccel(sd.nodes, sd.ordered_prims, sd.tri_verts, base_flags, sd.tri_uvs)
    orc.set_alpha_textures(sd.float_textures, sd.prim_alpha_tex)
    rays = _rays_at_cube(1 << 15)
    oh, _, _ = orc.intersect(rays)
    sh = rays.copy()
    sh["tmax"] = 0.999
    oo, _ = orc.occluded(sh)
    plain = oracle.OracleAccel(sd.nodes, sd.ordered_prims, sd.tri_verts, base_flags, sd.tri_uvs)
    assert (plain.intersect(rays)[0]["prim"] != oh["prim"]).mean() > 0.005  # the mask matters on this ray set
    import torch
    n = rays.shape[0]
    d_r = torch.from_numpy(rays.view(np.float32).reshape(-1, 8)).cuda()
    d_s = torch.from_numpy(sh.view(np.float32).reshape(-1, 8)).cuda()
    st = torch.cuda.current_stream().cuda_stream
    for variant in (0, 1, 2, 3, 4, 5):
        d_h = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
        d_o = torch.zeros(n, dtype=torch.uint8, device="cuda")
        acc.intersect_batch_device(d_r.data_ptr(), n, d_h.data_ptr(), st, variant)
        acc.occluded_batch_device(d_s.data_ptr(), n, d_o.data_ptr(), st, variant)
        torch.cuda.synchronize()
        h = d_h.cpu().numpy().view(gpu.HIT_DTYPE).reshape(-1)
        assert np.array_equal(h["prim"], oh["prim"]), variant
        assert np.array_equal(h["t"].view(np.uint32), oh["t"].view(np.uint32)), variant
        assert np.array_equal(h["b0"].view(np.uint32), oh["b0"].view(np.uint32)) and np.array_equal(h["b1"].view(np.uint32), oh["b1"].view(np.uint32))
        assert np.array_equal(d_o.cpu().numpy(), oo), variant
    # the host-buffer entry points (b200pt_intersect_batch / b200pt_occluded_batch) run the default kernels
    h = acc.intersect_batch(rays)
    assert np.array_equal(h["prim"], oh["prim"]) and np.array_equal(acc.occluded_batch(sh), oo)


@pytest.mark.gpu
@pytest.mark.parametrize("name,integrator,light", [("dots", "path", "all"), ("dots", "whitted", "point"), ("checkerboard", "directlighting", "all"), ("imagemap", "path", "area")])
def test_alpha_masked_render_matches_oracle(gpu, oracle, name, integrator, light):
    sd = alpha_mask_scene(TEXTURES[name], shadow_texture=TEXTURES["checkerboard"] if name == "imagemap" else None, res=64, spp=4, integrator=integrator, light=light)
    integ = gpu.PathIntegrator(sd)
    osc = oracle.OracleScene(sd)
    img = integ.render()
    ref, stats, _ = osc.render()
    assert ss.rel_rmse(img, ref) <= 1e-3
    assert [int(x) for x in integ.ray_counts()] == [int(x) for x in stats[:3]]
    rng = np.random.Generator(np.random.PCG64(1))
    ps = np.stack([rng.integers(0, 64, 1500), rng.integers(0, 64, 1500), rng.integers(0, 4, 1500)], axis=1).astype(np.int32)
    li, rays = integ.li(ps)
    oli = osc.li(ps)
    assert rays.tobytes() == osc.camera_rays(ps).tobytes()
    assert (li.view(np.uint32) == oli.view(np.uint32)).all(1).mean() >= 0.99


@pytest.mark.gpu
def test_alpha_masked_instanced_object(gpu, oracle):
    """An object with an alpha texture instanced twice (two-level walk: the texture test sits in the object-level accept path)."""
    from pbrt_v3_rs_b200 import workloads as wl
    from pbrt_v3_rs_b200.scene import SceneDescription
    sd = SceneDescription()
    m = sd.add_material(type="matte", Kd=(0.5, 0.4, 0.3))
    sd.add_mesh(wl.ground_quad(), m)
    t = sd.add_float_texture(**TEXTURES["dots"])
    tv = CUBE_P[CUBE_IDX].reshape(-1, 9) * F32(0.5)
    obj = sd.add_object(tv, m, uv=CUBE_ST[CUBE_IDX].reshape(-1, 6), alpha=("texture", t))
    rng = np.random.Generator(np.random.PCG64(4))
    for k in range(2):
        M = wl.rigid_transform(rng, extent=0.1)
        M[:3, 3] += [k * 1.5 - 0.75, 0.2, 0.0]
        sd.add_instance(obj, M)
    sd.add_point_light((2, 4, -3), (40, 40, 40))
    sd.add_infinite_light((0.5, 0.5, 0.5))
    sd.camera.update(eye=(0.0, 2.5, -5.0), look=(0.0, 0.0, 0.0), up=(0, 1, 0), fov=40.0)
    sd.film.update(xresolution=48, yresolution=48)
    sd.sampler.update(pixelsamples=4)
    sd.integrator.update(maxdepth=4, lightsamplestrategy="uniform")
    integ = gpu.PathIntegrator(sd)
    osc = oracle.OracleScene(sd)
    img = integ.render()
    ref, stats, _ = osc.render()
    assert ss.rel_rmse(img, ref) <= 1e-3
    assert [int(x) for x in integ.ray_counts()] == [int(x) for x in stats[:3]]
