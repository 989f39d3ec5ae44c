"""Textured "Kd" (matte / plastic): the 2-D checkerboard spectrum texture with its closed-form box filter over the footprint
the camera ray's differentials give (textures/src/checkerboard_2d.rs, core/src/interaction/surface_interaction.rs:203-277,
cameras/src/perspective_camera.rs:144-204).  The reference pins this code through its own renders
(tests/test_reference_renders.py: the ground quad of six shipped scenes); here:

  CPU  the oracle's texture function against an independent numpy restatement and hand-checkable values; the oracle's
       footprint against an analytic one (pinhole camera over a plane); scene-file loader == Python mirror
  GPU  CUDA path == oracle: images, ray counts and per-sample radiance for path / whitted / directlighting, with a lens,
       instanced geometry, plastic, Oren-Nayar, a texture that is black in one check, aamode none, null-material pass-through"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
import scenes_small as ss

F32 = np.float32


def _pkg():
    import __graft_entry__ as ge
    return ge.load_package()


def np_bump_int(x):
    h = np.floor(F32(x / F32(2)))
    return F32(h + F32(F32(2) * max(F32(F32(F32(x / F32(2)) - h) - F32(0.5)), F32(0))))


def np_checkerboard(t, u, v, der):
    """checkerboard_2d.rs:62-98 over uv_2d.rs:44-50 in f32, written from the reference independently of oracle_texture.h."""
    su, sv, du, dv = (F32(t.get(k, d)) for k, d in (("uscale", 1), ("vscale", 1), ("udelta", 0), ("vdelta", 0)))
    t1, t2 = np.array(t.get("tex1", (1, 1, 1)), F32), np.array(t.get("tex2", (0, 0, 0)), F32)
    dudx, dvdx, dudy, dvdy = (F32(x) for x in der)
    s, tt = F32(F32(su * F32(u)) + du), F32(F32(sv * F32(v)) + dv)
    point = t1 if (int(np.floor(s)) + int(np.floor(tt))) % 2 == 0 else t2
    if t.get("aamode", "closedform") == "none":
        return point
    ds = max(abs(F32(su * dudx)), abs(F32(su * dudy)))
    dt = max(abs(F32(sv * dvdx)), abs(F32(sv * dvdy)))
    s0, s1, t0, t1_ = F32(s - ds), F32(s + ds), F32(tt - dt), F32(tt + dt)
    if np.floor(s0) == np.floor(s1) and np.floor(t0) == np.floor(t1_):
        return point
    sint = F32(F32(np_bump_int(s1) - np_bump_int(s0)) / F32(F32(2) * ds))
    tint = F32(F32(np_bump_int(t1_) - np_bump_int(t0)) / F32(F32(2) * dt))
    area2 = F32(0.5) if (ds > 1 or dt > 1) else F32(F32(sint + tint) - F32(F32(F32(2) * sint) * tint))
    return (t1 * F32(F32(1) - area2) + t2 * area2).astype(F32)


def _eval(ol, pkg, t, u, v, der):
    keep = []
    arr = pkg.spectrum_texture_array([t], keep)
    d = np.array(der, dtype=F32)
    out = np.zeros(3, dtype=F32)
    ol.lib().orc_spectrum_texture_evaluate(C.cast(arr, C.c_void_p), float(u), float(v), ol._p(d), ol._p(out))
    return out


def test_oracle_checkerboard_matches_numpy_restatement():
    import oracle_lib as ol
    pkg = _pkg()
    rng = np.random.Generator(np.random.PCG64(11))
    t = dict(type="checkerboard", uscale=24.0, vscale=17.0, udelta=0.25, tex1=(0.3, 0.2, 0.1), tex2=(0.8, 0.9, 1.0))
    n_filtered = 0
    for _ in range(3000):
        u, v = rng.uniform(-1.5, 1.5, 2)
        der = rng.normal(0, 1, 4) * 10.0 ** rng.uniform(-5, -0.5)
        got, want = _eval(ol, pkg, t, u, v, der), np_checkerboard(t, u, v, der)
        assert got.tobytes() == want.tobytes(), (u, v, der, got, want)
        n_filtered += int(not (np.array_equal(want, np.array(t["tex1"], F32)) or np.array_equal(want, np.array(t["tex2"], F32))))
    assert n_filtered > 300  # the closed-form branch was exercised
    tn = dict(t, aamode="none")
    for _ in range(200):
        u, v = rng.uniform(-1.5, 1.5, 2)
        assert _eval(ol, pkg, tn, u, v, (0.1, 0.2, 0.3, 0.4)).tobytes() == np_checkerboard(tn, u, v, (0.1, 0.2, 0.3, 0.4)).tobytes()


def test_oracle_checkerboard_hand_values():
    import oracle_lib as ol
    pkg = _pkg()
    t = dict(type="checkerboard", uscale=2.0, vscale=2.0, tex1=(1, 1, 1), tex2=(0, 0, 0))
    z = (0, 0, 0, 0)
    assert _eval(ol, pkg, t, 0.1, 0.1, z).tolist() == [1, 1, 1]      # (0, 0): even
    assert _eval(ol, pkg, t, 0.6, 0.1, z).tolist() == [0, 0, 0]      # (1, 0): odd
    assert _eval(ol, pkg, t, -0.1, 0.1, z).tolist() == [0, 0, 0]     # (-1, 0): Rust's % keeps the sign, -1 != 0 -> tex2
    assert _eval(ol, pkg, t, 0.6, 0.6, z).tolist() == [1, 1, 1]
    # a footprint wider than a check in s or t: the filter gives up and returns the mean (checkerboard_2d.rs:90-94)
    assert _eval(ol, pkg, t, 0.3, 0.3, (0.6, 0, 0, 0)).tolist() == [0.5, 0.5, 0.5]
    # footprint straddling the s = 1 border symmetrically, t inside one check: half of each colour
    got = _eval(ol, pkg, t, 0.5, 0.25, (0.1, 0, 0, 0.01))
    assert np.allclose(got, 0.5, atol=1e-6)
    c = dict(type="constant", value=(0.25, 0.5, 0.75))
    assert _eval(ol, pkg, c, 0.3, 0.9, (1, 1, 1, 1)).tolist() == [0.25, 0.5, 0.75]


def _textured_scene(wl, integrator="path", aamode="closedform", lens=0.0, res=32, spp=8, material="matte", sampler="halton", black_check=False,
                    instanced=False, null_cover=False, glass=False, camera="perspective"):
    from pbrt_v3_rs_b200.scene import SceneDescription
    sd = SceneDescription()
    tex = sd.add_spectrum_texture("checkerboard", uscale=12.0, vscale=12.0, tex1=(0.0, 0.0, 0.0) if black_check else (0.3, 0.25, 0.2), tex2=(0.8, 0.85, 0.9),
                                  aamode=aamode)
    tex2 = sd.add_spectrum_texture("checkerboard", uscale=16.0, vscale=8.0, udelta=0.5, tex1=(0.9, 0.2, 0.1), tex2=(0.1, 0.3, 0.9), aamode=aamode)
    if material == "matte":
        ground = sd.add_material(type="matte", Kd=("texture", tex))
    elif material == "oren_nayar":
        ground = sd.add_material(type="matte", Kd=("texture", tex), sigma=25.0)
    else:
        ground = sd.add_material(type="plastic", Kd=("texture", tex), roughness=0.2)
    ball = sd.add_material(type="plastic", Kd=("texture", tex2)) if not glass else sd.add_material(type="glass", eta=1.5)
    gq = wl.ground_quad()
    guv = np.array([[0, 0, 1, 1, 1, 0], [0, 0, 0, 1, 1, 1]], dtype=F32)  # ground_quad(): (p0, p2, p1), (p0, p3, p2)
    sd.add_mesh(gq, ground, uv=guv)
    tv, uv, nrm = wl.displaced_sphere(24, 12, with_attrs=True)
    if instanced:
        obj = sd.add_object(tv, ball, uv=uv, normals=nrm)
        m = np.eye(4, dtype=F32)
        m[:3, :3] = np.array([[0.8, 0.0, 0.6], [0.0, 1.0, 0.0], [-0.6, 0.0, 0.8]], dtype=F32) * F32(0.9)
        m[:3, 3] = (0.4, 0.1, 0.3)
        sd.add_instance(obj, m)
    else:
        sd.add_mesh(tv, ball, uv=uv, normals=nrm)
    if null_cover:  # a material-less quad in front of the camera: the ray behind it is re-spawned and carries no differentials
        q = np.array([[-3, -3, -2.5], [3, -3, -2.5], [3, 3, -2.5], [-3, 3, -2.5]], dtype=F32)
        sd.add_mesh(np.stack([np.concatenate([q[0], q[1], q[2]]), np.concatenate([q[0], q[2], q[3]])]), -1)
    # delta lights only: every transcendental on the path is then exact on the device (an infinite light's acosf / atan2f would
    # move radiance VALUES by an ulp), so per-sample radiance can be required bit for bit
    sd.add_distant_light((1.1, 1.0, 0.9), (-0.3, 1.0, -0.5))
    sd.add_point_light((1.5, 3.0, -3.0), (30, 30, 30))
    sd.camera.update(eye=(0.0, 1.2, -4.0), look=(0.0, -0.4, 0.0), up=(0, 1, 0), fov=50.0, lensradius=lens, focaldistance=4.0, type=camera)
    if camera == "orthographic":
        sd.camera.update(screenwindow=(-2.5, 2.5, -2.5, 2.5))
    sd.film.update(xresolution=res, yresolution=res)
    sd.sampler.update(type=sampler, pixelsamples=spp)
    sd.integrator.update(name=integrator, maxdepth=4, lightsamplestrategy="power")
    return sd


def test_oracle_footprint_of_a_pinhole_over_a_plane():
    """A 2.5 cm checkerboard seen from 1 m up through a 32-pixel, 60-degree pinhole: a pixel's footprint (>= 3.3 cm) is
    wider than a check, so the closed form returns the mean of the two colours everywhere (checkerboard_2d.rs:90-94) - the
    image equals the render with Kd = 0.5.  At 16 spp the differentials shrink by sqrt(16) (sampler_integrator.rs:358): the
    near rows resolve the checks again.  Point sampling ("aamode" "none") sees pure black / white."""
    import oracle_lib as ol
    _pkg()
    from pbrt_v3_rs_b200.scene import SceneDescription

    def render(kd, spp, aamode="closedform"):
        sd = SceneDescription()
        if kd is None:
            kd = ("texture", sd.add_spectrum_texture("checkerboard", uscale=4000.0, vscale=4000.0, tex1=(1, 1, 1), tex2=(0, 0, 0), aamode=aamode))
        g = sd.add_material(type="matte", Kd=kd)
        half = 50.0
        p = np.array([[-half, 0, -half], [half, 0, -half], [half, 0, half], [-half, 0, half]], dtype=F32)
        sd.add_mesh(np.stack([np.concatenate([p[0], p[2], p[1]]), np.concatenate([p[0], p[3], p[2]])]), g, uv=np.array([[0, 0, 1, 1, 1, 0], [0, 0, 0, 1, 1, 1]], dtype=F32))
        sd.add_point_light((0.0, 5.0, 2.0), (100, 100, 100))
        sd.camera.update(eye=(0.0, 1.0, 0.0), look=(0.0, 0.0, 3.0), up=(0, 1, 0), fov=60.0)
        sd.film.update(xresolution=32, yresolution=32)
        sd.sampler.update(type="halton", pixelsamples=spp)
        sd.integrator.update(name="whitted", maxdepth=1)
        return ol.OracleScene(sd).render()[0][20:, :, 0]  # rows that see the ground

    half_grey = render((0.5, 0.5, 0.5), 1)
    assert np.all(half_grey > 0)
    assert np.array_equal(render(None, 1), half_grey)
    ps = render(None, 1, "none") / (2 * half_grey)
    # pure black or white (a pixel whose neighbour's sample sits exactly on the shared border holds two samples, film_tile.rs:73-76)
    assert ((ps == 0) | (ps == 1)).mean() > 0.9 and (ps == 0).any() and (ps == 1).any()
    r16 = render(None, 16) / render((0.5, 0.5, 0.5), 16)
    assert np.abs(r16[-4:] - 1).max() > 0.05 and np.abs(r16[:2] - 1).max() < 1e-5  # near rows resolve checks, far rows stay filtered


def test_oracle_specular_children_carry_differentials():
    """Whitted through a flat glass slab onto the 2.5 cm checkerboard: the ground is reached by a twice-refracted ray whose
    differentials come from specular_transmit (sampler_integrator.rs:164-227).  Their footprint is still wider than a check,
    so the image equals the render with Kd = 0.5 bit for bit; a child without differentials would point-sample black / white."""
    import oracle_lib as ol
    _pkg()
    from pbrt_v3_rs_b200.scene import SceneDescription

    def quad(y, up):
        p = np.array([[-50, y, -50], [50, y, -50], [50, y, 50], [-50, y, 50]], dtype=F32)
        return np.stack([np.concatenate([p[0], p[2], p[1]]), np.concatenate([p[0], p[3], p[2]])]) if up else np.stack([np.concatenate([p[0], p[1], p[2]]), np.concatenate([p[0], p[2], p[3]])])

    def render(kd, spp, aamode="closedform"):
        sd = SceneDescription()
        if kd is None:
            kd = ("texture", sd.add_spectrum_texture("checkerboard", uscale=4000.0, vscale=4000.0, tex1=(1, 1, 1), tex2=(0, 0, 0), aamode=aamode))
        g = sd.add_material(type="matte", Kd=kd)
        glass = sd.add_material(type="glass", eta=1.5)
        sd.add_mesh(quad(0.0, True), g, uv=np.array([[0, 0, 1, 1, 1, 0], [0, 0, 0, 1, 1, 1]], dtype=F32))
        sd.add_mesh(quad(0.5, True), glass)    # slab: top face (normal up) ...
        sd.add_mesh(quad(0.3, False), glass)   # ... and bottom face (normal down)
        sd.add_point_light((0.0, 0.2, 1.5), (10, 10, 10))  # below the slab: the ground's shadow rays are free
        sd.camera.update(eye=(0.0, 1.0, 0.0), look=(0.0, 0.0, 3.0), up=(0, 1, 0), fov=60.0)
        sd.film.update(xresolution=32, yresolution=32)
        sd.sampler.update(type="halton", pixelsamples=spp)
        sd.integrator.update(name="whitted", maxdepth=5)
        return ol.OracleScene(sd).render()[0][20:, :, 0]

    half_grey = render((0.5, 0.5, 0.5), 1)
    assert (half_grey > 0).mean() > 0.9
    assert np.array_equal(render(None, 1), half_grey)
    ps = render(None, 1, "none")
    lit = half_grey > 0
    assert (ps[lit] == 0).mean() > 0.2 and (ps[lit] > 1.5 * half_grey[lit]).mean() > 0.2  # point sampling: black or white checks


def test_loader_reads_spectrum_textures_like_the_mirror(tmp_path):
    import oracle_lib as ol
    pkg = _pkg()
    path = tmp_path / "t.pbrt"
    path.write_text('''LookAt 0 5 3  0 0 0  0 0 1
Camera "perspective" "float fov" 60
Sampler "halton" "integer pixelsamples" 4
Integrator "path" "integer maxdepth" 3
Film "image" "string filename" "x.pfm" "integer xresolution" [24] "integer yresolution" [24]
WorldBegin
  LightSource "point" "rgb I" [40 40 40] "point from" [-5 0 5]
  AttributeBegin
    Texture "grey" "spectrum" "constant" "rgb value" [.25 .5 .75]
    Texture "checks" "color" "checkerboard" "float uscale" [8] "float vscale" [6] "float udelta" [.5] "rgb tex1" [.3 .2 .1] "texture tex2" "grey"
    Material "plastic" "texture Kd" "checks" "rgb Ks" [.1 .1 .1]
    Shape "trianglemesh" "point P" [ -4 -4 0   4 -4 0   4 4 0   -4 4 0 ] "float st" [ 0 0   1 0   1 1   0 1 ] "integer indices" [ 0 1 2   0 2 3 ]
    Material "matte" "texture Kd" "grey"
    Shape "trianglemesh" "point P" [ -1 -1 1   1 -1 1   1 1 1   -1 1 1 ] "integer indices" [ 0 1 2   0 2 3 ]
  AttributeEnd
WorldEnd
''')
    loaded = pkg.load_pbrt(str(path))
    d = loaded.to_desc()
    assert d.n_spectrum_textures == 2 and d.n_materials == 2
    kd_tex = np.ctypeslib.as_array(C.cast(d.material_kd_tex, C.POINTER(C.c_int32)), (2,))
    assert kd_tex.tolist() == [1, -1]  # the constant texture is folded into the matte's Kd
    T = C.cast(d.spectrum_textures, C.POINTER(pkg.SpectrumTexture))[1]
    assert (T.type, T.su, T.sv, T.du, T.aa_closedform) == (pkg.STEX_CHECKERBOARD, 8.0, 6.0, 0.5, 1)
    assert np.allclose(list(T.tex1), (0.3, 0.2, 0.1)) and np.allclose(list(T.tex2), (0.25, 0.5, 0.75))
    M = C.cast(d.materials, C.POINTER(pkg.Material))
    assert np.allclose(list(M[1].kd), (0.25, 0.5, 0.75))
    img = ol.OracleScene(loaded).render()[0]
    assert np.isfinite(img).all() and img.max() > 0


# ---- GPU parity ---------------------------------------------------------------------------------------------------
def _pairs(res, spp):
    return np.array([(x, y, s) for y in range(res) for x in range(res) for s in range(spp)], dtype=np.int32)


CASES = [dict(), dict(material="plastic"), dict(material="oren_nayar"), dict(lens=0.08), dict(aamode="none"), dict(black_check=True), dict(instanced=True),
         dict(null_cover=True), dict(sampler="sobol"), dict(sampler="02sequence"), dict(integrator="whitted"), dict(integrator="directlighting"),
         dict(integrator="whitted", instanced=True, lens=0.05), dict(integrator="whitted", aamode="none", glass=True),
         # specular children carry the reflected / refracted differentials (sampler_integrator.rs:108-125, 164-227); the instanced ball
         # has vertex normals, whose dndu / dndv go through the instance's transform_normal
         dict(integrator="whitted", glass=True), dict(integrator="directlighting", glass=True, instanced=True), dict(integrator="whitted", glass=True, lens=0.05),
         # the other two cameras: their rays and their differentials (orthographic_camera.rs:95-150; Camera::generate_ray_differential's
         # finite differences for the environment camera, core/src/camera.rs:29-78)
         dict(camera="orthographic"), dict(camera="orthographic", lens=0.05), dict(camera="environment"), dict(camera="environment", integrator="whitted", glass=True)]


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join("%s=%s" % kv for kv in c.items()) or "default")
def test_textured_kd_matches_oracle(gpu, oracle, case):
    from pbrt_v3_rs_b200 import workloads as wl
    sd = _textured_scene(wl, **case)
    if case.get("sampler") == "02sequence":
        sd.sampler.update(dimensions=16)
    integ = gpu.PathIntegrator(sd)
    img = integ.render()
    osc = oracle.OracleScene(sd)
    ref, stats, _ = osc.render()
    assert np.isfinite(img).all()
    r = ss.rel_rmse(img, ref)
    assert r <= 1e-3, "relative RMSE %.3e" % r
    rc = integ.ray_counts()
    assert [int(v) for v in rc[:3]] == [int(v) for v in stats[:3]]
    ps = _pairs(32, 2)
    li, _ = integ.li(ps)
    oli = osc.li(ps)
    same = (li.view(np.uint32) == oli.view(np.uint32)).all(1)
    if case.get("lens", 0.0) > 0.0:
        floor = 0.0    # the thin lens' ray origins are equal to ~1e-6 only (tests/test_render_gpu.py::test_camera_rays_bit_exact): statistical check below
    elif case.get("integrator", "path") != "path":
        # tree integrators: nodes below depth 0 multiply the throughput on the way down, the reference on the way back up
        #  - every sample that sees the glass ball differs by an ulp
        floor = 0.8 if case.get("glass") else 0.95
    else:
        floor = 0.995
    assert same.mean() >= floor, "only %.4f of the per-sample radiances are bit-identical" % same.mean()
    assert np.isclose(li, oli, rtol=2e-3, atol=1e-5).all(1).mean() >= 0.999


@pytest.mark.gpu
def test_textures_show_in_the_image(gpu):
    """Sanity: the texture changes the picture (a constant-Kd scene differs), and closedform differs from point sampling."""
    from pbrt_v3_rs_b200 import workloads as wl
    a = gpu.PathIntegrator(_textured_scene(wl)).render()
    b = gpu.PathIntegrator(_textured_scene(wl, aamode="none")).render()
    assert ss.rel_rmse(a, b) > 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("integrator", ["path", "whitted"])
def test_wave_splitting_keeps_the_textured_image(gpu, integrator, monkeypatch):
    """The camera-ray differentials are per path of a WAVE (and, for the tree integrators, travel on the per-path stack): cutting
    the render into several waves must not change a bit of a textured image."""
    from pbrt_v3_rs_b200 import workloads as wl

    def film(log2):
        if log2 is None:
            monkeypatch.delenv("B200PT_WAVE_LOG2", raising=False)
        else:
            monkeypatch.setenv("B200PT_WAVE_LOG2", str(log2))
        sd = _textured_scene(wl, integrator=integrator, res=64, spp=8, glass=integrator == "whitted")
        return gpu.PathIntegrator(sd).render_rows()

    one, many = film(None), film(13)  # 32 768 paths: one wave vs four
    assert np.array_equal(one, many)
