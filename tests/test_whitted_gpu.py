"""GPU parity tests of the wavefront WhittedIntegrator (k_shade_whitted / k_resolve_whitted) against the oracle's
restatement of integrators/src/whitted.rs:60-126.  Same gates as the path integrator: per-pixel relative RMSE <= 1e-3 at
equal spp with the same Halton sequence, identical ray counts; per-sample radiance agrees to f32 rounding."""
import numpy as np
import pytest

import scenes_small as ss

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _pairs(res, spp):
    return np.array([(x, y, s) for y in range(res) for x in range(res) for s in range(spp)], dtype=np.int32)


def _whitted(wl, name, light, **kw):
    sd = ss.one_material_scene(wl, ss.MATERIALS[name], light=light, **kw)
    sd.integrator.update(name="whitted")
    return sd


@pytest.mark.parametrize("name", ["matte", "oren_nayar", "plastic", "glass", "rough_glass", "metal"])
@pytest.mark.parametrize("light", ["infinite", "point", "area", "all"])
def test_whitted_li_per_sample_matches_oracle(gpu, oracle, name, light):
    from pbrt_v3_rs_b200 import workloads as wl
    sd = _whitted(wl, name, light, res=16, spp=4, maxdepth=5)
    integ = gpu.PathIntegrator(sd)
    ps = _pairs(16, 4)
    li, rays = integ.li(ps)
    osc = oracle.OracleScene(sd)
    oli = osc.li(ps)
    assert rays.tobytes() == osc.camera_rays(ps).tobytes()
    assert np.isfinite(li).all()
    close = np.isclose(li, oli, rtol=1e-4, atol=1e-6).all(1)
    assert close.mean() >= 0.99, "only %.4f of the samples agree" % close.mean()
    if name != "glass":  # depth-0 nodes only: the radiance is formed in the reference's order, bit for bit
        exact = (li.view(np.uint32) == oli.view(np.uint32)).all(1)
        assert exact.mean() >= (0.9 if light in ("infinite", "all") else 0.999), exact.mean()  # acosf / atan2f of the environment lookup differ by an ulp


@pytest.mark.parametrize("name,light,filt,maxdepth", [("glass", "all", "box", 5), ("glass", "infinite", "box", 8), ("plastic", "all", "gaussian", 5),
                                                        ("matte", "area", "box", 5), ("glass", "point", "box", 2)])
def test_whitted_image_rel_rmse_and_ray_counts(gpu, oracle, name, light, filt, maxdepth):
    from pbrt_v3_rs_b200 import workloads as wl
    sd = _whitted(wl, name, light, res=48, spp=8, maxdepth=maxdepth, nu=60, nv=30, filt=filt)
    integ = gpu.PathIntegrator(sd)
    img = integ.render()
    ref, stats, _ = oracle.OracleScene(sd).render()
    assert img.shape == ref.shape and np.isfinite(img).all()
    r = ss.rel_rmse(img, ref)
    assert r <= TOL, "relative RMSE %.3e > %.0e" % (r, TOL)
    rc = integ.ray_counts()
    assert rc[0] == stats[0]
    assert int(rc[1]) == int(stats[1])  # closest-hit rays: the whole specular trees
    assert int(rc[2]) == int(stats[2])  # shadow rays (placeholder slots are not counted)


def _many_lights_scene(wl, n_strips, maxdepth):
    from pbrt_v3_rs_b200.scene import SceneDescription
    sd = SceneDescription()
    g = sd.add_material(type="matte", Kd=(0.5, 0.5, 0.5))
    gl = sd.add_material(type="glass", eta=1.5)
    sd.add_mesh(wl.ground_quad(), g)
    obj = sd.add_object(wl.displaced_sphere(24, 12, radius=0.5), gl)
    for k in range(4):
        M = np.eye(4, dtype=np.float32)
        M[:3, 3] = [(k % 2 - 0.5) * 1.6, -0.6, (k // 2 - 0.5) * 1.6]
        sd.add_instance(obj, M)
    lm = sd.add_material(type="matte", Kd=(0, 0, 0))
    xs = np.linspace(-2, 2, 5, dtype=np.float32)
    tris = []
    for i in range(4):
        for j in range(n_strips):
            z0, z1 = -2 + 0.5 * j, -2 + 0.5 * (j + 1)
            a, b, c, d = [xs[i], 3.0, z0], [xs[i + 1], 3.0, z0], [xs[i + 1], 3.0, z1], [xs[i], 3.0, z1]
            tris += [a + b + c, a + c + d]
    sd.add_mesh(np.array(tris, dtype=np.float32), lm, area_light=dict(L=(6, 6, 6)))
    sd.camera.update(eye=(0.0, 1.5, -4.5), look=(0.0, -0.4, 0.0), up=(0, 1, 0), fov=40.0)
    sd.film.update(xresolution=40, yresolution=40)
    sd.sampler.update(type="halton", pixelsamples=4)
    sd.integrator.update(name="whitted", maxdepth=maxdepth)
    return sd


def test_whitted_many_lights_and_instances(gpu, oracle):
    """Every light is sampled at every hit (whitted.rs:89): an emissive mesh of 16 triangles = 16 lights, over an instanced
    scene of glass spheres.  7 tree nodes x (2 x 16 + 4) dimensions stay inside the Halton sampler's 1000."""
    from pbrt_v3_rs_b200 import workloads as wl
    sd = _many_lights_scene(wl, 2, 3)
    integ = gpu.PathIntegrator(sd)
    img = integ.render()
    ref, stats, _ = oracle.OracleScene(sd).render()
    assert ss.rel_rmse(img, ref) <= TOL
    assert integ.ray_counts()[0] == stats[0] and int(integ.ray_counts()[1]) == int(stats[1])


def test_whitted_too_many_dimensions_fails_like_the_reference(gpu):
    """64 lights x 15 tree nodes need ~2000 sampler dimensions: the reference's HaltonSampler asserts at 1000
    (samplers/src/halton.rs:106-110); the device reports the same condition instead of reading past its tables."""
    from pbrt_v3_rs_b200 import workloads as wl
    sd = _many_lights_scene(wl, 8, 4)
    with pytest.raises(gpu.B200PTError, match="sampler dimensions"):
        gpu.PathIntegrator(sd).render()


def test_whitted_rejects_the_zero_two_sampler(gpu):
    from pbrt_v3_rs_b200 import workloads as wl
    sd = _whitted(wl, "matte", "point", res=8, spp=4)
    sd.sampler.update(type="02sequence", dimensions=32)
    with pytest.raises(gpu.B200PTError):
        gpu.PathIntegrator(sd).render()


def test_whitted_scene_file(gpu, oracle, tmp_path):
    """Integrator "whitted" through the scene-file loader (what every shipped reference scene selects)."""
    scene = tmp_path / "w.pbrt"
    scene.write_text('''
LookAt 0 2 -5  0 0 0  0 1 0
Camera "perspective" "float fov" [40]
Film "image" "integer xresolution" [32] "integer yresolution" [32] "string filename" ["w.pfm"]
Sampler "halton" "integer pixelsamples" [4]
Integrator "whitted" "integer maxdepth" [4]
WorldBegin
LightSource "point" "rgb I" [40 40 40] "point from" [2 4 -3]
AttributeBegin
  Material "glass" "float index" [1.5]
  Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-1 0 0  1 0 0  1 1.5 0  -1 1.5 0]
AttributeEnd
Material "matte" "rgb Kd" [0.5 0.4 0.3]
Shape "trianglemesh" "integer indices" [0 2 1 0 3 2] "point P" [-6 0 -6  6 0 -6  6 0 6  -6 0 6]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-6 0 3  6 0 3  6 5 3  -6 5 3]
WorldEnd
''')
    ls = gpu.load_pbrt(str(scene))
    assert ls.to_desc().integrator.type == gpu.INTEGRATOR_WHITTED
    integ = gpu.PathIntegrator(ls)
    img = integ.render()
    ref = oracle.OracleScene(ls).render()[0]
    assert img.mean() > 0 and ss.rel_rmse(img, ref) <= TOL


# ---- DirectLightingIntegrator (integrators/src/direct_lighting.rs:82-146) --------------------------------------------

def _direct(wl, name, light, strategy, **kw):
    sd = ss.one_material_scene(wl, ss.MATERIALS[name], light=light, **kw)
    sd.integrator.update(name="directlighting", strategy=strategy)
    return sd


@pytest.mark.parametrize("strategy", ["all", "one"])
@pytest.mark.parametrize("name", ["matte", "plastic", "glass", "rough_glass", "metal"])
@pytest.mark.parametrize("light", ["infinite", "point", "area", "all"])
def test_directlighting_li_per_sample_matches_oracle(gpu, oracle, name, light, strategy):
    from pbrt_v3_rs_b200 import workloads as wl
    sd = _direct(wl, name, light, strategy, res=16, spp=4, maxdepth=5)
    integ = gpu.PathIntegrator(sd)
    ps = _pairs(16, 4)
    li, rays = integ.li(ps)
    osc = oracle.OracleScene(sd)
    oli = osc.li(ps)
    assert rays.tobytes() == osc.camera_rays(ps).tobytes()
    assert np.isfinite(li).all()
    close = np.isclose(li, oli, rtol=2e-3, atol=1e-5).all(1)
    # the MIS half samples directions with sin / cos: an ulp there can flip which triangle the MIS ray hits
    assert close.mean() >= 0.999, "only %.4f of the samples agree" % close.mean()
    assert abs(li.mean() - oli.mean()) <= 0.02 * max(oli.mean(), 1e-3)


@pytest.mark.parametrize("name,light,strategy,filt,maxdepth", [("glass", "all", "all", "box", 5), ("plastic", "all", "one", "box", 5),
                                                                 ("matte", "area", "all", "gaussian", 3), ("metal", "infinite", "one", "box", 5)])
def test_directlighting_image_rel_rmse_and_ray_counts(gpu, oracle, name, light, strategy, filt, maxdepth):
    from pbrt_v3_rs_b200 import workloads as wl
    sd = _direct(wl, name, light, strategy, res=48, spp=8, maxdepth=maxdepth, nu=60, nv=30, filt=filt)
    integ = gpu.PathIntegrator(sd)
    img = integ.render()
    ref, stats, _ = oracle.OracleScene(sd).render()
    r = ss.rel_rmse(img, ref)
    assert r <= TOL, "relative RMSE %.3e > %.0e" % (r, TOL)
    rc = integ.ray_counts()
    assert rc[0] == stats[0]
    assert int(rc[1]) == int(stats[1]) and int(rc[2]) == int(stats[2])


def test_directlighting_scene_file(gpu, oracle, tmp_path):
    scene = tmp_path / "d.pbrt"
    scene.write_text('''
LookAt 0 2 -5  0 0 0  0 1 0
Camera "perspective" "float fov" [40]
Film "image" "integer xresolution" [32] "integer yresolution" [32] "string filename" ["d.pfm"]
Sampler "halton" "integer pixelsamples" [4]
Integrator "directlighting" "integer maxdepth" [3] "string strategy" ["one"]
WorldBegin
LightSource "point" "rgb I" [40 40 40] "point from" [2 4 -3]
LightSource "infinite" "rgb L" [0.5 0.5 0.6]
Material "matte" "rgb Kd" [0.5 0.4 0.3]
Shape "trianglemesh" "integer indices" [0 2 1 0 3 2] "point P" [-6 0 -6  6 0 -6  6 0 6  -6 0 6]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-1 0 0  1 0 0  1 1.5 0  -1 1.5 0]
WorldEnd
''')
    ls = gpu.load_pbrt(str(scene))
    d = ls.to_desc()
    assert d.integrator.type == gpu.INTEGRATOR_DIRECT and d.integrator.direct_strategy == gpu.DIRECT_ONE
    img = gpu.PathIntegrator(ls).render()
    ref = oracle.OracleScene(ls).render()[0]
    assert img.mean() > 0 and ss.rel_rmse(img, ref) <= TOL


@pytest.mark.parametrize("integrator", ["whitted", "directlighting"])
def test_glass_slab_closed_form_on_the_device(gpu, integrator):
    """L = Le * ((1 - F)^2 + (1 - F)^2 F^2) through a glass slab at normal incidence, F = 0.04 (see tests/test_whitted_cpu.py)."""
    Le = np.array(ss.SLAB_LE)
    F = ((1.5 - 1.0) / (1.5 + 1.0)) ** 2
    for maxdepth, expect in ((3, (1 - F) ** 2), (5, (1 - F) ** 2 * (1 + F ** 2)), (2, 0.0)):
        li, _ = gpu.PathIntegrator(ss.glass_slab_scene(maxdepth, integrator)).li(np.array([(1, 1, 0)], dtype=np.int32))
        assert np.allclose(li[0].astype(np.float64), Le * expect, rtol=1e-5, atol=1e-7), (maxdepth, li[0])
