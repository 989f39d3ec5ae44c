"""SpatialLightDistribution (core/src/light_distrib/spatial.rs), the path integrator's default "lightsamplestrategy".
Deterministic restatement: every lookup sees its voxel's distribution (the reference's lock-free table returns None ->
uniform sampling to threads that race a voxel's first computation; with one thread it never does)."""
import numpy as np
import pytest

import scenes_small as ss


def _scene(wl, strategy, res=24, spp=8, maxdepth=4, name="matte"):
    sd = ss.one_material_scene(wl, ss.MATERIALS[name], light="all", res=res, spp=spp, maxdepth=maxdepth, strategy="uniform")
    sd.integrator.update(lightsamplestrategy=strategy)
    return sd


def test_spatial_is_unbiased_and_differs_from_uniform(pkg, oracle):
    """Same expectation as uniform light sampling, different estimator: the images agree in the mean but not per pixel."""
    from pbrt_v3_rs_b200 import workloads as wl
    a = oracle.OracleScene(_scene(wl, "spatial", spp=64)).render()[0]
    b = oracle.OracleScene(_scene(wl, "uniform", spp=64)).render()[0]
    assert np.isfinite(a).all() and (a >= 0).all()
    assert abs(a.mean() - b.mean()) <= 0.03 * b.mean()
    assert not np.allclose(a, b, rtol=1e-3)


def test_spatial_is_thread_count_independent(pkg, oracle):
    from pbrt_v3_rs_b200 import workloads as wl
    a = oracle.OracleScene(_scene(wl, "spatial")).render(nthreads=1)[0]
    b = oracle.OracleScene(_scene(wl, "spatial")).render(nthreads=7)[0]
    assert np.array_equal(a, b)


def test_single_light_forces_uniform(pkg, oracle):
    """create_light_sample_distribution (light_distrib/mod.rs:59-70): one light => uniform, whatever was asked for."""
    from pbrt_v3_rs_b200 import workloads as wl
    def img(strategy):
        sd = ss.one_material_scene(wl, ss.MATERIALS["matte"], light="point", res=16, spp=4, strategy="uniform")
        sd.integrator.update(lightsamplestrategy=strategy)
        return oracle.OracleScene(sd).render()[0]
    assert np.array_equal(img("spatial"), img("uniform"))


def test_spatial_prefers_the_near_light(pkg, oracle):
    """Two equal point lights far apart over a floor: near each light its voxel distribution is dominated by that light
    (contribution ~ I / d^2), so the per-sample noise of a one-light-per-vertex estimator drops against uniform picking."""
    from pbrt_v3_rs_b200.scene import SceneDescription
    def render(strategy):
        sd = SceneDescription()
        m = sd.add_material(type="matte", Kd=(0.5, 0.5, 0.5))
        q = np.array([[-8, 0, -2], [8, 0, -2], [8, 0, 2], [-8, 0, 2]], dtype=np.float32)
        sd.add_mesh(np.stack([np.concatenate([q[0], q[2], q[1]]), np.concatenate([q[0], q[3], q[2]])]), m)
        sd.add_point_light((-6.0, 2.0, 0.0), (4, 4, 4))
        sd.add_point_light((6.0, 2.0, 0.0), (4, 4, 4))
        sd.camera.update(eye=(0.0, 14.0, 0.0), look=(0.0, 0.0, 0.0), up=(0, 0, 1), fov=60.0)
        sd.film.update(xresolution=32, yresolution=8)
        sd.sampler.update(type="halton", pixelsamples=64)
        sd.integrator.update(maxdepth=1, lightsamplestrategy=strategy)
        sc = oracle.OracleScene(sd)
        ps = np.array([(x, y, s) for y in range(8) for x in range(32) for s in range(64)], dtype=np.int32)
        return sc.li(ps).reshape(8, 32, 64, 3)[..., 0]
    u, s = render("uniform"), render("spatial")
    assert u.mean() > 0 and abs(u.mean() - s.mean()) <= 0.05 * u.mean()
    assert s.var(axis=2).mean() < 0.7 * u.var(axis=2).mean()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["matte", "plastic", "glass", "metal"])
def test_spatial_gpu_li_matches_oracle(gpu, oracle, name):
    from pbrt_v3_rs_b200 import workloads as wl
    sd = _scene(wl, "spatial", res=16, spp=4, maxdepth=5, name=name)
    integ = gpu.PathIntegrator(sd)
    ps = np.array([(x, y, s) for y in range(16) for x in range(16) for s in range(4)], dtype=np.int32)
    li, _ = integ.li(ps)
    oli = oracle.OracleScene(sd).li(ps)
    close = np.isclose(li, oli, rtol=2e-3, atol=1e-5).all(1)
    assert close.mean() >= 0.999, close.mean()
    li2, _ = integ.li(ps)  # second call: every voxel is already in the table
    assert np.array_equal(li, li2)


@pytest.mark.gpu
@pytest.mark.parametrize("name,filt", [("matte", "box"), ("plastic", "gaussian"), ("glass", "box")])
def test_spatial_gpu_image_rel_rmse(gpu, oracle, name, filt):
    from pbrt_v3_rs_b200 import workloads as wl
    sd = ss.one_material_scene(wl, ss.MATERIALS[name], light="all", res=48, spp=16, maxdepth=5, nu=60, nv=30, strategy="uniform", filt=filt)
    sd.integrator.update(lightsamplestrategy="spatial")
    integ = gpu.PathIntegrator(sd)
    img = integ.render()
    ref, stats, _ = oracle.OracleScene(sd).render()
    r = ss.rel_rmse(img, ref)
    assert r <= 1e-3, r
    rc = integ.ray_counts()
    assert rc[0] == stats[0] and int(rc[1]) == int(stats[1]) and int(rc[2]) == int(stats[2])


@pytest.mark.gpu
def test_spatial_gpu_instanced_scene_and_scene_file_default(gpu, oracle, tmp_path):
    """The scene-file default (no "lightsamplestrategy") is spatial (path.rs:314); instanced geometry uses world-space p."""
    scene = tmp_path / "s.pbrt"
    scene.write_text('''
LookAt 0 2 -5  0 0 0  0 1 0
Camera "perspective" "float fov" [40]
Film "image" "integer xresolution" [32] "integer yresolution" [32] "string filename" ["s.pfm"]
Sampler "halton" "integer pixelsamples" [8]
Integrator "path" "integer maxdepth" [3]
WorldBegin
LightSource "point" "rgb I" [40 40 40] "point from" [2 4 -3]
LightSource "point" "rgb I" [10 20 40] "point from" [-3 1 -1]
LightSource "infinite" "rgb L" [0.3 0.3 0.4]
Material "matte" "rgb Kd" [0.5 0.4 0.3]
Shape "trianglemesh" "integer indices" [0 2 1 0 3 2] "point P" [-6 0 -6  6 0 -6  6 0 6  -6 0 6]
ObjectBegin "wall"
  Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-1 0 0  1 0 0  1 1.5 0  -1 1.5 0]
  Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-1 0 0.5  1 0 0.5  1 1.5 0.5  -1 1.5 0.5]
ObjectEnd
AttributeBegin
  Translate 1.5 0 1
  ObjectInstance "wall"
AttributeEnd
AttributeBegin
  Translate -1.5 0 0
  Rotate 40 0 1 0
  ObjectInstance "wall"
AttributeEnd
WorldEnd
''')
    ls = gpu.load_pbrt(str(scene))
    assert ls.to_desc().integrator.light_strategy == gpu.LIGHTS_SPATIAL
    img = gpu.PathIntegrator(ls).render()
    ref = oracle.OracleScene(ls).render()[0]
    assert img.mean() > 0 and ss.rel_rmse(img, ref) <= 1e-3
