"""CPU tests of the oracle's renderer: properties the reference algorithm guarantees."""
import numpy as np
import pytest

import scenes_small as ss


def test_render_is_thread_count_independent_with_box_filter(pkg, oracle):
    from pbrt_v3_rs_b200 import workloads as wl
    sd = ss.one_material_scene(wl, ss.MATERIALS["plastic"], light="all", res=20, spp=4, strategy="power")
    a = oracle.OracleScene(sd).render(nthreads=1)[0]
    b = oracle.OracleScene(sd).render(nthreads=5)[0]
    assert np.array_equal(a, b)  # each pixel belongs to exactly one tile (box filter r = 0.5)


def test_no_lights_is_black_and_emitter_is_seen_directly(pkg, oracle):
    from pbrt_v3_rs_b200 import workloads as wl
    from pbrt_v3_rs_b200.scene import SceneDescription
    sd = ss.one_material_scene(wl, ss.MATERIALS["matte"], light="none", res=12, spp=2)
    assert not oracle.OracleScene(sd).render()[0].any()
    # a camera looking straight at a one-sided emitter sees exactly L (path.rs:123-127)
    sd = SceneDescription()
    m = sd.add_material(type="matte", Kd=(0, 0, 0))
    q = np.array([[-5, -5, 2], [5, -5, 2], [5, 5, 2], [-5, 5, 2]], dtype=np.float32)
    tris = np.stack([np.concatenate([q[0], q[2], q[1]]), np.concatenate([q[0], q[3], q[2]])])
    sd.add_mesh(tris, m, area_light=dict(L=(3, 2, 1)))
    sd.camera.update(eye=(0, 0, -1), look=(0, 0, 1), fov=30.0)
    sd.film.update(xresolution=8, yresolution=8)
    sd.sampler.update(pixelsamples=2)
    img = oracle.OracleScene(sd).render()[0]
    assert np.allclose(img, np.array([3, 2, 1], np.float32), rtol=1e-5)


def test_white_furnace_bound(pkg, oracle):
    """Energy conservation: a matte object under a constant environment L never reflects more than L."""
    from pbrt_v3_rs_b200 import workloads as wl
    sd = ss.one_material_scene(wl, dict(type="matte", Kd=(1, 1, 1)), light="none", res=16, spp=16, maxdepth=8)
    sd.add_infinite_light((1, 1, 1))
    img = oracle.OracleScene(sd).render()[0]
    assert img.max() <= 1.0 + 0.35 and img.min() >= 0.0  # MC noise allowance at 16 spp
    assert 0.5 < img.mean() <= 1.0 + 1e-3


@pytest.mark.parametrize("name", ["matte", "plastic", "glass", "metal", "rough_glass", "oren_nayar"])
def test_every_material_renders_finite(pkg, oracle, name):
    from pbrt_v3_rs_b200 import workloads as wl
    sd = ss.one_material_scene(wl, ss.MATERIALS[name], light="all", res=16, spp=4, strategy="power")
    img, stats, _ = oracle.OracleScene(sd).render()
    assert np.isfinite(img).all() and (img >= 0).all() and img.mean() > 0
    assert stats[0] == 16 * 16 * 4  # one camera ray per pixel sample


def test_gaussian_filter_and_crop_window(pkg, oracle):
    from pbrt_v3_rs_b200 import workloads as wl
    sd = ss.one_material_scene(wl, ss.MATERIALS["matte"], res=20, spp=2, filt="gaussian")
    sd.film["cropwindow"] = (0.25, 0.75, 0.1, 0.9)
    sc = oracle.OracleScene(sd)
    assert sc.shape() == (16, 10)
    img, stats, _ = sc.render(nthreads=1)
    assert np.isfinite(img).all() and img.mean() > 0
    assert stats[0] == (10 + 4) * (16 + 4) * 2  # sample bounds extend 2 px beyond the crop on each side


def test_li_batch_equals_render_accumulation(pkg, oracle):
    """Per-sample li values pushed through FilmTile::add_sample (film_tile.rs:62-108, box filter) reproduce the
    rendered image — including the reference's quirk that a sample with a jitter of exactly 0 also lands in the
    neighbouring pixel (window [ceil(p - 0.5 - r), floor(p - 0.5 + r)])."""
    import ctypes as C
    from pbrt_v3_rs_b200 import workloads as wl
    res, spp = 8, 4
    sd = ss.one_material_scene(wl, ss.MATERIALS["matte"], res=res, spp=spp)
    sc = oracle.OracleScene(sd)
    img = sc.render(nthreads=1)[0]
    ps = np.array([(x, y, s) for y in range(res) for x in range(res) for s in range(spp)], dtype=np.int32)
    li = sc.li(ps, nthreads=1).reshape(res, res, spp, 3)
    acc = np.zeros((res, res, 3), np.float64)
    wsum = np.zeros((res, res), np.float64)
    u = np.zeros((spp, 2), np.float32)
    doubles = 0
    for y in range(res):
        for x in range(res):
            oracle.lib().orc_halton_pixel(spp, res, res, x, y, 2, u.ctypes.data_as(C.c_void_p))
            for s in range(spp):
                dx, dy = np.float32(x) + u[s, 0] - np.float32(0.5), np.float32(y) + u[s, 1] - np.float32(0.5)
                for yy in range(max(int(np.ceil(dy - 0.5)), 0), min(int(np.floor(dy + 0.5)) + 1, res)):
                    for xx in range(max(int(np.ceil(dx - 0.5)), 0), min(int(np.floor(dx + 0.5)) + 1, res)):
                        acc[yy, xx] += li[y, x, s]
                        wsum[yy, xx] += 1
                        doubles += (xx, yy) != (x, y)
    assert doubles > 0  # the quirk is exercised at this resolution
    assert np.allclose(acc / wsum[..., None], img, rtol=2e-4, atol=1e-6)  # RGB->XYZ->RGB matrices are inverse only to ~1e-5


def _nine_spheres(wl, instanced, res=48, spp=8):
    from pbrt_v3_rs_b200.scene import SceneDescription
    sd = SceneDescription()
    m = sd.add_material(type="matte", Kd=(0.6, 0.6, 0.6))
    g = sd.add_material(type="matte", Kd=(0.4, 0.4, 0.4))
    sd.add_mesh(wl.ground_quad(), g)
    sph = wl.displaced_sphere(24, 12, radius=0.45)
    obj = sd.add_object(sph, m) if instanced else None
    rng = np.random.default_rng(4)
    for k in range(9):
        a = rng.uniform(0, 2 * np.pi)
        c, s_ = np.cos(a), np.sin(a)
        M = np.eye(4, dtype=np.float32)
        M[:3, :3] = rng.uniform(0.6, 1.3) * np.array([[c, 0, s_], [0, 1, 0], [-s_, 0, c]], dtype=np.float32)
        M[:3, 3] = [(k % 3 - 1) * 1.2, -0.6 + 0.3 * (k // 3), (k // 3 - 1) * 1.2]
        if instanced:
            sd.add_instance(obj, M)
        else:
            v = sph.reshape(-1, 3) @ M[:3, :3].T + M[:3, 3]
            sd.add_mesh(v.reshape(-1, 9).astype(np.float32), m)
    sd.add_infinite_light((1.0, 1.0, 1.0))
    sd.camera.update(eye=(0.0, 2.5, -5.0), look=(0.0, -0.3, 0.0), up=(0, 1, 0), fov=40.0)
    sd.film.update(xresolution=res, yresolution=res)
    sd.sampler.update(pixelsamples=spp)
    sd.integrator.update(maxdepth=3, lightsamplestrategy="uniform")
    return sd


def test_instanced_scene_agrees_with_baked_geometry(pkg, oracle):
    """TransformedPrimitive (transformed_primitive.rs:43-73): rendering instances of one object must agree with the same
    triangles transformed to world space up to rounding (a few pixels flip at silhouettes)."""
    from pbrt_v3_rs_b200 import workloads as wl
    a, sa, _ = oracle.OracleScene(_nine_spheres(wl, True)).render()
    b, sb, _ = oracle.OracleScene(_nine_spheres(wl, False)).render()
    assert sa[0] == sb[0]
    assert abs(a.mean() - b.mean()) <= 2e-3 * b.mean()
    close = np.isclose(a, b, rtol=1e-3, atol=1e-4).all(2)
    assert close.mean() >= 0.97, close.mean()


def test_orthographic_and_environment_camera_rays_closed_form(pkg, oracle):
    """OrthographicCamera (orthographic_camera.rs:64-93): all rays share the direction camera_to_world((0, 0, 1)), their origins
    lie on the film plane through the eye, one screen-window step per pixel.  EnvironmentCamera (environment_camera.rs:38-53):
    origin at the eye, direction (sin t cos p, cos t, sin t sin p) with t = pi y / yres, p = 2 pi x / xres in camera space."""
    from pbrt_v3_rs_b200.scene import SceneDescription
    import numpy as np

    def scene(kind):
        sd = SceneDescription()
        m = sd.add_material(type="matte")
        sd.add_mesh(np.array([[0, 0, 5, 1, 0, 5, 0, 1, 5]], dtype=np.float32), m)
        sd.add_point_light((0, 0, 0), (1, 1, 1))
        sd.camera.update(type=kind, eye=(1.0, 2.0, -3.0), look=(1.0, 2.0, 0.0), up=(0, 1, 0), screenwindow=(-2.0, 2.0, -1.0, 1.0))
        sd.film.update(xresolution=8, yresolution=4)
        sd.sampler.update(type="halton", pixelsamples=1, samplepixelcenter=True)
        return sd

    ps = np.array([(x, y, 0) for y in range(4) for x in range(8)], dtype=np.int32)
    rays = oracle.OracleScene(scene("orthographic")).camera_rays(ps)
    o, d = rays["o"].astype(np.float64), rays["d"].astype(np.float64)
    assert np.allclose(d, [0, 0, 1], atol=1e-6)  # looking down +z
    # pbrt's look_at is left-handed: raster x runs towards world -x here... the film plane is z = -3; pixel centres (x + .5, y + .5)
    assert np.allclose(o[:, 2], -3.0, atol=1e-5)
    xs, ys = o[:, 0].reshape(4, 8), o[:, 1].reshape(4, 8)
    assert np.allclose(np.abs(np.diff(xs, axis=1)), 4.0 / 8, atol=1e-5) and np.allclose(np.abs(np.diff(ys, axis=0)), 2.0 / 4, atol=1e-5)
    assert np.allclose(xs.mean(), 1.0, atol=1e-5) and np.allclose(ys.mean(), 2.0, atol=1e-5)  # centred on the eye
    rays = oracle.OracleScene(scene("environment")).camera_rays(ps)
    o, d = rays["o"].astype(np.float64), rays["d"].astype(np.float64)
    assert np.allclose(o, [1.0, 2.0, -3.0], atol=1e-5)
    t = np.pi * (ps[:, 1] + 0.5) / 4.0
    p = 2 * np.pi * (ps[:, 0] + 0.5) / 8.0
    cam = np.stack([np.sin(t) * np.cos(p), np.cos(t), np.sin(t) * np.sin(p)], axis=1)
    # camera space -> world for this look_at: z forward (+z world), y up; x = right-handedness of pbrt's look_at (left-handed: x -> -x)
    assert np.allclose(np.abs(d[:, 1]), np.abs(cam[:, 1]), atol=1e-5) and np.allclose(np.abs(d[:, 2]), np.abs(cam[:, 2]), atol=1e-5)
    assert np.allclose(np.abs(d[:, 0]), np.abs(cam[:, 0]), atol=1e-5) and np.allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-5)
    assert np.allclose(d[:, 1], cam[:, 1], atol=1e-5) and np.allclose(d[:, 2], cam[:, 2], atol=1e-5)  # y and z keep their sign
