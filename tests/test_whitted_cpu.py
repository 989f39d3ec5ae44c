"""CPU tests of the oracle's WhittedIntegrator restatement (integrators/src/whitted.rs:60-126,
core/src/integrator/sampler_integrator.rs:79-238): hand-checkable radiances and structural properties."""
import numpy as np
import pytest

import scenes_small as ss


def _floor_scene(kd=(0.5, 0.25, 0.125), light_pos=(0.0, 2.0, 0.0), intensity=(8.0, 8.0, 8.0), res=9):
    """A big matte floor at y = 0 seen from straight above, one point light: every quantity has a closed form."""
    from pbrt_v3_rs_b200.scene import SceneDescription
    sd = SceneDescription()
    m = sd.add_material(type="matte", Kd=kd)
    q = np.array([[-50, 0, -50], [50, 0, -50], [50, 0, 50], [-50, 0, 50]], dtype=np.float32)
    sd.add_mesh(np.stack([np.concatenate([q[0], q[2], q[1]]), np.concatenate([q[0], q[3], q[2]])]), m)
    sd.add_point_light(light_pos, intensity)
    sd.camera.update(eye=(0.0, 5.0, 0.0), look=(0.0, 0.0, 0.0), up=(0, 0, 1), fov=20.0)
    sd.film.update(xresolution=res, yresolution=res)
    sd.sampler.update(type="halton", pixelsamples=1)
    sd.integrator.update(name="whitted", maxdepth=5)
    return sd


def test_point_light_on_matte_floor_closed_form(pkg, oracle):
    """L = (Kd / pi) * I / d^2 * cos(theta)  (whitted.rs:89-112 with LambertianReflection::f and PointLight::sample_li),
    evaluated in float64 at the point where the oracle's own camera ray meets the floor."""
    kd, I, h = np.array([0.5, 0.25, 0.125]), 8.0, 2.0
    sd = _floor_scene(kd=tuple(kd), light_pos=(0.0, h, 0.0), intensity=(I, I, I))
    sc = oracle.OracleScene(sd)
    ps = np.array([(x, y, 0) for y in range(9) for x in range(9)], dtype=np.int32)
    li = sc.li(ps, nthreads=1).astype(np.float64)
    rays = sc.camera_rays(ps)
    o, d = rays["o"].astype(np.float64), rays["d"].astype(np.float64)
    t = -o[:, 1] / d[:, 1]
    p = o + t[:, None] * d  # floor hit
    to_light = np.array([0.0, h, 0.0]) - p
    d2 = (to_light ** 2).sum(1)
    cos = to_light[:, 1] / np.sqrt(d2)
    expect = (kd[None, :] / np.pi) * (I / d2)[:, None] * cos[:, None]
    assert np.allclose(li, expect, rtol=2e-5)
    assert np.argmax(li[:, 0]) == 40  # brightest under the light (centre pixel)


def test_spot_light_on_matte_floor_closed_form(pkg, oracle):
    """SpotLight (lights/src/spot.rs:62-107): L = (Kd / pi) * I * falloff / d^2 * cos(theta), falloff = 1 inside
    coneangle - conedeltaangle, ((cos a - cos total) / (cos start - cos total))^4 in the penumbra, 0 outside - float64 closed
    form at the points where the oracle's camera rays meet the floor; the loader's light equals the mirror's."""
    kd, I, h = np.array([0.5, 0.25, 0.125]), 8.0, 2.0
    sd = _floor_scene(kd=tuple(kd), res=33)
    sd.lights.clear()
    sd.camera.update(fov=60.0)
    sd.add_spot_light((I, I, I), (0.0, h, 0.0), (0.5, 0.0, 0.2), coneangle=35.0, conedeltaangle=15.0)
    sc = oracle.OracleScene(sd)
    ps = np.array([(x, y, 0) for y in range(33) for x in range(33)], dtype=np.int32)
    li = sc.li(ps, nthreads=1).astype(np.float64)
    rays = sc.camera_rays(ps)
    o, d = rays["o"].astype(np.float64), rays["d"].astype(np.float64)
    p = o + (-o[:, 1] / d[:, 1])[:, None] * d
    to_light = np.array([0.0, h, 0.0]) - p
    d2 = (to_light ** 2).sum(1)
    axis = np.array([0.5, -h, 0.2]); axis /= np.linalg.norm(axis)
    cos_a = (-to_light / np.sqrt(d2)[:, None]) @ axis
    ct, cs = np.cos(np.deg2rad(35.0)), np.cos(np.deg2rad(20.0))
    fall = np.where(cos_a < ct, 0.0, np.where(cos_a >= cs, 1.0, ((cos_a - ct) / (cs - ct)) ** 4))
    expect = (kd[None, :] / np.pi) * (I * fall / d2)[:, None] * (to_light[:, 1] / np.sqrt(d2))[:, None]
    edge = np.abs(cos_a - ct) < 1e-4  # the cone's rim: f32 vs f64 may fall on different sides
    assert np.allclose(li[~edge], expect[~edge], rtol=2e-4, atol=1e-6)
    assert (fall == 0).any() and (fall == 1).any() and ((fall > 0) & (fall < 1)).sum() > 20


def test_shadowed_point_is_black_and_emitter_is_seen(pkg, oracle):
    from pbrt_v3_rs_b200.scene import SceneDescription
    sd = _floor_scene()
    # an occluder between the light (y = 2) and the floor: the floor under it is black (visibility.unoccluded == false)
    m = sd.add_material(type="matte", Kd=(0.5, 0.5, 0.5))
    q = np.array([[-30, 1, -30], [30, 1, -30], [30, 1, 30], [-30, 1, 30]], dtype=np.float32)
    sd.add_mesh(np.stack([np.concatenate([q[0], q[2], q[1]]), np.concatenate([q[0], q[3], q[2]])]), m)
    sd.camera.update(eye=(0.0, 0.5, 0.0), look=(0.0, 0.0, 0.0), up=(0, 0, 1), fov=20.0)
    img = oracle.OracleScene(sd).render(nthreads=1)[0]
    assert not img.any()
    # an emitter seen directly contributes Le once (whitted.rs:87), also under Whitted
    sd = SceneDescription()
    m = sd.add_material(type="matte", Kd=(0, 0, 0))
    q = np.array([[-5, -5, 2], [5, -5, 2], [5, 5, 2], [-5, 5, 2]], dtype=np.float32)
    sd.add_mesh(np.stack([np.concatenate([q[0], q[2], q[1]]), np.concatenate([q[0], q[3], q[2]])]), m, area_light=dict(L=(3, 2, 1)))
    sd.camera.update(eye=(0, 0, -1), look=(0, 0, 1), fov=30.0)
    sd.film.update(xresolution=8, yresolution=8)
    sd.sampler.update(pixelsamples=2)
    sd.integrator.update(name="whitted")
    img = oracle.OracleScene(sd).render()[0]
    assert np.allclose(img, np.array([3, 2, 1], np.float32), rtol=1e-5)


def test_glass_spawns_reflection_and_transmission_trees(pkg, oracle):
    """Glass without multiple lobes = SpecularReflection + SpecularTransmission (glass.rs:112-120): with maxdepth 1 the
    sphere shows direct light only (nothing for glass: its lobes are deltas, f = 0), with maxdepth 5 the trees add
    the environment seen through / mirrored by the sphere."""
    from pbrt_v3_rs_b200 import workloads as wl
    imgs, stats = {}, {}
    for md in (1, 2, 5):
        sd = ss.one_material_scene(wl, ss.MATERIALS["glass"], light="all", res=24, spp=2, maxdepth=md)
        sd.integrator.update(name="whitted")
        imgs[md], stats[md], _ = oracle.OracleScene(sd).render()
        assert np.isfinite(imgs[md]).all() and (imgs[md] >= 0).all()
    assert stats[1][1] == stats[1][0]  # maxdepth 1: camera rays only
    assert stats[2][1] > stats[1][1] and stats[5][1] > stats[2][1]  # deeper trees trace more closest-hit rays
    assert imgs[5].mean() > imgs[1].mean()
    # a binary tree: at most 2^maxdepth - 1 closest-hit rays per camera sample
    assert stats[5][1] <= stats[5][0] * (2 ** 5 - 1)


@pytest.mark.parametrize("name", ["matte", "plastic", "metal", "rough_glass"])
def test_non_specular_materials_are_direct_lighting_only(pkg, oracle, name):
    """No delta lobes -> specular_reflect / specular_transmit return zero: one closest-hit ray per sample, and every
    light of the scene is sampled at every hit (one shadow ray per light with a non-black contribution)."""
    from pbrt_v3_rs_b200 import workloads as wl
    sd = ss.one_material_scene(wl, ss.MATERIALS[name], light="all", res=16, spp=2, maxdepth=5)
    sd.integrator.update(name="whitted")
    img, stats, _ = oracle.OracleScene(sd).render()
    assert stats[1] == stats[0]
    assert stats[2] <= 4 * stats[0]  # infinite + point + two area-light triangles
    assert np.isfinite(img).all() and img.mean() > 0


# ---- DirectLightingIntegrator (integrators/src/direct_lighting.rs:82-146) --------------------------------------------

@pytest.mark.parametrize("strategy", ["all", "one"])
def test_directlighting_point_light_closed_form(pkg, oracle, strategy):
    """One delta light: estimate_direct = f * |cos| * Li / 1 for both strategies (uniform pick of one light has pdf 1)."""
    kd, I, h = np.array([0.5, 0.25, 0.125]), 8.0, 2.0
    sd = _floor_scene(kd=tuple(kd), light_pos=(0.0, h, 0.0), intensity=(I, I, I))
    sd.integrator.update(name="directlighting", strategy=strategy)
    sc = oracle.OracleScene(sd)
    ps = np.array([(x, y, 0) for y in range(9) for x in range(9)], dtype=np.int32)
    li = sc.li(ps, nthreads=1).astype(np.float64)
    rays = sc.camera_rays(ps)
    o, d = rays["o"].astype(np.float64), rays["d"].astype(np.float64)
    p = o + (-o[:, 1] / d[:, 1])[:, None] * d
    to_light = np.array([0.0, h, 0.0]) - p
    d2 = (to_light ** 2).sum(1)
    expect = (kd[None, :] / np.pi) * (I / d2)[:, None] * (to_light[:, 1] / np.sqrt(d2))[:, None]
    assert np.allclose(li, expect, rtol=2e-5)


def test_directlighting_all_equals_whitted_for_delta_lights_and_one_converges_to_all(pkg, oracle):
    """With only point lights the MIS half of estimate_direct is skipped, so "all" is Whitted's light loop with two
    get_2d() per light instead of one (the values are unused by point lights): identical images.  "one" picks a light
    uniformly and divides by 1/n: same expectation."""
    from pbrt_v3_rs_b200 import workloads as wl
    def scene(name, **kw):
        sd = ss.one_material_scene(wl, ss.MATERIALS["plastic"], light="point", res=20, spp=64, maxdepth=3)
        sd.add_point_light((-2.0, 2.5, -2.0), (10, 12, 14))
        sd.integrator.update(name=name, **kw)
        return oracle.OracleScene(sd).render()[0]
    w, a, o = scene("whitted"), scene("directlighting", strategy="all"), scene("directlighting", strategy="one")
    assert np.allclose(a, w, rtol=1e-5, atol=1e-7)
    assert abs(o.mean() - a.mean()) <= 0.03 * a.mean()


def test_directlighting_area_light_all_vs_path_depth1(pkg, oracle):
    """maxdepth-1 path tracing of a matte scene = emission + one uniform_sample_one_light; "one" direct lighting draws
    the same estimator (uniform light pick) -> equal expectation."""
    from pbrt_v3_rs_b200 import workloads as wl
    sd = ss.one_material_scene(wl, ss.MATERIALS["matte"], light="area", res=16, spp=128, maxdepth=1, strategy="uniform")
    p = oracle.OracleScene(sd).render()[0]
    sd.integrator.update(name="directlighting", strategy="one")
    d = oracle.OracleScene(sd).render()[0]
    assert abs(p.mean() - d.mean()) <= 0.03 * p.mean()


def test_whitted_glass_slab_normal_incidence_closed_form(pkg, oracle):
    """A camera ray hits a glass slab (eta 1.5) head-on; behind the slab a one-sided emitter (Le), in front of it
    nothing.  With F = ((eta - 1) / (eta + 1))^2 = 0.04 at normal incidence, SpecularTransmission carries
    (1 - F) * eta_i^2 / eta_t^2 per interface (specular_transmission.rs:70-77; the two eta factors cancel over the slab)
    and SpecularReflection carries F.  Paths that reach the emitter within maxdepth 5 (tree depth counts ray segments):
    T T (3 segments) and T R R T (5 segments):  L = Le * ((1-F)^2 + (1-F)^2 F^2)."""
    Le = np.array(ss.SLAB_LE)
    F = ((1.5 - 1.0) / (1.5 + 1.0)) ** 2
    for maxdepth, expect in ((3, (1 - F) ** 2), (4, (1 - F) ** 2), (5, (1 - F) ** 2 * (1 + F ** 2)), (2, 0.0)):
        sd = ss.glass_slab_scene(maxdepth)
        sc = oracle.OracleScene(sd)
        li = sc.li(np.array([(1, 1, 0)], dtype=np.int32), nthreads=1)[0].astype(np.float64)
        # the centre ray is exactly (0, 0, 1): normal incidence.  The emitter adds its own direct light on the glass: none
        # (delta lobes have f = 0), so only the specular trees carry radiance.
        assert np.allclose(li, Le * expect, rtol=1e-5, atol=1e-7), (maxdepth, li, Le * expect)


@pytest.mark.parametrize("integrator", ["whitted", "path", "directlighting"])
def test_mirror_reflects_an_emitter_closed_form(pkg, oracle, integrator):
    """MirrorMaterial (materials/src/mirror.rs): SpecularReflection(Kr, FresnelNoOp) - a camera ray that leaves a mirror floor
    towards a one-sided emitter returns exactly Kr * Le (f = Kr / |cos|, pdf = 1, times |cos|), for all three integrators."""
    from pbrt_v3_rs_b200.scene import SceneDescription
    sd = SceneDescription()
    mirror = sd.add_material(type="mirror", Kr=(0.9, 0.8, 0.7))
    black = sd.add_material(type="matte", Kd=(0, 0, 0))
    q = np.array([[-50, 0, -50], [50, 0, -50], [50, 0, 50], [-50, 0, 50]], dtype=np.float32)
    sd.add_mesh(np.stack([np.concatenate([q[0], q[2], q[1]]), np.concatenate([q[0], q[3], q[2]])]), mirror)
    e = np.array([[-40, -1, 6], [40, -1, 6], [40, 60, 6], [-40, 60, 6]], dtype=np.float32)  # faces -z: towards the reflected rays
    sd.add_mesh(np.stack([np.concatenate([e[0], e[2], e[1]]), np.concatenate([e[0], e[3], e[2]])]), black, area_light=dict(L=(3, 2, 1)))
    sd.camera.update(eye=(0.0, 1.0, -1.0), look=(0.0, 0.0, 0.0), up=(0, 1, 0), fov=20.0)
    sd.film.update(xresolution=5, yresolution=5)
    sd.sampler.update(type="halton", pixelsamples=2)
    sd.integrator.update(name=integrator, maxdepth=4)
    img = oracle.OracleScene(sd).render(nthreads=1)[0]
    assert np.allclose(img, np.array([0.9 * 3, 0.8 * 2, 0.7 * 1]), rtol=1e-5)


def test_projection_light_on_matte_floor_closed_form(pkg, oracle):
    """ProjectionLight (lights/src/projection.rs:115-171): a point light whose intensity is the image at the perspective
    projection of the light-space direction - here a 2 x 1 image (left half red, right half green; aspect 2 -> screen window
    [-2, 2] x [-1, 1]), the light 2 above a matte floor looking straight down with fov 90: floor point (x, z) sees the texel at
    s = (x' / d + 2) / 4, t = (y' / d + 1) / 2 (bilinear over the two texels with repeat wrap), black outside the window."""
    kd, I, h = np.array([0.5, 0.25, 0.125]), 8.0, 2.0
    sd = _floor_scene(kd=tuple(kd), res=33)
    sd.lights.clear()
    sd.camera.update(fov=100.0)
    img = np.array([[[1, 0, 0], [0, 1, 0]]], dtype=np.float32)
    # light space: +z = world -y (down), x = world x, y = world z
    l2w = np.array([[1, 0, 0, 0], [0, 0, -1, h], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.float32)
    sd.add_projection_light((I, I, I), image=img, light_to_world=l2w, fov=90.0)
    sc = oracle.OracleScene(sd)
    ps = np.array([(x, y, 0) for y in range(33) for x in range(33)], dtype=np.int32)
    li = sc.li(ps, nthreads=1).astype(np.float64)
    rays = sc.camera_rays(ps)
    o, d = rays["o"].astype(np.float64), rays["d"].astype(np.float64)
    p = o + (-o[:, 1] / d[:, 1])[:, None] * d
    to_light = np.array([0.0, h, 0.0]) - p
    d2 = (to_light ** 2).sum(1)
    wl = np.stack([p[:, 0], p[:, 2], np.full(len(p), h)], axis=1)  # light-space direction towards the floor point (z = depth)
    px, py = wl[:, 0] / wl[:, 2], wl[:, 1] / wl[:, 2]  # fov 90: inv_tan = 1
    inside = (np.abs(px) <= 2) & (np.abs(py) <= 1)
    s = (px + 2) / 4
    # MIPMap::lookup_triangle(st, 0) over 2 x 1 texels, repeat wrap: texel centres at s = .25 (red) and .75 (green)
    fs = s * 2 - 0.5
    s0 = np.floor(fs); ds = fs - s0
    tex = np.array([[1, 0, 0], [0, 1, 0]], dtype=np.float64)
    c = tex[(s0.astype(int)) % 2] * (1 - ds)[:, None] + tex[(s0.astype(int) + 1) % 2] * ds[:, None]
    expect = (kd[None, :] / np.pi) * (I * c / d2[:, None]) * (to_light[:, 1] / np.sqrt(d2))[:, None] * inside[:, None]
    edge = (np.abs(np.abs(px) - 2) < 1e-3) | (np.abs(np.abs(py) - 1) < 1e-3)
    assert np.allclose(li[~edge], expect[~edge], rtol=3e-4, atol=1e-6)
    assert inside.sum() > 100 and (~inside).sum() > 100
