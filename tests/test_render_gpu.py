"""GPU parity tests of the wavefront path tracer against the oracle (call through the C ABI).

Gates (north_star / SURVEY.md §8d): camera rays bit-identical; per-pixel relative RMSE <= 1e-3 vs the
oracle image at equal spp with the same Halton sequence.  Shading uses CUDA's sinf/cosf/acosf/atan2f,
which may differ from the host libm by an ulp, so radiance parity is statistical, not bit-exact."""
import numpy as np
import pytest

import scenes_small as ss

pytestmark = pytest.mark.gpu
TOL = 1e-3  # north_star: per-pixel relative RMSE <= 1e-3


def _pairs(res, spp):
    return np.array([(x, y, s) for y in range(res) for x in range(res) for s in range(spp)], dtype=np.int32)


def test_camera_rays_bit_exact(gpu, oracle):
    from pbrt_v3_rs_b200 import workloads as wl
    for lens in (0.0, 0.05):
        sd = ss.one_material_scene(wl, ss.MATERIALS["matte"], res=16, spp=4)
        sd.camera.update(lensradius=lens, focaldistance=4.0)
        integ = gpu.PathIntegrator(sd)
        ps = _pairs(16, 4)
        _, rays = integ.li(ps)
        orays = oracle.OracleScene(sd).camera_rays(ps)
        if lens == 0.0:
            assert rays.tobytes() == orays.tobytes()
        else:  # thin lens: concentric_sample_disk uses sin/cos
            assert np.allclose(rays["o"], orays["o"], atol=1e-6) and np.allclose(rays["d"], orays["d"], atol=1e-6)


@pytest.mark.parametrize("name", list(ss.MATERIALS))
@pytest.mark.parametrize("light", ["infinite", "point", "area"])
def test_li_per_sample_matches_oracle(gpu, oracle, name, light):
    from pbrt_v3_rs_b200 import workloads as wl
    sd = ss.one_material_scene(wl, ss.MATERIALS[name], light=light, res=16, spp=4, maxdepth=6)
    integ = gpu.PathIntegrator(sd)
    ps = _pairs(16, 4)
    li, _ = integ.li(ps)
    oli = oracle.OracleScene(sd).li(ps)
    assert np.isfinite(li).all()
    close = np.isclose(li, oli, rtol=2e-3, atol=1e-5).all(1)
    # an ulp of difference in a sampled direction can flip which triangle a later bounce hits: allow a few paths
    assert close.mean() >= 0.999, "only %.4f of the paths agree" % close.mean()
    assert abs(li.mean() - oli.mean()) <= 0.02 * max(oli.mean(), 1e-3)


@pytest.mark.parametrize("name,light,strategy,filt", [("matte", "infinite", "uniform", "box"), ("plastic", "all", "power", "box"),
                                                        ("glass", "all", "power", "box"), ("metal", "all", "uniform", "box"),
                                                        ("matte", "all", "power", "gaussian"), ("plastic", "all", "power", "mitchell"),
                                                        ("matte", "all", "uniform", "sinc"), ("matte", "all", "uniform", "triangle")])
def test_image_rel_rmse(gpu, oracle, name, light, strategy, filt):
    from pbrt_v3_rs_b200 import workloads as wl
    sd = ss.one_material_scene(wl, ss.MATERIALS[name], light=light, res=48, spp=16, maxdepth=5, nu=60, nv=30, strategy=strategy, filt=filt)
    integ = gpu.PathIntegrator(sd)
    img = integ.render()
    ref, stats, _ = oracle.OracleScene(sd).render()
    assert img.shape == ref.shape and np.isfinite(img).all()
    r = ss.rel_rmse(img, ref)
    assert r <= TOL, "relative RMSE %.3e > %.0e" % (r, TOL)
    rc = integ.ray_counts()
    assert rc[0] == stats[0]  # same number of camera rays
    assert int(rc[1]) == int(stats[1]) and int(rc[2]) == int(stats[2])


def test_c1_config_small_and_row_shards(gpu, oracle):
    """C1 (plymesh-variant, 16 spp, maxdepth 5) at reduced mesh/resolution; rendering the image as row shards
    and summing the films (the multi-GPU decomposition) gives the same image as one call."""
    from pbrt_v3_rs_b200 import workloads as wl
    sd = wl.scene_c1(nu=80, nv=80, res=96, spp=16)
    integ = gpu.PathIntegrator(sd)
    full = integ.render_rows()
    ref = oracle.OracleScene(sd).render()[0]
    assert ss.rel_rmse(integ.resolve(full), ref) <= TOL
    parts = sum(integ.render_rows(a, b) for a, b in ((0, 31), (31, 64), (64, 96)))
    # identical up to the rounding of summing per-shard XYZ (a zero-jitter sample of the next shard's first row
    # also lands in this shard's last row)
    assert np.allclose(parts, full, rtol=1e-6, atol=1e-6)
    assert np.array_equal(parts[..., 3], full[..., 3])


def test_interleaved_shards_sum_to_full_image(gpu):
    """The multi-GPU decomposition, emulated on one device: shards of interleaved row bands sum to the full film
    (box filter: disjoint rows; gaussian: aprons add up)."""
    import torch
    from pbrt_v3_rs_b200 import workloads as wl
    for filt in ("box", "gaussian"):
        sd = ss.one_material_scene(wl, ss.MATERIALS["plastic"], light="all", res=40, spp=4, strategy="power", filt=filt)
        integ = gpu.PathIntegrator(sd)
        full = integ.render_rows()
        for n in (2, 3):
            acc = torch.zeros((40, 40, 4), dtype=torch.float32, device="cuda")
            for r in range(n):
                part = torch.zeros_like(acc)
                integ.render_shard_device(r, n, part.data_ptr(), band_rows=8)
                acc += part
            assert np.allclose(acc.cpu().numpy(), full, rtol=2e-6, atol=1e-6), (filt, n)


def test_c3_config_small(gpu, oracle):
    from pbrt_v3_rs_b200 import workloads as wl
    sd = wl.scene_c3(nu=40, nv=40, xres=160, yres=90, spp=16, maxdepth=8)
    integ = gpu.PathIntegrator(sd)
    img = integ.render()
    ref = oracle.OracleScene(sd).render()[0]
    assert ss.rel_rmse(img, ref) <= TOL


def test_pixel_bounds_and_crop(gpu, oracle):
    from pbrt_v3_rs_b200 import workloads as wl
    sd = ss.one_material_scene(wl, ss.MATERIALS["matte"], res=32, spp=4)
    sd.film["cropwindow"] = (0.25, 0.75, 0.0, 0.5)
    sd.integrator["pixelbounds"] = (10, 2, 20, 12)
    integ = gpu.PathIntegrator(sd)
    img = integ.render()
    ref = oracle.OracleScene(sd).render()[0]
    assert img.shape == ref.shape == (16, 16, 3)
    assert ss.rel_rmse(img, ref) <= TOL
    assert not img[12:].any() and img[2:12, 2:12].any()


def test_unsupported_inputs_fail_loudly(gpu):
    from pbrt_v3_rs_b200 import workloads as wl
    sd = ss.one_material_scene(wl, ss.MATERIALS["matte"], res=8, spp=2)
    sd.sampler["type"] = "02sequence"
    sd.integrator.update(name="whitted")  # the tree integrators draw a data-dependent number of dimensions: halton / sobol only
    with pytest.raises(gpu.B200PTError):
        gpu.PathIntegrator(sd).preprocess()


@pytest.mark.parametrize("name,light,spp,dims,res", [("matte", "all", 4, 4, 40), ("glass", "area", 3, 4, 37), ("plastic", "infinite", 2, 0, 24), ("metal", "point", 4, 7, 33)])
def test_zerotwo_default_dimensions_tile_sequential(gpu, oracle, name, light, spp, dims, res):
    """The reference's default "dimensions" = 4 (samplers/src/zero_two_sequence.rs:157): past the pre-generated slots
    get_1d / get_2d draw from the tile sampler's RNG inside li() (pixel_sampler.rs:88-110), so every tile is one
    sequential stream.  The tile-sequential mode must reproduce the reference's streams: images, ray counts and
    per-sample radiance (fresh tile stream per queried sample, the oracle's sampler_at convention) against the oracle,
    incl. clipped tiles (res not a multiple of 16) and pixel counts rounded up to a power of two."""
    from pbrt_v3_rs_b200 import workloads as wl
    sd = ss.one_material_scene(wl, ss.MATERIALS[name], light=light, res=res, spp=spp, maxdepth=5, strategy="power")
    sd.sampler.update(type="02sequence", dimensions=dims)
    integ = gpu.PathIntegrator(sd)
    osc = oracle.OracleScene(sd)
    spp2 = 1 << (spp - 1).bit_length()
    ps = np.array([(x, y, s) for y in (0, 3, 15, 16, res - 1) for x in (0, 15, 16, 31, res - 1) if x < res for s in range(spp2)], dtype=np.int32)
    li, rays = integ.li(ps)
    assert rays.tobytes() == osc.camera_rays(ps).tobytes()
    oli = osc.li(ps)
    close = np.isclose(li, oli, rtol=2e-3, atol=1e-5).all(1)
    assert close.mean() >= 0.999, close.mean()
    img = integ.render()
    ref, stats, _ = osc.render()
    assert ss.rel_rmse(img, ref) <= TOL
    assert [int(x) for x in integ.ray_counts()] == [int(x) for x in stats[:3]]
    # a row shard renders whole tiles but keeps only its rows: the two halves sum to the image
    film_a = integ.render_rows(0, 20)
    film_b = integ.render_rows(20, res)
    whole = integ.render_rows(0, res)
    assert np.array_equal((film_a + film_b).view(np.uint32), whole.view(np.uint32))
    # multi-GPU shards must be whole tile rows (a tile is one sequential stream): bands of 16 rows work and every ray is
    # traced once, bands of 8 are refused
    import torch
    h, w = integ.film_shape()
    parts, rays = [], np.zeros(3, dtype=np.int64)
    for r in range(2):
        part = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
        integ.render_shard_device(r, 2, part.data_ptr(), band_rows=16)
        torch.cuda.synchronize()
        parts.append(part.cpu().numpy())
        rays += np.array([int(x) for x in integ.ray_counts()], dtype=np.int64)
    assert np.array_equal((parts[0] + parts[1]).view(np.uint32), whole.view(np.uint32))
    assert rays.tolist() == [int(x) for x in stats[:3]]
    with pytest.raises(gpu.B200PTError, match="multiple of 16"):
        integ.render_shard_device(0, 2, torch.zeros((h, w, 4), dtype=torch.float32, device="cuda").data_ptr(), band_rows=8)


@pytest.mark.parametrize("name,light,spp", [("matte", "infinite", 8), ("plastic", "all", 4), ("glass", "area", 6)])
def test_zerotwo_sampler_matches_oracle(gpu, oracle, name, light, spp):
    """(0,2)-sequence sampler (samplers/src/zero_two_sequence.rs) with "dimensions" large enough that li() never
    draws from the tile RNG: per-sample radiance and image parity, incl. pixelsamples rounded up to a power of two
    and images larger than one 16x16 tile (per-tile PCG32 streams)."""
    from pbrt_v3_rs_b200 import workloads as wl
    sd = ss.one_material_scene(wl, ss.MATERIALS[name], light=light, res=40, spp=spp, maxdepth=5, strategy="power")
    sd.sampler.update(type="02sequence", dimensions=32)
    integ = gpu.PathIntegrator(sd)
    osc = oracle.OracleScene(sd)
    spp2 = 1 << (spp - 1).bit_length()
    ps = np.array([(x, y, s) for y in (0, 3, 15, 16, 39) for x in (0, 15, 16, 31, 39) for s in range(spp2)], dtype=np.int32)
    li, rays = integ.li(ps)
    assert rays.tobytes() == osc.camera_rays(ps).tobytes()
    oli = osc.li(ps)
    close = np.isclose(li, oli, rtol=2e-3, atol=1e-5).all(1)
    assert close.mean() >= 0.999
    img = integ.render()
    ref, stats, _ = osc.render()
    assert ss.rel_rmse(img, ref) <= TOL
    assert integ.ray_counts()[0] == stats[0] == 40 * 40 * spp2


def _instanced_scene(wl, mat, res=48, spp=4, lights="all"):
    from pbrt_v3_rs_b200.scene import SceneDescription
    sd = SceneDescription()
    m = sd.add_material(**mat)
    g = sd.add_material(type="matte", Kd=(0.4, 0.4, 0.4))
    sd.add_mesh(wl.ground_quad(), g)
    obj = sd.add_object(wl.displaced_sphere(24, 12, radius=0.45), m)
    obj2 = sd.add_object(wl.displaced_sphere(12, 6, radius=0.3, seed=9), g, reverse_orientation=True)
    rng = np.random.Generator(np.random.PCG64(4))
    for k in range(9):
        M = wl.rigid_transform(rng, extent=0.1)
        M[:3, 3] += [(k % 3 - 1) * 1.2, -0.4 + 0.3 * (k // 3), (k // 3 - 1) * 1.2]
        sd.add_instance(obj if k % 2 == 0 else obj2, M)
    sd.add_instance(obj, np.eye(4, dtype=np.float32))  # identity transform (TransformedPrimitive skips the hit transform)
    if lights in ("all", "infinite"):
        sd.add_infinite_light((1.0, 1.0, 1.0))
    if lights in ("all", "point"):
        sd.add_point_light((2, 4, -3), (40, 40, 40))
    sd.camera.update(eye=(0.0, 2.5, -5.0), look=(0.0, -0.3, 0.0), up=(0, 1, 0), fov=40.0)
    sd.film.update(xresolution=res, yresolution=res)
    sd.sampler.update(pixelsamples=spp)
    sd.integrator.update(maxdepth=5, lightsamplestrategy="power")
    return sd


@pytest.mark.parametrize("name", ["matte", "plastic", "glass"])
def test_instancing_two_level_bvh_matches_oracle(gpu, oracle, name):
    """TransformedPrimitive (static transform) + per-object BVHAccel: per-sample radiance, image and ray counts."""
    from pbrt_v3_rs_b200 import workloads as wl
    sd = _instanced_scene(wl, ss.MATERIALS[name])
    integ = gpu.PathIntegrator(sd)
    osc = oracle.OracleScene(sd)
    ps = _pairs(48, 4)[::3]
    li, rays = integ.li(ps)
    assert rays.tobytes() == osc.camera_rays(ps).tobytes()
    oli = osc.li(ps)
    close = np.isclose(li, oli, rtol=2e-3, atol=1e-5).all(1)
    assert close.mean() >= 0.999, close.mean()
    img = integ.render()
    ref, stats, _ = osc.render()
    assert ss.rel_rmse(img, ref) <= TOL
    rc = integ.ray_counts()
    assert rc[0] == stats[0] and int(rc[1]) == int(stats[1]) and int(rc[2]) == int(stats[2])


def test_instancing_point_light_only_is_bit_exact(gpu, oracle):
    from pbrt_v3_rs_b200 import workloads as wl
    sd = _instanced_scene(wl, ss.MATERIALS["matte"], lights="point")
    ps = _pairs(48, 4)
    li, _ = gpu.PathIntegrator(sd).li(ps)
    oli = oracle.OracleScene(sd).li(ps)
    same = (li.view(np.uint32) == oli.view(np.uint32)).all(1)
    assert same.mean() >= 0.999, same.mean()


@pytest.mark.parametrize("name", ["matte", "plastic", "rough_glass", "glass", "metal"])
def test_vertex_normals_and_uvs_match_oracle(gpu, oracle, name):
    """Meshes with "uv" and "N" (triangle.rs:384-394, 631-721): shading frame from the interpolated normal, geometric
    normal face-forwarded to it, dpdu from the uvs.  Point light only => every transcendental on the path is exact."""
    from pbrt_v3_rs_b200 import workloads as wl
    sd = ss.one_material_scene(wl, ss.MATERIALS[name], light="point", res=24, spp=4, maxdepth=5, smooth=True)
    integ = gpu.PathIntegrator(sd)
    osc = oracle.OracleScene(sd)
    ps = _pairs(24, 4)
    li, _ = integ.li(ps)
    oli = osc.li(ps)
    same = (li.view(np.uint32) == oli.view(np.uint32)).all(1)
    assert same.mean() >= 0.995, same.mean()
    sd2 = ss.one_material_scene(wl, ss.MATERIALS[name], light="all", res=32, spp=8, maxdepth=5, strategy="power", smooth=True)
    img = gpu.PathIntegrator(sd2).render()
    ref = oracle.OracleScene(sd2).render()[0]
    assert ss.rel_rmse(img, ref) <= TOL
    # and the attributes matter: the faceted version of the same scene is a different image
    flat = gpu.PathIntegrator(ss.one_material_scene(wl, ss.MATERIALS[name], light="all", res=32, spp=8, maxdepth=5, strategy="power")).render()
    assert ss.rel_rmse(img, flat) > 10 * TOL


def test_vertex_attributes_on_instanced_objects_and_flags(gpu, oracle):
    """uv / N / S on an object instance (shading normal goes through transform_surface_interaction), reverse_orientation
    with normals, tangents, and a second mesh without attributes in the same scene."""
    from pbrt_v3_rs_b200 import workloads as wl
    from pbrt_v3_rs_b200.scene import SceneDescription
    sd = SceneDescription()
    m = sd.add_material(type="plastic")
    g = sd.add_material(type="matte", Kd=(0.4, 0.4, 0.4))
    sd.add_mesh(wl.ground_quad(), g)
    tv, uv, nrm = wl.displaced_sphere(24, 12, radius=0.5, with_attrs=True)
    rng = np.random.Generator(np.random.PCG64(5))
    tan = rng.normal(size=nrm.shape).astype(np.float32)
    obj = sd.add_object(tv, m, uv=uv, normals=nrm, tangents=tan)
    obj2 = sd.add_object(tv, g, reverse_orientation=True, normals=nrm)
    for k in range(4):
        M = wl.rigid_transform(rng, extent=0.1)
        M[:3, 3] += [(k % 2 - 0.5) * 1.6, -0.3, (k // 2 - 0.5) * 1.6]
        sd.add_instance(obj if k % 2 == 0 else obj2, M)
    tv2, uv2, nrm2 = wl.displaced_sphere(16, 8, radius=0.4, center=(0.0, 0.9, 0.0), with_attrs=True)
    sd.add_mesh(tv2, m, uv=uv2, tangents=rng.normal(size=nrm2.shape).astype(np.float32), reverse_orientation=True)
    sd.add_point_light((2, 4, -3), (40, 40, 40))
    sd.camera.update(eye=(0.0, 2.5, -5.0), look=(0.0, 0.0, 0.0), up=(0, 1, 0), fov=40.0)
    sd.film.update(xresolution=32, yresolution=32)
    sd.sampler.update(pixelsamples=4)
    sd.integrator.update(maxdepth=4, lightsamplestrategy="uniform")
    ps = _pairs(32, 4)
    li, _ = gpu.PathIntegrator(sd).li(ps)
    oli = oracle.OracleScene(sd).li(ps)
    same = (li.view(np.uint32) == oli.view(np.uint32)).all(1)
    assert same.mean() >= 0.995, same.mean()
    assert li.any()


def _envmap_scene(wl, mat, size, spp=8, res=32, extra_point=False, strategy="uniform"):
    from pbrt_v3_rs_b200.scene import SceneDescription
    sd = SceneDescription()
    m = sd.add_material(**mat)
    g = sd.add_material(type="matte", Kd=(0.4, 0.4, 0.4))
    sd.add_mesh(wl.displaced_sphere(24, 12), m)
    sd.add_mesh(wl.ground_quad(), g)
    # pbrt scenes stand the lat-long map upright: Rotate -90 1 0 0 (scenes/materials/matte.pbrt:17-19)
    c, s_ = np.cos(np.deg2rad(-90.0)), np.sin(np.deg2rad(-90.0))
    rot = np.array([[1, 0, 0, 0], [0, c, -s_, 0], [0, s_, c, 0], [0, 0, 0, 1]], dtype=np.float32)
    sd.add_infinite_light((1.0, 0.9, 0.8), image=wl.sky_image(*size), light_to_world=rot)
    if extra_point:
        sd.add_point_light((1.5, 3.0, -3.0), (30, 30, 30))
    sd.camera.update(eye=(0.0, 1.2, -4.0), look=(0.0, -0.1, 0.0), up=(0, 1, 0), fov=40.0)
    sd.film.update(xresolution=res, yresolution=res)
    sd.sampler.update(type="halton", pixelsamples=spp)
    sd.integrator.update(maxdepth=5, lightsamplestrategy=strategy)
    return sd


@pytest.mark.parametrize("name,size,extra,strategy", [("matte", (16, 8), False, "uniform"), ("plastic", (20, 9), True, "power"),
                                                       ("glass", (64, 32), False, "uniform"), ("metal", (64, 4), True, "power")])
def test_image_mapped_infinite_light_matches_oracle(gpu, oracle, name, size, extra, strategy):
    """LightSource "infinite" with a "mapname" (SURVEY §8f rank 1): MIPMap level-0 lookups (le, sample_li) and the
    Distribution2D over the importance image (sample_li, pdf_li), incl. a non-power-of-two map through the resampler,
    a 16:1 map whose importance image reads pyramid levels > 0, and the power light-sampling strategy."""
    from pbrt_v3_rs_b200 import workloads as wl
    sd = _envmap_scene(wl, ss.MATERIALS[name], size, extra_point=extra, strategy=strategy)
    integ = gpu.PathIntegrator(sd)
    osc = oracle.OracleScene(sd)
    ps = _pairs(32, 8)[::5]
    li, rays = integ.li(ps)
    oli = osc.li(ps)
    assert rays.tobytes() == osc.camera_rays(ps).tobytes()
    close = np.isclose(li, oli, rtol=2e-3, atol=1e-5).all(1)
    assert close.mean() >= 0.999, close.mean()
    img = integ.render()
    ref, stats, _ = osc.render()
    assert np.isfinite(img).all() and ss.rel_rmse(img, ref) <= TOL
    rc = integ.ray_counts()
    assert rc[0] == stats[0] and int(rc[1]) == int(stats[1]) and int(rc[2]) == int(stats[2])
    # the map matters: the same scene lit by the constant light is a different image
    sd.lights[0].pop("image")
    assert ss.rel_rmse(gpu.PathIntegrator(sd).render(), img) > 10 * TOL


@pytest.mark.parametrize("with_instances", [False, True])
def test_scene_file_renders_like_the_oracle(gpu, oracle, tmp_path, with_instances):
    """.pbrt + PLY + PFM environment map read by the host loader, rendered through b200pt_scene_create / render and
    written as PFM: same image as the oracle renders from the same description."""
    import test_scene_loader_cpu as tl
    from pbrt_v3_rs_b200 import workloads as wl
    path, _ = tl._build_pair(tmp_path, gpu, wl, with_instances)
    ld = gpu.load_pbrt(path)
    integ = gpu.PathIntegrator(ld)
    img = integ.render()
    ref, stats, _ = oracle.OracleScene(ld).render()
    assert img.shape == ref.shape == (16, 20, 3) and img.any()
    assert ss.rel_rmse(img, ref) <= TOL
    assert integ.ray_counts()[0] == stats[0]
    out = str(tmp_path / ld.output)
    gpu.write_pfm(out, img)
    assert np.array_equal(gpu.read_pfm(out), img)


@pytest.mark.parametrize("integrator,strategy", [("path", "power"), ("path", "spatial"), ("whitted", "uniform"), ("directlighting", "uniform")])
def test_wave_splitting_does_not_change_the_image(gpu, integrator, strategy, monkeypatch):
    """A render is cut into waves of at most 2^k paths (B200PT_WAVE_LOG2, default: as large as the render needs); every
    (pixel, sample) is independent, so the film must not depend on where the cuts fall."""
    from pbrt_v3_rs_b200 import workloads as wl
    def film(log2):
        if log2 is None:
            monkeypatch.delenv("B200PT_WAVE_LOG2", raising=False)
        else:
            monkeypatch.setenv("B200PT_WAVE_LOG2", str(log2))
        sd = ss.one_material_scene(wl, ss.MATERIALS["glass"], light="all", res=96, spp=16, maxdepth=4, strategy="uniform")
        sd.integrator.update(name=integrator, lightsamplestrategy=strategy)
        return gpu.PathIntegrator(sd).render_rows()
    one, many = film(None), film(16)  # 147 456 paths: one wave vs three
    assert np.array_equal(one, many)


def _emissive_smooth_scene(wl, integrator="path", inward=False, twosided=False):
    """A small emissive sphere whose mesh carries vertex normals ("N"): Triangle::sample orients the sampled normal with
    the interpolated one (triangle.rs:931-937) and the hit's geometric normal is face-forwarded to the shading normal."""
    from pbrt_v3_rs_b200.scene import SceneDescription
    sd = SceneDescription()
    g = sd.add_material(type="matte", Kd=(0.5, 0.5, 0.5))
    e = sd.add_material(type="matte", Kd=(0.0, 0.0, 0.0))
    p = sd.add_material(type="plastic")
    sd.add_mesh(wl.ground_quad(), g)
    sd.add_mesh(wl.displaced_sphere(16, 8, radius=0.6, center=(-1.2, -0.6, 0.0)), p)
    tv, uv, nrm = wl.displaced_sphere(8, 4, radius=0.35, amplitude=0.0, center=(0.6, 0.4, -0.3), with_attrs=True)
    if inward:
        nrm = -nrm  # normals pointing into the sphere: the emitting side flips with them
    sd.add_mesh(tv, e, area_light=dict(L=(12, 11, 9), twosided=twosided), uv=uv, normals=nrm)
    sd.camera.update(eye=(0.0, 1.2, -4.0), look=(0.0, -0.2, 0.0), up=(0, 1, 0), fov=40.0)
    sd.film.update(xresolution=40, yresolution=32)
    sd.sampler.update(type="halton", pixelsamples=8)
    sd.integrator.update(name=integrator, maxdepth=4, lightsamplestrategy="power")
    return sd


@pytest.mark.parametrize("integrator,inward,twosided", [("path", False, False), ("path", True, False), ("path", False, True),
                                                       ("whitted", False, False), ("directlighting", True, False)])
def test_area_light_on_a_mesh_with_vertex_normals(gpu, oracle, integrator, inward, twosided):
    from pbrt_v3_rs_b200 import workloads as wl
    sd = _emissive_smooth_scene(wl, integrator, inward, twosided)
    integ = gpu.PathIntegrator(sd)
    osc = oracle.OracleScene(sd)
    ps = np.array([(x, y, s) for y in range(0, 32, 2) for x in range(0, 40, 2) for s in range(4)], dtype=np.int32)
    li, _ = integ.li(ps)
    oli = osc.li(ps)
    close = np.isclose(li, oli, rtol=2e-3, atol=1e-5).all(1)
    assert close.mean() >= 0.999, close.mean()
    img = integ.render()
    ref, stats, _ = osc.render()
    assert ss.rel_rmse(img, ref) <= TOL
    assert (img.mean() > 0.01) == (twosided or not inward)  # inward-facing normals: the sphere emits into itself only
    rc = integ.ray_counts()
    assert rc[0] == stats[0] and int(rc[1]) == int(stats[1])


@pytest.mark.parametrize("filt,integrator", [("box", "path"), ("gaussian", "path"), ("box", "whitted")])
def test_memory_budget_does_not_change_the_image(gpu, filt, integrator):
    """b200pt_scene_set_memory_budget: the wave state is bounded by a byte budget and the render is cut into as many
    waves as that needs; the film keeps running sums between waves (k_film), in the same pixel-major / sample order
    (core/src/film/film_tile.rs:62-108), so every bit of the film is independent of the budget - also with a filter
    wider than a pixel, where a wave boundary falls between samples that reach the same pixel."""
    from pbrt_v3_rs_b200 import workloads as wl
    sd = ss.one_material_scene(wl, ss.MATERIALS["plastic"], light="all", res=80, spp=24, maxdepth=4, strategy="power", filt=filt)
    sd.integrator.update(name=integrator)
    integ = gpu.PathIntegrator(sd)
    full = integ.render_rows()
    rc = integ.ray_counts().copy()
    assert full[..., 3].min() > 0
    integ.set_memory_budget(8 << 20)      # 8 MiB: ~16 K paths per wave -> about ten waves for 153 600+ samples
    small = integ.render_rows()
    assert np.array_equal(small, full)
    assert np.array_equal(integ.ray_counts(), rc)
    integ.set_memory_budget(0)
    assert np.array_equal(integ.render_rows(), full)


def test_null_material_is_passed_through(gpu, oracle):
    """Material "none" (api/src/graphics_state.rs:336 -> None): PathIntegrator::li re-spawns the ray through the surface
    without counting a bounce or drawing samples (integrators/src/path.rs:141-150); an emissive primitive without a
    material still emits.  Point light + area light => every transcendental on the path is exact: bit parity."""
    from pbrt_v3_rs_b200 import workloads as wl
    from pbrt_v3_rs_b200.scene import SceneDescription

    def build(with_null):
        sd = SceneDescription()
        m = sd.add_material(type="plastic")
        g = sd.add_material(type="matte", Kd=(0.4, 0.4, 0.4))
        sd.add_mesh(wl.displaced_sphere(24, 12), m)
        sd.add_mesh(wl.ground_quad(), g)
        if with_null:  # a shell around the sphere and a curtain in front of the camera, both without a material
            sd.add_mesh(wl.displaced_sphere(16, 8, radius=1.6, amplitude=0.0), -1)
            q = np.array([[-3, -1, -2.5], [3, -1, -2.5], [3, 3, -2.5], [-3, 3, -2.5]], dtype=np.float32)
            sd.add_mesh(np.stack([np.concatenate([q[0], q[1], q[2]]), np.concatenate([q[0], q[2], q[3]])]), -1)
        lq = np.array([[-1.5, 3.0, -1.0], [1.5, 3.0, -1.0], [1.5, 3.0, 1.0], [-1.5, 3.0, 1.0]], dtype=np.float32)
        sd.add_mesh(np.stack([np.concatenate([lq[0], lq[1], lq[2]]), np.concatenate([lq[0], lq[2], lq[3]])]), -1 if with_null else g,
                    area_light=dict(L=(15, 15, 15), twosided=True))
        sd.add_point_light((1.5, 2.5, -3.0), (30, 30, 30))
        sd.camera.update(eye=(0.0, 1.2, -4.0), look=(0.0, -0.1, 0.0), up=(0, 1, 0), fov=40.0)
        sd.film.update(xresolution=32, yresolution=32)
        sd.sampler.update(type="halton", pixelsamples=4)
        sd.integrator.update(maxdepth=5, lightsamplestrategy="power")
        return sd
    sd = build(True)
    integ = gpu.PathIntegrator(sd)
    osc = oracle.OracleScene(sd)
    ps = _pairs(32, 4)
    li, rays = integ.li(ps)
    oli = osc.li(ps)
    assert rays.tobytes() == osc.camera_rays(ps).tobytes()
    same = (li.view(np.uint32) == oli.view(np.uint32)).all(1)
    assert same.mean() >= 0.995, same.mean()
    img = integ.render()
    ref, stats, _ = osc.render()
    assert ss.rel_rmse(img, ref) <= TOL
    assert [int(x) for x in integ.ray_counts()] == [int(x) for x in stats[:3]]
    # the surfaces are really crossed: more closest-hit rays than the same scene without them, and shadow rays still see them as occluders
    plain = gpu.PathIntegrator(build(False))
    plain.render()
    assert int(integ.ray_counts()[1]) > int(plain.ray_counts()[1])
    # the tree integrators do not take primitives without a material
    sd.integrator.update(name="whitted")
    with pytest.raises(gpu.B200PTError):
        gpu.PathIntegrator(sd).preprocess()


def test_deep_paths_continue_in_a_new_control_segment(gpu, oracle):
    """maxdepth above the 16 iterations one segment of control blocks holds: the bounce loop reads the surviving queue
    size back once per segment and continues (mirror-like metal box: paths really get that deep)."""
    from pbrt_v3_rs_b200 import workloads as wl
    sd = ss.one_material_scene(wl, ss.MATERIALS["glass"], light="point", res=24, spp=4, maxdepth=40)
    shell = sd.add_material(type="matte", Kd=(0.9, 0.9, 0.9))
    sd.add_mesh(wl.displaced_sphere(16, 8, radius=9.0, amplitude=0.0), shell)  # a closed room: no path escapes
    sd.integrator.update(rrthreshold=0.0)  # no Russian roulette: every path lives until maxdepth
    integ = gpu.PathIntegrator(sd)
    osc = oracle.OracleScene(sd)
    ps = _pairs(24, 4)
    li, _ = integ.li(ps)
    oli = osc.li(ps)
    same = (li.view(np.uint32) == oli.view(np.uint32)).all(1)
    assert same.mean() >= 0.995, same.mean()
    img = integ.render()
    ref, stats, _ = osc.render()
    assert ss.rel_rmse(img, ref) <= TOL
    assert [int(x) for x in integ.ray_counts()] == [int(x) for x in stats[:3]]
    assert int(stats[1]) > 30 * int(stats[0])  # the paths really are that long


def test_path_depth_beyond_the_sampler_tables_is_refused(gpu):
    from pbrt_v3_rs_b200 import workloads as wl
    sd = ss.one_material_scene(wl, ss.MATERIALS["matte"], res=8, spp=2, maxdepth=125)  # 5 + 8 * 125 > 1000 Halton dimensions
    with pytest.raises(gpu.B200PTError):
        gpu.PathIntegrator(sd).preprocess()
    sd.integrator.update(maxdepth=124)
    gpu.PathIntegrator(sd).preprocess()


def test_spatial_row_pool_exhaustion_is_reported(gpu, monkeypatch):
    """SpatialLightDistribution rows are handed out on first touch; a pool too small for the voxels a render reaches is an
    error, not a silently different image."""
    from pbrt_v3_rs_b200 import workloads as wl
    sd = ss.one_material_scene(wl, ss.MATERIALS["matte"], light="all", res=48, spp=4, strategy="spatial")
    full = gpu.PathIntegrator(sd).render_rows()
    assert full[..., 3].min() > 0
    monkeypatch.setenv("B200PT_SPATIAL_BUDGET", "1024")  # -> the minimum pool of 64 rows
    integ = gpu.PathIntegrator(sd)
    with pytest.raises(gpu.B200PTError, match="spatial"):
        integ.render_rows()


def test_render_multi_on_the_visible_devices(gpu, oracle):
    """b200pt_multi_*: the scene replicated on every visible GPU of this process, interleaved row bands, bands gathered on
    the first device (NCCL send / recv; ncclReduce for the gaussian filter).  Same film as one device renders (box:
    bit-identical, every sample is taken by exactly one device; gaussian: up to the rounding of summing shard films)."""
    import torch
    from pbrt_v3_rs_b200 import workloads as wl
    n = torch.cuda.device_count()
    for filt in ("box", "gaussian"):
        sd = ss.one_material_scene(wl, ss.MATERIALS["plastic"], light="all", res=72, spp=8, maxdepth=4, strategy="power", filt=filt)
        single = gpu.PathIntegrator(sd)
        full = single.render_rows()
        rc = [int(x) for x in single.ray_counts()]
        for devs in sorted({1, n}):
            multi = gpu.MultiGPURender(sd, list(range(devs)))
            film = multi.render_rows(band_rows=8)
            info = multi.info()
            if filt == "box":
                assert np.array_equal(film, full), (filt, devs)
            else:
                assert np.allclose(film, full, rtol=2e-6, atol=1e-6), (filt, devs)
            assert [int(x) for x in info["rays"]] == rc or filt == "gaussian"
            assert int(info["rays"][0]) == rc[0]
            multi.close()
    assert gpu.current_device() == 0


@pytest.mark.gpu
@pytest.mark.parametrize("name,light,integrator,strategy", [("matte", "distant", "path", "uniform"), ("plastic", "all+distant", "path", "power"), ("glass", "all+distant", "path", "spatial"),
                                                            ("matte", "all+distant", "whitted", "uniform"), ("plastic", "all+distant", "directlighting", "uniform")])
def test_distant_light_matches_oracle(gpu, oracle, name, light, integrator, strategy):
    """DistantLight (lights/src/distant.rs): a delta-direction light - sample_li with pdf 1 towards w_light, visibility tested
    up to p + w_light * 2 * world_radius, power L * pi * r^2 in the power / spatial distributions, no BSDF-sampled MIS ray."""
    from pbrt_v3_rs_b200 import workloads as wl
    sd = ss.one_material_scene(wl, ss.MATERIALS[name], light=light, res=40, spp=8, maxdepth=4, strategy=strategy)
    sd.integrator.update(name=integrator)
    integ = gpu.PathIntegrator(sd)
    osc = oracle.OracleScene(sd)
    ps = _pairs(40, 8)
    li, rays = integ.li(ps)
    oli = osc.li(ps)
    assert rays.tobytes() == osc.camera_rays(ps).tobytes()
    assert np.isclose(li, oli, rtol=2e-3, atol=1e-5).all(1).mean() >= 0.999
    img = integ.render()
    ref, stats, _ = osc.render()
    assert img.mean() > 0.01 and ss.rel_rmse(img, ref) <= TOL
    assert [int(x) for x in integ.ray_counts()] == [int(x) for x in stats[:3]]



@pytest.mark.gpu
@pytest.mark.parametrize("name,light,integrator,strategy", [("matte", "spot", "path", "uniform"), ("plastic", "all+spot", "path", "power"), ("glass", "all+spot", "path", "spatial"),
                                                            ("matte", "all+spot", "whitted", "uniform"), ("plastic", "all+spot", "directlighting", "uniform")])
def test_spot_light_matches_oracle(gpu, oracle, name, light, integrator, strategy):
    """SpotLight (lights/src/spot.rs): a point light times falloff(-wi) = smooth step ^ 4 between cos(coneangle) and
    cos(coneangle - conedeltaangle) in light space; power I * 2 pi * (1 - (cos_start + cos_total) / 2) in the distributions."""
    from pbrt_v3_rs_b200 import workloads as wl
    sd = ss.one_material_scene(wl, ss.MATERIALS[name], light=light, res=40, spp=8, maxdepth=4, strategy=strategy)
    sd.integrator.update(name=integrator)
    integ = gpu.PathIntegrator(sd)
    osc = oracle.OracleScene(sd)
    ps = _pairs(40, 8)
    li, rays = integ.li(ps)
    oli = osc.li(ps)
    assert rays.tobytes() == osc.camera_rays(ps).tobytes()
    assert np.isclose(li, oli, rtol=2e-3, atol=1e-5).all(1).mean() >= 0.999
    if light == "spot":  # no infinite light: every transcendental is exact, radiance must agree bit for bit
        assert (li.view(np.uint32) == oli.view(np.uint32)).all(1).mean() >= 0.995
    img = integ.render()
    ref, stats, _ = osc.render()
    assert img.mean() > 0.01 and ss.rel_rmse(img, ref) <= TOL
    assert [int(x) for x in integ.ray_counts()] == [int(x) for x in stats[:3]]


@pytest.mark.gpu
@pytest.mark.parametrize("camera", ["orthographic", "environment"])
def test_other_cameras_rays_bit_exact_and_images(gpu, oracle, camera):
    """OrthographicCamera (origin on the film plane, direction +z) and EnvironmentCamera (direction from the film position in
    spherical coordinates, glibc's sinf / cosf): camera rays bit-identical, image within the gate."""
    from pbrt_v3_rs_b200 import workloads as wl
    sd = ss.one_material_scene(wl, ss.MATERIALS["plastic"], light="all", res=24, spp=4, maxdepth=4)
    sd.camera.update(type=camera, screenwindow=(-2.0, 2.0, -2.0, 2.0))
    integ = gpu.PathIntegrator(sd)
    osc = oracle.OracleScene(sd)
    ps = _pairs(24, 4)
    li, rays = integ.li(ps)
    assert rays.tobytes() == osc.camera_rays(ps).tobytes()
    assert np.isclose(li, osc.li(ps), rtol=2e-3, atol=1e-5).all(1).mean() >= 0.999
    img = integ.render()
    ref, stats, _ = osc.render()
    assert img.mean() > 0.01 and ss.rel_rmse(img, ref) <= TOL
    assert [int(x) for x in integ.ray_counts()] == [int(x) for x in stats[:3]]


@pytest.mark.gpu
@pytest.mark.parametrize("integrator,with_map", [("whitted", True), ("path", True), ("directlighting", False)])
def test_goniometric_light_matches_oracle(gpu, oracle, integrator, with_map):
    """GonioPhotometricLight (lights/src/goniometric.rs): a point light scaled by an image looked up at the spherical
    coordinates of the light-space direction (y / z swapped) - MIPMap::new over a non-power-of-two image (Lanczos resampling),
    lookup_triangle(st, 0); power 4 pi I lookup_triangle((.5, .5), .5) in the light distribution."""
    from pbrt_v3_rs_b200 import workloads as wl
    sd = ss.one_material_scene(wl, ss.MATERIALS["plastic"], light="area", res=32, spp=8, maxdepth=4, strategy="power")
    rng = np.random.default_rng(5)
    img = (rng.uniform(0.0, 1.0, size=(12, 20, 3)) ** 2).astype(np.float32) if with_map else None
    c, s_ = np.float32(np.cos(0.7)), np.float32(np.sin(0.7))
    m = np.array([[c, 0, s_, 1.0], [0, 1, 0, 3.0], [-s_, 0, c, -2.0], [0, 0, 0, 1]], dtype=np.float32)
    sd.add_goniometric_light((60, 55, 50), image=img, light_to_world=m)
    sd.integrator.update(name=integrator)
    integ = gpu.PathIntegrator(sd)
    osc = oracle.OracleScene(sd)
    ps = _pairs(32, 4)
    li, _ = integ.li(ps)
    assert np.isclose(li, osc.li(ps), rtol=2e-3, atol=1e-5).all(1).mean() >= 0.999
    img_g = integ.render()
    ref, stats, _ = osc.render()
    assert img_g.mean() > 0.01 and ss.rel_rmse(img_g, ref) <= TOL
    assert [int(x) for x in integ.ray_counts()] == [int(x) for x in stats[:3]]


@pytest.mark.gpu
@pytest.mark.parametrize("integrator,with_map", [("whitted", True), ("path", True), ("directlighting", False)])
def test_projection_light_matches_oracle(gpu, oracle, integrator, with_map):
    """ProjectionLight (lights/src/projection.rs): a point light that projects an image through Transform::perspective(fov, 1e-3,
    1e30) - black behind the near plane and outside the screen window, MIPMap::lookup_triangle(st, 0) inside; power
    lookup * I * 2 pi (1 - cos_total_width) in the light distribution."""
    from pbrt_v3_rs_b200 import workloads as wl
    sd = ss.one_material_scene(wl, ss.MATERIALS["plastic"], light="area", res=32, spp=8, maxdepth=4, strategy="power")
    rng = np.random.default_rng(6)
    img = (rng.uniform(0.0, 1.0, size=(10, 24, 3)) ** 2).astype(np.float32) if with_map else None
    # the light above the scene looking down (+z of the light = world -y), slightly rotated about its axis
    c, s_ = np.float32(np.cos(0.4)), np.float32(np.sin(0.4))
    m = np.array([[c, -s_, 0, 0.3], [0, 0, -1, 3.5], [s_, c, 0, -0.5], [0, 0, 0, 1]], dtype=np.float32)
    sd.add_projection_light((80, 75, 70), image=img, light_to_world=m, fov=60.0)
    sd.integrator.update(name=integrator)
    integ = gpu.PathIntegrator(sd)
    osc = oracle.OracleScene(sd)
    ps = _pairs(32, 4)
    li, _ = integ.li(ps)
    assert np.isclose(li, osc.li(ps), rtol=2e-3, atol=1e-5).all(1).mean() >= 0.999
    img_g = integ.render()
    ref, stats, _ = osc.render()
    assert img_g.mean() > 0.01 and ss.rel_rmse(img_g, ref) <= TOL
    assert [int(x) for x in integ.ray_counts()] == [int(x) for x in stats[:3]]
