"""BASELINE.json's path-tracing configurations at FULL size (C1, C3, C4, C5), through the C ABI, against the oracle.

Per config (core/src/integrator/sampler_integrator.rs:243-415 is the loop being replaced):
  * the whole image renders, is finite and non-trivial;
  * 2 048 random (pixel, sample) triples of the full scene: camera rays bit-identical, per-sample radiance within 2e-3
    on EVERY sample and bit-identical on the stated fraction (scenes without an infinite light are bit-exact because
    sin/cos are glibc-exact on the device, csrc/libm_exact.cuh; an infinite light adds acosf/atan2f of the environment
    lookup, which may differ from glibc by an ulp in the radiance VALUE, never in the geometry);
  * an oracle render of a centre crop (the whole image for C1) vs the same crop on the GPU: per-pixel relative RMSE
    <= 1e-3 (north_star) and IDENTICAL camera / closest-hit / shadow ray counts.
The oracle needs seconds per crop; C4 (10 M triangles, 256 spp) builds its scene twice and takes about a minute."""
import numpy as np
import pytest

import scenes_small as ss

pytestmark = pytest.mark.gpu
TOL = 1e-3  # north_star: per-pixel relative RMSE <= 1e-3


def _check_config(gpu, oracle, sd, crop, bit_min, n_li=2048, expect_rays=None):
    integ = gpu.PathIntegrator(sd)
    film = integ.render_rows()
    img = integ.resolve(film)
    rc_full = integ.ray_counts()
    h, w = integ.film_shape()
    spp = sd.sampler["pixelsamples"]
    assert np.isfinite(img).all() and img.mean() > 1e-3
    assert int(rc_full[0]) == h * w * spp
    if expect_rays is not None:  # the full-size counts the oracle-checked round-1 runs recorded (profiles/r1_config_*_full.json)
        assert [int(x) for x in rc_full] == list(expect_rays)
    osc = oracle.OracleScene(sd)
    rng = np.random.Generator(np.random.PCG64(11))
    ps = np.stack([rng.integers(0, w, n_li), rng.integers(0, h, n_li), rng.integers(0, spp, n_li)], axis=1).astype(np.int32)
    li, rays = integ.li(ps)
    oli = osc.li(ps)
    assert rays.tobytes() == osc.camera_rays(ps).tobytes()
    close = np.isclose(li, oli, rtol=2e-3, atol=1e-5).all(1).mean()
    bit = (li.view(np.uint32) == oli.view(np.uint32)).all(1).mean()
    assert close == 1.0, "only %.4f of the sampled radiances agree within 2e-3" % close
    assert bit >= bit_min, "only %.4f of the sampled radiances are bit-identical (expected >= %.3f)" % (bit, bit_min)
    del integ, osc
    if crop < 1.0:
        lo, hi = 0.5 - crop / 2, 0.5 + crop / 2
        sd.film["cropwindow"] = (lo, hi, lo, hi)
    integ2 = gpu.PathIntegrator(sd)
    gimg = integ2.render()
    ref, stats, _ = oracle.OracleScene(sd).render()
    rel = ss.rel_rmse(gimg, ref)
    assert gimg.shape == ref.shape and rel <= TOL, "relative RMSE %.3e > %.0e" % (rel, TOL)
    rc = integ2.ray_counts()
    assert [int(x) for x in rc] == [int(x) for x in stats[:3]], "ray counts differ from the oracle's: %s vs %s" % (rc, stats[:3])
    return bit, rel


def test_c1_full(gpu, oracle):
    """C1: scenes/shapes/plymesh.pbrt stand-in, 217 802 triangles, path maxdepth 5, 400x400 @ 16 spp, infinite light."""
    from pbrt_v3_rs_b200 import workloads as wl
    _check_config(gpu, oracle, wl.scene_c1(), crop=1.0, bit_min=0.90, expect_rays=(2560000, 8626267, 1537067))


def test_c3_full(gpu, oracle):
    """C3: 871 204 triangles, matte / plastic / glass / metal, area + point light, maxdepth 8, 1920x1080 @ 64 spp."""
    from pbrt_v3_rs_b200 import workloads as wl
    _check_config(gpu, oracle, wl.scene_c3(), crop=0.1, bit_min=0.999, expect_rays=(132710400, 267922855, 101808588))


def test_c5_full(gpu, oracle):
    """C5: TransformedPrimitive two-level BVH, 100 352 002 instanced triangles, infinite light, 1920x1080 @ 128 spp."""
    from pbrt_v3_rs_b200 import workloads as wl
    _check_config(gpu, oracle, wl.scene_c5(), crop=0.05, bit_min=0.90, n_li=1024, expect_rays=(265420800, 1388427731, 298320541))


def test_c4_full(gpu, oracle):
    """C4: 10 018 804 triangles, 1920x1080 @ 256 spp (the configuration BASELINE shards over 2 / 4 / 8 GPUs)."""
    from pbrt_v3_rs_b200 import workloads as wl
    _check_config(gpu, oracle, wl.scene_c4(), crop=0.06, bit_min=0.90, expect_rays=(530841600, 1370728745, 435980554))
