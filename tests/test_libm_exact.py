"""The device sin/cos (pbrt-v3-rs_b200/csrc/libm_exact.cuh) must reproduce the host libm bit for bit:
the reference's f32::sin/cos resolve to glibc's sinf/cosf, and sampled directions feed ray geometry."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_build_matches_glibc_on_2e7_inputs(tmp_path):
    exe = str(tmp_path / "libm_exact_check")
    subprocess.run(["g++", "-O2", "-mfma", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tests", "libm_exact_check.cpp"), "-lm"], check=True)
    out = subprocess.run([exe, "20000000"], capture_output=True, text=True)
    n, bad_s, bad_c = (int(x) for x in out.stdout.split())
    assert (bad_s, bad_c) == (0, 0), "sinf/cosf differ from the host libm on %d / %d of %d inputs" % (bad_s, bad_c, n)


@pytest.mark.gpu
def test_sampled_directions_bit_exact_on_device(gpu, oracle):
    """Matte + area/point lights only (no acosf/atan2f on the path): every per-sample radiance is bit-identical."""
    import scenes_small as ss
    from pbrt_v3_rs_b200 import workloads as wl
    for name in ("matte", "plastic", "glass", "metal"):
        sd = ss.one_material_scene(wl, ss.MATERIALS[name], light="area", res=24, spp=8, maxdepth=6)
        sd.add_point_light((1.5, 3.0, -3.0), (30, 30, 30))
        ps = np.array([(x, y, s) for y in range(24) for x in range(24) for s in range(8)], dtype=np.int32)
        li, _ = gpu.PathIntegrator(sd).li(ps)
        oli = oracle.OracleScene(sd).li(ps)
        same = (li.view(np.uint32) == oli.view(np.uint32)).all(1)
        assert same.mean() >= 0.999, "%s: only %.5f of the samples are bit-identical" % (name, same.mean())
