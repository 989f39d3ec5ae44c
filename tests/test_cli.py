"""b200pt_render: the C++ command-line driver over the C ABI (pbrt-v3-rs_b200/cli/b200pt_render.cpp), the counterpart for
this path of the reference's `pbrt-v3-rs` binary."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "pbrt-v3-rs_b200", "b200pt_render")

SCENE = '''
LookAt 0 2 -5  0 0 0  0 1 0
Camera "perspective" "float fov" [40]
Film "image" "integer xresolution" [40] "integer yresolution" [30] "string filename" ["cli_out.exr"]
Sampler "halton" "integer pixelsamples" [4]
Integrator "%s" "integer maxdepth" [3]
WorldBegin
LightSource "point" "rgb I" [40 40 40] "point from" [2 4 -3]
LightSource "infinite" "rgb L" [0.4 0.4 0.5]
Material "matte" "rgb Kd" [0.5 0.4 0.3]
Shape "trianglemesh" "integer indices" [0 2 1 0 3 2] "point P" [-6 0 -6  6 0 -6  6 0 6  -6 0 6]
AttributeBegin
  Material "glass"
  Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-1 0 0  1 0 0  1 1.5 0  -1 1.5 0]
AttributeEnd
WorldEnd
'''


def test_cli_help_and_missing_scene(pkg):
    assert os.path.exists(CLI), "b200pt_render is not built (python pbrt-v3-rs_b200/build.py)"
    r = subprocess.run([CLI, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "usage: b200pt_render" in r.stdout
    r = subprocess.run([CLI], capture_output=True, text=True)
    assert r.returncode == 2 and "no scene file" in r.stderr


def test_cli_fails_loudly_without_a_device(pkg, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    (tmp_path / "s.pbrt").write_text(SCENE % "path")
    r = subprocess.run([CLI, str(tmp_path / "s.pbrt")], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("integrator", ["path", "whitted", "directlighting"])
def test_cli_renders_the_same_image_as_the_library_calls(gpu, tmp_path, integrator):
    path = tmp_path / "s.pbrt"
    path.write_text(SCENE % integrator)
    out = tmp_path / "img.pfm"
    r = subprocess.run([CLI, "--outfile", str(out), str(path)], capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr
    assert "integrator %s" % integrator in r.stdout and "wrote" in r.stdout
    img = gpu.read_pfm(str(out))
    ref = gpu.PathIntegrator(gpu.load_pbrt(str(path))).render()
    assert img.shape == (30, 40, 3) and img.mean() > 0
    assert np.array_equal(img, ref)
    # default output name: the Film's "filename"; an extension this path does not write (.exr) becomes .pfm
    r = subprocess.run([CLI, str(path)], capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0 and (tmp_path / "cli_out.pfm").exists()
    # .png: the reference's 8-bit sRGB encode (core/src/image_io.rs:384-390)
    png = tmp_path / "img.png"
    r = subprocess.run([CLI, "--outfile", str(png), str(path)], capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr
    from PIL import Image
    got = np.array(Image.open(str(png))).astype(np.int32)
    v = np.asarray(ref, dtype=np.float32)
    g = np.where(v <= np.float32(0.0031308), np.float32(12.92) * v, np.float32(1.055) * np.power(np.maximum(v, 0), np.float32(1 / 2.4)) - np.float32(0.055))
    want = np.clip(np.float32(255.0) * g + np.float32(0.5), 0, 255).astype(np.uint8).astype(np.int32)
    assert np.abs(got - want).max() <= 1 and (got == want).mean() > 0.999  # powf rounding may move a value sitting on a .5 boundary


@pytest.mark.gpu
def test_cli_multi_device_render_from_one_cpp_process(gpu, tmp_path):
    """b200pt_multi_create / _render / _info / _destroy driven from C++ (--devices): one process, every visible GPU,
    bands gathered on the first device; the image equals the single-device render bit for bit (box filter: every
    sample is taken once, by exactly one device)."""
    import torch
    path = tmp_path / "s.pbrt"
    path.write_text(SCENE % "path")
    ref = gpu.PathIntegrator(gpu.load_pbrt(str(path))).render()
    n = torch.cuda.device_count()
    for devs in sorted({1, min(2, n), n}):
        out = tmp_path / ("multi%d.pfm" % devs)
        r = subprocess.run([CLI, "--devices", ",".join(str(k) for k in range(devs)), "--outfile", str(out), str(path)], capture_output=True, text=True, cwd=str(tmp_path))
        assert r.returncode == 0, r.stderr
        assert "%d devices, band gather" % devs in r.stdout
        assert np.array_equal(gpu.read_pfm(str(out)), ref), devs
