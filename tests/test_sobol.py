"""SobolSampler (samplers/src/sobol.rs, core/src/low_discrepency.rs:1770-1845): oracle KATs, the derived pixel <-> index
tables, and GPU parity."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import scenes_small as ss

REF_TABLES = "/root/reference/core/src/sobol_matrices.rs"


def _m32(pkg):
    return pkg.sobol_matrices_32()


def test_generator_matrices_are_the_sobol_sequence(pkg):
    """Dimension 0 is the van der Corput sequence (identity matrix: bit reversal), dimension 1 the Pascal-triangle matrix of
    the (0,2)-sequence (column j = binomial(j, i) mod 2) - checkable without any reference file."""
    m = _m32(pkg).reshape(1024, 52)
    assert [int(x) for x in m[0, :32]] == [1 << (31 - j) for j in range(32)] and not m[0, 32:].any()
    for j in range(32):
        col = 0
        for i in range(j + 1):
            if (j & i) == i:  # Lucas: C(j, i) odd iff i is a submask of j
                col |= 1 << (31 - i)
        assert int(m[1, j]) == col, j


@pytest.mark.parametrize("which", ["oracle", "product"])
def test_interval_tables_match_the_reference_literals(pkg, oracle, which):
    """VD_C_SOBOL_MATRICES / VD_C_SOBOL_MATRICES_INV are derived from dimensions 0 and 1 (oracle/oracle_sobol.h,
    csrc/host_sampler.cpp: two independent derivations); with the reference mounted they are compared with its literal
    tables, otherwise with each other and with the defining property."""
    m32 = _m32(pkg)
    oracle.lib().orc_set_sobol_matrices(m32.ctypes.data_as(C.c_void_p), m32.size)

    def derived(m, src):
        a, b = np.zeros(52, np.uint64), np.zeros(52, np.uint64)
        if src == "oracle":
            oracle.lib().orc_sobol_interval_tables(m, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p))
        else:
            pkg.lib().b200pt_sobol_interval_tables.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
            assert pkg.lib().b200pt_sobol_interval_tables(m32.ctypes.data_as(C.c_void_p), m, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p)) == 0
        return [int(x) for x in a], [int(x) for x in b]

    for m in (1, 2, 5, 11, 25):
        assert derived(m, "oracle") == derived(m, "product")
    if os.path.exists(REF_TABLES):
        text = open(REF_TABLES).read()

        def table(name):
            start = text.index("pub const " + name)
            body = re.sub(r"//[^\n]*", "", text[text.index("= [", start) + 3:text.index("];\n", start)])
            return [[int(t, 16) if t.startswith("0x") else int(t) for t in re.findall(r"0x[0-9a-fA-F]+|\d+", r)] for r in re.findall(r"\[([^\[\]]*)\]", body)]
        vdc, inv = table("VD_C_SOBOL_MATRICES"), table("VD_C_SOBOL_MATRICES_INV")
        for m in range(1, 26):
            a, b = derived(m, which)
            assert a == vdc[m - 1] and b == inv[m - 1], m


def _np_sobol(m32, index, dim):
    v = 0
    row = m32.reshape(1024, 52)[dim]
    i = 0
    while index:
        if index & 1:
            v ^= int(row[i])
        index >>= 1
        i += 1
    return min(np.float32(v) * np.float32(2.0 ** -32), np.float32(1.0) - np.float32(2.0 ** -24))


def test_oracle_sampler_values_positions_and_stratification(pkg, oracle):
    m32 = _m32(pkg)
    oracle.lib().orc_set_sobol_matrices(m32.ctypes.data_as(C.c_void_p), m32.size)
    sb = np.array([0, 0, 24, 20], dtype=np.int32)  # resolution 32, m = 5
    spp, ndim = 4, 9
    seen = set()
    for (px, py) in [(0, 0), (1, 0), (7, 13), (23, 19), (16, 16)]:
        out = np.zeros((spp, ndim), np.float32)
        idx = np.zeros(spp, np.uint64)
        n = oracle.lib().orc_sobol_pixel(spp, sb.ctypes.data_as(C.c_void_p), px, py, ndim, out.ctypes.data_as(C.c_void_p), idx.ctypes.data_as(C.c_void_p))
        assert n == spp
        for s in range(spp):
            i = int(idx[s])
            assert i not in seen  # every (pixel, sample) owns a distinct point of the global sequence
            seen.add(i)
            # the sample's first two dimensions, scaled to the 32 x 32 grid, fall into its pixel
            x, y = float(_np_sobol(m32, i, 0)) * 32, float(_np_sobol(m32, i, 1)) * 32
            assert int(x) == px and int(y) == py
            assert np.isclose(out[s, 0], x - px, atol=1e-5) and np.isclose(out[s, 1], y - py, atol=1e-5)
            for d in range(2, ndim):
                assert out[s, d] == _np_sobol(m32, i, d)
        # (0,2)-sequence: the 4 samples of a pixel are stratified 2x2, 4x1 and 1x4 inside it
        u = out[:, :2]
        assert len({(int(a * 2), int(b * 2)) for a, b in u}) == 4
        assert len({int(a * 4) for a, _ in u}) == 4 and len({int(b * 4) for _, b in u}) == 4


def test_non_power_of_two_spp_is_rounded_up_and_render_agrees_with_halton(pkg, oracle):
    from pbrt_v3_rs_b200 import workloads as wl
    sd = ss.one_material_scene(wl, ss.MATERIALS["matte"], light="all", res=16, spp=24, strategy="power")
    h = oracle.OracleScene(sd).render()[0]
    sd.sampler.update(type="sobol")
    img, stats, _ = oracle.OracleScene(sd).render()
    assert stats[0] == 16 * 16 * 32  # sobol.rs:28-37: 24 -> 32 samples per pixel
    assert np.isfinite(img).all() and abs(img.mean() - h.mean()) <= 0.05 * h.mean()


@pytest.mark.gpu
@pytest.mark.parametrize("integrator", ["path", "whitted", "directlighting"])
def test_sobol_gpu_matches_oracle(gpu, oracle, integrator):
    from pbrt_v3_rs_b200 import workloads as wl
    sd = ss.one_material_scene(wl, ss.MATERIALS["glass" if integrator != "path" else "plastic"], light="all", res=24, spp=4, maxdepth=4, strategy="power")
    sd.sampler.update(type="sobol")
    sd.integrator.update(name=integrator)
    integ = gpu.PathIntegrator(sd)
    osc = oracle.OracleScene(sd)
    ps = np.array([(x, y, s) for y in range(24) for x in range(24) for s in range(4)], dtype=np.int32)
    li, rays = integ.li(ps)
    assert rays.tobytes() == osc.camera_rays(ps).tobytes()  # film position from dimensions 0 / 1 incl. the pixel offset and clamp
    oli = osc.li(ps)
    close = np.isclose(li, oli, rtol=2e-3, atol=1e-5).all(1)
    assert close.mean() >= 0.97, close.mean()
    img = integ.render()
    ref, stats, _ = osc.render()
    assert ss.rel_rmse(img, ref) <= 1e-3
    assert integ.ray_counts()[0] == stats[0]


@pytest.mark.gpu
def test_sobol_scene_file_and_crop_window(gpu, oracle, tmp_path):
    scene = tmp_path / "s.pbrt"
    scene.write_text('''
LookAt 0 2 -5  0 0 0  0 1 0
Camera "perspective" "float fov" [40]
Film "image" "integer xresolution" [48] "integer yresolution" [32] "float cropwindow" [0.25 0.75 0.2 0.9] "string filename" ["s.pfm"]
Sampler "sobol" "integer pixelsamples" [6]
Integrator "path" "integer maxdepth" [3] "string lightsamplestrategy" ["uniform"]
WorldBegin
LightSource "point" "rgb I" [40 40 40] "point from" [2 4 -3]
LightSource "infinite" "rgb L" [0.4 0.4 0.5]
Material "matte" "rgb Kd" [0.5 0.4 0.3]
Shape "trianglemesh" "integer indices" [0 2 1 0 3 2] "point P" [-6 0 -6  6 0 -6  6 0 6  -6 0 6]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-1 0 0  1 0 0  1 1.5 0  -1 1.5 0]
WorldEnd
''')
    ls = gpu.load_pbrt(str(scene))
    assert ls.to_desc().sampler.type == gpu.SAMPLER_SOBOL
    integ = gpu.PathIntegrator(ls)
    img = integ.render()
    ref, stats, _ = oracle.OracleScene(ls).render()
    assert img.shape == ref.shape and ss.rel_rmse(img, ref) <= 1e-3
    assert integ.ray_counts()[0] == stats[0] == img.shape[0] * img.shape[1] * 8  # 6 -> 8 samples per pixel
