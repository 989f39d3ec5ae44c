"""Host-side scene ingestion (b200pt_load_pbrt, PLY, PFM — SURVEY.md §8f rank 3), no GPU needed.

The same scene is described twice — as a pbrt-v3 scene file (+ binary / ascii PLY meshes + a PFM environment map) read
by the C++ loader, and through the Python SceneDescription mirror — and the two b200pt_scene_desc must agree: geometry,
flags, materials, lights, film, sampler and integrator byte for byte, camera matrices to a few ulps (the two paths use
different tan() implementations), BVH node arrays identical.  The oracle then renders both to the same image."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

import scenes_small as ss


def _index_mesh(tri, uv=None, nrm=None):
    """De-indexed (n, 9) triangles -> (P, indices[, uv, N]) with one vertex per corner (keeps corner attributes exact)."""
    P = tri.reshape(-1, 3)
    idx = np.arange(P.shape[0], dtype=np.int32)
    return P, idx, None if uv is None else uv.reshape(-1, 2), None if nrm is None else nrm.reshape(-1, 3)


def _write_ply(path, P, idx, N=None, UV=None, binary=True, quads=False):
    n_face = len(idx) // (4 if quads else 3)
    props = ["property float x", "property float y", "property float z"]
    if N is not None:
        props += ["property float nx", "property float ny", "property float nz"]
    if UV is not None:
        props += ["property float u", "property float v"]
    hdr = ["ply", "format %s 1.0" % ("binary_little_endian" if binary else "ascii"), "comment test mesh", "element vertex %d" % len(P)] + props + \
          ["element face %d" % n_face, "property list uchar int vertex_indices", "end_header"]
    cols = [P] + ([N] if N is not None else []) + ([UV] if UV is not None else [])
    V = np.concatenate(cols, axis=1).astype("<f4")
    k = 4 if quads else 3
    with open(path, "wb") as f:
        f.write(("\n".join(hdr) + "\n").encode())
        if binary:
            f.write(V.tobytes())
            for i in range(n_face):
                f.write(struct.pack("<B%di" % k, k, *[int(x) for x in idx[k * i:k * i + k]]))
        else:
            for row in V:
                f.write((" ".join(repr(float(x)) for x in row) + "\n").encode())
            for i in range(n_face):
                f.write(("%d %s\n" % (k, " ".join(str(int(x)) for x in idx[k * i:k * i + k]))).encode())


def _fl(a):
    return " ".join(repr(float(np.float32(x))) for x in np.asarray(a).reshape(-1))


def _build_pair(tmp_path, pkg, wl, with_instances=False):
    """Returns (path of the .pbrt file, equivalent SceneDescription)."""
    from pbrt_v3_rs_b200.scene import SceneDescription
    sd = SceneDescription()
    img = wl.sky_image(16, 8)
    pkg.write_pfm(str(tmp_path / "sky.pfm"), img)
    tv, uv, nrm = wl.displaced_sphere(12, 6, radius=0.8, with_attrs=True)
    quad = wl.ground_quad()
    lq = np.array([[-1.5, 3.0, -1.0], [1.5, 3.0, -1.0], [1.5, 3.0, 1.0], [-1.5, 3.0, 1.0]], dtype=np.float32)
    lt = np.stack([np.concatenate([lq[0], lq[1], lq[2]]), np.concatenate([lq[0], lq[2], lq[3]])])
    tv2 = wl.displaced_sphere(8, 4, radius=0.3, seed=5)

    P, idx, UV, N = _index_mesh(tv, uv, nrm)
    _write_ply(str(tmp_path / "sphere.ply"), P, idx, N=N, UV=UV, binary=True)
    P2, idx2, _, _ = _index_mesh(tv2)
    _write_ply(str(tmp_path / "small.ply"), P2, idx2, binary=False)

    m_plastic = sd.add_material(type="plastic", Kd=(0.3, 0.2, 0.1), Ks=(0.2, 0.2, 0.2), roughness=0.05)
    m_matte = sd.add_material(type="matte", Kd=(0.4, 0.4, 0.4))
    m_glass = sd.add_material(type="glass", eta=1.4)
    m_black = sd.add_material(type="matte", Kd=(0.0, 0.0, 0.0))
    m_metal = sd.add_material(type="metal", roughness=0.02)
    sd.add_mesh(tv, m_plastic, uv=uv, normals=nrm)
    sd.add_mesh(quad, m_matte)
    sd.add_mesh(lt, m_black, area_light=dict(L=(15 * 0.5, 15 * 0.5, 15 * 0.5)))
    lines = ['LookAt 0 1.2 -4  0 -0.1 0  0 1 0', 'Camera "perspective" "float fov" [ 40 ]',
             'Film "image" "integer xresolution" [20] "integer yresolution" [16] "string filename" "out.pfm"',
             'Sampler "halton" "integer pixelsamples" 4', 'PixelFilter "box"',
             'Integrator "path" "integer maxdepth" [4] "string lightsamplestrategy" "power"',
             'Accelerator "bvh" "integer maxnodeprims" [4]', '# a comment', 'WorldBegin',
             'MakeNamedMaterial "shiny" "string type" "plastic" "rgb Kd" [0.3 0.2 0.1] "rgb Ks" [0.2 0.2 0.2] "float roughness" 0.05',
             'AttributeBegin', '  NamedMaterial "shiny"', '  Shape "plymesh" "string filename" "sphere.ply"', 'AttributeEnd',
             'AttributeBegin', '  Material "matte" "rgb Kd" [0.4 0.4 0.4]',
             '  Shape "trianglemesh" "integer indices" [%s] "point P" [%s]' % (" ".join(str(i) for i in range(6)), _fl(quad)), 'AttributeEnd',
             'AttributeBegin', '  Material "glass" "float eta" 1.4', 'AttributeEnd',   # defined but unused: material index 2 stays reserved
             'AttributeBegin', '  Material "matte" "rgb Kd" [0 0 0]', '  AreaLightSource "diffuse" "rgb L" [15 15 15] "rgb scale" [0.5 0.5 0.5]',
             '  Shape "trianglemesh" "integer indices" [0 1 2 3 4 5] "point P" [%s]' % _fl(lt), 'AttributeEnd']
    if with_instances:
        obj = sd.add_object(tv2, m_metal)
        lines += ['ObjectBegin "ball"', '  Material "metal" "float roughness" 0.02', '  Shape "plymesh" "string filename" "small.ply"', 'ObjectEnd']
        for k, (tx, sc) in enumerate([(-1.2, 1.0), (1.3, 0.7)]):
            lines += ['AttributeBegin', '  Translate %r 0.9 -0.4' % tx, '  Rotate 30 0 1 0', '  Scale %r %r %r' % (sc, sc, sc), '  ObjectInstance "ball"', 'AttributeEnd']
            T = np.eye(4, dtype=np.float32); T[0, 3], T[1, 3], T[2, 3] = tx, 0.9, -0.4
            c, s_ = np.float32(np.cos(np.float32(np.deg2rad(np.float32(30))))), np.float32(np.sin(np.float32(np.deg2rad(np.float32(30)))))
            R = np.array([[c, 0, s_, 0], [0, 1, 0, 0], [-s_, 0, c, 0], [0, 0, 0, 1]], dtype=np.float32)
            S = np.diag([sc, sc, sc, 1]).astype(np.float32)
            sd.add_instance(obj, (T @ R @ S).astype(np.float32))
    else:
        sd.add_material(type="metal", roughness=0.02)  # keep material tables aligned (not needed without instances)
        sd.materials.pop()
    lines += ['LightSource "point" "point from" [1.5 3 -3] "rgb I" [30 30 30]',
              'AttributeBegin', '  Rotate -90 1 0 0', '  LightSource "infinite" "rgb L" [1 0.9 0.8] "string mapname" "sky.pfm"', 'AttributeEnd', 'WorldEnd']
    # lights in file order: area (2 triangles), point, infinite -> same order in the mirror
    sd.add_point_light((1.5, 3.0, -3.0), (30, 30, 30))
    c, s_ = np.float32(np.cos(np.float32(np.deg2rad(np.float32(-90))))), np.float32(np.sin(np.float32(np.deg2rad(np.float32(-90)))))
    rot = np.array([[1, 0, 0, 0], [0, c, -s_, 0], [0, s_, c, 0], [0, 0, 0, 1]], dtype=np.float32)
    sd.add_infinite_light((1.0, 0.9, 0.8), image=img, light_to_world=rot)
    sd.camera.update(eye=(0.0, 1.2, -4.0), look=(0.0, -0.1, 0.0), up=(0, 1, 0), fov=40.0)
    sd.film.update(xresolution=20, yresolution=16)
    sd.sampler.update(type="halton", pixelsamples=4)
    sd.integrator.update(maxdepth=4, lightsamplestrategy="power")
    path = tmp_path / "scene.pbrt"
    path.write_text("\n".join(lines) + "\n")
    return str(path), sd


def _arr(ptr, n, dtype):
    if not ptr or n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n * np.dtype(dtype).itemsize,)).view(dtype).copy()


@pytest.mark.parametrize("with_instances", [False, True])
def test_loader_matches_python_mirror_and_oracle_renders_agree(tmp_path, pkg, oracle, with_instances):
    from pbrt_v3_rs_b200 import workloads as wl
    path, sd = _build_pair(tmp_path, pkg, wl, with_instances)
    ld = pkg.load_pbrt(path)
    a, b = ld.to_desc(), sd.to_desc()
    assert ld.output == "out.pfm"
    assert a.n_prims == b.n_prims and a.n_top_tris == b.n_top_tris and a.n_lights == b.n_lights and a.n_instances == b.n_instances
    n = a.n_prims
    assert _arr(a.tri_verts, 9 * n, np.float32).tobytes() == _arr(b.tri_verts, 9 * n, np.float32).tobytes()
    assert np.array_equal(_arr(a.prim_flags, n, np.uint32), _arr(b.prim_flags, n, np.uint32))
    assert np.array_equal(_arr(a.prim_light, n, np.int32), _arr(b.prim_light, n, np.int32))
    fa, fb = _arr(a.prim_flags, n, np.uint32), _arr(b.prim_flags, n, np.uint32)
    has_uv = (fa & 16) != 0
    assert has_uv.sum() == 2 * 12 * 6
    ua, ub = _arr(a.tri_uvs, 6 * n, np.float32).reshape(n, 6), _arr(b.tri_uvs, 6 * n, np.float32).reshape(n, 6)
    na, nb = _arr(a.tri_normals, 9 * n, np.float32).reshape(n, 9), _arr(b.tri_normals, 9 * n, np.float32).reshape(n, 9)
    assert ua[has_uv].tobytes() == ub[has_uv].tobytes() and na[has_uv].tobytes() == nb[has_uv].tobytes()
    # materials referenced by the primitives are the same records (indices may differ: the file defines them in its own order)
    ma = np.ctypeslib.as_array(C.cast(a.materials, C.POINTER(C.c_uint8)), shape=(a.n_materials * C.sizeof(pkg.Material),)).reshape(a.n_materials, -1)
    mb = np.ctypeslib.as_array(C.cast(b.materials, C.POINTER(C.c_uint8)), shape=(b.n_materials * C.sizeof(pkg.Material),)).reshape(b.n_materials, -1)
    pa, pb = _arr(a.prim_material, n, np.int32), _arr(b.prim_material, n, np.int32)
    assert np.array_equal(ma[pa], mb[pb])
    # lights: type, position, radiance, primitive; the infinite light's transform and map
    for i in range(a.n_lights):
        la, lb = C.cast(a.lights, C.POINTER(pkg.Light))[i], C.cast(b.lights, C.POINTER(pkg.Light))[i]
        assert la.type == lb.type and list(la.L) == list(lb.L) and la.prim == lb.prim and list(la.pos) == list(lb.pos)
        if la.type == pkg.LIGHT_INFINITE:
            assert np.allclose(list(la.light_to_world), list(lb.light_to_world), atol=1e-7) and (la.map_width, la.map_height) == (16, 8)
            assert np.array_equal(_arr(la.map_rgb, 16 * 8 * 3, np.float32), _arr(lb.map_rgb, 16 * 8 * 3, np.float32))
    assert bytes(a.film) == bytes(b.film) and bytes(a.sampler) == bytes(b.sampler) and bytes(a.integrator) == bytes(b.integrator)
    assert np.allclose(list(a.camera.camera_to_world), list(b.camera.camera_to_world), atol=1e-6)
    assert np.allclose(list(a.camera.raster_to_camera), list(b.camera.raster_to_camera), rtol=1e-5, atol=1e-7)
    # same BVH (the loader calls the same host builder over the same bounds)
    assert a.n_nodes == b.n_nodes
    assert _arr(a.nodes, a.n_nodes, pkg.NODE_DTYPE).tobytes() == _arr(b.nodes, b.n_nodes, pkg.NODE_DTYPE).tobytes()
    if with_instances:
        ia, ib = C.cast(a.instances, C.POINTER(pkg.Instance)), C.cast(b.instances, C.POINTER(pkg.Instance))
        for i in range(a.n_instances):
            assert ia[i].object == ib[i].object and np.allclose(list(ia[i].instance_to_world), list(ib[i].instance_to_world), atol=1e-6)
            # the file's inverse is the PRODUCT of the directives' inverses (transform.rs:640-660), the mirror inverts numerically
            assert np.allclose(list(ia[i].world_to_instance), list(ib[i].world_to_instance), atol=1e-5)
    # the oracle renders both descriptions to (nearly) the same image
    ra = oracle.OracleScene(ld).render(nthreads=4)[0]
    rb = oracle.OracleScene(sd).render(nthreads=4)[0]
    assert ra.shape == rb.shape == (16, 20, 3) and np.isfinite(ra).all() and ra.any()
    assert ss.rel_rmse(ra, rb) <= 2e-2


def test_ply_variants_pfm_roundtrip_and_transform_stack(tmp_path, pkg):
    from pbrt_v3_rs_b200 import workloads as wl
    # PFM round trip (top row first in memory, bottom row first on disk)
    img = wl.sky_image(8, 4)
    pkg.write_pfm(str(tmp_path / "a.pfm"), img)
    assert np.array_equal(pkg.read_pfm(str(tmp_path / "a.pfm")), img)
    raw = open(tmp_path / "a.pfm", "rb").read()
    assert raw.startswith(b"PF\n8 4\n-1.0\n") and np.frombuffer(raw[-8 * 4 * 12:], dtype="<f4")[:3].tolist() == img[3, 0].tolist()
    # quads split as (0 1 2)(3 0 2) (plymesh.rs:211-245); ascii == binary; CTM = Translate * Scale applied to P; ReverseOrientation
    P = np.float32([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]])
    for binary in (True, False):
        _write_ply(str(tmp_path / "q.ply"), P, np.int32([0, 1, 2, 3]), binary=binary, quads=True)
        (tmp_path / "s.pbrt").write_text('Camera "perspective"\nIntegrator "path" "string lightsamplestrategy" "uniform"\nWorldBegin\nTranslate 1 2 3\nScale 2 2 -2\n'
                                         'ReverseOrientation\nShape "plymesh" "string filename" "q.ply"\nWorldEnd\n')
        ld = pkg.load_pbrt(str(tmp_path / "s.pbrt"))  # owns the arrays the desc points into
        d = ld.to_desc()
        v = _arr(d.tri_verts, 18, np.float32).reshape(2, 3, 3)
        assert np.array_equal(v[0], [[1, 2, 3], [3, 2, 3], [3, 4, 3]]) and np.array_equal(v[1], [[1, 4, 3], [1, 2, 3], [3, 4, 3]])
        # Scale(2,2,-2) swaps handedness, ReverseOrientation is on: flip = reverse ^ swaps = false, bit 8 (reverse) set
        assert _arr(d.prim_flags, 2, np.uint32).tolist() == [8, 8]
        assert d.n_materials == 1 and d.film.xres == 1280 and d.sampler.spp == 16 and d.integrator.max_depth == 5
    # unsupported input fails loudly, with the directive named
    (tmp_path / "u.pbrt").write_text('Integrator "bdpt"\nWorldBegin\nShape "trianglemesh" "integer indices" [0 1 2] "point P" [0 0 0 1 0 0 0 1 0]\nWorldEnd\n')
    with pytest.raises(pkg.B200PTError, match="bdpt"):
        pkg.load_pbrt(str(tmp_path / "u.pbrt"))
    (tmp_path / "t.pbrt").write_text('WorldBegin\nTexture "x" "spectrum" "marble"\nWorldEnd\n')  # spectrum textures: constant and checkerboard only (test_spectrum_textures.py)
    with pytest.raises(pkg.B200PTError, match="Texture"):
        pkg.load_pbrt(str(tmp_path / "t.pbrt"))
    (tmp_path / "sph.pbrt").write_text('WorldBegin\nShape "sphere"\nWorldEnd\n')
    with pytest.raises(pkg.B200PTError, match="sphere"):
        pkg.load_pbrt(str(tmp_path / "sph.pbrt"))
    with pytest.raises(pkg.B200PTError):
        pkg.load_pbrt(str(tmp_path / "missing.pbrt"))


@pytest.mark.parametrize("name,params,kw", [("triangle", '"float xwidth" 1.5 "float ywidth" 2.5', dict(radius=(1.5, 2.5))), ("mitchell", '"float B" 0.2 "float C" 0.4', dict(B=0.2, Cm=0.4)),
                                            ("mitchell", "", {}), ("sinc", '"float tau" 2.5', dict(tau=2.5)), ("sinc", "", {}), ("gaussian", "", {}), ("box", "", {})])
def test_loader_filter_tables_match_the_mirror_and_closed_forms(tmp_path, pkg, name, params, kw):
    """PixelFilter box / gaussian / triangle / mitchell / sinc (filters/src/*.rs) sampled at Film::new's 16x16 table points
    (film/mod.rs:113-125): loader == mirror, and spot values against float64 closed forms."""
    from pbrt_v3_rs_b200.scene import filter_table
    (tmp_path / "f.pbrt").write_text('Camera "perspective"\nPixelFilter "%s" %s\nFilm "image" "integer xresolution" [8] "integer yresolution" [8]\nWorldBegin\n'
                                     'Shape "trianglemesh" "integer indices" [0 1 2] "point P" [0 0 1 1 0 1 0 1 1]\nLightSource "point"\nWorldEnd\n' % (name, params))
    loaded = pkg.load_pbrt(str(tmp_path / "f.pbrt"))  # owns the description's memory
    d = loaded.to_desc()
    tab, (rx, ry) = filter_table(name, kw.pop("radius", None), **kw)
    got = np.array(list(d.film.filter_table), dtype=np.float32)
    assert (d.film.filter_radius[0], d.film.filter_radius[1]) == (rx, ry)
    assert np.allclose(got, tab, rtol=2e-6, atol=1e-7)  # sinc: numpy's sinf vs libm's
    x, y = (3 + 0.5) * rx / 16.0, (9 + 0.5) * ry / 16.0  # table entry [9][3]
    if name == "triangle":
        want = max(0.0, rx - x) * max(0.0, ry - y)
    elif name == "sinc":
        tau = kw.get("tau", 3.0)
        sinc = lambda v: 1.0 if abs(v) < 1e-5 else np.sin(np.pi * abs(v)) / (np.pi * abs(v))
        want = sinc(x) * sinc(x / tau) * sinc(y) * sinc(y / tau)
    elif name == "mitchell":
        B, Cc = kw.get("B", 1 / 3), kw.get("Cm", 1 / 3)

        def m1(v):
            v = abs(2 * v)
            if v > 1:
                return ((-B - 6 * Cc) * v ** 3 + (6 * B + 30 * Cc) * v ** 2 + (-12 * B - 48 * Cc) * v + (8 * Cc + 24 * Cc)) / 6  # the reference's constant term (pbrt has 8 B + 24 C)
            return ((12 - 9 * B - 6 * Cc) * v ** 3 + (-18 + 12 * B + 6 * Cc) * v ** 2 + (6 - 2 * B)) / 6
        want = m1(x / rx) * m1(y / ry)
    else:
        return
    assert np.isclose(got[9 * 16 + 3], want, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("camera", ["orthographic", "environment"])
def test_loader_spot_light_and_other_cameras_match_the_mirror(tmp_path, pkg, oracle, camera):
    """LightSource "spot" (the ignored "conedelta" included), Camera "orthographic" / "environment" and a checkerboard Kd: the
    loader's description equals the mirror's (light record bit for bit under an identity CTM) and the oracle renders agree."""
    from pbrt_v3_rs_b200 import workloads as wl
    from pbrt_v3_rs_b200.scene import SceneDescription
    quad = wl.ground_quad()
    (tmp_path / "s.pbrt").write_text('\n'.join([
        'LookAt 0 1.2 -4  0 -0.1 0  0 1 0', 'Camera "%s"%s' % (camera, ' "float screenwindow" [-3 3 -3 3]' if camera == "orthographic" else ""),
        'Film "image" "integer xresolution" [20] "integer yresolution" [16] "string filename" "o.png"', 'Sampler "halton" "integer pixelsamples" 4',
        'Integrator "whitted" "integer maxdepth" 3', 'WorldBegin',
        'LightSource "spot" "rgb I" [10 20 30] "rgb scale" [2 2 2] "point from" [1 3 -2] "point to" [0 -0.5 0] "float coneangle" 28 "float conedeltaangle" 12 "float conedelta" 99',
        'Texture "c" "spectrum" "checkerboard" "float uscale" 6 "float vscale" 6 "rgb tex1" [.2 .3 .4] "rgb tex2" [.8 .7 .6]',
        'Material "matte" "texture Kd" "c"',
        'Shape "trianglemesh" "integer indices" [0 1 2 3 4 5] "point P" [%s] "float st" [0 0 1 1 1 0 0 0 0 1 1 1]' % _fl(quad), 'WorldEnd']) + '\n')
    sd = SceneDescription()
    t = sd.add_spectrum_texture("checkerboard", uscale=6.0, vscale=6.0, tex1=(0.2, 0.3, 0.4), tex2=(0.8, 0.7, 0.6))
    sd.add_mesh(quad, sd.add_material(type="matte", Kd=("texture", t)), uv=np.array([[0, 0, 1, 1, 1, 0], [0, 0, 0, 1, 1, 1]], dtype=np.float32))
    sd.add_spot_light((20, 40, 60), (1.0, 3.0, -2.0), (0.0, -0.5, 0.0), coneangle=28.0, conedeltaangle=12.0)
    sd.camera.update(eye=(0.0, 1.2, -4.0), look=(0.0, -0.1, 0.0), up=(0, 1, 0), type=camera, fov=90.0, screenwindow=(-3, 3, -3, 3) if camera == "orthographic" else None)
    sd.film.update(xresolution=20, yresolution=16)
    sd.sampler.update(type="halton", pixelsamples=4)
    sd.integrator.update(name="whitted", maxdepth=3)
    ld = pkg.load_pbrt(str(tmp_path / "s.pbrt"))
    a, b = ld.to_desc(), sd.to_desc()
    la, lb = C.cast(a.lights, C.POINTER(pkg.Light))[0], C.cast(b.lights, C.POINTER(pkg.Light))[0]
    assert la.type == lb.type == pkg.LIGHT_SPOT and list(la.L) == list(lb.L) == [20.0, 40.0, 60.0] and list(la.pos) == list(lb.pos)
    assert (la.cos_total_width, la.cos_falloff_start) == (lb.cos_total_width, lb.cos_falloff_start)
    assert np.isclose(la.cos_total_width, np.cos(np.deg2rad(28.0)), atol=1e-7) and np.isclose(la.cos_falloff_start, np.cos(np.deg2rad(16.0)), atol=1e-7)
    wa, wb = np.array(list(la.world_to_light), np.float32).reshape(4, 4), np.array(list(lb.world_to_light), np.float32).reshape(4, 4)
    assert wa[:3, :3].tobytes() == wb[:3, :3].tobytes() and np.allclose(wa, wb, atol=1e-6)
    assert a.camera.type == b.camera.type == {"orthographic": pkg.CAMERA_ORTHOGRAPHIC, "environment": pkg.CAMERA_ENVIRONMENT}[camera]
    assert np.allclose(list(a.camera.camera_to_world), list(b.camera.camera_to_world), atol=1e-6)
    if camera == "orthographic":
        assert np.allclose(list(a.camera.raster_to_camera), list(b.camera.raster_to_camera), rtol=1e-5, atol=1e-7)
    assert a.n_spectrum_textures == b.n_spectrum_textures == 1
    assert bytes(C.cast(a.spectrum_textures, C.POINTER(pkg.SpectrumTexture))[0]) == bytes(C.cast(b.spectrum_textures, C.POINTER(pkg.SpectrumTexture))[0])
    ra, rb = oracle.OracleScene(ld).render(nthreads=2)[0], oracle.OracleScene(sd).render(nthreads=2)[0]
    assert ra.shape == rb.shape == (16, 20, 3) and ra.any() and ss.rel_rmse(ra, rb) <= 2e-2


@pytest.mark.parametrize("mode", ["RGB", "RGBA", "L", "LA", "P"])
def test_loader_decodes_8_bit_pngs_like_the_reference(tmp_path, pkg, mode):
    """read_8_bit (core/src/image_io.rs:192-218): image::open(..).into_rgb8(), value / 255, no gamma - for every 8-bit PNG
    colour type (the decoder inflates with zlib and undoes all five scanline filters; PIL is the independent decoder)."""
    from PIL import Image
    rng = np.random.default_rng(3)
    a = np.clip(np.cumsum(rng.integers(-9, 10, (23, 31, 3)), axis=1) + 128, 0, 255).astype(np.uint8)  # smooth rows: the encoder picks varied filters
    Image.fromarray(a).convert(mode).save(str(tmp_path / "m.png"))
    (tmp_path / "s.pbrt").write_text('Camera "perspective"\nWorldBegin\nLightSource "goniometric" "string mapname" "m.png"\n'
                                     'Shape "trianglemesh" "integer indices" [0 1 2] "point P" [0 0 1 1 0 1 0 1 1]\nWorldEnd\n')
    ld = pkg.load_pbrt(str(tmp_path / "s.pbrt"))
    light = C.cast(ld.to_desc().lights, C.POINTER(pkg.Light))[0]
    assert (light.type, light.map_width, light.map_height) == (pkg.LIGHT_GONIOMETRIC, 31, 23)
    got = _arr(light.map_rgb, 23 * 31 * 3, np.float32).reshape(23, 31, 3)
    want = np.array(Image.open(str(tmp_path / "m.png")).convert("RGB")).astype(np.float32) / np.float32(255)
    assert np.array_equal(got, want)


def test_loader_projection_light_matches_the_mirror(tmp_path, pkg, oracle):
    """LightSource "projection" with a PNG map under Translate / Rotate: type, position, fov, the image and cos_total_width (z of
    the normalised screen corner through the numerically inverted perspective matrix, projection.rs:90-95) equal the mirror's."""
    from PIL import Image
    from pbrt_v3_rs_b200 import workloads as wl
    from pbrt_v3_rs_b200.scene import SceneDescription
    rng = np.random.default_rng(9)
    px = rng.integers(0, 256, (6, 14, 3)).astype(np.uint8)
    Image.fromarray(px).save(str(tmp_path / "slide.png"))
    quad = wl.ground_quad()
    (tmp_path / "s.pbrt").write_text('\n'.join([
        'LookAt 0 1.2 -4  0 -0.1 0  0 1 0', 'Camera "perspective" "float fov" 40', 'Film "image" "integer xresolution" [20] "integer yresolution" [16] "string filename" "o.png"',
        'Sampler "halton" "integer pixelsamples" 4', 'Integrator "whitted"', 'WorldBegin',
        'AttributeBegin', 'Translate 0.5 3 0', 'Rotate 90 1 0 0', 'LightSource "projection" "rgb I" [30 30 30] "float fov" 50 "string mapname" "slide.png"', 'AttributeEnd',
        'Material "matte"', 'Shape "trianglemesh" "integer indices" [0 1 2 3 4 5] "point P" [%s]' % _fl(quad), 'WorldEnd']) + '\n')
    sd = SceneDescription()
    sd.add_mesh(quad, sd.add_material(type="matte"))
    c, s_ = np.float32(np.cos(np.float32(np.deg2rad(np.float32(90))))), np.float32(np.sin(np.float32(np.deg2rad(np.float32(90)))))
    T = np.eye(4, dtype=np.float32); T[0, 3], T[1, 3] = 0.5, 3.0
    R = np.array([[1, 0, 0, 0], [0, c, -s_, 0], [0, s_, c, 0], [0, 0, 0, 1]], dtype=np.float32)
    sd.add_projection_light((30, 30, 30), image=px.astype(np.float32) / np.float32(255), light_to_world=(T @ R).astype(np.float32), fov=50.0)
    sd.camera.update(eye=(0.0, 1.2, -4.0), look=(0.0, -0.1, 0.0), up=(0, 1, 0), fov=40.0)
    sd.film.update(xresolution=20, yresolution=16)
    sd.sampler.update(type="halton", pixelsamples=4)
    sd.integrator.update(name="whitted")
    ld = pkg.load_pbrt(str(tmp_path / "s.pbrt"))
    a, b = ld.to_desc(), sd.to_desc()
    la, lb = C.cast(a.lights, C.POINTER(pkg.Light))[0], C.cast(b.lights, C.POINTER(pkg.Light))[0]
    assert la.type == lb.type == pkg.LIGHT_PROJECTION and la.fov == lb.fov == 50.0 and (la.map_width, la.map_height) == (lb.map_width, lb.map_height) == (14, 6)
    assert np.allclose(list(la.pos), list(lb.pos), atol=1e-6) and np.allclose(list(la.world_to_light), list(lb.world_to_light), atol=1e-6)
    assert np.array_equal(_arr(la.map_rgb, 6 * 14 * 3, np.float32), _arr(lb.map_rgb, 6 * 14 * 3, np.float32))
    assert np.isclose(la.cos_total_width, lb.cos_total_width, rtol=1e-6)
    # closed form: corner (aspect, 1) of the screen window at fov 50 -> direction (a t, t, 1) with t = tan(25 deg)
    t = np.tan(np.deg2rad(25.0)); asp = 14.0 / 6.0
    assert np.isclose(la.cos_total_width, 1.0 / np.sqrt((asp * t) ** 2 + t ** 2 + 1.0), rtol=1e-5)
    ra, rb = oracle.OracleScene(ld).render(nthreads=2)[0], oracle.OracleScene(sd).render(nthreads=2)[0]
    assert ra.any() and ss.rel_rmse(ra, rb) <= 2e-2
