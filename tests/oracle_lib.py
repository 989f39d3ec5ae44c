"""ctypes binding of oracle/_build/liboracle.so — TEST INFRASTRUCTURE ONLY.

The oracle is the CPU restatement of the reference algorithm (oracle/*.h).  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; the product package never does.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_build", "liboracle.so")

RAY_DTYPE = np.dtype([("o", "<f4", 3), ("tmax", "<f4"), ("d", "<f4", 3), ("time", "<f4")])
HIT_DTYPE = np.dtype([("t", "<f4"), ("prim", "<u4"), ("b0", "<f4"), ("b1", "<f4")])
DIAG_DTYPE = np.dtype([("b2", "<f4"), ("det", "<f4"), ("min_e_abs", "<f4"), ("second_t", "<f4")])
NODE_DTYPE = np.dtype([("bounds", "<f4", 6), ("offset", "<u4"), ("n_primitives", "<u2"), ("axis", "u1"), ("pad", "u1")])

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        vp, i64 = C.c_void_p, C.c_int64
        L.orc_radical_inverse.restype = C.c_float
        L.orc_radical_inverse.argtypes = [C.c_int, C.c_uint64]
        L.orc_scrambled_radical_inverse.restype = C.c_float
        L.orc_scrambled_radical_inverse.argtypes = [C.c_int, C.c_uint64]
        L.orc_gamma.restype = C.c_float
        L.orc_next_float_up.restype = C.c_float
        L.orc_next_float_up.argtypes = [C.c_float]
        L.orc_next_float_down.restype = C.c_float
        L.orc_next_float_down.argtypes = [C.c_float]
        L.orc_pcg32_stream.argtypes = [C.c_uint64, C.c_uint64, C.c_int, vp, C.c_int]
        L.orc_pcg32_floats.argtypes = [C.c_uint64, vp, C.c_int]
        L.orc_pcg32_bounded.restype = C.c_uint32
        L.orc_pcg32_bounded.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_int]
        L.orc_halton_index.restype = C.c_uint64
        L.orc_halton_pixel.argtypes = [C.c_int] * 6 + [vp]
        L.orc_bvh_build_sah.restype = i64
        L.orc_bvh_build_sah.argtypes = [vp, i64, C.c_int, vp, vp]
        L.orc_set_sobol_matrices.argtypes = [vp, i64]
        L.orc_sobol_interval_tables.argtypes = [C.c_int, vp, vp]
        L.orc_sobol_pixel.argtypes = [C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, vp]
        L.orc_sobol_pixel.restype = C.c_int
        L.orc_bvh_build_hlbvh.restype = i64
        L.orc_bvh_build_hlbvh.argtypes = [vp, i64, C.c_int, vp, vp, vp]
        L.orc_triangle_bounds.argtypes = [vp, i64, vp]
        L.orc_accel_create.restype = vp
        L.orc_accel_create.argtypes = [vp, i64, vp, vp, vp, i64]
        L.orc_accel_destroy.argtypes = [vp]
        L.orc_accel_set_uvs.argtypes = [vp, vp, i64]
        L.orc_accel_set_alpha.argtypes = [vp, vp, C.c_int, vp, vp, i64]
        L.orc_float_texture_evaluate.argtypes = [vp, vp, C.c_float, C.c_float]
        L.orc_float_texture_evaluate.restype = C.c_float
        L.orc_spectrum_texture_evaluate.argtypes = [vp, C.c_float, C.c_float, vp, vp]
        L.orc_spectrum_texture_evaluate.restype = None
        L.orc_noise_3d.argtypes = [vp, C.c_float, C.c_float, C.c_float]
        L.orc_noise_3d.restype = C.c_float
        L.orc_envmap_prepare.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp]
        L.orc_envmap_lookup.argtypes = [vp, C.c_int, C.c_int, vp, C.c_float, vp]
        L.orc_triangle_geometry.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, vp]
        L.orc_intersect_batch.argtypes = [vp, vp, i64, vp, vp, vp, C.c_int]
        L.orc_occluded_batch.argtypes = [vp, vp, i64, vp, vp, C.c_int]
        L.orc_triangle_intersect.argtypes = [vp, vp, vp]
        L.orc_bounds_intersect.argtypes = [vp, vp]
        L.orc_scene_create.restype = vp
        L.orc_scene_create.argtypes = [vp]
        L.orc_scene_destroy.argtypes = [vp]
        L.orc_render.restype = C.c_double
        L.orc_render.argtypes = [vp, vp, vp, C.c_int]
        L.orc_li_batch.argtypes = [vp, vp, i64, vp, C.c_int]
        L.orc_camera_rays.argtypes = [vp, vp, i64, vp]
        L.orc_camera_matrices.argtypes = [vp, vp, vp, C.c_float, C.c_int, C.c_int, vp, vp, vp]
        L.orc_matrix_inverse.argtypes = [vp, vp]
        for f in (L.orc_coordinate_system, L.orc_cross):
            f.argtypes = [vp, vp, vp]
        L.orc_normalize.argtypes = [vp, vp]
        L.orc_offset_ray_origin.argtypes = [vp] * 5
        L.orc_halton_permutations.argtypes = [vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _pkg():
    """The product package, for its ctypes struct definitions only (b200pt_float_texture mirrors include/b200pt.h)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import __graft_entry__ as ge
    return ge.load_package()


def ncpu():
    return os.cpu_count() or 1


def triangle_bounds(tri_verts):
    v = np.ascontiguousarray(tri_verts, dtype=np.float32).reshape(-1, 9)
    out = np.empty((v.shape[0], 6), dtype=np.float32)
    lib().orc_triangle_bounds(_p(v), v.shape[0], _p(out))
    return out


def build_bvh_sah(prim_bounds, max_prims_in_node=4):
    pb = np.ascontiguousarray(prim_bounds, dtype=np.float32).reshape(-1, 6)
    n = pb.shape[0]
    nodes = np.zeros(max(2 * n - 1, 1), dtype=NODE_DTYPE)
    ordered = np.zeros(max(n, 1), dtype=np.uint32)
    nn = lib().orc_bvh_build_sah(_p(pb), n, max_prims_in_node, _p(nodes), _p(ordered))
    return nodes[:nn].copy(), ordered[:n].copy()


def build_bvh_hlbvh(prim_bounds, max_prims_in_node=4, with_codes=False):
    """oracle restatement of BVHAccel::new(.., SplitMethod::HLBVH); raises where the reference panics."""
    pb = np.ascontiguousarray(prim_bounds, dtype=np.float32).reshape(-1, 6)
    n = pb.shape[0]
    nodes = np.zeros(max(2 * n - 1, 1), dtype=NODE_DTYPE)
    ordered = np.zeros(max(n, 1), dtype=np.uint32)
    codes = np.zeros(max(n, 1), dtype=np.uint32)
    nn = lib().orc_bvh_build_hlbvh(_p(pb), n, max_prims_in_node, _p(nodes), _p(ordered), _p(codes))
    if nn < 0:
        raise RuntimeError("the reference asserts on this input")
    return (nodes[:nn].copy(), ordered[:n].copy()) + ((codes[:n].copy(),) if with_codes else ())


class OracleAccel:
    def __init__(self, nodes, ordered, tri_verts, flags=None, tri_uvs=None):
        self.nodes = np.ascontiguousarray(nodes)
        self.ordered = np.ascontiguousarray(ordered, dtype=np.uint32)
        self.verts = np.ascontiguousarray(tri_verts, dtype=np.float32).reshape(-1, 9)
        self.flags = None if flags is None else np.ascontiguousarray(flags, dtype=np.uint32)
        if tri_uvs is not None:
            fl = np.zeros(self.verts.shape[0], dtype=np.uint32) if self.flags is None else self.flags
            self.flags = fl | np.uint32(16)  # PRIM_HAS_UV
        self.h = lib().orc_accel_create(_p(self.nodes), len(self.nodes), _p(self.ordered), _p(self.verts), _p(self.flags), self.verts.shape[0])
        self.uv = None
        if tri_uvs is not None:
            uv = self.uv = np.ascontiguousarray(tri_uvs, dtype=np.float32).reshape(-1, 6)
            lib().orc_accel_set_uvs(self.h, _p(uv), uv.shape[0])

    def set_alpha_textures(self, textures, prim_alpha_tex):
        """Same arguments as BVHAccel.set_alpha_textures; the flags gain PRIM_ALPHA_TEXTURE where an index is >= 0."""
        pkg = _pkg()
        pat = np.ascontiguousarray(prim_alpha_tex, dtype=np.int32).reshape(-1, 2)
        fl = np.zeros(self.verts.shape[0], dtype=np.uint32) if self.flags is None else self.flags.copy()
        fl[(pat >= 0).any(1)] |= np.uint32(128)
        self.flags = fl
        lib().orc_accel_destroy(self.h)
        self.h = lib().orc_accel_create(_p(self.nodes), len(self.nodes), _p(self.ordered), _p(self.verts), _p(self.flags), self.verts.shape[0])
        if getattr(self, "uv", None) is not None:
            lib().orc_accel_set_uvs(self.h, _p(self.uv), self.uv.shape[0])
        keep = []
        arr = pkg.float_texture_array(list(textures), keep)
        perm = pkg.noise_perm()
        lib().orc_accel_set_alpha(self.h, C.cast(arr, C.c_void_p), len(textures), _p(pat), _p(perm), self.verts.shape[0])

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_accel_destroy(self.h)
            self.h = None

    def intersect(self, rays, nthreads=None, counters=True, diag=True):
        r = np.ascontiguousarray(rays).view(RAY_DTYPE).reshape(-1)
        n = r.shape[0]
        hits = np.empty(n, dtype=HIT_DTYPE)
        dg = np.empty(n, dtype=DIAG_DTYPE) if diag else None
        ct = np.empty((n, 2), dtype=np.uint32) if counters else None
        lib().orc_intersect_batch(self.h, _p(r), n, _p(hits), _p(dg), _p(ct), nthreads or ncpu())
        return hits, dg, ct

    def occluded(self, rays, nthreads=None, counters=True):
        r = np.ascontiguousarray(rays).view(RAY_DTYPE).reshape(-1)
        n = r.shape[0]
        out = np.empty(n, dtype=np.uint8)
        ct = np.empty((n, 2), dtype=np.uint32) if counters else None
        lib().orc_occluded_batch(self.h, _p(r), n, _p(out), _p(ct), nthreads or ncpu())
        return out, ct


def triangle_geometry(verts9, b, uv=None, normals=None, tangents=None, flip=False, reverse=False):
    """Triangle::intersect's hit geometry -> dict of 3-vectors, or None for a degenerate hit."""
    f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float32).reshape(-1)
    v, bb, out = f(verts9), f(b), np.zeros(21, dtype=np.float32)
    u, n, s_ = f(uv), f(normals), f(tangents)
    if not lib().orc_triangle_geometry(_p(v), _p(bb), _p(u), _p(n), _p(s_), int(flip), int(reverse), _p(out)):
        return None
    names = ["p", "p_error", "n", "dpdu", "dpdv", "shading_n", "shading_dpdu"]
    return {k: out[3 * i:3 * i + 3].copy() for i, k in enumerate(names)}


def envmap_prepare(image, L=(1.0, 1.0, 1.0)):
    img = None if image is None else np.ascontiguousarray(image, dtype=np.float32)
    h, w = (0, 0) if img is None else img.shape[:2]
    Lf = np.asarray(L, dtype=np.float32)
    size = np.zeros(4, dtype=np.int32)
    lib().orc_envmap_prepare(_p(img), w, h, _p(Lf), _p(size), None, None, None)
    lvl0 = np.zeros((size[1], size[0], 3), dtype=np.float32)
    imp = np.zeros((size[3], size[2]), dtype=np.float32)
    pw = np.zeros(3, dtype=np.float32)
    lib().orc_envmap_prepare(_p(img), w, h, _p(Lf), _p(size), _p(lvl0), _p(imp), _p(pw))
    return lvl0, imp, pw


def envmap_lookup(image, st, width):
    img = np.ascontiguousarray(image, dtype=np.float32)
    out = np.zeros(3, dtype=np.float32)
    s2 = np.asarray(st, dtype=np.float32)
    lib().orc_envmap_lookup(_p(img), img.shape[1], img.shape[0], _p(s2), float(width), _p(out))
    return out


class OracleScene:
    """Oracle render of a product SceneDescription (same b200pt_scene_desc bytes)."""

    def __init__(self, scene_description):
        self.sd = scene_description
        self.desc = scene_description.to_desc()
        if self.desc.sobol_matrices_32:  # the Sobol sampler's generator matrices are data the caller supplies
            lib().orc_set_sobol_matrices(self.desc.sobol_matrices_32, 1024 * 52)
        self.h = lib().orc_scene_create(C.addressof(self.desc))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_scene_destroy(self.h)
            self.h = None

    def shape(self):
        c = self.desc.film.crop
        return (c[3] - c[1], c[2] - c[0])

    def render(self, nthreads=None):
        h, w = self.shape()
        rgb = np.empty((h, w, 3), dtype=np.float32)
        stats = np.zeros(4, dtype=np.uint64)
        secs = lib().orc_render(self.h, _p(rgb), _p(stats), nthreads or ncpu())
        return rgb, stats, secs

    def li(self, pixel_sample, nthreads=None):
        ps = np.ascontiguousarray(pixel_sample, dtype=np.int32).reshape(-1, 3)
        out = np.empty((ps.shape[0], 3), dtype=np.float32)
        lib().orc_li_batch(self.h, _p(ps), ps.shape[0], _p(out), nthreads or ncpu())
        return out

    def camera_rays(self, pixel_sample):
        ps = np.ascontiguousarray(pixel_sample, dtype=np.int32).reshape(-1, 3)
        out = np.empty(ps.shape[0], dtype=RAY_DTYPE)
        lib().orc_camera_rays(self.h, _p(ps), ps.shape[0], _p(out))
        return out


def exempt_mask(hits, diag):
    """Rays exempt from bit-exact comparison: hits within 4 ulp of a triangle edge
    (min |e_i| tiny relative to |det|) or with a second accepted candidate within 2 ulp in t."""
    t = hits["t"]
    with np.errstate(invalid="ignore"):
        return _exempt(hits, diag, t)


def _exempt(hits, diag, t):
    edge = diag["min_e_abs"] <= 4 * np.finfo(np.float32).eps * np.abs(diag["det"])
    tie = np.abs(diag["second_t"] - t) <= 2 * np.spacing(np.abs(t))
    return (hits["prim"] != 0xFFFFFFFF) & (edge | tie)
