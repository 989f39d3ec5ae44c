"""pbrt-v3-rs_b200 — B200-native ray-intersection + path-integration hot path.

Host-side mirror (Python, over the C ABI in include/b200pt.h) of the reference
interfaces this path replaces:

  * ``BVHAccel``        accelerators/src/bvh/mod.rs:22-360 (``new``, ``from_params``,
                        ``world_bound``, ``intersect``, ``intersect_p``)
  * ``PathIntegrator``  integrators/src/path.rs:22-326 (``from_params``, ``preprocess``,
                        ``render``, ``li``)

The directory name is not a valid Python identifier; load it with
``__graft_entry__.load_package()`` which registers it as ``pbrt_v3_rs_b200``.

All compute goes through ``libb200pt.so`` (hand-written sm_100a CUDA).  There is
no CPU fallback: if the library is missing or no sm_100 device is bound, calls
raise ``B200PTError``.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200pt.so")

MISS = 0xFFFFFFFF
PRIM_FLIP_NORMAL, PRIM_ALPHA_ZERO, PRIM_SHADOW_ALPHA_ZERO = 1, 2, 4
PRIM_REVERSE_ORIENTATION, PRIM_HAS_UV, PRIM_HAS_NORMALS, PRIM_HAS_TANGENTS = 8, 16, 32, 64
MAT_MATTE, MAT_PLASTIC, MAT_GLASS, MAT_METAL, MAT_MIRROR = 0, 1, 2, 3, 4
LIGHT_POINT, LIGHT_AREA, LIGHT_INFINITE, LIGHT_DISTANT, LIGHT_SPOT, LIGHT_GONIOMETRIC, LIGHT_PROJECTION = 0, 1, 2, 3, 4, 5, 6
SAMPLER_HALTON, SAMPLER_ZEROTWO, SAMPLER_SOBOL = 0, 1, 2
LIGHTS_UNIFORM, LIGHTS_POWER, LIGHTS_SPATIAL = 0, 1, 2
INTEGRATOR_PATH, INTEGRATOR_WHITTED, INTEGRATOR_DIRECT = 0, 1, 2
DIRECT_ALL, DIRECT_ONE = 0, 1

# numpy views of the 32-byte ray, 16-byte hit and 32-byte node records
RAY_DTYPE = np.dtype([("o", "<f4", 3), ("tmax", "<f4"), ("d", "<f4", 3), ("time", "<f4")])
HIT_DTYPE = np.dtype([("t", "<f4"), ("prim", "<u4"), ("b0", "<f4"), ("b1", "<f4")])
NODE_DTYPE = np.dtype([("bounds", "<f4", 6), ("offset", "<u4"), ("n_primitives", "<u2"), ("axis", "u1"), ("pad", "u1")])
assert RAY_DTYPE.itemsize == 32 and HIT_DTYPE.itemsize == 16 and NODE_DTYPE.itemsize == 32


def sobol_matrices_32():
    """SOBOL_MATRICES_32 of core/src/sobol_matrices.rs (1024 x 52 u32), from pbrt-v3-rs_b200/data (tools/extract_sobol_matrices.py)."""
    global _sobol
    if _sobol is None:
        _sobol = np.fromfile(os.path.join(_HERE, "data", "sobol_matrices_32.bin"), dtype="<u4")
        assert _sobol.size == 1024 * 52
    return _sobol


_sobol = None


class B200PTError(RuntimeError):
    pass


class Material(C.Structure):
    _fields_ = [("type", C.c_int32), ("kd", C.c_float * 3), ("ks", C.c_float * 3), ("kt", C.c_float * 3),
                ("eta", C.c_float * 3), ("k", C.c_float * 3), ("sigma", C.c_float), ("urough", C.c_float),
                ("vrough", C.c_float), ("remap_roughness", C.c_int32)]


class Light(C.Structure):
    _fields_ = [("type", C.c_int32), ("pos", C.c_float * 3), ("L", C.c_float * 3), ("prim", C.c_int32),
                ("two_sided", C.c_int32), ("light_to_world", C.c_float * 16), ("world_to_light", C.c_float * 16),
                ("map_rgb", C.c_void_p), ("map_width", C.c_int32), ("map_height", C.c_int32),
                ("cos_total_width", C.c_float), ("cos_falloff_start", C.c_float), ("fov", C.c_float), ("pad_", C.c_int32)]


CAMERA_PERSPECTIVE, CAMERA_ORTHOGRAPHIC, CAMERA_ENVIRONMENT = 0, 1, 2


class Camera(C.Structure):
    _fields_ = [("raster_to_camera", C.c_float * 16), ("camera_to_world", C.c_float * 16), ("lens_radius", C.c_float),
                ("focal_distance", C.c_float), ("shutter_open", C.c_float), ("shutter_close", C.c_float), ("type", C.c_int32)]


class Film(C.Structure):
    _fields_ = [("xres", C.c_int32), ("yres", C.c_int32), ("crop", C.c_int32 * 4), ("filter_radius", C.c_float * 2),
                ("filter_table", C.c_float * 256), ("scale", C.c_float), ("max_sample_luminance", C.c_float)]


class Sampler(C.Structure):
    _fields_ = [("type", C.c_int32), ("spp", C.c_int32), ("sample_at_center", C.c_int32), ("dimensions", C.c_int32)]


class Integrator(C.Structure):
    _fields_ = [("max_depth", C.c_int32), ("rr_threshold", C.c_float), ("pixel_bounds", C.c_int32 * 4),
                ("light_strategy", C.c_int32), ("type", C.c_int32), ("direct_strategy", C.c_int32)]


class Object(C.Structure):
    _fields_ = [("nodes", C.c_void_p), ("n_nodes", C.c_int64), ("ordered_prims", C.c_void_p), ("first_prim", C.c_int64), ("n_prims", C.c_int64)]


class Instance(C.Structure):
    _fields_ = [("object", C.c_int32), ("instance_to_world", C.c_float * 16), ("world_to_instance", C.c_float * 16)]


class FloatTexture(C.Structure):
    _fields_ = [("type", C.c_int32), ("su", C.c_float), ("sv", C.c_float), ("du", C.c_float), ("dv", C.c_float), ("value", C.c_float * 2),
                ("wrap", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("texels", C.c_void_p)]


TEX_CONSTANT, TEX_CHECKERBOARD, TEX_DOTS, TEX_IMAGEMAP = 0, 1, 2, 3
PRIM_ALPHA_TEXTURE = 128


def noise_perm():
    """NOISE_PERM[0..256) (tools/extract_noise_perm.py)."""
    return np.fromfile(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "noise_perm.bin"), dtype=np.uint8)


def float_texture_array(textures, keep):
    """ctypes array of b200pt_float_texture from the mirror's texture dicts (scene.SceneDescription.add_float_texture)."""
    arr = (FloatTexture * max(1, len(textures)))()
    for k, t in enumerate(textures):
        T = arr[k]
        T.type = {"constant": TEX_CONSTANT, "checkerboard": TEX_CHECKERBOARD, "dots": TEX_DOTS, "imagemap": TEX_IMAGEMAP}[t["type"]]
        T.su, T.sv, T.du, T.dv = t.get("uscale", 1.0), t.get("vscale", 1.0), t.get("udelta", 0.0), t.get("vdelta", 0.0)
        if t["type"] == "constant":
            T.value[:] = (t.get("value", 1.0), 0.0)
        elif t["type"] == "checkerboard":  # checkerboard_2d.rs:127-128 defaults
            T.value[:] = (t.get("tex1", 1.0), t.get("tex2", 0.0))
        elif t["type"] == "dots":
            # dots.rs:82-86: Self::new(inside, outside, map) feeds new(outside_dot, inside_dot, mapping)
            T.value[:] = (t.get("inside", 1.0), t.get("outside", 0.0))
        else:
            tex = np.ascontiguousarray(t["texels"], dtype=np.float32)
            keep.append(tex)
            T.height, T.width = tex.shape
            T.texels = tex.ctypes.data_as(C.c_void_p)
            T.wrap = {"repeat": 0, "black": 1, "clamp": 2}[t.get("wrap", "repeat")]
    keep.append(arr)
    return arr


class SpectrumTexture(C.Structure):
    _fields_ = [("type", C.c_int32), ("su", C.c_float), ("sv", C.c_float), ("du", C.c_float), ("dv", C.c_float), ("tex1", C.c_float * 3),
                ("tex2", C.c_float * 3), ("aa_closedform", C.c_int32)]


STEX_CONSTANT, STEX_CHECKERBOARD = 0, 1


def spectrum_texture_array(textures, keep):
    """ctypes array of b200pt_spectrum_texture from the mirror's dicts (scene.SceneDescription.add_spectrum_texture)."""
    arr = (SpectrumTexture * max(1, len(textures)))()
    for k, t in enumerate(textures):
        T = arr[k]
        T.type = {"constant": STEX_CONSTANT, "checkerboard": STEX_CHECKERBOARD}[t["type"]]
        T.su, T.sv, T.du, T.dv = t.get("uscale", 1.0), t.get("vscale", 1.0), t.get("udelta", 0.0), t.get("vdelta", 0.0)
        if t["type"] == "constant":
            T.tex1[:] = t.get("value", (1.0, 1.0, 1.0))
        else:  # checkerboard_2d.rs:127-128, 134: tex1 = 1, tex2 = 0, aamode closedform
            T.tex1[:] = t.get("tex1", (1.0, 1.0, 1.0))
            T.tex2[:] = t.get("tex2", (0.0, 0.0, 0.0))
        T.aa_closedform = 0 if t.get("aamode", "closedform") == "none" else 1
    keep.append(arr)
    return arr


class SceneDesc(C.Structure):
    _fields_ = [("nodes", C.c_void_p), ("n_nodes", C.c_int64), ("ordered_prims", C.c_void_p), ("tri_verts", C.c_void_p),
                ("prim_flags", C.c_void_p), ("prim_material", C.c_void_p), ("prim_light", C.c_void_p), ("n_prims", C.c_int64),
                ("materials", C.c_void_p), ("n_materials", C.c_int32), ("lights", C.c_void_p), ("n_lights", C.c_int32),
                ("camera", Camera), ("film", Film), ("sampler", Sampler), ("integrator", Integrator),
                ("n_top_tris", C.c_int64), ("objects", C.c_void_p), ("n_objects", C.c_int32), ("instances", C.c_void_p), ("n_instances", C.c_int32),
                ("tri_uvs", C.c_void_p), ("tri_normals", C.c_void_p), ("tri_tangents", C.c_void_p), ("sobol_matrices_32", C.c_void_p),
                ("float_textures", C.c_void_p), ("n_float_textures", C.c_int32), ("prim_alpha_tex", C.c_void_p), ("noise_perm", C.c_void_p),
                ("spectrum_textures", C.c_void_p), ("n_spectrum_textures", C.c_int32), ("material_kd_tex", C.c_void_p)]


_lib = None


def lib():
    """Loads libb200pt.so; fails loudly when the CUDA extension is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200PTError("libb200pt.so is not built (run `python pbrt-v3-rs_b200/build.py`); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
    L.b200pt_last_error.restype = C.c_char_p
    L.b200pt_device_l2_bytes.restype = i64
    L.b200pt_launch_count.restype = i64
    L.b200pt_envmap_prepare.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp]
    L.b200pt_load_pbrt.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.b200pt_loaded_scene_desc.argtypes = [vp]
    L.b200pt_loaded_scene_desc.restype = C.POINTER(SceneDesc)
    L.b200pt_loaded_scene_output.argtypes = [vp]
    L.b200pt_loaded_scene_output.restype = C.c_char_p
    L.b200pt_loaded_scene_free.argtypes = [vp]
    L.b200pt_loaded_scene_free.restype = None
    L.b200pt_write_pfm.argtypes = [C.c_char_p, vp, i32, i32]
    L.b200pt_read_pfm.argtypes = [C.c_char_p, vp, vp]
    L.b200pt_write_png.argtypes = [C.c_char_p, vp, i32, i32]
    L.b200pt_write_image.argtypes = [C.c_char_p, vp, i32, i32]
    L.b200pt_init.argtypes = [C.c_int]
    L.b200pt_bvh_build_sah.argtypes = [vp, i64, C.c_int, vp, C.POINTER(i64), vp]
    L.b200pt_triangle_bounds.argtypes = [vp, i64, vp]
    L.b200pt_bvh_build_hlbvh.argtypes = [vp, i64, C.c_int, vp, C.POINTER(i64), vp]
    L.b200pt_bvh_build_hlbvh_gpu.argtypes = [vp, i64, C.c_int, vp, C.POINTER(i64), vp]
    L.b200pt_hlbvh_morton_codes.argtypes = [vp, i64, vp]
    L.b200pt_bvh_build_sah_gpu.argtypes = [vp, i64, C.c_int, vp, C.POINTER(i64), vp]
    L.b200pt_bvh_build_sah_device.argtypes = [vp, i64, C.c_int, vp, C.POINTER(i64), vp, vp]
    L.b200pt_triangle_bounds_device.argtypes = [vp, i64, vp, vp]
    L.b200pt_bvh_build_hlbvh_device.argtypes = [vp, i64, C.c_int, vp, C.POINTER(i64), vp, vp]
    L.b200pt_accel_create.argtypes = [vp, i64, vp, vp, vp, i64, C.POINTER(vp)]
    L.b200pt_accel_create_uv.argtypes = [vp, i64, vp, vp, vp, vp, i64, C.POINTER(vp)]
    L.b200pt_accel_create_device.argtypes = [vp, i64, vp, C.c_int, vp, C.POINTER(vp)]
    L.b200pt_accel_set_alpha_textures.argtypes = [vp, vp, i32, vp, vp, vp, vp]
    L.b200pt_accel_download.argtypes = [vp, vp, C.POINTER(i64), vp]
    L.b200pt_accel_destroy.argtypes = [vp]
    L.b200pt_accel_destroy.restype = None
    L.b200pt_accel_world_bound.argtypes = [vp, vp]
    L.b200pt_accel_intersect1.argtypes = [vp, vp, vp]
    L.b200pt_accel_occluded1.argtypes = [vp, vp, vp]
    L.b200pt_intersect_batch.argtypes = [vp, vp, i64, vp]
    L.b200pt_occluded_batch.argtypes = [vp, vp, i64, vp]
    L.b200pt_intersect_batch_device.argtypes = [vp, vp, i64, vp, vp, C.c_int]
    L.b200pt_occluded_batch_device.argtypes = [vp, vp, i64, vp, vp, C.c_int]
    L.b200pt_count_work_device.argtypes = [vp, vp, i64, C.c_int, vp, vp]
    L.b200pt_scene_create.argtypes = [C.POINTER(SceneDesc), C.POINTER(vp)]
    L.b200pt_scene_destroy.argtypes = [vp]
    L.b200pt_scene_destroy.restype = None
    L.b200pt_render_rows.argtypes = [vp, i32, i32, vp]
    L.b200pt_render_rows_device.argtypes = [vp, i32, i32, vp, vp]
    L.b200pt_render_shard_device.argtypes = [vp, i32, i32, i32, vp, vp]
    L.b200pt_band_owner.argtypes = [i32, i32]
    L.b200pt_band_owner.restype = i32
    L.b200pt_render_shard_device_raw.argtypes = [vp, i32, i32, i32, vp, vp]
    L.b200pt_film_finish_device.argtypes = [vp, i64, vp, vp]
    L.b200pt_film_resolve.argtypes = [C.POINTER(Film), vp, vp]
    L.b200pt_li_batch.argtypes = [vp, vp, i64, vp, vp]
    L.b200pt_scene_ray_counts.argtypes = [vp, vp]
    L.b200pt_scene_set_memory_budget.argtypes = [vp, C.c_uint64]
    L.b200pt_set_device.argtypes = [C.c_int]
    L.b200pt_multi_create.argtypes = [C.POINTER(SceneDesc), vp, i32, C.POINTER(vp)]
    L.b200pt_multi_render.argtypes = [vp, i32, vp]
    L.b200pt_multi_info.argtypes = [vp, vp, C.POINTER(C.c_double), C.POINTER(i32)]
    L.b200pt_multi_film_device.argtypes = [vp]
    L.b200pt_multi_film_device.restype = vp
    L.b200pt_multi_destroy.argtypes = [vp]
    L.b200pt_multi_destroy.restype = None
    L.b200pt_render_multi.argtypes = [C.POINTER(SceneDesc), vp, i32, i32, vp]
    _lib = L
    return L


def _check(rc, what):
    if rc != 0:
        raise B200PTError("%s failed (%d): %s" % (what, rc, lib().b200pt_last_error().decode()))


_inited = None


def init(device=0):
    """b200pt_init: binds the calling thread (and, the first time, the process default) to the sm_100 device.
    Raises when there is none.  Handles remember their device, so init(other) later does not move existing ones."""
    global _inited
    if _inited == device:
        return
    _check(lib().b200pt_init(device), "b200pt_init")
    _inited = device


def current_device():
    return int(lib().b200pt_current_device())


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def triangle_bounds(tri_verts):
    """Triangle::world_bound for n triangles (shapes/src/triangle.rs:427-431)."""
    v = np.ascontiguousarray(tri_verts, dtype=np.float32).reshape(-1, 9)
    out = np.empty((v.shape[0], 6), dtype=np.float32)
    _check(lib().b200pt_triangle_bounds(_ptr(v), v.shape[0], _ptr(out)), "b200pt_triangle_bounds")
    return out


GPU_BUILD_MIN_PRIMS = 4096  # below this the sequential host build is faster than ~200 kernel launches


def build_bvh_sah(prim_bounds, max_prims_in_node=4, where="auto"):
    """BVHAccel::new(.., SplitMethod::SAH) (mod.rs:43-153): returns (nodes, ordered_prims).

    where="host": the sequential C++ builder; where="gpu": the level-parallel CUDA builder (csrc/bvh_build.cu).
    Both return the same bytes, so "auto" only picks the faster one: the GPU once a device is bound (init()) and
    the input has at least GPU_BUILD_MIN_PRIMS primitives."""
    pb = np.ascontiguousarray(prim_bounds, dtype=np.float32).reshape(-1, 6)
    n = pb.shape[0]
    if where == "auto":
        where = "gpu" if (_inited is not None and n >= GPU_BUILD_MIN_PRIMS) else "host"
    nodes = np.zeros(max(2 * n - 1, 1), dtype=NODE_DTYPE)
    ordered = np.zeros(max(n, 1), dtype=np.uint32)
    nn = C.c_int64(0)
    if where == "gpu":
        init(_inited if _inited is not None else 0)
        fn, name = lib().b200pt_bvh_build_sah_gpu, "b200pt_bvh_build_sah_gpu"
    elif where == "host":
        fn, name = lib().b200pt_bvh_build_sah, "b200pt_bvh_build_sah"
    else:
        raise ValueError("where must be 'host' or 'gpu'")
    _check(fn(_ptr(pb), n, int(max_prims_in_node), _ptr(nodes), C.byref(nn), _ptr(ordered)), name)
    return nodes[:nn.value].copy(), ordered[:n].copy()


def build_bvh_hlbvh(prim_bounds, max_prims_in_node=4, where="host"):
    """BVHAccel::new(.., SplitMethod::HLBVH) (hlbvh.rs:33-449): returns (nodes, ordered_prims); where = "host" | "gpu" | "auto"
    (same bytes; "auto" = the GPU once a device is bound and the input has at least GPU_BUILD_MIN_PRIMS primitives)."""
    pb = np.ascontiguousarray(prim_bounds, dtype=np.float32).reshape(-1, 6)
    n = pb.shape[0]
    nodes = np.zeros(max(2 * n - 1, 1), dtype=NODE_DTYPE)
    ordered = np.zeros(max(n, 1), dtype=np.uint32)
    nn = C.c_int64(0)
    if where == "auto":
        where = "gpu" if (_inited is not None and n >= GPU_BUILD_MIN_PRIMS) else "host"
    if where == "gpu":
        init(_inited if _inited is not None else 0)
        fn, name = lib().b200pt_bvh_build_hlbvh_gpu, "b200pt_bvh_build_hlbvh_gpu"
    else:
        fn, name = lib().b200pt_bvh_build_hlbvh, "b200pt_bvh_build_hlbvh"
    _check(fn(_ptr(pb), n, int(max_prims_in_node), _ptr(nodes), C.byref(nn), _ptr(ordered)), name)
    return nodes[:nn.value].copy(), ordered[:n].copy()


def hlbvh_morton_codes(prim_bounds):
    pb = np.ascontiguousarray(prim_bounds, dtype=np.float32).reshape(-1, 6)
    codes = np.zeros(max(pb.shape[0], 1), dtype=np.uint32)
    _check(lib().b200pt_hlbvh_morton_codes(_ptr(pb), pb.shape[0], _ptr(codes)), "b200pt_hlbvh_morton_codes")
    return codes[:pb.shape[0]]


class LoadedScene:
    """A scene file read by the host-side loader (b200pt_load_pbrt): duck-types SceneDescription.to_desc() for
    PathIntegrator, so ``PathIntegrator(load_pbrt("scene.pbrt")).render()`` is the reference's `pbrt scene.pbrt`."""

    def __init__(self, path):
        h = C.c_void_p()
        _check(lib().b200pt_load_pbrt(os.fsencode(path), C.byref(h)), "b200pt_load_pbrt")
        self._h = h
        self.path = path

    def to_desc(self):
        return lib().b200pt_loaded_scene_desc(self._h).contents

    @property
    def output(self):
        return lib().b200pt_loaded_scene_output(self._h).decode()

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.b200pt_loaded_scene_free(self._h)
            self._h = None

    __del__ = close


def load_pbrt(path):
    return LoadedScene(path)


def write_pfm(path, rgb):
    a = np.ascontiguousarray(rgb, dtype=np.float32)
    _check(lib().b200pt_write_pfm(os.fsencode(path), _ptr(a), a.shape[1], a.shape[0]), "b200pt_write_pfm")


def write_image(path, rgb):
    """write_image (core/src/image_io.rs): .png = the reference's 8-bit sRGB encode, .pfm = float."""
    a = np.ascontiguousarray(rgb, dtype=np.float32)
    _check(lib().b200pt_write_image(os.fsencode(path), _ptr(a), a.shape[1], a.shape[0]), "b200pt_write_image")


def read_pfm(path):
    size = np.zeros(2, dtype=np.int32)
    _check(lib().b200pt_read_pfm(os.fsencode(path), None, _ptr(size)), "b200pt_read_pfm")
    out = np.zeros((size[1], size[0], 3), dtype=np.float32)
    _check(lib().b200pt_read_pfm(os.fsencode(path), _ptr(out), _ptr(size)), "b200pt_read_pfm")
    return out


def envmap_prepare(image, L=(1.0, 1.0, 1.0)):
    """Host-side InfiniteAreaLight::new preparation (no GPU): -> (level0 (h0, w0, 3), importance (2 h0, 2 w0), power_lookup (3,))."""
    img = None if image is None else np.ascontiguousarray(image, dtype=np.float32)
    h, w = (0, 0) if img is None else img.shape[:2]
    Lf = np.asarray(L, dtype=np.float32)
    size = np.zeros(4, dtype=np.int32)
    _check(lib().b200pt_envmap_prepare(_ptr(img), w, h, _ptr(Lf), _ptr(size), None, None, None), "b200pt_envmap_prepare")
    lvl0 = np.zeros((size[1], size[0], 3), dtype=np.float32)
    imp = np.zeros((size[3], size[2]), dtype=np.float32)
    pw = np.zeros(3, dtype=np.float32)
    _check(lib().b200pt_envmap_prepare(_ptr(img), w, h, _ptr(Lf), _ptr(size), _ptr(lvl0), _ptr(imp), _ptr(pw)), "b200pt_envmap_prepare")
    return lvl0, imp, pw


class BVHAccel:
    """Device-resident BVHAccel (accelerators/src/bvh/mod.rs).

    ``BVHAccel.from_params({"splitmethod": "sah", "maxnodeprims": 4}, tri_verts)``
    mirrors ``impl From<(&ParamSet, &[ArcPrimitive])> for BVHAccel`` (mod.rs:339-360)
    for triangle primitives given as an (n, 9) float32 array of world-space vertices.
    """

    def __init__(self, tri_verts, nodes, ordered_prims, prim_flags=None, tri_uvs=None):
        init(_inited if _inited is not None else 0)
        self.tri_verts = np.ascontiguousarray(tri_verts, dtype=np.float32).reshape(-1, 9)
        self.nodes = np.ascontiguousarray(nodes, dtype=NODE_DTYPE)
        self.ordered_prims = np.ascontiguousarray(ordered_prims, dtype=np.uint32)
        self.prim_flags = None if prim_flags is None else np.ascontiguousarray(prim_flags, dtype=np.uint32)
        self.tri_uvs = None if tri_uvs is None else np.ascontiguousarray(tri_uvs, dtype=np.float32).reshape(-1, 6)
        if self.tri_uvs is not None:  # every primitive of a mesh with "uv"/"st" uses them
            fl = np.zeros(self.tri_verts.shape[0], dtype=np.uint32) if self.prim_flags is None else self.prim_flags
            self.prim_flags = fl | np.uint32(PRIM_HAS_UV)
        h = C.c_void_p()
        _check(lib().b200pt_accel_create_uv(_ptr(self.nodes), len(self.nodes), _ptr(self.ordered_prims), _ptr(self.tri_verts), _ptr(self.tri_uvs),
                                            _ptr(self.prim_flags), self.tri_verts.shape[0], C.byref(h)), "b200pt_accel_create_uv")
        self._h = h

    @classmethod
    def from_params(cls, params, tri_verts, prim_flags=None, tri_uvs=None):
        split = params.get("splitmethod", "sah")
        if split not in ("sah", "hlbvh"):
            # middle / equal are outside this path
            raise B200PTError("BVHAccel: splitmethod 'sah' and 'hlbvh' are built on this path, got %r" % split)
        max_prims = int(params.get("maxnodeprims", 4)) & 0xFF
        tv = np.ascontiguousarray(tri_verts, dtype=np.float32).reshape(-1, 9)
        init(_inited if _inited is not None else 0)
        nodes, ordered = (build_bvh_sah if split == "sah" else build_bvh_hlbvh)(triangle_bounds(tv), max_prims, where="auto")
        return cls(tv, nodes, ordered, prim_flags, tri_uvs)

    @classmethod
    def from_device_triangles(cls, d_tri_verts_ptr, n_prims, max_prims_in_node=4, d_prim_flags_ptr=None, stream=0, download=False):
        """BVHAccel::new for triangles already in HBM (b200pt_accel_create_device): bounds, SAH build and traversal records
        never leave the device.  download=True also fetches nodes / ordered_prims (what the host builders return)."""
        init(_inited if _inited is not None else 0)
        self = cls.__new__(cls)
        h = C.c_void_p()
        _check(lib().b200pt_accel_create_device(d_tri_verts_ptr, int(n_prims), d_prim_flags_ptr, int(max_prims_in_node) & 0xFF, stream, C.byref(h)),
               "b200pt_accel_create_device")
        self._h = h
        self.tri_verts = self.prim_flags = self.tri_uvs = None
        self.nodes = self.ordered_prims = None
        if download:
            nn = C.c_int64(0)
            _check(lib().b200pt_accel_download(h, None, C.byref(nn), None), "b200pt_accel_download")
            self.nodes = np.zeros(nn.value, dtype=NODE_DTYPE)
            self.ordered_prims = np.zeros(int(n_prims), dtype=np.uint32)
            _check(lib().b200pt_accel_download(h, _ptr(self.nodes), C.byref(nn), _ptr(self.ordered_prims)), "b200pt_accel_download")
        return self

    def set_alpha_textures(self, textures, prim_alpha_tex):
        """Alpha masks of the meshes (shapes/src/triangle.rs:278-312): ``textures`` = texture dicts as taken by
        scene.SceneDescription.add_float_texture, ``prim_alpha_tex`` = (n, 2) indices (alpha, shadowalpha) or -1 per primitive.
        Primitives with an index >= 0 get B200PT_PRIM_ALPHA_TEXTURE."""
        pat = np.ascontiguousarray(prim_alpha_tex, dtype=np.int32).reshape(-1, 2)
        fl = np.zeros(self.tri_verts.shape[0], dtype=np.uint32) if self.prim_flags is None else self.prim_flags.copy()
        fl[(pat >= 0).any(1)] |= np.uint32(PRIM_ALPHA_TEXTURE)
        self.prim_flags, self.prim_alpha_tex, self.float_textures = fl, pat, list(textures)
        # the triangle records carry the flags: rebuild the accelerator with them, then attach the textures
        lib().b200pt_accel_destroy(self._h)
        h = C.c_void_p()
        _check(lib().b200pt_accel_create_uv(_ptr(self.nodes), len(self.nodes), _ptr(self.ordered_prims), _ptr(self.tri_verts), _ptr(self.tri_uvs),
                                            _ptr(self.prim_flags), self.tri_verts.shape[0], C.byref(h)), "b200pt_accel_create_uv")
        self._h = h
        keep = []
        arr = float_texture_array(self.float_textures, keep)
        perm = noise_perm()
        _check(lib().b200pt_accel_set_alpha_textures(self._h, C.cast(arr, C.c_void_p), len(self.float_textures), _ptr(pat), _ptr(self.tri_uvs), _ptr(self.prim_flags),
                                                     _ptr(perm)), "b200pt_accel_set_alpha_textures")

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            try:
                _lib.b200pt_accel_destroy(self._h)
            except Exception:  # interpreter shutdown
                pass
            self._h = None

    __del__ = close

    @property
    def handle(self):
        return self._h

    def world_bound(self):
        out = np.empty(6, dtype=np.float32)
        _check(lib().b200pt_accel_world_bound(self._h, _ptr(out)), "b200pt_accel_world_bound")
        return out

    def intersect(self, ray):
        """Primitive::intersect for one ray record; lowers ray['tmax'] on a hit and returns the hit record or None."""
        r = np.ascontiguousarray(ray, dtype=RAY_DTYPE).reshape(1)
        h = np.zeros(1, dtype=HIT_DTYPE)
        _check(lib().b200pt_accel_intersect1(self._h, _ptr(r), _ptr(h)), "b200pt_accel_intersect1")
        if hasattr(ray, "dtype") and ray.dtype == RAY_DTYPE:
            ray["tmax"] = r["tmax"][0]
        return None if h["prim"][0] == MISS else h[0]

    def intersect_p(self, ray):
        r = np.ascontiguousarray(ray, dtype=RAY_DTYPE).reshape(1)
        o = np.zeros(1, dtype=np.uint8)
        _check(lib().b200pt_accel_occluded1(self._h, _ptr(r), _ptr(o)), "b200pt_accel_occluded1")
        return bool(o[0])

    def intersect_batch(self, rays):
        """Closest hit for a HOST array of rays (copies in/out inside the call)."""
        r = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        h = np.empty(r.shape[0], dtype=HIT_DTYPE)
        _check(lib().b200pt_intersect_batch(self._h, _ptr(r), r.shape[0], _ptr(h)), "b200pt_intersect_batch")
        return h

    def occluded_batch(self, rays):
        r = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        o = np.empty(r.shape[0], dtype=np.uint8)
        _check(lib().b200pt_occluded_batch(self._h, _ptr(r), r.shape[0], _ptr(o)), "b200pt_occluded_batch")
        return o

    def intersect_batch_device(self, d_rays_ptr, n, d_hits_ptr, stream=0, variant=0):
        _check(lib().b200pt_intersect_batch_device(self._h, d_rays_ptr, n, d_hits_ptr, stream, variant),
               "b200pt_intersect_batch_device")

    def occluded_batch_device(self, d_rays_ptr, n, d_out_ptr, stream=0, variant=0):
        _check(lib().b200pt_occluded_batch_device(self._h, d_rays_ptr, n, d_out_ptr, stream, variant),
               "b200pt_occluded_batch_device")


def count_work_device(accel, d_rays_ptr, n, any_hit=False, d_per_ray_ptr=None):
    """(N_node, N_tri) totals of the reference-order walk over a device ray batch (roofline accounting)."""
    tot = np.zeros(2, dtype=np.uint64)
    _check(lib().b200pt_count_work_device(accel.handle, d_rays_ptr, n, 1 if any_hit else 0, _ptr(tot), d_per_ray_ptr),
           "b200pt_count_work_device")
    return int(tot[0]), int(tot[1])


def launch_count():
    return int(lib().b200pt_launch_count())


from .scene import MultiGPURender, PathIntegrator, SceneDescription  # noqa: E402,F401
