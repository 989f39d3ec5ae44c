"""Scene description + PathIntegrator mirror over the C ABI.

``SceneDescription`` gathers what the reference's ``api`` crate holds when it
reaches ``pbrt_world_end`` (api/src/lib.rs:447-507): the primitive list (here:
triangles), materials, lights, camera, film, sampler and integrator parameters,
keyed by the same ParamSet names.  ``to_desc()`` flattens it into the
``b200pt_scene_desc`` of include/b200pt.h.
"""
import ctypes as C
import math

import numpy as np

F32 = np.float32


# ---- host-side scene setup math (the reference does this in `api`, above the boundary) --
def _m4_inverse(m):
    """Matrix4x4::inverse (core/src/geometry/matrix4x4.rs:55-123), float32 Gauss-Jordan."""
    minv = np.array(m, dtype=F32).copy()
    indxc, indxr, ipiv = [0] * 4, [0] * 4, [0] * 4
    for i in range(4):
        irow = icol = 0
        big = F32(0)
        for j in range(4):
            if ipiv[j] != 1:
                for k in range(4):
                    if ipiv[k] == 0 and abs(minv[j, k]) >= big:
                        big, irow, icol = abs(minv[j, k]), j, k
        ipiv[icol] += 1
        if irow != icol:
            minv[[irow, icol]] = minv[[icol, irow]]
        indxr[i], indxc[i] = irow, icol
        pivinv = F32(1) / minv[icol, icol]
        minv[icol, icol] = F32(1)
        minv[icol, :] = minv[icol, :] * pivinv
        for j in range(4):
            if j != icol:
                save = minv[j, icol]
                minv[j, icol] = F32(0)
                minv[j, :] = minv[j, :] - minv[icol, :] * save
    for j in range(3, -1, -1):
        if indxr[j] != indxc[j]:
            minv[:, [indxr[j], indxc[j]]] = minv[:, [indxc[j], indxr[j]]]
    return minv


def _m4_mul(a, b):
    r = np.zeros((4, 4), dtype=F32)
    for i in range(4):
        for j in range(4):
            r[i, j] = F32(F32(F32(a[i, 0] * b[0, j]) + F32(a[i, 1] * b[1, j])) + F32(a[i, 2] * b[2, j])) + F32(a[i, 3] * b[3, j])
    return r


def _norm(v):
    v = np.asarray(v, dtype=F32)
    l2 = F32(F32(v[0] * v[0] + v[1] * v[1]) + v[2] * v[2])
    return v * (F32(1) / np.sqrt(l2, dtype=F32))


def _cross(a, b):
    return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]], dtype=F32)


def look_at_camera_to_world(eye, look, up):
    """Transform::look_at (core/src/geometry/transform.rs:191-214): returns camera_to_world."""
    eye, look, up = (np.asarray(x, dtype=F32) for x in (eye, look, up))
    d = _norm(look - eye)
    right = _norm(_cross(_norm(up), d))
    new_up = _cross(d, right)
    m = np.eye(4, dtype=F32)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = right, new_up, d, eye
    return m


def perspective_raster_to_camera(fov, xres, yres, screen_window=None, orthographic=False):
    """PerspectiveCamera::new / OrthographicCamera::new + ProjectiveCameraData::new (cameras/src/perspective_camera.rs:35-75,
    orthographic_camera.rs:24-41, core/src/camera.rs:276-306): raster_to_camera = inverse(camera_to_screen) * inverse(screen_to_raster)."""
    n, f = F32(1e-2), F32(1000.0)
    persp = np.eye(4, dtype=F32)
    persp[2, 2] = f / (f - n)
    persp[2, 3] = -f * n / (f - n)
    persp[3, 2] = F32(1)
    persp[3, 3] = F32(0)
    inv_tan = F32(1) / F32(math.tan(float(F32(fov) * F32(math.pi / 180.0)) / 2.0))
    # camera_to_screen = scale(inv_tan) * persp; Transform products keep m_inv = rhs.m_inv * lhs.m_inv (transform.rs:644-656)
    c2s_inv = _m4_mul(_m4_inverse(persp), np.diag([F32(1) / inv_tan, F32(1) / inv_tan, F32(1), F32(1)]).astype(F32))
    if orthographic:  # Transform::orthographic(0, 1) = scale(1, 1, 1) * translate(0, 0, -0): its stored inverse is the identity
        c2s_inv = np.eye(4, dtype=F32)
    frame = F32(xres) / F32(yres)
    if screen_window is None:
        sw = [-frame, frame, F32(-1), F32(1)] if frame > 1 else [F32(-1), F32(1), F32(-1) / frame, F32(1) / frame]
    else:
        sw = [F32(x) for x in screen_window]
    s2 = np.diag([F32(1) / (sw[1] - sw[0]), F32(1) / (sw[2] - sw[3]), F32(1), F32(1)]).astype(F32)
    # screen_to_raster = scale(res) * scale(1/window) * translate(-window origin); inverse of the product = product of the stored inverses in reverse order (transform.rs:644-656)
    s1i = np.diag([F32(1) / F32(xres), F32(1) / F32(yres), F32(1), F32(1)]).astype(F32)
    s2i = np.diag([F32(1) / s2[0, 0], F32(1) / s2[1, 1], F32(1), F32(1)]).astype(F32)
    tri = np.eye(4, dtype=F32)
    tri[0, 3], tri[1, 3] = sw[0], sw[3]
    r2s = _m4_mul(tri, _m4_mul(s2i, s1i))
    return _m4_mul(c2s_inv, r2s)


def filter_table(kind="box", radius=None, alpha=2.0, B=1.0 / 3.0, Cm=1.0 / 3.0, tau=3.0):
    """Film::new filter table (core/src/film/mod.rs:113-125), 16x16 floats, for the reference's five filters."""
    if radius is None:
        radius = {"box": (0.5, 0.5), "sinc": (4.0, 4.0)}.get(kind, (2.0, 2.0))
    rx, ry = F32(radius[0]), F32(radius[1])
    tab = np.zeros(256, dtype=F32)
    if kind == "box":
        tab[:] = 1.0
    elif kind == "gaussian":  # filters/src/gaussian.rs
        a = F32(alpha)
        ex, ey = np.exp(-a * rx * rx, dtype=F32), np.exp(-a * ry * ry, dtype=F32)
        k = 0
        for y in range(16):
            for x in range(16):
                px = (F32(x) + F32(0.5)) * rx * F32(1.0 / 16.0)
                py = (F32(y) + F32(0.5)) * ry * F32(1.0 / 16.0)
                gx = max(F32(0), np.exp(-a * px * px, dtype=F32) - ex)
                gy = max(F32(0), np.exp(-a * py * py, dtype=F32) - ey)
                tab[k] = gx * gy
                k += 1
    elif kind in ("triangle", "mitchell", "sinc"):  # filters/src/{triangle,mitchell,sinc}.rs
        b, c, ta, pi = F32(B), F32(Cm), F32(tau), F32(3.14159265358979323846)

        def mitchell_1d(x):  # mitchell.rs:41-57 (incl. its 8 C + 24 C constant term); np.float32 scalars round after every operation
            x = abs(F32(2) * F32(x))
            if x > 1:
                return ((-b - F32(6) * c) * x * x * x + (F32(6) * b + F32(30) * c) * x * x + (F32(-12) * b - F32(48) * c) * x + (F32(8) * c + F32(24) * c)) * (F32(1) / F32(6))
            return ((F32(12) - F32(9) * b - F32(6) * c) * x * x * x + (F32(-18) + F32(12) * b + F32(6) * c) * x * x + (F32(6) - F32(2) * b)) * (F32(1) / F32(6))

        def sinc(x):  # sinc.rs:72-79
            x = abs(F32(x))
            return F32(1) if x < F32(1e-5) else F32(np.sin(F32(pi * x), dtype=F32) / F32(pi * x))

        def windowed_sinc(x, radius):
            x = abs(F32(x))
            return F32(0) if x > radius else F32(sinc(x) * sinc(F32(x / ta)))

        k = 0
        for y in range(16):
            for x in range(16):
                px = (F32(x) + F32(0.5)) * rx * F32(1.0 / 16.0)
                py = (F32(y) + F32(0.5)) * ry * F32(1.0 / 16.0)
                if kind == "triangle":
                    tab[k] = max(F32(0), F32(rx - abs(px))) * max(F32(0), F32(ry - abs(py)))
                elif kind == "mitchell":
                    tab[k] = mitchell_1d(F32(px * F32(F32(1) / rx))) * mitchell_1d(F32(py * F32(F32(1) / ry)))
                else:
                    tab[k] = windowed_sinc(px, rx) * windowed_sinc(py, ry)
                k += 1
    else:
        raise ValueError("filter %r is not one of the reference's filters (box, gaussian, triangle, mitchell, sinc)" % kind)
    return tab, (float(rx), float(ry))


class SceneDescription:
    """Flat scene: triangles + per-primitive material/light + camera/film/sampler/integrator params."""

    def __init__(self):
        self.tri_verts = np.zeros((0, 9), dtype=F32)
        self.prim_material = np.zeros(0, dtype=np.int32)
        self.prim_light = np.zeros(0, dtype=np.int32)
        self.prim_flags = np.zeros(0, dtype=np.uint32)
        self.prim_alpha_tex = np.zeros((0, 2), dtype=np.int32)  # per primitive: alpha / shadowalpha float-texture index or -1
        self.float_textures = []  # dicts, see add_float_texture
        self.spectrum_textures = []  # dicts, see add_spectrum_texture
        self.materials = []   # list of dicts
        self.lights = []      # list of dicts
        self.camera = dict(eye=(0, 0, -5), look=(0, 0, 0), up=(0, 1, 0), fov=45.0, lensradius=0.0, focaldistance=1e6,
                           shutteropen=0.0, shutterclose=1.0)
        self.film = dict(xresolution=64, yresolution=64, filter="box", scale=1.0, maxsampleluminance=float("inf"))
        self.sampler = dict(type="halton", pixelsamples=16, samplepixelcenter=False, dimensions=4)
        self.integrator = dict(name="path", maxdepth=5, rrthreshold=1.0, lightsamplestrategy="uniform", pixelbounds=None)
        self.accel_params = dict(splitmethod="sah", maxnodeprims=4)
        self.nodes = None
        self.ordered_prims = None
        self.tri_uvs = self.tri_normals = self.tri_tangents = None  # optional (n, 6) / (n, 9) / (n, 9), rows of zeros for meshes without
        self.objects = []     # dicts: tri_verts, material, flags, nodes, ordered
        self.instances = []   # (object id, instance_to_world, world_to_instance)
        self._keep = []

    # -- scene assembly, mirroring Api::pbrt_shape / pbrt_light_source (api/src/lib.rs:751-811) --
    def add_material(self, **m):
        self.materials.append(m)
        return len(self.materials) - 1

    @staticmethod
    def _mesh_flags(reverse_orientation, swaps_handedness, alpha, shadowalpha, uv, normals, tangents):
        return ((1 if bool(reverse_orientation) != bool(swaps_handedness) else 0) | (2 if alpha == 0.0 else 0) | (4 if shadowalpha == 0.0 else 0) |
                (8 if reverse_orientation else 0) | (16 if uv is not None else 0) | (32 if normals is not None else 0) | (64 if tangents is not None else 0))

    @staticmethod
    def _attr(a, n, width, old, n_old):
        """Appends an optional per-triangle attribute block (zeros where a mesh has none)."""
        if a is None and old is None:
            return None
        base = old if old is not None else np.zeros((n_old, width), dtype=F32)
        blk = np.zeros((n, width), dtype=F32) if a is None else np.ascontiguousarray(a, dtype=F32).reshape(n, width)
        return np.concatenate([base, blk])

    def add_float_texture(self, type, **params):
        """Texture "name" "float" "<type>" (api/src/lib.rs make_float_texture) for use as a mesh's alpha / shadowalpha:
        type "constant" (value), "checkerboard" (tex1, tex2, uscale, vscale, udelta, vdelta), "dots" (inside, outside - the
        scene-file parameter names -, uscale, ...), "imagemap" (texels = level-0 (h, w) float array, wrap).  Returns its index."""
        assert type in ("constant", "checkerboard", "dots", "imagemap")
        self.float_textures.append(dict(type=type, **params))
        return len(self.float_textures) - 1

    def add_spectrum_texture(self, type, **params):
        """Texture "name" "spectrum" "<type>" for use as a matte / plastic material's Kd (``Kd=("texture", index)``): type
        "constant" (value) or "checkerboard" (tex1, tex2 - RGB constants -, uscale, vscale, udelta, vdelta, aamode
        "closedform" | "none"; textures/src/checkerboard_2d.rs).  Returns its index."""
        assert type in ("constant", "checkerboard")
        self.spectrum_textures.append(dict(type=type, **params))
        return len(self.spectrum_textures) - 1

    def _alpha_pair(self, alpha, shadowalpha):
        """(constant alpha, constant shadowalpha, texture index pair): a texture index is given as ("texture", k)."""
        pair = [-1, -1]
        consts = [1.0, 1.0]
        for c, a in enumerate((alpha, shadowalpha)):
            if isinstance(a, tuple) and a[0] == "texture":
                t = self.float_textures[a[1]]
                if t["type"] == "constant":
                    consts[c] = float(t.get("value", 1.0))
                else:
                    pair[c] = int(a[1])
            else:
                consts[c] = float(a)
        return consts[0], consts[1], pair

    def add_mesh(self, tri_verts, material, area_light=None, reverse_orientation=False, alpha=1.0, shadowalpha=1.0, uv=None, normals=None,
                 tangents=None, swaps_handedness=False):
        """One GeometricPrimitive per triangle; with ``area_light={'L': (r,g,b), 'twosided': False}`` one
        DiffuseAreaLight per triangle (api/src/lib.rs:783-803).  Optional de-indexed vertex attributes in world space:
        ``uv`` (n, 6), ``normals`` (n, 9), ``tangents`` (n, 9) — the mesh's "uv"/"st", "N", "S" (triangle.rs:384-394, 631-721)."""
        tv = np.ascontiguousarray(tri_verts, dtype=F32).reshape(-1, 9)
        n0 = self.tri_verts.shape[0]
        n = tv.shape[0]
        self.tri_uvs = self._attr(uv, n, 6, self.tri_uvs, n0)
        self.tri_normals = self._attr(normals, n, 9, self.tri_normals, n0)
        self.tri_tangents = self._attr(tangents, n, 9, self.tri_tangents, n0)
        self.tri_verts = np.concatenate([self.tri_verts, tv])
        self.prim_material = np.concatenate([self.prim_material, np.full(n, material, dtype=np.int32)])
        alpha, shadowalpha, tex_pair = self._alpha_pair(alpha, shadowalpha)
        flags = self._mesh_flags(reverse_orientation, swaps_handedness, alpha, shadowalpha, uv, normals, tangents) | (128 if max(tex_pair) >= 0 else 0)
        self.prim_flags = np.concatenate([self.prim_flags, np.full(n, flags, dtype=np.uint32)])
        self.prim_alpha_tex = np.concatenate([self.prim_alpha_tex, np.tile(np.array(tex_pair, dtype=np.int32), (n, 1))])
        pl = np.full(n, -1, dtype=np.int32)
        if area_light is not None:
            for i in range(n):
                pl[i] = len(self.lights)
                self.lights.append(dict(type="diffuse", L=tuple(area_light.get("L", (1, 1, 1))), prim=n0 + i,
                                        twosided=bool(area_light.get("twosided", False))))
        self.prim_light = np.concatenate([self.prim_light, pl])
        return n0

    def add_point_light(self, pos, I):
        self.lights.append(dict(type="point", pos=tuple(pos), L=tuple(I)))

    def add_spot_light(self, I, p_from, p_to, coneangle=30.0, conedeltaangle=5.0):
        """LightSource "spot" (lights/src/spot.rs) under an identity CTM: intensity ``I`` from ``p_from`` towards ``p_to``,
        full intensity inside coneangle - conedeltaangle degrees, falling to zero at coneangle.  world_to_light's vector part
        is the dir_to_z rotation (rows du, dv, dir of coordinate_system(dir), spot.rs:213-224); the cosines are libm's cosf
        of f32 radians, as the reference's f32::cos."""
        f, t = np.asarray(p_from, dtype=F32), np.asarray(p_to, dtype=F32)
        d = (t - f).astype(F32)
        d = (d * (F32(1.0) / np.sqrt(F32(F32(d[0] * d[0]) + F32(d[1] * d[1])) + F32(d[2] * d[2]), dtype=F32))).astype(F32)
        if abs(d[0]) > abs(d[1]):
            k = F32(1.0) / np.sqrt(F32(d[0] * d[0]) + F32(d[2] * d[2]), dtype=F32)
            du = np.array([-d[2] * k, F32(0.0) * k, d[0] * k], dtype=F32)
        else:
            k = F32(1.0) / np.sqrt(F32(d[1] * d[1]) + F32(d[2] * d[2]), dtype=F32)
            du = np.array([F32(0.0) * k, d[2] * k, -d[1] * k], dtype=F32)
        dv = np.array([F32(d[1] * du[2]) - F32(d[2] * du[1]), F32(d[2] * du[0]) - F32(d[0] * du[2]), F32(d[0] * du[1]) - F32(d[1] * du[0])], dtype=F32)
        w2l = np.eye(4, dtype=F32)
        w2l[0, :3], w2l[1, :3], w2l[2, :3] = du, dv, d
        for r in range(3):  # dir_to_z * Translate(-from) (Transform::mul keeps the product of the stored inverses)
            w2l[r, 3] = F32(F32(F32(F32(w2l[r, 0] * -f[0]) + F32(w2l[r, 1] * -f[1])) + F32(w2l[r, 2] * -f[2])) + F32(0.0))
        l2w = np.eye(4, dtype=F32)
        l2w[:3, :3] = w2l[:3, :3].T
        l2w[:3, 3] = f
        cosf = C.CDLL("libm.so.6").cosf
        cosf.restype, cosf.argtypes = C.c_float, [C.c_float]
        rad = F32(3.14159265358979323846) / F32(180.0)
        self.lights.append(dict(type="spot", L=tuple(I), pos=tuple(float(c) for c in f), light_to_world=l2w.reshape(-1), world_to_light=w2l.reshape(-1),
                                cos_total_width=cosf(float(F32(coneangle) * rad)), cos_falloff_start=cosf(float(F32(F32(coneangle) - F32(conedeltaangle)) * rad))))

    def add_goniometric_light(self, I, image=None, light_to_world=None):
        """LightSource "goniometric" (lights/src/goniometric.rs): a point light at light_to_world(0) whose intensity in a
        direction is scaled by ``image`` ((h, w, 3) floats, top row first; None = 1) at the direction's spherical coordinates."""
        m = np.eye(4, dtype=F32) if light_to_world is None else np.asarray(light_to_world, dtype=F32).reshape(4, 4)
        img = None if image is None else np.ascontiguousarray(image, dtype=F32)
        self.lights.append(dict(type="goniometric", L=tuple(I), pos=tuple(float(c) for c in m[:3, 3]), light_to_world=m.reshape(-1),
                                world_to_light=_m4_inverse(m).reshape(-1), image=img))

    def add_projection_light(self, I, image=None, light_to_world=None, fov=45.0):
        """LightSource "projection" (lights/src/projection.rs): a point light at light_to_world(0) that projects ``image``
        ((h, w, 3) floats, top row first; None = white) along the light's +z with the given field of view."""
        m = np.eye(4, dtype=F32) if light_to_world is None else np.asarray(light_to_world, dtype=F32).reshape(4, 4)
        img = None if image is None else np.ascontiguousarray(image, dtype=F32)
        aspect = F32(1.0) if img is None else F32(img.shape[1]) / F32(img.shape[0])
        cx, cy = (aspect, F32(1.0)) if aspect > 1 else (F32(1.0), F32(1.0) / aspect)
        # w_corner = normalize(screen_to_light(p_max)), projection.rs:92-95, with light_projection = Transform::perspective(fov, 1e-3, 1e30)
        n, f = F32(1e-3), F32(1e30)
        persp = np.eye(4, dtype=F32)
        persp[2, 2] = f / (f - n); persp[2, 3] = -f * n / (f - n); persp[3, 2] = F32(1); persp[3, 3] = F32(0)
        inv_tan = F32(1) / F32(math.tan(float(F32(fov) * F32(math.pi / 180.0)) / 2.0))
        s2l = _m4_mul(_m4_inverse(persp), np.diag([F32(1) / inv_tan, F32(1) / inv_tan, F32(1), F32(1)]).astype(F32))
        q = s2l @ np.array([cx, cy, 0.0, 1.0], dtype=F32)
        wc = (q[:3] / q[3]).astype(F32)
        cos_total = float(wc[2] / np.sqrt(np.sum(wc * wc, dtype=F32), dtype=F32))
        self.lights.append(dict(type="projection", L=tuple(I), pos=tuple(float(c) for c in m[:3, 3]), light_to_world=m.reshape(-1),
                                world_to_light=_m4_inverse(m).reshape(-1), image=img, fov=float(fov), cos_total_width=cos_total))

    def add_distant_light(self, L, w_light):
        """LightSource "distant" (lights/src/distant.rs): radiance ``L`` arriving from direction ``w_light`` (towards the
        light, world space; normalised here the way Vector3::normalize does: v * (1 / |v|))."""
        w = np.asarray(w_light, dtype=F32)
        inv = F32(1.0) / np.sqrt(F32(w[0] * w[0] + w[1] * w[1]) + F32(w[2] * w[2]), dtype=F32)
        self.lights.append(dict(type="distant", L=tuple(L), pos=tuple(float(F32(inv * c)) for c in w)))

    def add_infinite_light(self, L, image=None, light_to_world=None):
        """LightSource "infinite": ``L`` (times "scale"), optional environment ``image`` = decoded "mapname" as an
        (h, w, 3) float32 array (top row first, lat-long), optional 4x4 ``light_to_world``."""
        l = dict(type="infinite", L=tuple(L))
        if image is not None:
            l["image"] = np.ascontiguousarray(image, dtype=F32)
            assert l["image"].ndim == 3 and l["image"].shape[2] == 3
        if light_to_world is not None:
            m = np.asarray(light_to_world, dtype=F32).reshape(4, 4)
            l["light_to_world"], l["world_to_light"] = m.reshape(-1), _m4_inverse(m).reshape(-1)
        self.lights.append(l)

    # -- flattening --
    # -- instancing: ObjectBegin/ObjectEnd + ObjectInstance (api/src/lib.rs:880-987) --
    def add_object(self, tri_verts, material, reverse_orientation=False, uv=None, normals=None, tangents=None, alpha=1.0, shadowalpha=1.0):
        """Defines a named object (its triangles get their own BVHAccel); returns the object id."""
        tv = np.ascontiguousarray(tri_verts, dtype=F32).reshape(-1, 9)
        alpha, shadowalpha, tex_pair = self._alpha_pair(alpha, shadowalpha)
        self.objects.append(dict(tri_verts=tv, material=material,
                                 flags=self._mesh_flags(reverse_orientation, False, alpha, shadowalpha, uv, normals, tangents) | (128 if max(tex_pair) >= 0 else 0),
                                 nodes=None, ordered=None, uv=uv, normals=normals, tangents=tangents, alpha_tex=tex_pair))
        return len(self.objects) - 1

    def add_instance(self, obj, instance_to_world):
        """ObjectInstance: a TransformedPrimitive with a static, affine instance-to-world matrix (4x4, row-major)."""
        m = np.asarray(instance_to_world, dtype=F32).reshape(4, 4)
        self.instances.append((obj, m, _m4_inverse(m)))

    def build_accel(self, builder):
        """builder(prim_bounds, max_prims) -> (nodes, ordered).  Uses the product's host SAH builder by default.
        With objects, every object gets its own BVH and the scene aggregate is built over the top-level triangles
        followed by one primitive per instance (TransformedPrimitive::world_bound = transform_bounds of the object's root)."""
        from . import build_bvh_sah, triangle_bounds  # noqa
        builder = builder or build_bvh_sah
        mp = self.accel_params["maxnodeprims"]
        bounds = triangle_bounds(self.tri_verts)
        for o in self.objects:
            o["nodes"], o["ordered"] = builder(triangle_bounds(o["tri_verts"]), mp)
        ib = np.zeros((len(self.instances), 6), dtype=F32)
        for k, (obj, m, _) in enumerate(self.instances):
            b = self.objects[obj]["nodes"][0]["bounds"]
            corners = [(b[0], b[1], b[2]), (b[3], b[1], b[2]), (b[0], b[4], b[2]), (b[0], b[1], b[5]), (b[0], b[4], b[5]), (b[3], b[4], b[2]),
                       (b[3], b[1], b[5]), (b[3], b[4], b[5])]  # Transform::transform_bounds order, transform.rs:552-561
            pts = np.array([[F32(F32(F32(m[r, 0] * c[0]) + F32(m[r, 1] * c[1])) + F32(m[r, 2] * c[2])) + m[r, 3] for r in range(3)] for c in corners], dtype=F32)
            ib[k, :3], ib[k, 3:] = pts.min(0), pts.max(0)
        self.nodes, self.ordered_prims = builder(np.concatenate([bounds, ib]) if len(ib) else bounds, mp)

    def sample_bounds(self):
        xres, yres = self.film["xresolution"], self.film["yresolution"]
        _, (rx, ry) = filter_table(self.film["filter"], self.film.get("radius"))
        return (int(math.floor(0 + 0.5 - rx)), int(math.floor(0 + 0.5 - ry)), int(math.ceil(xres - 0.5 + rx)), int(math.ceil(yres - 0.5 + ry)))

    def to_desc(self):
        from . import Camera, Film, Integrator, Light, Material, Sampler, SceneDesc
        from . import (DIRECT_ALL, DIRECT_ONE, INTEGRATOR_DIRECT, INTEGRATOR_PATH, INTEGRATOR_WHITTED, LIGHT_AREA, LIGHT_DISTANT, LIGHT_INFINITE, LIGHT_POINT, LIGHT_SPOT, LIGHTS_POWER, LIGHTS_SPATIAL, LIGHTS_UNIFORM, MAT_GLASS, MAT_MATTE, MAT_METAL,
                       MAT_PLASTIC, SAMPLER_HALTON, SAMPLER_SOBOL, SAMPLER_ZEROTWO)
        if self.nodes is None:
            self.build_accel(None)
        d = SceneDesc()
        keep = self._keep = []

        def arr(a, dt):
            a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
            return a.ctypes.data_as(C.c_void_p)

        from . import Instance, Object
        tv, pf, pm, pl = [self.tri_verts], [self.prim_flags], [self.prim_material], [self.prim_light]
        n_top = self.tri_verts.shape[0]
        objs = (Object * max(1, len(self.objects)))()
        first = n_top
        for k, o in enumerate(self.objects):
            n = o["tri_verts"].shape[0]
            tv.append(o["tri_verts"])
            pf.append(np.full(n, o["flags"], dtype=np.uint32))
            pm.append(np.full(n, o["material"], dtype=np.int32))
            pl.append(np.full(n, -1, dtype=np.int32))  # no area lights inside object instances (as in pbrt)
            objs[k].nodes, objs[k].n_nodes = arr(o["nodes"], o["nodes"].dtype), len(o["nodes"])
            objs[k].ordered_prims = arr(o["ordered"], np.uint32)
            objs[k].first_prim, objs[k].n_prims = first, n
            first += n
        insts = (Instance * max(1, len(self.instances)))()
        for k, (obj, m, minv) in enumerate(self.instances):
            insts[k].object = obj
            insts[k].instance_to_world[:] = m.reshape(-1)
            insts[k].world_to_instance[:] = minv.reshape(-1)
        keep.extend([objs, insts])
        d.nodes, d.n_nodes = arr(self.nodes, self.nodes.dtype), len(self.nodes)
        d.ordered_prims = arr(self.ordered_prims, np.uint32)
        all_tv = np.concatenate(tv)
        d.tri_verts = arr(all_tv, F32)
        d.prim_flags = arr(np.concatenate(pf), np.uint32)
        d.prim_material = arr(np.concatenate(pm), np.int32)
        d.prim_light = arr(np.concatenate(pl), np.int32)
        d.n_prims = all_tv.shape[0]
        d.n_top_tris = n_top
        # optional vertex attributes: top-level blocks first, then one block per object (zeros where a mesh has none)
        for field, width, top, key in (("tri_uvs", 6, self.tri_uvs, "uv"), ("tri_normals", 9, self.tri_normals, "normals"),
                                       ("tri_tangents", 9, self.tri_tangents, "tangents")):
            if top is None and all(o.get(key) is None for o in self.objects):
                continue
            blocks = [top if top is not None else np.zeros((n_top, width), dtype=F32)]
            for o in self.objects:
                n = o["tri_verts"].shape[0]
                blocks.append(np.zeros((n, width), dtype=F32) if o.get(key) is None else np.ascontiguousarray(o[key], dtype=F32).reshape(n, width))
            setattr(d, field, arr(np.concatenate(blocks), F32))
        if self.float_textures:
            from . import float_texture_array, noise_perm
            blocks = [self.prim_alpha_tex.reshape(-1, 2)]
            for o in self.objects:
                blocks.append(np.tile(np.array(o.get("alpha_tex", (-1, -1)), dtype=np.int32), (o["tri_verts"].shape[0], 1)))
            d.prim_alpha_tex = arr(np.concatenate(blocks), np.int32)
            d.float_textures = C.cast(float_texture_array(self.float_textures, keep), C.c_void_p)
            d.n_float_textures = len(self.float_textures)
            d.noise_perm = arr(noise_perm(), np.uint8)
        d.objects, d.n_objects = C.cast(objs, C.c_void_p), len(self.objects)
        d.instances, d.n_instances = C.cast(insts, C.c_void_p), len(self.instances)

        mats = (Material * max(1, len(self.materials)))()
        kd_tex = np.full(max(1, len(self.materials)), -1, dtype=np.int32)
        for i, m in enumerate(self.materials):
            t = m["type"]
            M = mats[i]
            if isinstance(m.get("Kd"), tuple) and len(m["Kd"]) == 2 and m["Kd"][0] == "texture":
                kd_tex[i] = int(m["Kd"][1])
                m = dict(m, Kd=(0.5, 0.5, 0.5) if t == "matte" else (0.25, 0.25, 0.25))
            M.remap_roughness = 1 if m.get("remaproughness", True) else 0
            if t == "matte":     # materials/src/matte.rs:77-84
                M.type = MAT_MATTE
                M.kd[:] = m.get("Kd", (0.5, 0.5, 0.5))
                M.sigma = m.get("sigma", 0.0)
            elif t == "plastic":  # plastic.rs:104-113
                M.type = MAT_PLASTIC
                M.kd[:] = m.get("Kd", (0.25, 0.25, 0.25))
                M.ks[:] = m.get("Ks", (0.25, 0.25, 0.25))
                M.urough = M.vrough = m.get("roughness", 0.1)
            elif t == "glass":   # glass.rs:130-146
                M.type = MAT_GLASS
                M.ks[:] = m.get("Kr", (1, 1, 1))
                M.kt[:] = m.get("Kt", (1, 1, 1))
                M.eta[0] = m.get("eta", m.get("index", 1.5))
                M.urough, M.vrough = m.get("uroughness", 0.0), m.get("vroughness", 0.0)
            elif t == "mirror":  # mirror.rs:62-70
                M.type = 4
                M.ks[:] = m.get("Kr", (0.9, 0.9, 0.9))
            elif t == "metal":   # metal.rs:109-133; the default copper SPD->RGB conversion is done by the caller
                M.type = MAT_METAL
                M.eta[:] = m.get("eta", (0.19999069, 0.92208463, 1.09987593))
                M.k[:] = m.get("k", (3.90463543, 2.44763327, 2.13765264))
                r = m.get("roughness", 0.01)
                M.urough, M.vrough = m.get("uroughness", r), m.get("vroughness", r)
            else:
                raise ValueError("material %r is outside this path" % t)
        keep.append(mats)
        d.materials, d.n_materials = C.cast(mats, C.c_void_p), len(self.materials)
        if self.spectrum_textures and (kd_tex >= 0).any():
            from . import spectrum_texture_array
            d.spectrum_textures = C.cast(spectrum_texture_array(self.spectrum_textures, keep), C.c_void_p)
            d.n_spectrum_textures = len(self.spectrum_textures)
            d.material_kd_tex = arr(kd_tex, np.int32)

        lights = (Light * max(1, len(self.lights)))()
        ident = np.eye(4, dtype=F32).reshape(-1)
        for i, l in enumerate(self.lights):
            Lt = lights[i]
            Lt.prim = -1
            Lt.L[:] = l["L"]
            Lt.light_to_world[:] = l.get("light_to_world", ident)
            Lt.world_to_light[:] = l.get("world_to_light", ident)
            if l["type"] == "point":
                Lt.type = LIGHT_POINT
                Lt.pos[:] = l["pos"]
            elif l["type"] == "distant":
                Lt.type = LIGHT_DISTANT
                Lt.pos[:] = l["pos"]
            elif l["type"] in ("goniometric", "projection"):
                Lt.type = 5 if l["type"] == "goniometric" else 6
                Lt.pos[:] = l["pos"]
                if l["type"] == "projection":
                    Lt.fov, Lt.cos_total_width = l["fov"], l["cos_total_width"]
                if l.get("image") is not None:
                    keep.append(l["image"])
                    Lt.map_rgb = l["image"].ctypes.data_as(C.c_void_p)
                    Lt.map_height, Lt.map_width = l["image"].shape[:2]
            elif l["type"] == "spot":
                Lt.type = LIGHT_SPOT
                Lt.pos[:] = l["pos"]
                Lt.cos_total_width, Lt.cos_falloff_start = l["cos_total_width"], l["cos_falloff_start"]
            elif l["type"] == "diffuse":
                Lt.type = LIGHT_AREA
                Lt.prim = l["prim"]
                Lt.two_sided = 1 if l.get("twosided") else 0
            elif l["type"] == "infinite":
                Lt.type = LIGHT_INFINITE
                if l.get("image") is not None:
                    keep.append(l["image"])
                    Lt.map_rgb = l["image"].ctypes.data_as(C.c_void_p)
                    Lt.map_height, Lt.map_width = l["image"].shape[:2]
            else:
                raise ValueError("light %r is outside this path" % l["type"])
        keep.append(lights)
        d.lights, d.n_lights = C.cast(lights, C.c_void_p), len(self.lights)

        cam = self.camera
        xres, yres = self.film["xresolution"], self.film["yresolution"]
        d.camera.camera_to_world[:] = look_at_camera_to_world(cam["eye"], cam["look"], cam["up"]).reshape(-1)
        ctype = cam.get("type", "perspective")
        d.camera.type = {"perspective": 0, "orthographic": 1, "environment": 2}[ctype]
        d.camera.raster_to_camera[:] = perspective_raster_to_camera(cam["fov"], xres, yres, cam.get("screenwindow"), orthographic=ctype == "orthographic").reshape(-1)
        d.camera.lens_radius, d.camera.focal_distance = (0.0 if ctype == "environment" else cam["lensradius"]), cam["focaldistance"]
        d.camera.shutter_open, d.camera.shutter_close = cam["shutteropen"], cam["shutterclose"]

        tab, (rx, ry) = filter_table(self.film["filter"], self.film.get("radius"))
        d.film.xres, d.film.yres = xres, yres
        crop = self.film.get("cropwindow", (0.0, 1.0, 0.0, 1.0))  # film/mod.rs:101-110
        d.film.crop[:] = [int(math.ceil(xres * crop[0])), int(math.ceil(yres * crop[2])), int(math.ceil(xres * crop[1])), int(math.ceil(yres * crop[3]))]
        d.film.filter_radius[:] = [rx, ry]
        d.film.filter_table[:] = tab
        d.film.scale = self.film.get("scale", 1.0)
        d.film.max_sample_luminance = self.film.get("maxsampleluminance", float("inf"))

        stype = self.sampler["type"]
        if stype not in ("halton", "02sequence", "lowdiscrepancy", "sobol"):
            raise ValueError("Sampler %r is outside this path (halton, 02sequence, sobol)" % stype)
        d.sampler.type = {"halton": SAMPLER_HALTON, "sobol": SAMPLER_SOBOL}.get(stype, SAMPLER_ZEROTWO)
        if stype == "sobol":
            from . import sobol_matrices_32
            m32 = sobol_matrices_32()
            keep.append(m32)
            d.sobol_matrices_32 = m32.ctypes.data_as(C.c_void_p)
        d.sampler.spp = self.sampler["pixelsamples"]
        d.sampler.sample_at_center = 1 if self.sampler.get("samplepixelcenter") else 0
        d.sampler.dimensions = self.sampler.get("dimensions", 4)

        d.integrator.max_depth = self.integrator["maxdepth"]
        d.integrator.rr_threshold = self.integrator["rrthreshold"]
        sb = [int(math.floor(d.film.crop[0] + 0.5 - rx)), int(math.floor(d.film.crop[1] + 0.5 - ry)),
              int(math.ceil(d.film.crop[2] - 0.5 + rx)), int(math.ceil(d.film.crop[3] - 0.5 + ry))]
        pb = self.integrator.get("pixelbounds")
        if pb:  # path.rs:296-311
            sb = [max(sb[0], pb[0]), max(sb[1], pb[1]), min(sb[2], pb[2]), min(sb[3], pb[3])]
        d.integrator.pixel_bounds[:] = sb
        strat = self.integrator.get("lightsamplestrategy", "uniform")
        if strat not in ("uniform", "power", "spatial"):  # path.rs:314-324 falls back to spatial for unknown names
            raise ValueError("lightsamplestrategy %r (uniform, power, spatial)" % strat)
        d.integrator.light_strategy = {"uniform": LIGHTS_UNIFORM, "power": LIGHTS_POWER, "spatial": LIGHTS_SPATIAL}[strat]
        name = self.integrator.get("name", "path")
        if name not in ("path", "whitted", "directlighting"):
            raise ValueError("Integrator %r is outside this path (path, whitted, directlighting)" % name)
        d.integrator.type = {"path": INTEGRATOR_PATH, "whitted": INTEGRATOR_WHITTED, "directlighting": INTEGRATOR_DIRECT}[name]
        d.integrator.direct_strategy = DIRECT_ONE if self.integrator.get("strategy", "all") == "one" else DIRECT_ALL  # direct_lighting.rs:161-169
        return d


class PathIntegrator:
    """integrators/src/path.rs PathIntegrator over the CUDA wavefront path tracer.

    ``PathIntegrator.from_params(params, sampler_params, camera/film...)`` is folded into the
    SceneDescription; ``preprocess`` uploads the scene (Scene::new + Integrator::preprocess),
    ``render`` runs Integrator::render and returns the RGB image (what Film::write_image stores).
    """

    def __init__(self, scene_description):
        self.sd = scene_description
        self._h = None

    def preprocess(self):
        from . import _check, init, lib, _inited
        init(_inited if _inited is not None else 0)
        d = self.sd.to_desc()
        self._desc = d
        h = C.c_void_p()
        _check(lib().b200pt_scene_create(C.byref(d), C.byref(h)), "b200pt_scene_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            try:
                from . import lib
                lib().b200pt_scene_destroy(self._h)
            except Exception:  # interpreter shutdown: the import machinery may already be gone
                pass
            self._h = None

    __del__ = close

    @property
    def handle(self):
        return self._h

    def film_shape(self):
        c = self._desc.film.crop
        return (c[3] - c[1], c[2] - c[0])

    def filter_radius(self):
        return tuple(self._desc.film.filter_radius)

    def render_rows(self, row_begin=0, row_end=None):
        """Renders pixel rows [row_begin,row_end) of the cropped window; returns (H, W, 4) XYZ+weight."""
        from . import _check, _ptr, lib
        if self._h is None:
            self.preprocess()
        h, w = self.film_shape()
        row_end = h if row_end is None else row_end
        film = np.zeros((h, w, 4), dtype=F32)
        _check(lib().b200pt_render_rows(self._h, row_begin, row_end, _ptr(film)), "b200pt_render_rows")
        return film

    def render_shard_device(self, shard, n_shards, d_film_ptr, band_rows=8, stream=0):
        """Renders shard `shard` of `n_shards` (interleaved row bands) into a device film (H*W*4 float32, full window)."""
        from . import _check, lib
        if self._h is None:
            self.preprocess()
        _check(lib().b200pt_render_shard_device(self._h, shard, n_shards, band_rows, d_film_ptr, stream), "b200pt_render_shard_device")

    def render_shard_device_raw(self, shard, n_shards, d_film_ptr, band_rows=8, stream=0):
        """The shard as running sums {filter-weighted RGB, weight}: combine shard films in this space, then film_finish_device."""
        from . import _check, lib
        if self._h is None:
            self.preprocess()
        _check(lib().b200pt_render_shard_device_raw(self._h, shard, n_shards, band_rows, d_film_ptr, stream), "b200pt_render_shard_device_raw")

    @staticmethod
    def film_finish_device(d_film_ptr, n_pix, stream=0):
        """Film::merge_film_tile's RGB -> XYZ on an assembled film of running sums, in place."""
        from . import _check, lib
        _check(lib().b200pt_film_finish_device(d_film_ptr, n_pix, d_film_ptr, stream), "b200pt_film_finish_device")

    def render_rows_device(self, row_begin, row_end, d_film_ptr, stream=0):
        from . import _check, lib
        if self._h is None:
            self.preprocess()
        _check(lib().b200pt_render_rows_device(self._h, row_begin, row_end, d_film_ptr, stream), "b200pt_render_rows_device")

    def resolve(self, film_xyzw):
        from . import _check, _ptr, lib
        h, w = film_xyzw.shape[:2]
        rgb = np.empty((h, w, 3), dtype=F32)
        f = np.ascontiguousarray(film_xyzw, dtype=F32)
        _check(lib().b200pt_film_resolve(C.byref(self._desc.film), _ptr(f), _ptr(rgb)), "b200pt_film_resolve")
        return rgb

    def render(self):
        """Integrator::render: the whole image as (H, W, 3) RGB float32."""
        return self.resolve(self.render_rows())

    def li(self, pixel_sample):
        """Integrator::li for explicit (x, y, sample) triples -> (n,3) radiance, (n,) camera rays."""
        from . import RAY_DTYPE, _check, _ptr, lib
        if self._h is None:
            self.preprocess()
        ps = np.ascontiguousarray(pixel_sample, dtype=np.int32).reshape(-1, 3)
        out = np.empty((ps.shape[0], 3), dtype=F32)
        rays = np.empty(ps.shape[0], dtype=RAY_DTYPE)
        _check(lib().b200pt_li_batch(self._h, _ptr(ps), ps.shape[0], _ptr(out), _ptr(rays)), "b200pt_li_batch")
        return out, rays

    def ray_counts(self):
        from . import _check, _ptr, lib
        c = np.zeros(3, dtype=np.uint64)
        _check(lib().b200pt_scene_ray_counts(self._h, _ptr(c)), "b200pt_scene_ray_counts")
        return c

    def set_memory_budget(self, n_bytes):
        """Bytes of device memory the wave state of a render may take (0 = default); results do not depend on it."""
        from . import _check, lib
        if self._h is None:
            self.preprocess()
        _check(lib().b200pt_scene_set_memory_budget(self._h, int(n_bytes)), "b200pt_scene_set_memory_budget")


class MultiGPURender:
    """Integrator::render on several GPUs of this process (b200pt_multi_*): the scene replicated per device, interleaved
    row bands, the bands gathered on devices[0] over NVLink (NCCL send / recv, or one ncclReduce for wide filters)."""

    def __init__(self, scene_description, devices):
        from . import _check, lib
        self.sd = scene_description
        self._desc = scene_description.to_desc()
        self.devices = np.ascontiguousarray(devices, dtype=np.int32)
        h = C.c_void_p()
        _check(lib().b200pt_multi_create(C.byref(self._desc), self.devices.ctypes.data_as(C.c_void_p), len(self.devices), C.byref(h)), "b200pt_multi_create")
        self._h = h

    def film_shape(self):
        c = self._desc.film.crop
        return (c[3] - c[1], c[2] - c[0])

    def render_rows(self, band_rows=8, download=True):
        """-> (H, W, 4) XYZ + weight film (host), or None with download=False (the film stays on devices[0])."""
        from . import _check, _ptr, lib
        h, w = self.film_shape()
        film = np.zeros((h, w, 4), dtype=F32) if download else None
        _check(lib().b200pt_multi_render(self._h, band_rows, _ptr(film)), "b200pt_multi_render")
        return film

    def info(self):
        from . import _check, _ptr, lib
        rays = np.zeros(3, dtype=np.uint64)
        ms, nccl = C.c_double(0.0), C.c_int32(0)
        _check(lib().b200pt_multi_info(self._h, _ptr(rays), C.byref(ms), C.byref(nccl)), "b200pt_multi_info")
        return {"rays": rays, "gather_ms": ms.value, "nccl": bool(nccl.value)}

    def close(self):
        if getattr(self, "_h", None):
            try:
                from . import lib
                lib().b200pt_multi_destroy(self._h)
            except Exception:
                pass
            self._h = None

    __del__ = close
