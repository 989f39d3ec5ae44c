"""Synthetic workloads of BASELINE.json: procedural meshes, the
fixed ray sets of the ray-cast microbench (C2) and the path-tracing scenes
(C1, C3).  Pure numpy, deterministic, independent of the oracle: the CPU oracle
and the CUDA path both consume the bytes produced here.
"""
import numpy as np

from .scene import SceneDescription

F32 = np.float32
U32 = np.uint32
U64 = np.uint64


# ---- small deterministic generators -------------------------------------------------
def _hash_u32(x):
    """Integer hash (lowbias32) on uint32 arrays."""
    x = x.astype(U32)
    x ^= x >> U32(16)
    x *= U32(0x7FEB352D)
    x ^= x >> U32(15)
    x *= U32(0x846CA68B)
    x ^= x >> U32(16)
    return x


def _lattice(ix, iy, iz, seed):
    h = _hash_u32(ix.astype(U32) * U32(73856093) ^ iy.astype(U32) * U32(19349663) ^ iz.astype(U32) * U32(83492791) ^ U32(seed))
    return (h >> U32(8)).astype(F32) * F32(1.0 / 16777216.0)


def value_noise(p, freq, seed):
    """Trilinear value noise with smoothstep weights; p: (n,3) float32."""
    q = p.astype(F32) * F32(freq) + F32(1000.0)
    i = np.floor(q).astype(np.int64)
    f = (q - i.astype(F32)).astype(F32)
    w = f * f * (F32(3.0) - F32(2.0) * f)
    out = np.zeros(p.shape[0], dtype=F32)
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                c = _lattice(i[:, 0] + dx, i[:, 1] + dy, i[:, 2] + dz, seed)
                wx = w[:, 0] if dx else F32(1) - w[:, 0]
                wy = w[:, 1] if dy else F32(1) - w[:, 1]
                wz = w[:, 2] if dz else F32(1) - w[:, 2]
                out += c * wx * wy * wz
    return out


def displaced_sphere(nu, nv, radius=1.0, amplitude=0.15, base_freq=8.0, seed=1, center=(0.0, 0.0, 0.0), with_attrs=False):
    """nu x nv lat-long grid -> 2*nu*nv triangles (the pole rows yield zero-area
    triangles, which the reference rejects with det == 0), radially displaced by 4
    octaves of value noise.  Returns (n_tris, 9) float32; with_attrs=True also returns de-indexed
    "uv" (n_tris, 6) = (u / 2 pi, v / pi) and "N" (n_tris, 9) = 1.5 x the undisplaced radial direction
    (smooth shading normals, deliberately not unit length: the reference does not renormalise N)."""
    u = (np.arange(nu, dtype=F32) / F32(nu)) * F32(2.0 * np.pi)
    v = (np.arange(nv + 1, dtype=F32) / F32(nv)) * F32(np.pi)
    uu, vv = np.meshgrid(u, v)  # (nv+1, nu)
    d = np.stack([np.sin(vv) * np.cos(uu), np.cos(vv), np.sin(vv) * np.sin(uu)], axis=-1).astype(F32).reshape(-1, 3)
    # exact poles so every vertex of a pole row coincides
    d[:nu] = np.array([0, 1, 0], dtype=F32)
    d[nv * nu:] = np.array([0, -1, 0], dtype=F32)
    disp = np.zeros(d.shape[0], dtype=F32)
    amp, fr = F32(amplitude), F32(base_freq)
    for o in range(4):
        disp += amp * (value_noise(d, fr, seed + o) - F32(0.5))
        amp *= F32(0.5)
        fr *= F32(2.0)
    p = (d * (F32(radius) + disp)[:, None] + np.asarray(center, dtype=F32)).astype(F32).reshape(nv + 1, nu, 3)
    i0 = np.arange(nu)
    i1 = (i0 + 1) % nu
    tris = np.empty((nv, nu, 2, 9), dtype=F32)
    a, b, c, e = p[:-1][:, i0], p[:-1][:, i1], p[1:][:, i0], p[1:][:, i1]
    tris[:, :, 0, 0:3], tris[:, :, 0, 3:6], tris[:, :, 0, 6:9] = a, c, b
    tris[:, :, 1, 0:3], tris[:, :, 1, 3:6], tris[:, :, 1, 6:9] = b, c, e
    if not with_attrs:
        return tris.reshape(-1, 9)
    dn = (F32(1.5) * d).astype(F32).reshape(nv + 1, nu, 3)
    na, nb, nc, ne = dn[:-1][:, i0], dn[:-1][:, i1], dn[1:][:, i0], dn[1:][:, i1]
    nrm = np.empty((nv, nu, 2, 9), dtype=F32)
    nrm[:, :, 0, 0:3], nrm[:, :, 0, 3:6], nrm[:, :, 0, 6:9] = na, nc, nb
    nrm[:, :, 1, 0:3], nrm[:, :, 1, 3:6], nrm[:, :, 1, 6:9] = nb, nc, ne
    uu0 = (np.arange(nu, dtype=F32) / F32(nu))[None, :].repeat(nv, 0)
    uu1 = ((np.arange(nu, dtype=F32) + F32(1)) / F32(nu))[None, :].repeat(nv, 0)   # no wrap at the seam
    vv0 = (np.arange(nv, dtype=F32) / F32(nv))[:, None].repeat(nu, 1)
    vv1 = ((np.arange(nv, dtype=F32) + F32(1)) / F32(nv))[:, None].repeat(nu, 1)
    uv = np.empty((nv, nu, 2, 6), dtype=F32)
    # vertex order matches the triangles: (a, c, b) and (b, c, e)
    uv[:, :, 0, 0], uv[:, :, 0, 1], uv[:, :, 0, 2], uv[:, :, 0, 3], uv[:, :, 0, 4], uv[:, :, 0, 5] = uu0, vv0, uu0, vv1, uu1, vv0
    uv[:, :, 1, 0], uv[:, :, 1, 1], uv[:, :, 1, 2], uv[:, :, 1, 3], uv[:, :, 1, 4], uv[:, :, 1, 5] = uu1, vv0, uu0, vv1, uu1, vv1
    return tris.reshape(-1, 9), uv.reshape(-1, 6), nrm.reshape(-1, 9)


def triangle_soup(n, edge=0.02, seed=2):
    """n small triangles with centroids uniform in [-1,1]^3 (stress variant)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    c = rng.uniform(-1, 1, size=(n, 1, 3)).astype(F32)
    off = rng.uniform(-edge, edge, size=(n, 3, 3)).astype(F32)
    return (c + off).reshape(n, 9).astype(F32)


def ground_quad(y=-1.3, half=6.0):
    """Two-triangle ground plane (scenes/shapes/plymesh.pbrt:26-33 analogue)."""
    p = np.array([[-half, y, -half], [half, y, -half], [half, y, half], [-half, y, half]], dtype=F32)
    return np.stack([np.concatenate([p[0], p[2], p[1]]), np.concatenate([p[0], p[3], p[2]])]).astype(F32)


def radical_inverse_2_3(idx):
    """Halton (2,3) points for integer indices (float64 math, rounded to f32)."""
    idx = np.asarray(idx, dtype=np.int64)
    out = []
    for base in (2, 3):
        i = idx.copy()
        f = np.ones(idx.shape, dtype=np.float64)
        r = np.zeros(idx.shape, dtype=np.float64)
        while np.any(i > 0):
            f /= base
            r += f * (i % base)
            i //= base
        out.append(r)
    return np.stack(out, axis=-1).astype(F32)


def _halton_dim(idx, base):
    i = np.asarray(idx, dtype=np.int64).copy()
    f = np.ones(i.shape, dtype=np.float64)
    r = np.zeros(i.shape, dtype=np.float64)
    while np.any(i > 0):
        f /= base
        r += f * (i % base)
        i //= base
    return np.minimum(r, 1.0 - 2.0 ** -24).astype(F32)


# ---- C2: ray sets ------------------------------------------------------------------
def primary_rays(width, height, eye=(0.0, 0.0, -3.5), fov_deg=38.0):
    """Pinhole camera looking down +z over a width x height grid of (non-square) pixels spanning a
    square fov, one ray per pixel with Halton(2,3) jitter, t_max = inf."""
    from . import RAY_DTYPE
    n = width * height
    ids = np.arange(n, dtype=np.int64)
    jit = radical_inverse_2_3(ids)
    x = (ids % width).astype(F32) + jit[:, 0]
    y = (ids // width).astype(F32) + jit[:, 1]
    th = F32(np.tan(np.deg2rad(fov_deg) / 2.0))
    sx = (F32(2.0) * x / F32(width) - F32(1.0)) * th
    sy = (F32(1.0) - F32(2.0) * y / F32(height)) * th
    d = np.stack([sx, sy, np.ones(n, dtype=F32)], axis=-1).astype(F32)
    d /= np.sqrt((d * d).sum(-1, dtype=F32))[:, None]
    rays = np.zeros(n, dtype=RAY_DTYPE)
    rays["o"] = np.asarray(eye, dtype=F32)
    rays["d"] = d
    rays["tmax"] = np.inf
    return rays


def primary_rays_lookat(width, height, eye, look, up=(0.0, 1.0, 0.0), fov_deg=40.0):
    """primary_rays through a look-at camera (right-handed basis from eye / look / up; the fov spans the shorter image side)."""
    rays = primary_rays(width, height, eye=(0.0, 0.0, 0.0), fov_deg=fov_deg)
    eye, look, up = (np.asarray(v, dtype=np.float64) for v in (eye, look, up))
    f = look - eye
    f /= np.linalg.norm(f)
    r = np.cross(up, f)
    r /= np.linalg.norm(r)
    u = np.cross(f, r)
    d = rays["d"].astype(np.float64)
    dw = d[:, 0:1] * r + d[:, 1:2] * u + d[:, 2:3] * f
    dw /= np.linalg.norm(dw, axis=1, keepdims=True)
    rays["d"] = dw.astype(F32)
    rays["o"] = eye.astype(F32)
    return rays


def bounce_rays(tri_verts, primary, hits, n, shuffle_seed=3):
    """Diffuse-bounce rays: origin = hit point of a primary ray nudged off the surface, direction =
    cosine-weighted about the geometric normal (facing the incoming side), then shuffled so that
    neighbouring threads are incoherent.  ``hits`` are closest-hit records of ``primary``."""
    from . import MISS, RAY_DTYPE
    hit_ids = np.nonzero(hits["prim"] != MISS)[0]
    if hit_ids.size == 0:
        raise ValueError("no primary ray hit the mesh")
    src = hit_ids[np.arange(n, dtype=np.int64) % hit_ids.size]
    h = hits[src]
    tv = tri_verts.reshape(-1, 3, 3)[h["prim"].astype(np.int64)]
    b0, b1 = h["b0"][:, None], h["b1"][:, None]
    p = (b0 * tv[:, 0] + b1 * tv[:, 1] + (F32(1) - b0 - b1) * tv[:, 2]).astype(F32)
    ng = np.cross(tv[:, 1] - tv[:, 0], tv[:, 2] - tv[:, 0]).astype(F32)
    ng /= np.maximum(np.sqrt((ng * ng).sum(-1, dtype=F32)), F32(1e-30))[:, None]
    din = primary["d"][src]
    flip = (ng * din).sum(-1) > 0
    ng[flip] = -ng[flip]
    ids = np.arange(n, dtype=np.int64)
    u0, u1 = _halton_dim(ids, 11), _halton_dim(ids, 13)
    r = np.sqrt(u0, dtype=F32)
    phi = F32(2.0 * np.pi) * u1
    lx, ly = r * np.cos(phi), r * np.sin(phi)
    lz = np.sqrt(np.maximum(F32(0), F32(1) - lx * lx - ly * ly), dtype=F32)
    helper = np.where(np.abs(ng[:, 0:1]) > 0.9, np.array([[0, 1, 0]], dtype=F32), np.array([[1, 0, 0]], dtype=F32))
    t = np.cross(helper, ng).astype(F32)
    t /= np.sqrt((t * t).sum(-1, dtype=F32))[:, None]
    bt = np.cross(ng, t).astype(F32)
    d = (lx[:, None] * t + ly[:, None] * bt + lz[:, None] * ng).astype(F32)
    rays = np.zeros(n, dtype=RAY_DTYPE)
    rays["o"] = p + ng * F32(1e-4)
    rays["d"] = d
    rays["tmax"] = np.inf
    perm = np.random.Generator(np.random.PCG64(shuffle_seed)).permutation(n)
    return rays[perm]


def shadow_rays(closest_rays, radius=2.5):
    """Any-hit set: same origins, d = target - o (un-normalised) with targets uniform on the
    radius-2.5 sphere, t_max = 1 - 1e-4 (core/src/interaction/mod.rs:212-223 shadow-ray form)."""
    from . import RAY_DTYPE
    n = closest_rays.shape[0]
    ids = np.arange(n, dtype=np.int64)
    u0, u1 = _halton_dim(ids, 17), _halton_dim(ids, 19)
    z = F32(1) - F32(2) * u0
    r = np.sqrt(np.maximum(F32(0), F32(1) - z * z), dtype=F32)
    phi = F32(2.0 * np.pi) * u1
    tgt = (F32(radius) * np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=-1)).astype(F32)
    rays = np.zeros(n, dtype=RAY_DTYPE)
    rays["o"] = closest_rays["o"]
    rays["d"] = tgt - closest_rays["o"]
    rays["tmax"] = F32(1.0) - F32(1e-4)
    return rays


def sky_image(width, height, seed=7, sun=(0.3, 0.25), sun_radius=0.04, sun_radiance=400.0):
    """Procedural HDR lat-long environment (height, width, 3) float32: a blue-to-white sky gradient over a brown
    ground, value-noise clouds and a small, very bright sun at (u, v) = ``sun`` — a stand-in for the EXR maps of the
    pbrt-v3 scenes (scenes/materials/matte.pbrt:18-19)."""
    v = ((np.arange(height, dtype=F32) + F32(0.5)) / F32(height))[:, None].repeat(width, 1)
    u = ((np.arange(width, dtype=F32) + F32(0.5)) / F32(width))[None, :].repeat(height, 0)
    sky = np.stack([F32(0.35) + F32(0.5) * v, F32(0.55) + F32(0.35) * v, F32(1.0) - F32(0.1) * v], axis=-1).astype(F32)
    ground = np.array([0.25, 0.2, 0.15], dtype=F32)
    img = np.where((v < 0.5)[..., None], sky, ground[None, None, :]).astype(F32)
    d = np.stack([np.cos(u * F32(2 * np.pi)), v * F32(2.0), np.sin(u * F32(2 * np.pi))], axis=-1).reshape(-1, 3).astype(F32)
    clouds = value_noise(d, 3.0, seed).reshape(height, width)
    img += (F32(0.6) * np.maximum(clouds - F32(0.55), F32(0)) * (v < 0.5))[..., None]
    du = np.minimum(np.abs(u - F32(sun[0])), F32(1) - np.abs(u - F32(sun[0])))
    r2 = du * du + (v - F32(sun[1])) ** 2
    img += (F32(sun_radiance) * np.exp(-r2 / F32(sun_radius * sun_radius)))[..., None] * np.array([1.0, 0.9, 0.7], dtype=F32)
    return img.astype(F32)


C2_FULL = dict(nu=1000, nv=500, width=4096, height=2048)       # 1 000 000 tris, 2^23 primary + 2^23 bounce
C2_SMALL = dict(nu=100, nv=50, width=128, height=64)           # 10 000 tris, 2^13 + 2^13 (CPU-test size)


def c2_mesh(cfg=C2_FULL):
    return displaced_sphere(cfg["nu"], cfg["nv"])


# ---- path-tracing scenes -------------------------------------------------------------
COPPER_ETA = (0.19999069, 0.92208463, 1.09987593)
COPPER_K = (3.90463543, 2.44763327, 2.13765264)


def scene_c1(nu=330, nv=330, res=400, spp=16, maxdepth=5):
    """C1: plymesh-variant — displaced sphere (2*nu*nv tris) + ground quad, matte, constant
    infinite light L=[1.2 1.2 1.1], PathIntegrator maxdepth 5, Halton, box filter."""
    sd = SceneDescription()
    m = sd.add_material(type="matte", Kd=(0.5, 0.5, 0.5))
    g = sd.add_material(type="matte", Kd=(0.4, 0.4, 0.4))
    sd.add_mesh(displaced_sphere(nu, nv), m)
    sd.add_mesh(ground_quad(), g)
    sd.add_infinite_light((1.2, 1.2, 1.1))
    sd.camera.update(eye=(0.0, 1.2, -4.0), look=(0.0, -0.1, 0.0), up=(0, 1, 0), fov=40.0)
    sd.film.update(xresolution=res, yresolution=res, filter="box")
    sd.sampler.update(type="halton", pixelsamples=spp)
    sd.integrator.update(maxdepth=maxdepth, lightsamplestrategy="uniform")
    return sd


def scene_c3(nu=330, nv=330, xres=1920, yres=1080, spp=64, maxdepth=8):
    """C3: four displaced spheres (matte / plastic / glass / metal), ground quad, a two-triangle
    diffuse area light and a point light, lightsamplestrategy "power"."""
    sd = SceneDescription()
    mats = [sd.add_material(type="matte", Kd=(0.5, 0.45, 0.4)), sd.add_material(type="plastic"),
            sd.add_material(type="glass", eta=1.5), sd.add_material(type="metal", eta=COPPER_ETA, k=COPPER_K, roughness=0.01)]
    g = sd.add_material(type="matte", Kd=(0.4, 0.4, 0.4))
    for i, mid in enumerate(mats):
        sd.add_mesh(displaced_sphere(nu, nv, seed=1 + 10 * i, center=(-3.6 + 2.4 * i, 0.0, 0.0)), mid)
    sd.add_mesh(ground_quad(y=-1.3, half=12.0), g)
    lq = np.array([[-2.0, 4.0, -1.0], [2.0, 4.0, -1.0], [2.0, 4.0, 1.0], [-2.0, 4.0, 1.0]], dtype=F32)
    light_tris = np.stack([np.concatenate([lq[0], lq[1], lq[2]]), np.concatenate([lq[0], lq[2], lq[3]])])
    lm = sd.add_material(type="matte", Kd=(0.0, 0.0, 0.0))
    sd.add_mesh(light_tris, lm, area_light=dict(L=(20, 20, 20)))
    sd.add_point_light((0.0, 3.0, -6.0), (50, 50, 50))
    sd.camera.update(eye=(0.0, 2.0, -9.0), look=(0.0, 0.0, 0.0), up=(0, 1, 0), fov=38.0)
    sd.film.update(xresolution=xres, yresolution=yres, filter="box")
    sd.sampler.update(type="halton", pixelsamples=spp)
    sd.integrator.update(maxdepth=maxdepth, lightsamplestrategy="power")
    return sd


def scene_c4(n_objects=46, nu=330, nv=330, xres=1920, yres=1080, spp=256, maxdepth=8):
    """C4: San-Miguel-scale stand-in — n_objects displaced spheres (46 x 217 800 = 10.0 M triangles) on an 8-wide grid
    with the four C3 materials in turn, ground quad, a two-triangle area light, a point light and a dim constant
    environment, lightsamplestrategy "power"; every triangle is a top-level primitive of ONE BVH (no instancing)."""
    sd = SceneDescription()
    mats = [sd.add_material(type="matte", Kd=(0.5, 0.45, 0.4)), sd.add_material(type="plastic"),
            sd.add_material(type="glass", eta=1.5), sd.add_material(type="metal", eta=COPPER_ETA, k=COPPER_K, roughness=0.01)]
    g = sd.add_material(type="matte", Kd=(0.4, 0.4, 0.4))
    cols = 8
    for i in range(n_objects):
        cx = (i % cols - (cols - 1) / 2.0) * 2.4
        cz = (i // cols) * 2.4
        sd.add_mesh(displaced_sphere(nu, nv, seed=1 + 10 * i, center=(cx, 0.0, cz)), mats[i % 4])
    sd.add_mesh(ground_quad(y=-1.3, half=30.0), g)
    lq = np.array([[-6.0, 6.0, 2.0], [6.0, 6.0, 2.0], [6.0, 6.0, 10.0], [-6.0, 6.0, 10.0]], dtype=F32)
    light_tris = np.stack([np.concatenate([lq[0], lq[1], lq[2]]), np.concatenate([lq[0], lq[2], lq[3]])])
    lm = sd.add_material(type="matte", Kd=(0.0, 0.0, 0.0))
    sd.add_mesh(light_tris, lm, area_light=dict(L=(20, 20, 20)))
    sd.add_point_light((0.0, 5.0, -8.0), (120, 120, 120))
    sd.add_infinite_light((0.3, 0.35, 0.45))
    sd.camera.update(eye=(0.0, 7.5, -13.0), look=(0.0, 0.0, 5.5), up=(0, 1, 0), fov=42.0)
    sd.film.update(xresolution=xres, yresolution=yres, filter="box")
    sd.sampler.update(type="halton", pixelsamples=spp)
    sd.integrator.update(maxdepth=maxdepth, lightsamplestrategy="power")
    return sd


def rigid_transform(rng, scale_range=(0.7, 1.3), extent=8.0, y_range=(-0.9, 1.5)):
    """Random rotation about a random axis, uniform scale and translation (instance_to_world, 4x4 row-major f32)."""
    axis = rng.normal(size=3)
    axis /= np.linalg.norm(axis)
    ang = rng.uniform(0, 2 * np.pi)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)
    M = np.eye(4, dtype=F32)
    M[:3, :3] = (rng.uniform(*scale_range) * R).astype(F32)
    M[:3, 3] = [rng.uniform(-extent, extent), rng.uniform(*y_range), rng.uniform(-extent, extent)]
    return M


def scene_c5(nu=224, nv=224, n_instances=1000, xres=1920, yres=1080, spp=128, maxdepth=5, seed=4):
    """C5: ecosys-style instancing — one displaced-sphere object (2*nu*nv ~ 100 K triangles) placed n_instances times
    with random rigid transforms (two-level BVH), a ground quad, constant infinite light L=[1 1 1]."""
    sd = SceneDescription()
    m = sd.add_material(type="matte", Kd=(0.45, 0.5, 0.35))
    g = sd.add_material(type="matte", Kd=(0.4, 0.4, 0.4))
    sd.add_mesh(ground_quad(y=-1.3, half=14.0), g)
    obj = sd.add_object(displaced_sphere(nu, nv, radius=0.35), m)
    rng = np.random.Generator(np.random.PCG64(seed))
    for _ in range(n_instances):
        sd.add_instance(obj, rigid_transform(rng))
    sd.add_infinite_light((1.0, 1.0, 1.0))
    sd.camera.update(eye=(0.0, 4.0, -13.0), look=(0.0, 0.0, 0.0), up=(0, 1, 0), fov=38.0)
    sd.film.update(xresolution=xres, yresolution=yres, filter="box")
    sd.sampler.update(type="halton", pixelsamples=spp)
    sd.integrator.update(maxdepth=maxdepth, lightsamplestrategy="uniform")
    return sd
