"""Small path-tracing scenes shared by the CPU and GPU tests."""
import numpy as np


def one_material_scene(wl, mat, light="infinite", res=24, spp=8, maxdepth=5, nu=24, nv=12, strategy="uniform", filt="box", smooth=False):
    from pbrt_v3_rs_b200.scene import SceneDescription
    sd = SceneDescription()
    m = sd.add_material(**mat)
    g = sd.add_material(type="matte", Kd=(0.4, 0.4, 0.4))
    if smooth:  # the mesh carries "uv" and "N" (triangle.rs:384-394, 631-721)
        tv, uv, nrm = wl.displaced_sphere(nu, nv, with_attrs=True)
        sd.add_mesh(tv, m, uv=uv, normals=nrm)
    else:
        sd.add_mesh(wl.displaced_sphere(nu, nv), m)
    sd.add_mesh(wl.ground_quad(), g)
    if light in ("infinite", "all"):
        sd.add_infinite_light((1.2, 1.2, 1.1))
    if light in ("point", "all"):
        sd.add_point_light((1.5, 3.0, -3.0), (30, 30, 30))
    if light in ("spot", "all+spot"):
        sd.add_spot_light((40, 38, 35), (1.0, 3.0, -2.0), (0.0, -0.5, 0.0), coneangle=28.0, conedeltaangle=12.0)
    if light == "all+spot":
        light = "all"
    if light in ("distant", "all+distant"):
        sd.add_distant_light((2.5, 2.4, 2.2), (-0.4, 1.0, -0.6))
    if light == "all+distant":
        light = "all"
        sd.add_infinite_light((1.2, 1.2, 1.1))
        sd.add_point_light((1.5, 3.0, -3.0), (30, 30, 30))
    if light in ("area", "all"):
        lq = np.array([[-1.5, 3.0, -1.0], [1.5, 3.0, -1.0], [1.5, 3.0, 1.0], [-1.5, 3.0, 1.0]], dtype=np.float32)
        lt = np.stack([np.concatenate([lq[0], lq[1], lq[2]]), np.concatenate([lq[0], lq[2], lq[3]])])
        lm = sd.add_material(type="matte", Kd=(0.0, 0.0, 0.0))
        sd.add_mesh(lt, lm, area_light=dict(L=(15, 15, 15)))
    sd.camera.update(eye=(0.0, 1.2, -4.0), look=(0.0, -0.1, 0.0), up=(0, 1, 0), fov=40.0)
    sd.film.update(xresolution=res, yresolution=res, filter=filt)
    sd.sampler.update(type="halton", pixelsamples=spp)
    sd.integrator.update(maxdepth=maxdepth, lightsamplestrategy=strategy)
    return sd


MATERIALS = {
    "matte": dict(type="matte", Kd=(0.5, 0.45, 0.4)),
    "oren_nayar": dict(type="matte", Kd=(0.5, 0.5, 0.5), sigma=30.0),
    "plastic": dict(type="plastic"),
    "glass": dict(type="glass", eta=1.5),
    "rough_glass": dict(type="glass", eta=1.5, uroughness=0.2, vroughness=0.1),
    "metal": dict(type="metal", roughness=0.01),
    "mirror": dict(type="mirror", Kr=(0.9, 0.85, 0.8)),
    "rough_metal": dict(type="metal", uroughness=0.3, vroughness=0.1, remaproughness=False),
}


def rel_rmse(a, b, eps=1e-2):
    """Per-pixel relative RMSE (north_star's image gate, <= 1e-3)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.sqrt(np.mean(((a - b) / (np.abs(b) + eps)) ** 2)))


SLAB_LE = (3.0, 2.0, 1.0)


def glass_slab_scene(maxdepth, integrator="whitted"):
    """A pencil of camera rays along +z through a glass slab (eta 1.5, faces at z = 1 and z = 1.5 with outward normals) onto
    a one-sided emitter at z = 3 that faces the camera.  The centre ray of the 3x3 image is exactly (0, 0, 1)."""
    from pbrt_v3_rs_b200.scene import SceneDescription

    def quad(z, s=4.0, flip=False):
        q = np.array([[-s, -s, z], [s, -s, z], [s, s, z], [-s, s, z]], dtype=np.float32)
        if flip:
            return np.stack([np.concatenate([q[0], q[2], q[1]]), np.concatenate([q[0], q[3], q[2]])])
        return np.stack([np.concatenate([q[0], q[1], q[2]]), np.concatenate([q[0], q[2], q[3]])])
    sd = SceneDescription()
    g = sd.add_material(type="glass", eta=1.5)
    e = sd.add_material(type="matte", Kd=(0, 0, 0))
    sd.add_mesh(quad(1.0, flip=True), g)    # front face, outward normal -z (towards the camera)
    sd.add_mesh(quad(1.5), g)               # back face, outward normal +z: the ray leaves the glass there
    sd.add_mesh(quad(3.0, flip=True), e, area_light=dict(L=SLAB_LE))
    sd.camera.update(eye=(0.0, 0.0, -2.0), look=(0.0, 0.0, 1.0), up=(0, 1, 0), fov=1.0)
    sd.film.update(xresolution=3, yresolution=3)
    sd.sampler.update(type="halton", pixelsamples=1, samplepixelcenter=True)
    sd.integrator.update(name=integrator, maxdepth=maxdepth)
    return sd


# ---- frozen per-sample radiances (tests/golden/li_small.npz) -------------------------------------------------------
LI_GOLDEN_CASES = {
    "path_halton_power": dict(integrator="path", sampler="halton", strategy="power"),
    "path_halton_spatial": dict(integrator="path", sampler="halton", strategy="spatial"),
    "path_zerotwo_uniform": dict(integrator="path", sampler="02sequence", strategy="uniform"),
    "path_sobol_power": dict(integrator="path", sampler="sobol", strategy="power"),
    "whitted_halton": dict(integrator="whitted", sampler="halton", strategy="uniform"),
    "whitted_sobol": dict(integrator="whitted", sampler="sobol", strategy="uniform"),
    "direct_all_halton": dict(integrator="directlighting", sampler="halton", strategy="uniform", direct="all"),
    "direct_one_halton": dict(integrator="directlighting", sampler="halton", strategy="uniform", direct="one"),
}


def li_golden_scene(wl, integrator, sampler, strategy, direct="all"):
    sd = one_material_scene(wl, MATERIALS["glass"], light="all", res=12, spp=4, maxdepth=4, nu=16, nv=8, strategy="uniform")
    sd.sampler.update(type=sampler)
    if sampler == "02sequence":
        sd.sampler.update(dimensions=16)
    sd.integrator.update(name=integrator, lightsamplestrategy=strategy, strategy=direct)
    return sd


def li_golden_pairs():
    return np.array([(x, y, s) for y in range(12) for x in range(12) for s in range(4)], dtype=np.int32)
