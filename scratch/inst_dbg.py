import sys, numpy as np, time
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import __graft_entry__ as ge
pkg = ge.load_package()
import oracle_lib as ol
from pbrt_v3_rs_b200 import workloads as wl
from pbrt_v3_rs_b200.scene import SceneDescription
def build(instanced, res=64, spp=8):
    sd = SceneDescription()
    m = sd.add_material(type="matte", Kd=(0.6,0.6,0.6)); g = sd.add_material(type="matte", Kd=(0.4,0.4,0.4))
    sd.add_mesh(wl.ground_quad(), g)
    sph = wl.displaced_sphere(24, 12, radius=0.45)
    obj = sd.add_object(sph, m) if instanced else None
    rng = np.random.default_rng(4)
    for k in range(9):
        a = rng.uniform(0, 2*np.pi); c, s_ = np.cos(a), np.sin(a); sc = rng.uniform(0.6, 1.3)
        M = np.eye(4, dtype=np.float32)
        M[:3,:3] = sc*np.array([[c,0,s_],[0,1,0],[-s_,0,c]], dtype=np.float32)
        M[:3,3] = [(k%3-1)*1.2, -0.6 + 0.3*(k//3), (k//3-1)*1.2]
        if instanced: sd.add_instance(obj, M)
        else:
            v = sph.reshape(-1,3) @ M[:3,:3].T + M[:3,3]
            sd.add_mesh(v.reshape(-1,9).astype(np.float32), m)
    sd.add_infinite_light((1.0,1.0,1.0))
    sd.camera.update(eye=(0.0, 2.5, -5.0), look=(0.0, -0.3, 0.0), up=(0,1,0), fov=40.0)
    sd.film.update(xresolution=res, yresolution=res); sd.sampler.update(pixelsamples=spp)
    sd.integrator.update(maxdepth=3, lightsamplestrategy="uniform")
    return sd
a = ol.OracleScene(build(True)).render()[0]
b = ol.OracleScene(build(False)).render()[0]
print('instanced mean %.4f  baked mean %.4f  black frac %.3f vs %.3f' % (a.mean(), b.mean(), (a.max(2)==0).mean(), (b.max(2)==0).mean()))
print(np.round(a[40,20:44,0],2)); print(np.round(b[40,20:44,0],2))
