import sys, time
sys.path.insert(0, ".")
import __graft_entry__ as ge
pkg = ge.load_package(); pkg.init(0)
from pbrt_v3_rs_b200 import workloads as wl
import numpy as np
t=time.time(); sd = wl.scene_c4(); print("scene_gen %.2f" % (time.time()-t))
pkg.build_bvh_sah(pkg.triangle_bounds(sd.tri_verts[:100000]), 4)  # warm the builder
t=time.time(); pb = pkg.triangle_bounds(sd.tri_verts); print("triangle_bounds(host) %.3f" % (time.time()-t))
t=time.time(); n,o = pkg.build_bvh_sah(pb, 4, where="gpu"); print("build_bvh_sah gpu (host in/out) %.3f" % (time.time()-t))
t=time.time(); sd.build_accel(None); print("sd.build_accel %.3f" % (time.time()-t))
t=time.time(); d = sd.to_desc(); print("to_desc %.3f" % (time.time()-t))
import ctypes as C
h = C.c_void_p()
t=time.time(); rc = pkg.lib().b200pt_scene_create(C.byref(d), C.byref(h)); print("scene_create %.3f rc=%d" % (time.time()-t, rc))
