import sys, time, numpy as np
sys.path.insert(0,'.')
import __graft_entry__ as ge
pkg = ge.load_package(); pkg.init(0)
from pbrt_v3_rs_b200 import workloads as wl
import torch
spp = int(sys.argv[1]); iters = int(sys.argv[2])
sd = wl.scene_c3(spp=spp)
integ = pkg.PathIntegrator(sd); integ.preprocess()
for it in range(iters):
    torch.cuda.synchronize(); t = time.time(); film = integ.render_rows(); dt = time.time()-t
    rc = integ.ray_counts()
    print('render %.3fs  samples/s %.3e  Mrays/s %.1f' % (dt, rc[0]/dt, (rc[1]+rc[2])/dt/1e6))
import os; os._exit(0)
