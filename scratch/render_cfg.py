import sys, time, numpy as np, json
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import __graft_entry__ as ge
pkg = ge.load_package(); pkg.init(0)
import oracle_lib as ol, scenes_small as ss
from pbrt_v3_rs_b200 import workloads as wl
import torch
which = sys.argv[1]
if which == 'c1': sd = wl.scene_c1()
elif which == 'c3': sd = wl.scene_c3(spp=int(sys.argv[2]) if len(sys.argv) > 2 else 64)
t = time.time(); integ = pkg.PathIntegrator(sd); integ.preprocess(); print('preprocess (incl BVH build) %.2fs' % (time.time()-t))
for it in range(3):
    torch.cuda.synchronize(); t = time.time(); film = integ.render_rows(); dt = time.time()-t
    rc = integ.ray_counts()
    print('render %.3fs  samples/s %.3e  rays: cam %d closest %d shadow %d  Mrays/s %.1f' % (dt, rc[0]/dt, rc[0], rc[1], rc[2], (rc[1]+rc[2])/dt/1e6))
img = integ.resolve(film)
np.save('gpurun_out/%s_gpu.npy' % which, img)
if which == 'c1' or '--oracle' in sys.argv:
    ref, stats, secs = ol.OracleScene(sd).render()
    print('oracle render %.2fs on %d threads, samples/s %.3e, stats %s' % (secs, ol.ncpu(), stats[0]/secs, stats))
    print('rel-RMSE', ss.rel_rmse(img, ref))
    np.save('gpurun_out/%s_ref.npy' % which, ref)
import os; os._exit(0)
