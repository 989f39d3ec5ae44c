import sys, time, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import __graft_entry__ as ge
ge.build_oracle(); pkg = ge.load_package()
import oracle_lib as ol
from pbrt_v3_rs_b200 import workloads as wl
def save(img, path):
    from PIL import Image
    x = np.clip(img, 0, 1) ** (1/2.2)
    Image.fromarray((x*255).astype(np.uint8)).save(path)
which = sys.argv[1]
if which == 'c1':
    sd = wl.scene_c1(nu=100, nv=100, res=200, spp=16)
else:
    sd = wl.scene_c3(nu=60, nv=60, xres=320, yres=180, spp=32)
t=time.time(); sc = ol.OracleScene(sd); print('setup', time.time()-t)
img, stats, secs = sc.render()
print('render secs', secs, 'stats', stats, 'mean', img.mean(), 'max', img.max(), 'nan', np.isnan(img).sum())
save(img, 'scratch/%s.png' % which)
