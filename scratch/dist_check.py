import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, '.')
import __graft_entry__ as ge
rank = int(os.environ.get('RANK', 0)); lr = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(lr)
dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
pkg = ge.load_package(); pkg.init(lr)
from pbrt_v3_rs_b200 import workloads as wl, multigpu
for filt in ('box', 'gaussian'):
    sd = wl.scene_c3(nu=60, nv=60, xres=320, yres=180, spp=8)
    sd.film['filter'] = filt
    integ = pkg.PathIntegrator(sd); integ.preprocess()
    film = multigpu.render_distributed(integ).cpu().numpy()
    full = integ.render_rows()
    ok = np.allclose(film, full, rtol=2e-6, atol=1e-6)
    print('rank', rank, filt, 'distributed == single-GPU film:', ok, float(np.abs(film - full).max()))
dist.barrier(); dist.destroy_process_group()
