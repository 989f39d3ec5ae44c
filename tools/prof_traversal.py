"""One launch of the default traversal kernels under `ncu --profile-from-start off` (the capture brackets the third pair).

  python tools/prof_traversal.py c2        C2: 1 M-triangle mesh (L2-resident), 2^24 closest + 2^24 any-hit rays
  python tools/prof_traversal.py c4        C4-rays: 10 M-triangle mesh (HBM-resident), 2^24 incoherent bounce rays, closest-hit
  python tools/prof_traversal.py c3        C3 render at 16 spp (the per-material shade kernels: ncu -k regex:k_shade)
  python tools/prof_traversal.py c5        C5: two-level walk (1 000 instances of a 100 K-triangle object), 2^22 primary + bounce rays
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import __graft_entry__ as ge  # noqa: E402
import bench  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
pkg = ge.load_package()
pkg.init(0)
from pbrt_v3_rs_b200 import workloads as wl  # noqa: E402

if which in ("c5", "c3"):
    # two-level scenes have no stand-alone accelerator handle: drive the kernel through a render (c3: the shade kernels)
    sd = wl.scene_c5(spp=8) if which == "c5" else wl.scene_c3(spp=16)
    integ = pkg.PathIntegrator(sd)
    integ.preprocess()
    film = torch.zeros((1080, 1920, 4), dtype=torch.float32, device="cuda")
    integ.render_rows_device(0, 1080, film.data_ptr(), 0)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    integ.render_rows_device(0, 1080, film.data_ptr(), 0)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    os._exit(0)

if which == "c4":
    sd = wl.scene_c4()
    tv = sd.tri_verts
    accel = pkg.BVHAccel.from_params({"splitmethod": "sah", "maxnodeprims": 4}, tv)
    prim = wl.primary_rays_lookat(4096, 4096, sd.camera["eye"], sd.camera["look"], sd.camera["up"], float(sd.camera["fov"]))
    hits = accel.intersect_batch(prim)
    closest = wl.bounce_rays(tv, prim, hits, prim.shape[0])
    shadow = wl.shadow_rays(closest)
else:
    w = bench.build_workload(pkg, False, 0)
    accel, closest, shadow = w['accel'], w['closest'], w['shadow']
n = closest.shape[0]
d_c = torch.from_numpy(closest.view(np.float32).reshape(-1, 8)).cuda()
d_s = torch.from_numpy(shadow.view(np.float32).reshape(-1, 8)).cuda()
d_h = torch.zeros((n, 4), dtype=torch.float32, device='cuda')
d_o = torch.zeros(n, dtype=torch.uint8, device='cuda')
for it in range(3):
    if it == 2:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()   # ncu --profile-from-start off: capture the third pair only
    accel.intersect_batch_device(d_c.data_ptr(), n, d_h.data_ptr(), 0, 0)
    accel.occluded_batch_device(d_s.data_ptr(), n, d_o.data_ptr(), 0, 0)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
os._exit(0)
