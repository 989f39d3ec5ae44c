import sys, os
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import __graft_entry__ as ge
gpu = ge.load_package()
gpu.init(0)
import scenes_small as ss
from pbrt_v3_rs_b200 import workloads as wl
import torch
n = torch.cuda.device_count()
sd = ss.one_material_scene(wl, ss.MATERIALS["plastic"], light="all", res=72, spp=8, maxdepth=4, strategy="power", filt="box")
single = gpu.PathIntegrator(sd)
full = single.render_rows()
for mode in ("nccl", "peer"):
    if mode == "peer": os.environ["B200PT_GATHER"] = "peer"
    multi = gpu.MultiGPURender(sd, list(range(n)))
    film = multi.render_rows(band_rows=8)
    d = np.abs(film - full).max(2)
    rows = np.nonzero(d.max(1) > 0)[0]
    print(mode, "rows differing:", rows.tolist(), "max", d.max(), "info", multi.info())
    if len(rows):
        y = rows[0]; x = int(np.argmax(d[y])); print(" at", y, x, film[y, x], full[y, x])
    multi.close()
# shard renders on device 1 directly
gpu.set_device(1) if hasattr(gpu, "set_device") else None
