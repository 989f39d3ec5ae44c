import sys, time
sys.path.insert(0, ".")
import __graft_entry__ as ge
pkg = ge.load_package(); pkg.init(0)
import torch
from pbrt_v3_rs_b200 import workloads as wl
which, n, r = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
sd = wl.scene_c3() if which == "c3" else wl.scene_c4()
integ = pkg.PathIntegrator(sd); integ.preprocess()
h, w = integ.film_shape()
film = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
ts = []
for _ in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    integ.render_shard_device_raw(r, n, film.data_ptr(), 8, 0)
    torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print(which, "shard", r, "of", n, " ".join("%.2f" % t for t in ts))
