"""3840x2160 @ 1024 spp on one GPU: 8.5e9 camera samples, rendered in 2^26-path waves whose film sums
carry over; prints time, peak device memory and the number of waves.  python tools/r2_4k.py [spp]"""
import sys
import time

sys.path.insert(0, ".")
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
pkg.init(0)
import torch  # noqa: E402
from pbrt_v3_rs_b200 import workloads as wl  # noqa: E402

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
sd = wl.scene_c3(xres=3840, yres=2160, spp=spp)
integ = pkg.PathIntegrator(sd)
integ.preprocess()
film = torch.zeros((2160, 3840, 4), dtype=torch.float32, device="cuda")
free0, total = torch.cuda.mem_get_info()
t0 = time.time()
integ.render_rows_device(0, 2160, film.data_ptr(), 0)
torch.cuda.synchronize()
dt = time.time() - t0
free1, _ = torch.cuda.mem_get_info()
rc = integ.ray_counts()
n = 3840 * 2160 * spp
print("3840x2160 @ %d spp: %.2f s, %.3e samples/s, %d camera / %d closest / %d shadow rays, %d waves of <= 2^26 paths, device memory in use after the render %.1f GB of %.0f GB (before: %.1f GB), film weight min %.0f max %.0f"
      % (spp, dt, n / dt, rc[0], rc[1], rc[2], -(-n // (1 << 26)), (total - free1) / 1e9, total / 1e9, (total - free0) / 1e9, float(film[..., 3].min()), float(film[..., 3].max())))
