"""The fixed cost of a wave: C3's scene rendered into a tiny film (a few thousand paths), so that every launch of the bounce
loop is (almost) empty and the time is launch + drain latency only.  python tools/r2_floor.py"""
import sys
import time

sys.path.insert(0, ".")
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
pkg.init(0)
import torch  # noqa: E402
from pbrt_v3_rs_b200 import workloads as wl  # noqa: E402

for xres, yres, spp in ((64, 36, 1), (256, 144, 4), (960, 540, 8), (1920, 1080, 8)):
    sd = wl.scene_c3(xres=xres, yres=yres, spp=spp)
    integ = pkg.PathIntegrator(sd)
    integ.preprocess()
    film = torch.zeros((yres, xres, 4), dtype=torch.float32, device="cuda")
    best = 1e9
    l0 = 0
    for _ in range(5):
        torch.cuda.synchronize()
        l0 = pkg.launch_count()
        t0 = time.perf_counter()
        integ.render_rows_device(0, yres, film.data_ptr(), 0)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print("%dx%d @ %d spp = %d paths: %.3f ms, %d launches" % (xres, yres, spp, xres * yres * spp, best * 1e3, pkg.launch_count() - l0))
    del integ
