"""One tiny C3 render (2 304 paths) for a launch list of the bounce loop's fixed cost: python tools/r2_floor_one.py"""
import sys
sys.path.insert(0, ".")
import __graft_entry__ as ge  # noqa: E402
pkg = ge.load_package()
pkg.init(0)
import torch  # noqa: E402
from pbrt_v3_rs_b200 import workloads as wl  # noqa: E402
sd = wl.scene_c3(xres=64, yres=36, spp=1)
integ = pkg.PathIntegrator(sd)
integ.preprocess()
film = torch.zeros((36, 64, 4), dtype=torch.float32, device="cuda")
integ.render_rows_device(0, 36, film.data_ptr(), 0)
torch.cuda.synchronize()
torch.cuda.profiler.start()
integ.render_rows_device(0, 36, film.data_ptr(), 0)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
