"""Strong-scaling estimate on ONE GPU: renders each of the N shards of C3 / C4 in turn (the work one rank of an N-GPU run
does) and prints per-shard times, their imbalance and the implied efficiency T(1) / (N * max shard).
  python tools/r2_shard_time.py c3 8 [band_rows]     (B200PT_BAND_ORDER=roundrobin for the old dealing order)"""
import sys
import time

sys.path.insert(0, ".")
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
pkg.init(0)
import torch  # noqa: E402
from pbrt_v3_rs_b200 import workloads as wl  # noqa: E402

which, n = sys.argv[1], int(sys.argv[2])
band_rows = int(sys.argv[3]) if len(sys.argv) > 3 else 8
sd = wl.scene_c3() if which == "c3" else wl.scene_c4()
integ = pkg.PathIntegrator(sd)
integ.preprocess()
h, w = integ.film_shape()
film = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")


def timed(fn, reps=2):
    best = 1e9
    for _ in range(reps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


full = timed(lambda: integ.render_rows_device(0, h, film.data_ptr(), 0))
shards = [timed(lambda r=r: integ.render_shard_device_raw(r, n, film.data_ptr(), band_rows, 0)) for r in range(n)]
print("%s full %.2f ms; %d shards (band_rows %d): %s ms; max %.2f mean %.2f (imbalance %.1f %%); ideal %.2f; efficiency by max shard %.3f, by mean shard %.3f"
      % (which, full, n, band_rows, " ".join("%.2f" % t for t in shards), max(shards), sum(shards) / n, 100 * (max(shards) / (sum(shards) / n) - 1), full / n,
         full / (n * max(shards)), full / sum(shards)))
