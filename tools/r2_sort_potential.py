"""How much would sorting the bounce rays buy?  C4 mesh (HBM-resident) / C2 mesh (L2-resident), 2^24 diffuse-bounce closest-hit
rays in three orders: shuffled (the microbench's), pixel order (what a render's first bounce looks like), sorted by a
Morton code of the origin (12 / 18 bits).  python tools/r2_sort_potential.py c4|c2"""
import sys
import numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as ge  # noqa: E402
pkg = ge.load_package()
pkg.init(0)
import torch  # noqa: E402
from pbrt_v3_rs_b200 import workloads as wl  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "c4"
if which == "c4":
    sd = wl.scene_c4()
    tv = sd.tri_verts
    prim = wl.primary_rays_lookat(4096, 4096, sd.camera["eye"], sd.camera["look"], sd.camera["up"], float(sd.camera["fov"]))
else:
    cfg = wl.C2_FULL
    tv = wl.c2_mesh(cfg)
    prim = wl.primary_rays(4096, 4096)
accel = pkg.BVHAccel.from_params({"splitmethod": "sah", "maxnodeprims": 4}, tv)
hits = accel.intersect_batch(prim)
shuffled = wl.bounce_rays(tv, prim, hits, prim.shape[0])
n = shuffled.shape[0]
inv = np.argsort(np.random.Generator(np.random.PCG64(3)).permutation(n))  # undo bounce_rays' shuffle
pixel_order = shuffled[inv]


def morton(o, bits):
    lo, hi = o.min(0), o.max(0)
    q = np.minimum(((o - lo) / np.maximum(hi - lo, 1e-20) * (1 << bits)).astype(np.int64), (1 << bits) - 1)
    code = np.zeros(o.shape[0], dtype=np.int64)
    for b in range(bits):
        for a in range(3):
            code |= ((q[:, a] >> b) & 1) << (3 * b + a)
    return code


def time_rays(r):
    d = torch.from_numpy(np.ascontiguousarray(r).view(np.float32).reshape(-1, 8)).cuda()
    h = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    for _ in range(2):
        accel.intersect_batch_device(d.data_ptr(), n, h.data_ptr(), 0, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        accel.intersect_batch_device(d.data_ptr(), n, h.data_ptr(), 0, 0)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5


for name, r in (("shuffled", shuffled), ("pixel order", pixel_order), ("morton 12 bits", shuffled[np.argsort(morton(shuffled["o"], 4), kind="stable")]),
                ("morton 18 bits", shuffled[np.argsort(morton(shuffled["o"], 6), kind="stable")]), ("morton 30 bits", shuffled[np.argsort(morton(shuffled["o"], 10), kind="stable")])):
    ms = time_rays(r)
    print("%s %-15s %.2f ms  %.0f Mrays/s" % (which, name, ms, n / ms / 1e3))
