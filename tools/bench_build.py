"""Times the GPU SAH builder (b200pt_bvh_build_sah_device, bounds already in HBM) against the host builder.
Usage: python tools/bench_build.py [n_triangles ...]   (default: C2's 1 M mesh and a 10 M soup)"""
import ctypes as C
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import __graft_entry__ as ge

pkg = ge.load_package()
from pbrt_v3_rs_b200 import workloads as wl

pkg.init(0)
L = pkg.lib()


def run(name, tv, host=True):
    n = tv.shape[0]
    d_tv = torch.from_numpy(tv).cuda()
    d_pb = torch.empty((n, 6), dtype=torch.float32, device="cuda")
    d_nodes = torch.empty((2 * n, 32), dtype=torch.uint8, device="cuda")
    d_ord = torch.empty(n, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    nn = C.c_int64(0)
    times = []
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = pkg.launch_count()
        e0.record()
        assert L.b200pt_triangle_bounds_device(d_tv.data_ptr(), n, d_pb.data_ptr(), st) == 0
        rc = L.b200pt_bvh_build_sah_device(d_pb.data_ptr(), n, 4, d_nodes.data_ptr(), C.byref(nn), d_ord.data_ptr(), st)
        assert rc == 0, L.b200pt_last_error()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
        launches = pkg.launch_count() - l0
    out = {"mesh": name, "triangles": n, "nodes": nn.value, "gpu_ms": min(times[1:]), "gpu_ms_all": times, "launches": launches,
           "mtris_per_s_gpu": n / min(times[1:]) / 1e3}
    rt = []
    for it in range(4):  # whole accelerator (bounds + build + traversal records) from device-resident triangles
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        acc = pkg.BVHAccel.from_device_triangles(d_tv.data_ptr(), n, 4, stream=st)
        torch.cuda.synchronize()
        rt.append((time.perf_counter() - t0) * 1e3)
        acc.close()
    out["accel_create_device_ms"] = min(rt[1:])
    ht = []
    for it in range(3):  # SplitMethod::HLBVH on the GPU (same buffers)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = L.b200pt_bvh_build_hlbvh_device(d_pb.data_ptr(), n, 4, d_nodes.data_ptr(), C.byref(nn), d_ord.data_ptr(), st)
        assert rc == 0, L.b200pt_last_error()
        e1.record()
        torch.cuda.synchronize()
        ht.append(e0.elapsed_time(e1))
    out["hlbvh_gpu_ms"] = min(ht[1:])
    if host:
        t0 = time.perf_counter()
        n2, o2 = pkg.build_bvh_hlbvh(d_pb.cpu().numpy(), 4)
        out["hlbvh_host_ms"] = (time.perf_counter() - t0) * 1e3
        out["hlbvh_identical"] = bool(d_nodes[:nn.value].cpu().numpy().tobytes() == n2.tobytes() and np.array_equal(d_ord.cpu().numpy().view(np.uint32), o2))
        # the SAH build again so that the comparison below sees its output
        assert L.b200pt_bvh_build_sah_device(d_pb.data_ptr(), n, 4, d_nodes.data_ptr(), C.byref(nn), d_ord.data_ptr(), st) == 0
        torch.cuda.synchronize()
    if host:
        pb = d_pb.cpu().numpy()
        t0 = time.perf_counter()
        n1, o1 = pkg.build_bvh_sah(pb, 4, where="host")
        out["host_ms"] = (time.perf_counter() - t0) * 1e3
        out["identical"] = bool(d_nodes[:nn.value].cpu().numpy().tobytes() == n1.tobytes() and np.array_equal(d_ord.cpu().numpy().view(np.uint32), o1))
        out["speedup"] = out["host_ms"] / out["gpu_ms"]
    print(json.dumps(out), flush=True)


sizes = [int(a) for a in sys.argv[1:]]
if not sizes:
    run("c2_displaced_sphere_1M", wl.c2_mesh(wl.C2_FULL))
    run("soup_10M", wl.triangle_soup(10_000_000))
else:
    for s in sizes:
        run("soup_%d" % s, wl.triangle_soup(s))
