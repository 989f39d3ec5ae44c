import sys, numpy as np, torch
sys.path.insert(0, '.')
import __graft_entry__ as ge
import bench
pkg = ge.load_package(); pkg.init(0)
w = bench.build_workload(pkg, False, 0)
accel = w['accel']; n = w['closest'].shape[0]; h = n // 2
def morton3(q, bits):
    k = np.zeros(q.shape[0], dtype=np.uint64)
    for b in range(bits):
        for a in range(3):
            k |= ((q[:, a] >> b) & 1).astype(np.uint64) << np.uint64(3 * b + a)
    return k
def sort_rays(rays, bits, octant):
    o = rays['o']; lo, hi = o.min(0), o.max(0)
    q = np.minimum(((o - lo) / (hi - lo + 1e-9) * (1 << bits)).astype(np.int64), (1 << bits) - 1)
    k = morton3(q, bits)
    if octant:
        d = rays['d']; oc = (d[:, 0] < 0).astype(np.uint64) | ((d[:, 1] < 0).astype(np.uint64) << np.uint64(1)) | ((d[:, 2] < 0).astype(np.uint64) << np.uint64(2))
        if octant == 1: k = (k << np.uint64(3)) | oc
        else: k = k | (oc << np.uint64(3 * bits))
    return rays[np.argsort(k, kind='stable')]
def run(name, rays, anyhit, v=0):
    m = rays.shape[0]
    d_r = torch.from_numpy(np.ascontiguousarray(rays).view(np.float32).reshape(-1, 8)).cuda()
    d_o = torch.zeros(m, dtype=torch.uint8, device='cuda') if anyhit else torch.zeros((m, 4), dtype=torch.float32, device='cuda')
    f = accel.occluded_batch_device if anyhit else accel.intersect_batch_device
    for _ in range(2): f(d_r.data_ptr(), m, d_o.data_ptr(), 0, v)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): f(d_r.data_ptr(), m, d_o.data_ptr(), 0, v)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print('%-34s %7.3f ms  %8.1f Mrays/s' % (name, ms, m / ms / 1e3), flush=True)
b = w['closest'][h:]; s = w['shadow']
run('bounce shuffled', b, False)
for bits, oc in [(4, 0), (5, 0), (5, 1), (5, 2), (6, 0), (7, 0), (7, 1), (10, 0)]:
    run('bounce morton %d bits octant %d' % (bits, oc), sort_rays(b, bits, oc), False)
run('shadow as generated', s, True)
run('shadow 2nd half as generated', s[h:], True)
for bits, oc in [(5, 0), (7, 0), (7, 1)]:
    run('shadow 2nd half morton %d oct %d' % (bits, oc), sort_rays(s[h:], bits, oc), True)
import os; os._exit(0)
