import sys, numpy as np, torch
sys.path.insert(0, '.')
import __graft_entry__ as ge
import bench
pkg = ge.load_package(); pkg.init(0)
w = bench.build_workload(pkg, False, 0)
accel = w['accel']; n = w['closest'].shape[0]; h = n // 2
sets = {'primary': w['closest'][:h], 'bounce': w['closest'][h:], 'shadow': w['shadow']}
variants = [int(v) for v in sys.argv[1:]] or [0, 2, 3]
ref = {}
for name, rays in sets.items():
    m = rays.shape[0]
    d_r = torch.from_numpy(np.ascontiguousarray(rays).view(np.float32).reshape(-1, 8)).cuda()
    anyhit = name == 'shadow'
    d_o = torch.zeros(m, dtype=torch.uint8, device='cuda') if anyhit else torch.zeros((m, 4), dtype=torch.float32, device='cuda')
    for v in variants:
        f = accel.occluded_batch_device if anyhit else accel.intersect_batch_device
        for _ in range(2): f(d_r.data_ptr(), m, d_o.data_ptr(), 0, v)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): f(d_r.data_ptr(), m, d_o.data_ptr(), 0, v)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        res = d_o.cpu().numpy().view(np.uint32 if not anyhit else np.uint8)
        if name not in ref: ref[name] = res
        ok = np.array_equal(ref[name], res)
        print('%-8s variant %d: %7.3f ms  %8.1f Mrays/s  %s' % (name, v, ms, m / ms / 1e3, 'same' if ok else 'DIFFERENT'))
import os; os._exit(0)
