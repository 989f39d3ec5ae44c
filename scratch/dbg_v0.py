import sys, numpy as np, torch
sys.path.insert(0, '.')
import __graft_entry__ as ge
import bench
pkg = ge.load_package(); pkg.init(0)
w = bench.build_workload(pkg, False, 0)
accel = w['accel']
for kind in ('closest', 'shadow'):
    rays = w[kind]; n = rays.shape[0]
    d_r = torch.from_numpy(rays.view(np.float32).reshape(-1, 8)).cuda()
    print(kind, 'nan rays', np.isnan(rays.view(np.float32)).sum(), 'zero dir comps', (rays['d'] == 0).sum())
    outs = {}
    for v in (2, 0, 0, 1):
        if kind == 'closest':
            d_h = torch.zeros((n, 4), dtype=torch.float32, device='cuda')
            accel.intersect_batch_device(d_r.data_ptr(), n, d_h.data_ptr(), 0, v)
            torch.cuda.synchronize()
            outs.setdefault(v, []).append(d_h.cpu().numpy().view(np.uint32).reshape(-1, 4))
        else:
            d_h = torch.zeros(n, dtype=torch.uint8, device='cuda')
            accel.occluded_batch_device(d_r.data_ptr(), n, d_h.data_ptr(), 0, v)
            torch.cuda.synchronize()
            outs.setdefault(v, []).append(d_h.cpu().numpy().reshape(-1, 1))
    hb = accel.intersect_batch(rays).view(np.uint32).reshape(-1, 4) if kind == 'closest' else accel.occluded_batch(rays).reshape(-1, 1)
    ref = outs[2][0]
    for name, h in (('v0a', outs[0][0]), ('v0b', outs[0][1]), ('v1', outs[1][0]), ('chunked', hb)):
        d = (h != ref).any(1)
        print(' ', name, 'diff rows', d.sum(), 'first', np.nonzero(d)[0][:6])
        for i in np.nonzero(d)[0][:3]: print('     ', i, h[i], ref[i], rays[i])
    if kind == 'closest':
        print('  nan in ref hits', np.isnan(ref.view(np.float32)).sum())
import os; os._exit(0)
