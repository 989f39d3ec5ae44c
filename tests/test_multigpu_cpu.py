"""N > 1 host logic on CPU (gloo, world_size 2): row-band sharding partitions the image and the one collective of
the path — a gather of the owned bands on rank 0 (box-sized filters), or a sum all-reduce of the per-rank films (wider
filters) — reassembles it."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, h, w, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import __graft_entry__ as ge
    ge.load_package()
    from pbrt_v3_rs_b200 import multigpu
    full = torch.from_numpy(np.random.default_rng(5).random((h, w, 4)).astype(np.float32))
    rows = multigpu.shard_rows(h, world, rank)
    part = torch.zeros_like(full)
    part[rows] = full[rows]
    out = multigpu.reduce_film(part.clone())
    ok = bool(torch.equal(out, full))
    # the gather of owned bands (box-sized filters): rank 0 ends up with the whole image, nobody ships more than its rows
    # plus one spill row per band (samples sitting exactly on a pixel boundary also land in the row above their band)
    spill_all = torch.from_numpy(np.random.default_rng(6).random((h, w, 4)).astype(np.float32))
    want = full.clone()
    for r in range(world):
        sr = multigpu.spill_rows(h, world, r)
        want[sr] += spill_all[sr]
    mine_spill = multigpu.spill_rows(h, world, rank)
    part2 = part.clone()
    part2[mine_spill] = spill_all[mine_spill]  # rows this rank does not own: they hold only its boundary contributions
    g = multigpu.BandGather(h, w, "cpu")
    got = g(part2.clone())
    ok = ok and (bool(torch.equal(got, want)) if rank == 0 else bool(torch.equal(got, part2)))
    ok = ok and g.mine.shape[0] == max(multigpu.shard_rows(h, world, r).size for r in range(world)) + max(multigpu.spill_rows(h, world, r).size for r in range(world))
    all_spill = np.concatenate([multigpu.spill_rows(h, world, r) for r in range(world)])
    # every band boundary between two different owners has exactly one spill row
    want_spill = [k * multigpu.BAND_ROWS - 1 for k in range(1, -(-h // multigpu.BAND_ROWS)) if multigpu.band_owner(k, world) != multigpu.band_owner(k - 1, world)]
    ok = ok and sorted(all_spill.tolist()) == want_spill
    q.put((rank, ok, int(rows.size)))
    dist.destroy_process_group()


def test_shard_rows_partition(pkg):
    from pbrt_v3_rs_b200 import multigpu
    for h in (1, 7, 8, 9, 90, 1080):
        for n in (1, 2, 3, 4, 8):
            allrows = np.concatenate([multigpu.shard_rows(h, n, r) for r in range(n)])
            assert sorted(allrows) == list(range(h))
    # 1080 rows in bands of 8 over 8 ranks: every rank gets 128..136 rows (load balance by interleaving)
    sizes = [multigpu.shard_rows(1080, 8, r).size for r in range(8)]
    assert max(sizes) - min(sizes) <= 8


def test_film_allreduce_gloo_world2(pkg):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 37, 16, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res)
    assert sum(n for _, _, n in res) == 37
