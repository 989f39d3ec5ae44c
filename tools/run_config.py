"""Full-size runs of BASELINE.json's path-tracing configurations (C1, C3, C4, C5) on one GPU:
render time / samples per second / rays per second, plus parity against the CPU oracle at the FULL scene size:

  * per-sample radiance (Integrator::li) on --li random (pixel, sample) triples, bit-compared and within tolerance;
  * an oracle render of a --crop fraction of the image (centre window) vs the same crop rendered on the GPU,
    per-pixel relative RMSE (the <= 1e-3 gate of north_star).

Usage (on the GPU box):  python tools/run_config.py c4 [--spp 256] [--li 4096] [--crop 0.08] [--reps 2] [--json out.json]
This is measurement tooling: it loads the oracle (tests/oracle_lib.py) as the checker only.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["c1", "c3", "c4", "c5"])
    ap.add_argument("--spp", type=int, default=None)
    ap.add_argument("--li", type=int, default=4096, help="number of (pixel, sample) triples checked against the oracle (0 = skip)")
    ap.add_argument("--crop", type=float, default=0.0, help="fraction of width/height of the centre window rendered by the oracle (0 = skip)")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--small", action="store_true", help="reduced geometry (CPU-sized smoke run of this tool)")
    ap.add_argument("--json", default=None)
    ap.add_argument("--integrator", default="path", choices=["path", "whitted", "directlighting"])
    ap.add_argument("--strategy", default="all", choices=["all", "one"], help="directlighting strategy")
    a = ap.parse_args()

    import __graft_entry__ as ge
    pkg = ge.load_package()
    pkg.init(0)
    import torch
    from pbrt_v3_rs_b200 import workloads as wl

    kw = {}
    if a.spp:
        kw["spp"] = a.spp
    if a.small:
        kw.update(dict(nu=40, nv=40))
        if a.config == "c5":
            kw["n_instances"] = 50
        if a.config == "c4":
            kw["n_objects"] = 6
    t0 = time.time()
    sd = {"c1": wl.scene_c1, "c3": wl.scene_c3, "c4": wl.scene_c4, "c5": wl.scene_c5}[a.config](**kw)
    t_gen = time.time() - t0
    sd.integrator.update(name=a.integrator, strategy=a.strategy)
    t0 = time.time()
    integ = pkg.PathIntegrator(sd)
    integ.preprocess()
    t_pre = time.time() - t0
    n_tris = int(sd.tri_verts.shape[0]) + sum(int(o["tri_verts"].shape[0]) for o in sd.objects)
    n_inst_tris = int(sd.tri_verts.shape[0]) + sum(int(sd.objects[o]["tri_verts"].shape[0]) for o, _, _ in sd.instances)
    out = {"config": a.config, "integrator": a.integrator + ("/" + a.strategy if a.integrator == "directlighting" else ""), "stored_triangles": n_tris, "instanced_triangles": n_inst_tris, "spp": sd.sampler["pixelsamples"], "scene_gen_s": t_gen, "preprocess_s": t_pre}
    print("%d stored / %d instanced triangles" % (n_tris, n_inst_tris))
    print("scene generated in %.1f s, preprocess (SAH builds + upload) %.1f s" % (t_gen, t_pre), flush=True)

    best = None
    for it in range(a.reps):
        torch.cuda.synchronize()
        t0 = time.time()
        film = integ.render_rows()
        dt = time.time() - t0
        rc = integ.ray_counts()
        print("render %d: %.3f s  %.3e samples/s  rays: camera %d closest %d shadow %d  %.1f Mrays/s" %
              (it, dt, rc[0] / dt, rc[0], rc[1], rc[2], (rc[1] + rc[2]) / dt / 1e6), flush=True)
        if best is None or dt < best:
            best = dt
            out.update(render_s=dt, samples=int(rc[0]), samples_per_s=rc[0] / dt, closest_rays=int(rc[1]), shadow_rays=int(rc[2]),
                       mrays_per_s=(int(rc[1]) + int(rc[2])) / dt / 1e6)
    img = integ.resolve(film)
    out["image_mean"] = float(img.mean())
    out["image_finite"] = bool(np.isfinite(img).all())

    if a.li > 0 or a.crop > 0:
        import oracle_lib as ol
        import scenes_small as ss
        t0 = time.time()
        osc = ol.OracleScene(sd)
        print("oracle scene built in %.1f s" % (time.time() - t0), flush=True)
    if a.li > 0:
        h, w = integ.film_shape()
        rng = np.random.Generator(np.random.PCG64(11))
        ps = np.stack([rng.integers(0, w, a.li), rng.integers(0, h, a.li), rng.integers(0, sd.sampler["pixelsamples"], a.li)], axis=1).astype(np.int32)
        li, rays = integ.li(ps)
        t0 = time.time()
        oli = osc.li(ps)
        t_li = time.time() - t0
        rays_same = rays.tobytes() == osc.camera_rays(ps).tobytes()
        bit = (li.view(np.uint32) == oli.view(np.uint32)).all(1).mean()
        close = np.isclose(li, oli, rtol=2e-3, atol=1e-5).all(1).mean()
        print("li parity on %d samples: camera rays identical %s, bit-identical radiance %.4f, within 2e-3 %.4f (oracle %.1f s)" %
              (a.li, rays_same, bit, close, t_li), flush=True)
        out.update(li_samples=a.li, li_camera_rays_identical=bool(rays_same), li_bit_identical_frac=float(bit), li_close_frac=float(close))
    if a.crop > 0:
        lo, hi = 0.5 - a.crop / 2, 0.5 + a.crop / 2
        sd.film["cropwindow"] = (lo, hi, lo, hi)
        integ2 = pkg.PathIntegrator(sd)
        gimg = integ2.render()
        osc2 = ol.OracleScene(sd)
        ref, stats, secs = osc2.render()
        rel = ss.rel_rmse(gimg, ref)
        rc = integ2.ray_counts()
        print("crop %dx%d: rel-RMSE %.3e; oracle %.1f s on %d threads = %.3e samples/s; rays gpu (%d, %d, %d) oracle (%d, %d, %d)" %
              (gimg.shape[1], gimg.shape[0], rel, secs, ol.ncpu(), stats[0] / secs, rc[0], rc[1], rc[2], stats[0], stats[1], stats[2]), flush=True)
        out.update(crop_shape=list(gimg.shape[:2]), crop_rel_rmse=float(rel), oracle_samples_per_s=float(stats[0] / secs), oracle_threads=ol.ncpu(),
                   crop_rays_gpu=[int(x) for x in rc], crop_rays_oracle=[int(x) for x in stats[:3]])
    print(json.dumps(out))
    if a.json:
        with open(a.json, "w") as f:
            json.dump(out, f)
    os._exit(0)


if __name__ == "__main__":
    main()
