"""Cost of a textured Kd: full-size C3 with constant materials vs the same scene with a checkerboard on the
ground quad and on the plastic sphere (the KM_TEX instantiations of the matte / plastic shade kernels + 48 B of camera-ray
differentials per path), median of 5 renders each, plus parity of the textured render's centre crop against the oracle.
Usage (GPU box): python tools/r2_textured_cost.py [--spp 64]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ap = argparse.ArgumentParser()
ap.add_argument("--spp", type=int, default=64)
a = ap.parse_args()
import __graft_entry__ as ge
pkg = ge.load_package()
pkg.init(0)
import torch
from pbrt_v3_rs_b200 import workloads as wl


def timed(sd):
    integ = pkg.PathIntegrator(sd)
    integ.render()
    ts = []
    for _ in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        integ.render()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts)), ts, integ


out = {"spp": a.spp}
sd = wl.scene_c3(spp=a.spp)
out["constant_ms"], out["constant_all_ms"], _ = timed(sd)
sd = wl.scene_c3(spp=a.spp)
t_ground = sd.add_spectrum_texture("checkerboard", uscale=48.0, vscale=48.0, tex1=(0.1, 0.1, 0.1), tex2=(0.8, 0.8, 0.8))
t_ball = sd.add_spectrum_texture("checkerboard", uscale=16.0, vscale=16.0, tex1=(0.9, 0.2, 0.1), tex2=(0.1, 0.3, 0.9))
sd.materials[4]["Kd"] = ("texture", t_ground)  # the ground quad's matte
sd.materials[1]["Kd"] = ("texture", t_ball)    # the plastic sphere
out["textured_ms"], out["textured_all_ms"], integ = timed(sd)
out["ratio"] = out["textured_ms"] / out["constant_ms"]
import oracle_lib
ps = np.random.default_rng(1).integers(0, [1920, 1080, a.spp], size=(2048, 3)).astype(np.int32)
li, _ = integ.li(ps)
oli = oracle_lib.OracleScene(sd).li(ps)
out["li_bit_identical_frac"] = float((li.view(np.uint32) == oli.view(np.uint32)).all(1).mean())
out["li_close_frac"] = float(np.isclose(li, oli, rtol=2e-3, atol=1e-5).all(1).mean())
print(json.dumps(out))
