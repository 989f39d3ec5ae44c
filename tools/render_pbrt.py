"""Renders a pbrt-v3 scene file through the CUDA path and writes the image as PFM:

    python tools/render_pbrt.py scene.pbrt [--out image.pfm] [--device 0]

The scene is read by the C++ host loader (b200pt_load_pbrt: the subset of the format listed in include/b200pt.h);
anything outside the path fails with the offending directive named.  The output name defaults to the Film's
"filename" with a .pfm extension."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("scene")
    ap.add_argument("--out", default=None)
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args()
    import __graft_entry__ as ge
    pkg = ge.load_package()
    pkg.init(a.device)
    t0 = time.time()
    ld = pkg.load_pbrt(a.scene)
    d = ld.to_desc()
    print("loaded %s: %d triangles, %d instances, %d lights, %dx%d @ %d spp (%.2f s)" %
          (a.scene, d.n_prims, d.n_instances, d.n_lights, d.film.xres, d.film.yres, d.sampler.spp, time.time() - t0))
    integ = pkg.PathIntegrator(ld)
    integ.preprocess()
    t0 = time.time()
    img = integ.render()
    dt = time.time() - t0
    rc = integ.ray_counts()
    print("rendered in %.3f s: %.3e samples/s, %.1f Mrays/s" % (dt, rc[0] / dt, (rc[1] + rc[2]) / dt / 1e6))
    out = a.out or os.path.splitext(ld.output)[0] + ".pfm"
    pkg.write_pfm(out, img)
    print("wrote", out)


if __name__ == "__main__":
    main()
