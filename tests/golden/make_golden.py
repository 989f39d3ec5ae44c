"""Regenerates the frozen oracle fixtures under tests/golden/ (run from the repo root).
The reference cannot be built here (no Rust toolchain), so these are frozen ORACLE outputs:
they detect drift, they do not pin the oracle to the reference."""
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402

ge.build_oracle()
ge.load_package()
import oracle_lib as ol  # noqa: E402
from pbrt_v3_rs_b200 import workloads as wl  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
L = ol.lib()
perm = np.zeros(L.orc_prime_total(), dtype=np.uint16)
L.orc_halton_permutations(perm.ctypes.data_as(C.c_void_p))
json.dump({"sha256": hashlib.sha256(perm.tobytes()).hexdigest(), "n": int(perm.size)}, open(os.path.join(HERE, "halton_perm_sha256.json"), "w"))

tv = wl.c2_mesh(wl.C2_SMALL)
nodes, ordered = ol.build_bvh_sah(ol.triangle_bounds(tv), 4)
json.dump({"sha256": hashlib.sha256(nodes.tobytes() + ordered.tobytes()).hexdigest(), "n_nodes": int(len(nodes))},
          open(os.path.join(HERE, "bvh_c2_small_sha256.json"), "w"))

# fixed ray set: closest-hit ids/t and any-hit flags of the small microbench
acc = ol.OracleAccel(nodes, ordered, tv)
cfg = wl.C2_SMALL
rays = wl.primary_rays(cfg["width"], cfg["height"])
hits, diag, ct = acc.intersect(rays)
br = wl.bounce_rays(tv, rays, hits, rays.shape[0])
bh, _, bct = acc.intersect(br)
sr = wl.shadow_rays(br)
occ, _ = acc.occluded(sr)
np.savez_compressed(os.path.join(HERE, "c2_small_hits.npz"), primary_prim=hits["prim"], primary_t=hits["t"], bounce_prim=bh["prim"],
                    bounce_t=bh["t"], occluded=np.packbits(occ), counters=np.stack([ct.sum(0), bct.sum(0)]))
print("golden fixtures written")

# per-sample radiances (Integrator::li) of one small scene for every integrator / sampler / light-strategy combination,
# plus the HLBVH tree of the small mesh: frozen so that a change of the oracle AND the device in the same direction is seen
import scenes_small as ss  # noqa: E402

LI_CASES = ss.LI_GOLDEN_CASES
li = {}
for name, kw in LI_CASES.items():
    sd = ss.li_golden_scene(wl, **kw)
    ps = ss.li_golden_pairs()
    li[name] = ol.OracleScene(sd).li(ps, nthreads=1)
np.savez_compressed(os.path.join(HERE, "li_small.npz"), **li)
hn, ho = ol.build_bvh_hlbvh(ol.triangle_bounds(tv), 4)
json.dump({"sha256": hashlib.sha256(hn.tobytes() + ho.tobytes()).hexdigest(), "n_nodes": int(len(hn))},
          open(os.path.join(HERE, "hlbvh_c2_small_sha256.json"), "w"))
print("integrator fixtures written:", ", ".join(sorted(li)))
