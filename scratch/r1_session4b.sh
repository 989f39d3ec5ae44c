python tools/run_config.py c3 --integrator whitted --li 2048 --crop 0.1 --json gpurun_out/r1_c3_whitted.json 2>&1 | grep "^render\|parity\|crop"
python tools/run_config.py c3 --integrator directlighting --strategy all --li 2048 --crop 0.1 --json gpurun_out/r1_c3_direct_all.json 2>&1 | grep "^render\|parity\|crop"
python - <<'PY'
import sys, json, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_package(); pkg.init(0)
from pbrt_v3_rs_b200 import workloads as wl
import oracle_lib as ol, scenes_small as ss
out = {}
for name, upd in (("spatial", dict(lightsamplestrategy="spatial")), ("sobol", dict())):
    sd = wl.scene_c3()
    sd.integrator.update(**upd)
    if name == "sobol": sd.sampler.update(type="sobol")
    integ = pkg.PathIntegrator(sd); integ.preprocess()
    best = 1e9
    for it in range(3):
        t0 = time.time(); film = integ.render_rows(); best = min(best, time.time() - t0)
    rc = integ.ray_counts()
    sd.film["cropwindow"] = (0.45, 0.55, 0.45, 0.55)
    g = pkg.PathIntegrator(sd).render()
    ref, stats, secs = ol.OracleScene(sd).render()
    out[name] = dict(render_s=best, samples_per_s=rc[0] / best, crop_rel_rmse=ss.rel_rmse(g, ref), rays=[int(x) for x in rc])
    print(name, out[name], flush=True)
json.dump(out, open("gpurun_out/r1_c3_spatial_sobol.json", "w"))
PY
