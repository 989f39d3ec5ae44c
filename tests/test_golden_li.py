"""Frozen per-sample radiances (tests/golden/li_small.npz, written by tests/golden/make_golden.py from the oracle): every
integrator / sampler / light-strategy combination on one small glass + matte scene.  The oracle must reproduce them bit for
bit (CPU), the device within the parity tolerance (GPU): a change that moves oracle and device together is still caught."""
import hashlib
import json
import os

import numpy as np
import pytest

import scenes_small as ss

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("case", sorted(ss.LI_GOLDEN_CASES))
def test_oracle_reproduces_frozen_radiances(pkg, oracle, case):
    from pbrt_v3_rs_b200 import workloads as wl
    gold = np.load(os.path.join(GOLD, "li_small.npz"))[case]
    li = oracle.OracleScene(ss.li_golden_scene(wl, **ss.LI_GOLDEN_CASES[case])).li(ss.li_golden_pairs(), nthreads=1)
    assert li.shape == gold.shape and li.mean() > 0
    assert np.array_equal(li.view(np.uint32), gold.view(np.uint32))


def test_hlbvh_tree_hash_is_frozen(pkg, oracle):
    from pbrt_v3_rs_b200 import workloads as wl
    tv = wl.c2_mesh(wl.C2_SMALL)
    pb = pkg.triangle_bounds(tv)
    gold = json.load(open(os.path.join(GOLD, "hlbvh_c2_small_sha256.json")))
    for build in (oracle.build_bvh_hlbvh, pkg.build_bvh_hlbvh):
        n, o = build(pb, 4)
        assert len(n) == gold["n_nodes"] and hashlib.sha256(n.tobytes() + o.tobytes()).hexdigest() == gold["sha256"]


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(ss.LI_GOLDEN_CASES))
def test_device_matches_frozen_radiances(gpu, case):
    from pbrt_v3_rs_b200 import workloads as wl
    gold = np.load(os.path.join(GOLD, "li_small.npz"))[case]
    li, _ = gpu.PathIntegrator(ss.li_golden_scene(wl, **ss.LI_GOLDEN_CASES[case])).li(ss.li_golden_pairs())
    close = np.isclose(li, gold, rtol=2e-3, atol=1e-5).all(1)
    assert close.mean() >= 0.97, "only %.4f of the samples agree with the frozen oracle values" % close.mean()
    assert abs(li.mean() - gold.mean()) <= 0.02 * gold.mean()
