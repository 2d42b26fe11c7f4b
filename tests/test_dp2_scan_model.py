"""The scan formulation of the pass-2 DP (dp2_scan_kernel) as stated by tools/dp2_scan_model.py,
against the one-point-at-a-time rules and the oracle on the CPU: identical back pointers and path
for every block size, on inputs built to hit the rare branches (NEAR / GAP points, followers,
restarts from a cluster best, binade crossings).  The CUDA kernel itself is checked in
test_gpu_parity.py."""
import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _stage_b_case(seed, **kw):
    spec = importlib.util.spec_from_file_location("_gpu_parity_cases", os.path.join(ROOT, "tests", "test_gpu_parity.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod._random_stage_b_case(seed, **kw)


@pytest.mark.parametrize("seed,kw,nb", [(3, dict(crossing=False), 1024), (4, dict(crossing=True), 64),
                                        (6, dict(crossing=False, n_a=14000, n_v=14000, n_cor=8), 1024)])
def test_scan_model_equals_sequential_rules_and_oracle(seed, kw, nb):
    import dp2_scan_model as S
    M = S
    from oracle import align_oracle as ao
    audio, video, plans, n_clusters = _stage_b_case(seed, **kw)
    pi, pj, pc, pq = ao.score_corridors(plans, audio, video)
    pk, cell, ro, flags = M.point_flags(plans, pi, pj, pc)
    b_seq, s_seq, _ = M.run_scalar(plans, pi, pj, pq, pk, ro, flags)
    b_scan, s_scan, cnt = S.run_scan(plans, pi, pj, pq, pk, ro, flags, NB=nb)
    assert b_seq == b_scan and s_seq.top == s_scan.top
    assert cnt["block_points"] + cnt["scalar_points"] + cnt["hard_points"] == len(pi)
    assert cnt["block_points"] > 0.5 * len(pi)
    want = ao.stage_b(plans, n_clusters, audio, video)["path"]
    path, p = [], s_scan.top[2]
    while p >= 0:
        path.append(p)
        p = b_scan[p][1]
    path.reverse()
    assert len(path) == len(want)
    np.testing.assert_array_equal(pi[path], want[:, 1].astype(np.int32))
    np.testing.assert_array_equal(pj[path], want[:, 0])
