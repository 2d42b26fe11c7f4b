"""The block formulation of the pass-2 DP (dp2_block_kernel) as stated by tools/dp2_block_model.py,
against the oracle on the CPU: same back pointers as the one-point-at-a-time rules, same path as
oracle.stage_b.  (The CUDA kernel itself is checked in test_gpu_parity.py.)"""
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


@pytest.mark.parametrize("seed,kw", [(3, dict(crossing=False)), (4, dict(crossing=True)),
                                     (6, dict(crossing=False, n_a=14000, n_v=14000, n_cor=8))])
def test_block_model_equals_sequential_rules_and_oracle(seed, kw):
    import dp2_block_model as M
    from oracle import align_oracle as ao
    audio, video, plans, n_clusters = _stage_b_case(seed, **kw)
    pi, pj, pc, pq = ao.score_corridors(plans, audio, video)
    pk, cell, ro, flags = M.point_flags(plans, pi, pj, pc)
    b_seq, s_seq, _ = M.run_scalar(plans, pi, pj, pq, pk, ro, flags)
    b_blk, s_blk, counters = M.run_blocks(plans, pi, pj, pq, pk, ro, flags, min_block=4)
    assert b_seq == b_blk and s_seq.top == s_blk.top
    assert counters["block_points"] + counters.get("fast_points", 0) + counters["scalar_points"] == len(pi)
    assert counters["block_points"] + counters.get("fast_points", 0) > 0
    # follow the back pointers from the frontier's best entry: the oracle's path
    want = ao.stage_b(plans, n_clusters, audio, video)["path"]
    path, p = [], s_blk.top[2]
    while p >= 0:
        path.append(p)
        p = b_blk[p][1]
    path.reverse()
    assert len(path) == len(want)
    np.testing.assert_array_equal(pi[path], want[:, 1].astype(np.int32))
    np.testing.assert_array_equal(pj[path], want[:, 0])
