#!/usr/bin/env python
"""Host-side cost of every C-ABI call for one pair (single thread), with results in pinned or
pageable memory.  Diagnostic for the batch path; run on a GPU box."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    from describealign_b200 import _cabi, api, build
    from profile_pair import cached_pair
    build.build()
    v, a = cached_pair("C2", 0, 1.0)
    import torch
    dv, da = torch.from_numpy(v).cuda(), torch.from_numpy(a).cuda()
    out = {}
    for rep in range(4):
        job = api.AlignJob()
        t0 = time.perf_counter()
        job.load_pcm_device((dv.data_ptr(), dv.shape[0], dv.shape[1]), (da.data_ptr(), da.shape[0], da.shape[1]))
        t1 = time.perf_counter()
        job.device_stage_a()
        t2 = time.perf_counter()
        job.host_stage()
        t3 = time.perf_counter()
        job.device_stage_b()
        t4 = time.perf_counter()
        out[f"rep{rep}"] = {"load_ms": 1e3 * (t1 - t0), "stage_a_total_ms": 1e3 * (t2 - t1), "host_fit_ms": 1e3 * (t3 - t2),
                            "stage_b_total_ms": 1e3 * (t4 - t3), "calls": dict(job.host_ms), "kernels": job.pair.timings()}
        job.close()
    # raw allocator cost
    t0 = time.perf_counter()
    arrs = [_cabi.pinned_empty(2_600_000, np.float64) for _ in range(8)]
    t1 = time.perf_counter()
    del arrs
    t2 = time.perf_counter()
    arrs = [_cabi.pinned_empty(2_600_000, np.float64) for _ in range(8)]
    t3 = time.perf_counter()
    out["pinned_alloc_ms"] = {"first_8x20MB": 1e3 * (t1 - t0), "free": 1e3 * (t2 - t1), "recycled_8x20MB": 1e3 * (t3 - t2)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
