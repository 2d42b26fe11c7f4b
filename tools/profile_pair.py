#!/usr/bin/env python
"""Run one synthetic pair through the CUDA path a few times (the command ncu wraps).

    python tools/profile_pair.py [--config C2] [--scale 1.0] [--reps 2] [--seed 0]

The generated PCM is cached under /tmp so that several ncu invocations in one gpurun call do not
regenerate it.  Prints per-kernel device times of the last repetition.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def cached_pair(config, seed, scale):
    from describealign_b200 import synth
    path = f"/tmp/dab_pair_{config}_{seed}_{scale}.npz"
    if os.path.isfile(path):
        d = np.load(path)
        return d["v"], d["a"]
    v, a = synth.config_pair(config, seed, scale)
    try:
        np.savez(path, v=v, a=a)
    except OSError:
        pass
    return v, a


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--dp2-impl", type=int, default=0)
    args = ap.parse_args()
    from describealign_b200 import api, build
    build.build()
    v, a = cached_pair(args.config, args.seed, args.scale)
    api.context().set_option("dp2_impl", args.dp2_impl)
    det = {}
    for _ in range(args.reps):
        det = {}
        api.align_pcm(v, a, details=det)
    print(json.dumps({"timings_ms": det["timings"], "stats": det["stats"]}))


if __name__ == "__main__":
    main()
