#!/usr/bin/env python
"""Time the jump search of --stretch_audio on the GPU next to its numpy restatement.

    python tools/bench_stretch.py [--minutes 5] [--channels 2]

One segment of the given length (float16, programme-like), stretched by 3 % (the 10 base jump distances) and by
700 samples (all 482 distances); prints one JSON object: seconds on the GPU (CUDA path incl. the host <-> device
copies of the call) and for the numpy oracle on one host core, plus the equality of the two results."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--minutes", type=float, default=5.0)
    ap.add_argument("--channels", type=int, default=2)
    ap.add_argument("--all-jumps-seconds", type=float, default=20.0)
    args = ap.parse_args()
    from describealign_b200 import build, stretch as st, synth
    from oracle import stretch_oracle as so
    build.build()
    v, _ = synth.make_pair(args.minutes * 60.0, 1.0, seed=3, ch=args.channels)
    x = (synth.as_reference_input(v).astype(np.float32) / 1.7).astype(np.float16)
    out = {"segment_minutes": args.minutes, "channels": args.channels, "cases": {}}
    for name, seg, jumps in (("base_10_jumps", x, st.jump_distances(50000)),
                             ("all_482_jumps", x[:, :int(args.all_jumps_seconds * 44100)], st.jump_distances(700))):
        st.best_jumps(seg[:, :4000], True, jumps)          # context, allocator warm-up
        t = time.perf_counter()
        loc, best = st.best_jumps(seg, True, jumps)
        t_gpu = time.perf_counter() - t
        t = time.perf_counter()
        oloc, obest = so.best_jumps(seg, True, jumps)
        t_cpu = time.perf_counter() - t
        out["cases"][name] = {"samples": int(seg.shape[1]), "jumps": len(jumps), "windows": int(seg.shape[1] // 512),
                              "gpu_s": t_gpu, "numpy_oracle_s_one_core": t_cpu, "speedup": t_cpu / t_gpu,
                              "identical": bool(np.array_equal(loc, oloc) and np.array_equal(best.view(np.int64), obest.view(np.int64)))}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
