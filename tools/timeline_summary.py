#!/usr/bin/env python
"""Summarise a bench.py --timeline dump: per device stage, how long it ran, how long the pair's stream
sat idle before it (host round trip, copy-engine queue, ...), and how many pairs were in it at once.

    python tools/timeline_summary.py gpurun_out/timeline.json > profiles/rN_timeline.txt
"""
import json
import sys

import numpy as np

STAGES = ["features_video", "features_audio", "prep_codes", "tables", "gate", "score", "dp1_trace", "corridors", "dp2_trace"]


def main(path):
    d = json.load(open(path))
    pairs = d["pairs"]
    print(f"# {path}: step {d['step_ms']:.1f} ms, {len(pairs)} pairs, host_input={d['host_input']} (True = end-to-end run, PCM uploaded inside the step)")
    print(f"{'stage':16s} {'dur mean':>9s} {'dur p50':>9s} {'dur max':>9s} | {'idle-before mean':>16s} {'max':>8s} | {'in flight avg':>13s} {'max':>4s}")
    ts = np.linspace(0, d["step_ms"], 400)
    for s in STAGES:
        dur, gap = [], []
        for p in pairs:
            if s not in p:
                continue
            a, b = p[s]
            dur.append(b - a)
            k = STAGES.index(s)
            prev = [p[q][1] for q in STAGES[:k] if q in p]
            if prev:
                gap.append(a - prev[-1])
        conc = [sum(1 for p in pairs if s in p and p[s][0] <= t < p[s][1]) for t in ts]
        print(f"{s:16s} {np.mean(dur):9.2f} {np.median(dur):9.2f} {np.max(dur):9.2f} | {np.mean(gap) if gap else 0:16.2f} {np.max(gap) if gap else 0:8.2f} | {np.mean(conc):13.2f} {np.max(conc):4d}")
    span = [p["slot_host_ms"][1] - p["slot_host_ms"][0] for p in pairs]
    print(f"# host time per pair (slot acquired -> results on host): mean {np.mean(span):.1f} ms, max {np.max(span):.1f} ms")


if __name__ == "__main__":
    main(sys.argv[1])
