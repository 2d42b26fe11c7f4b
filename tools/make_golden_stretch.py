#!/usr/bin/env python
"""Golden hashes for --stretch_audio resynthesis, made by the UNMODIFIED reference.

Authoring-container tool (needs /root/reference).  Runs the reference's own `replace_aligned_segments`
(describealign.py:229-416) on seeded synthetic tracks with hand-made alignment nodes that exercise the quadratic
resampling branch and both directions of the pitch-preserving stretch, and records the SHA-256 of the resulting
float16 array in tests/golden/stretch_ref.json.

usage: python tools/make_golden_stretch.py
"""
import contextlib
import hashlib
import io
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from describealign_b200 import synth  # noqa: E402
from oracle.ref_loader import load_reference  # noqa: E402

CASES = {
    # name: (make_pair kwargs, audio node times, video node times, no_pitch_correction)
    "mono_three_segments": (dict(video_s=40.0, offset_s=3.0, seed=5, ch=1), [3.0, 13.0, 23.5, 33.0], [0.0, 10.02, 20.0, 30.2], False),
    "stereo_three_segments": (dict(video_s=40.0, offset_s=3.0, seed=6, ch=2), [3.0, 13.0, 23.5, 33.0], [0.0, 10.02, 20.0, 30.2], False),
    "mono_small_offsets": (dict(video_s=30.0, offset_s=2.0, seed=7, ch=1), [2.0, 8.0, 14.03, 26.0], [0.0, 6.04, 12.0, 23.7], False),
    "mono_all_jump_distances": (dict(video_s=12.0, offset_s=2.0, seed=9, ch=1), [2.0, 4.482, 7.002, 10.0], [0.0, 2.5, 5.0, 8.0], False),
    "mono_no_pitch_correction": (dict(video_s=30.0, offset_s=2.0, seed=8, ch=1), [2.0, 14.0, 26.0], [0.0, 12.5, 23.9], True),
}


def case_arrays(kw):
    v, a = synth.make_pair(**kw)
    return synth.as_reference_input(v).copy(), synth.as_reference_input(a).copy()


def main():
    da = load_reference()
    out = {"reference": "julbean/describealign v2.0.8 replace_aligned_segments", "cases": {}}
    for name, (kw, xt, yt, npc) in CASES.items():
        va, aa = case_arrays(kw)
        before = va.copy()
        with contextlib.redirect_stdout(io.StringIO()):
            da.replace_aligned_segments(va, aa, np.array(xt), np.array(yt), npc)
        out["cases"][name] = {"make_pair": kw, "audio_times": xt, "video_times": yt, "no_pitch_correction": npc,
                              "samples_changed": int(np.sum(va != before)),
                              "sha256": hashlib.sha256(np.ascontiguousarray(va).tobytes()).hexdigest()}
        print(name, out["cases"][name]["samples_changed"], out["cases"][name]["sha256"][:16])
    with open(os.path.join(ROOT, "tests", "golden", "stretch_ref.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
