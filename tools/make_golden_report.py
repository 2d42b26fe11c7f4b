#!/usr/bin/env python
"""Golden text for the report writer / `setts` expression, made by the UNMODIFIED reference.

Authoring-container tool (needs /root/reference).  For every alignment fixture under tests/golden/
(align_<case>.npz: the nodes, similarity and median slope the reference's align() returned) it calls the
reference's own `plot_alignment` (describealign.py:159-227; matplotlib replaced by a mock, the text part is what
is kept) and `encode_fit_as_ffmpeg_expr` (describealign.py:419-435) and writes

    tests/golden/report_<case>.txt     the alignment-dependent lines of the report + the setts expression

usage: python tools/make_golden_report.py
"""
import glob
import os
import sys
import tempfile
from unittest import mock

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_loader import load_reference  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    da = load_reference()
    da.plt = mock.MagicMock()
    for npz in sorted(glob.glob(os.path.join(GOLD, "align_*.npz"))):
        case = os.path.basename(npz)[len("align_"):-len(".npz")]
        d = np.load(npz)
        audio_t, video_t = d["nodes_x"], d["nodes_y"]
        sim, med = float(d["similarity"]), float(d["median_slope"])
        path = d["path2"].copy()
        with tempfile.TemporaryDirectory() as tmp:
            base = os.path.join(tmp, "report")
            da.plot_alignment(base, path, audio_t, video_t, sim, med, False, False, "FFMPEG-COMMAND-PLACEHOLDER")
            with open(base + ".txt") as f:
                lines = f.read().splitlines()
        first = next(k for k, ln in enumerate(lines) if ln.startswith("Input file similarity"))
        last = max(k for k, ln in enumerate(lines) if ln.startswith("Rate change of"))
        keep = lines[first:last + 1]
        video_offset = video_t[0] - audio_t[0]
        setts = da.encode_fit_as_ffmpeg_expr(audio_t, video_t, video_offset)
        out = os.path.join(GOLD, f"report_{case}.txt")
        with open(out, "w") as f:
            f.write("\n".join(keep) + "\n")
            f.write("setts: " + setts + "\n")
        print(out, len(keep), "lines")


if __name__ == "__main__":
    main()
