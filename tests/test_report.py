"""Report writer / `setts` expression parity: the text a user reads, against golden files
written by the reference's own plot_alignment / encode_fit_as_ffmpeg_expr (tools/make_golden_report.py)."""
import os

import numpy as np
import pytest

from conftest import golden_pair_pcm

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _golden_text(name):
    with open(os.path.join(GOLD, f"report_{name}.txt")) as f:
        lines = f.read().splitlines()
    assert lines[-1].startswith("setts: ")
    return lines[:-1], lines[-1][len("setts: "):]


def _ours(nx, ny, sim, med):
    from describealign_b200 import report
    return report.report_lines(nx, ny, sim, med), report.setts_expression(nx, ny)


@pytest.mark.parametrize("name", ["pair_a", "pair_warp"])
def test_report_text_from_reference_nodes(golden_align, name):
    """The formatting itself: the reference's nodes in, the reference's text out."""
    data, _ = golden_align
    g = data[name]
    want_lines, want_setts = _golden_text(name)
    lines, setts = _ours(g["nodes_x"], g["nodes_y"], float(g["similarity"]), float(g["median_slope"]))
    assert lines == want_lines
    assert setts == want_setts


def test_report_helpers():
    from describealign_b200 import report
    assert report.str_from_time(0.0) == " 0:00:00.000"
    assert report.str_from_time(3725.0625) == " 1:02:05.062"     # round-half-even of the format spec
    assert report.str_from_time(36000 + 59.9996) == "10:00:60.000"  # the reference's own quirk: seconds are not carried
    # a single segment, rate 1: one clip term with a zero correction
    x, y = np.array([10.0, 70.0]), np.array([0.0, 60.0])
    assert report.setts_expression(x, y) == "TS+(0+clip(TS-10.0000/TB,0,60.0000/TB)*0.000000000)"
    assert report.report_lines(x, y, 50.0, 1.0)[2] == "Start Offset: 10.00 seconds"


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["pair_a", "pair_warp"])
def test_report_text_from_the_cuda_path(gpu_ctx, golden_align, name):
    """PCM in -> CUDA alignment -> report text and setts expression identical to the reference's."""
    from describealign_b200 import api
    _, meta = golden_align
    v, a = golden_pair_pcm(meta, name)
    nx, ny, sim, path, med = api.align_pcm(v, a)
    want_lines, want_setts = _golden_text(name)
    lines, setts = _ours(nx, ny, sim, med)
    assert lines == want_lines
    assert setts == want_setts
