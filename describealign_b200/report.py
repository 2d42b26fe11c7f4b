"""What the user reads of an alignment: the text report and the ffmpeg `setts` expression.

SURVEY.md section 8(f) N4.  The reference writes, next to every output file, a report whose
alignment-dependent lines are the similarity, the start offset, the median rate change and one
"Rate change of ..." line per segment between break-point nodes (describealign.py:211-224), and in
the default (video-stretch) mode hands ffmpeg a piece-wise linear timestamp expression built from the
same nodes (`encode_fit_as_ffmpeg_expr`, describealign.py:419-435).  The launcher leaves both to the
reference's own code (they are callers of the path, not part of it); these restatements exist so
that the parity surface a user sees - `.2f` percentages, `h:mm:ss.mmm` times, `.4f` clip bounds and
`.9f` slopes - is pinned by golden text files produced by the reference itself
(tools/make_golden_report.py, tests/golden/report_*.txt).

Inputs are what align() returns: audio_desc_times, video_times (seconds, float64), the similarity
percentage and the median slope.
"""
from __future__ import annotations

import numpy as np


def str_from_time(seconds) -> str:
    """`h:mm:ss.mmm` with the hour right-aligned in two columns (describealign.py:217-220)."""
    minutes, seconds = divmod(seconds, 60)
    hours, minutes = divmod(minutes, 60)
    return f"{hours:2.0f}:{minutes:02.0f}:{seconds:06.3f}"


def start_offset(audio_desc_times, video_times) -> float:
    """Seconds the description starts after the video (describealign.py:211, 1160): the report prints its negative."""
    return video_times[0] - audio_desc_times[0]


def report_lines(audio_desc_times, video_times, similarity_percent, median_slope) -> list[str]:
    """The alignment-dependent lines of the text report, in order (describealign.py:212-224)."""
    video_offset = start_offset(audio_desc_times, video_times)
    lines = [f"Input file similarity: {similarity_percent:.2f}%",
             "Main changes needed to video to align it to audio input:",
             f"Start Offset: {-video_offset:.2f} seconds",
             f"Median Rate Change: {(median_slope - 1.) * 100:.2f}%"]
    for i in range(len(video_times) - 1):
        slope = (video_times[i + 1] - video_times[i]) / (audio_desc_times[i + 1] - audio_desc_times[i])
        lines.append(f"Rate change of {(slope - 1.) * 100:8.1f}% from {str_from_time(video_times[i])} to "
                     f"{str_from_time(video_times[i + 1])} aligning with audio from "
                     f"{str_from_time(audio_desc_times[i])} to {str_from_time(audio_desc_times[i + 1])}")
    return lines


def setts_expression(audio_desc_times, video_times, video_offset=None) -> str:
    """ffmpeg `setts` expression that moves video frame timestamps onto the description's clock: one clip()
    term per segment, offsets with 4 decimals, rate corrections with 9 (describealign.py:419-435)."""
    x = np.asarray(audio_desc_times, dtype=np.float64)
    y = np.asarray(video_times, dtype=np.float64)
    if video_offset is None:
        video_offset = start_offset(x, y)
    dx, dy = np.diff(x), np.diff(y)
    slopes = dx / dy
    terms = ["TS", "+(0"]
    for i in range(len(x) - 1):
        terms.append(f"+clip(TS-{y[i] - video_offset:.4f}/TB,0,{max(0, dy[i]):.4f}/TB)*{slopes[i] - 1:.9f}")
    terms.append(")")
    return "".join(terms)
