"""`--stretch_audio` resynthesis: the description's audio stretched onto the video's clock.

Mirror of the reference's `replace_aligned_segments` (describealign.py:229-416), same signature and the same in-place
effect on `video_arr`.  Per aligned segment the reference either resamples the description with a quadratic
interpolator (small rate differences; scipy `interp1d`, kept as the same library call) or time-stretches it without
changing the pitch (`stretch`, :306-396): the segment is cut into 512-sample windows, for every window and every
candidate jump distance the position with the largest 512-sample Pearson correlation between the signal and itself
`jump` samples away is found, a small dynamic programme over (window, drift) picks where to jump, and the pieces are
cross-faded.  The correlation search is the expensive part - 10 to 482 jump distances times every sample - and is the
part that runs on the GPU here (`best_jumps`, csrc/stretch.cu): one thread per (piece, jump) reproduces the
reference's float64 running sums in its order, so locations and losses are bit-identical
(oracle/stretch_oracle.py is the numpy checker).  The dynamic programme over (window, drift) and its traceback are
native host code (dab_host_stretch_plan, csrc/host_stage.cpp; plan_jumps_numpy is its checker); the cross-fade
assembly is host numpy.
"""
from __future__ import annotations

import ctypes

import numpy as np
import scipy.interpolate
import scipy.signal

AUDIO_SAMPLE_RATE = 44100                  # describealign.py:31
MAX_RATE_RATIO_DIFF_ALIGN = .1             # :33
MIN_DURATION_TO_REPLACE_SECONDS = 2        # :34
JUST_NOTICEABLE_DIFF_IN_FREQ_RATIO = .005  # :35
MIN_STRETCH_OFFSET = 30                    # :36
WINDOW = 512
MAX_DRIFT = 512 * 3
BASE_JUMPS = (506, 451, 284, 410, 480, 379, 308, 430, 265, 494)


def best_jumps(segment: np.ndarray, negative: bool, jumps):
    """Per 512-sample window and jump distance: (best position int16 [windows, jumps], its Pearson correlation
    float64 [windows, jumps]).  Runs on the GPU (dab_stretch_best_jumps)."""
    from . import _cabi, api
    x = np.ascontiguousarray(segment, dtype=np.float16)
    ch, n = x.shape
    jumps = np.ascontiguousarray(list(jumps), dtype=np.int32)
    nw = n // WINDOW
    loc = np.zeros((nw, len(jumps)), dtype=np.int16)
    best = np.full((nw, len(jumps)), -np.inf, dtype=np.float64)
    ctx = api.context()
    ctx.check(_cabi.load().dab_stretch_best_jumps(ctx.handle, x.ctypes.data, int(ch), int(n), int(bool(negative)),
                                                  jumps.ctypes.data, int(len(jumps)), loc.ctypes.data, best.ctypes.data))
    return loc, best


def jump_distances(total_offset_samples: int):
    """The candidate jump distances (describealign.py:311-319): all of them for small or hard-to-reach offsets."""
    jumps = list(BASE_JUMPS)
    if abs(total_offset_samples) < 10000:
        if abs(total_offset_samples) > 1000:
            jumps.extend([MIN_STRETCH_OFFSET + int(o) for o in (2 ** np.arange(8)) - 1])
        else:
            jumps = list(range(MIN_STRETCH_OFFSET, WINDOW))
    return jumps


def plan_jumps(num_input_samples: int, num_output_samples: int, jumps, loc, best):
    """Dynamic programme over (window, drift) and its traceback (describealign.py:320-371): where in the input to
    jump, and by how much.  loc / best as returned by best_jumps.  Returns an int array (k, 2) of
    (input index, signed jump distance).  Native (dab_host_stretch_plan, csrc/host_stage.cpp); plan_jumps_numpy is its
    checker and the path for inputs whose index arithmetic leaves the arrays (it then raises what the reference raises)."""
    from . import _cabi
    nw = num_input_samples // WINDOW
    jl = np.ascontiguousarray(list(jumps), dtype=np.int32)
    loc_c = np.ascontiguousarray(loc, dtype=np.int16)
    best_c = np.ascontiguousarray(best, dtype=np.float64)
    if nw < 2 or loc_c.shape != (nw, len(jl)) or best_c.shape != (nw, len(jl)) or np.isnan(best_c).any():
        return plan_jumps_numpy(num_input_samples, num_output_samples, jumps, loc, best)
    at, dist = np.empty(nw, dtype=np.int64), np.empty(nw, dtype=np.int64)
    count = ctypes.c_int64(0)
    rc = _cabi.load().dab_host_stretch_plan(int(num_input_samples), int(num_output_samples), jl.ctypes.data, len(jl),
                                            loc_c.ctypes.data, best_c.ctypes.data, at.ctypes.data, dist.ctypes.data,
                                            ctypes.byref(count))
    if rc != 0:
        return plan_jumps_numpy(num_input_samples, num_output_samples, jumps, loc, best)
    chosen = np.array(list(zip(at[:count.value].tolist(), dist[:count.value].tolist())))
    if num_output_samples - num_input_samples > 0:
        chosen[:, 1] *= -1
    return chosen


def plan_jumps_numpy(num_input_samples: int, num_output_samples: int, jumps, loc, best):
    """The numpy statement of plan_jumps (describealign.py:320-371): checker of the native code."""
    width = MAX_DRIFT * 2 + 1
    total = num_output_samples - num_input_samples
    nw = num_input_samples // WINDOW

    def offset_at(w):
        return (total * min(nw - 1, max(0, w))) // (nw - 1)

    def offset_step(w):
        return abs(offset_at(w) - offset_at(w - 1))

    back = np.zeros((nw, width), dtype=np.int16)
    cum = np.zeros((3, width)) + np.inf
    cum[1:, MAX_DRIFT] = 0
    cols = np.arange(width)
    last_step = 0
    for w in range(nw):
        losses = 1 - best[w]
        step = offset_step(w)
        step2 = step + last_step
        cand = np.zeros((len(jumps) + 1, width)) + np.inf
        # no jump: the loss at the corresponding drift one window back
        cand[0, :width - step] = cum[(w - 1) % 3, step:]
        for k, jump in enumerate(jumps):
            cut = step2 - jump
            # a jump of this distance from two windows back (one window is skipped so that cross-fades never overlap)
            cand[k + 1, jump:width - max(0, cut)] = cum[(w - 2) % 3, step2:width + min(0, cut)] + losses[k]
        pick = np.argmin(cand, axis=0)
        back[w] = pick
        cum[w % 3] = cand[pick, cols]
        last_step = step
    drift = MAX_DRIFT
    chosen = []
    skip = False
    for w in range(nw - 1, -1, -1):
        drift += offset_step(w + 1)
        if skip:
            skip = False
            continue
        k = back[w, drift] - 1
        if k == -1:
            continue
        jump = jumps[k]
        chosen.append((w * WINDOW + loc[w, k].item(), jump))
        drift -= jump
        skip = True
    chosen = np.array(chosen[::-1])
    # longer output: jump backwards in the input (samples are repeated); shorter: forwards (samples are dropped)
    if total > 0:
        chosen[:, 1] *= -1
    return chosen


def stretch(segment, output):
    """Time-stretch `segment` (channels, n_in) into `output` (channels, n_out) in place, pitch preserved
    (describealign.py:306-396)."""
    n_in, n_out = segment.shape[1], output.shape[1]
    total = n_out - n_in
    jumps = jump_distances(total)
    loc, best = best_jumps(segment, total > 0, jumps)
    chosen = plan_jumps(n_in, n_out, jumps, loc, best)
    at, dist = chosen[:, 0], chosen[:, 1]
    in_starts = np.concatenate(([0], at + dist))
    in_ends = np.concatenate((at, [n_in]))
    out_ends = np.cumsum(in_ends - in_starts)
    out_starts = np.concatenate(([0], out_ends[:-1]))
    bump = scipy.signal.windows.hann(2 * WINDOW + 1)
    rise, fall = bump[:WINDOW], bump[WINDOW:-1]
    output[:, :WINDOW] = segment[:, :WINDOW]
    for i0, i1, o0, o1 in zip(in_starts, in_ends, out_starts, out_ends):
        output[:, o0:o0 + WINDOW] *= fall
        output[:, o0:o0 + WINDOW] += segment[:, i0:i0 + WINDOW] * rise
        output[:, o0 + WINDOW:o1 + WINDOW] = segment[:, i0 + WINDOW:i1 + WINDOW]


def resample_quadratic(audio_desc_arr, samples):
    """The description's waveform at fractional sample positions, by the reference's quadratic interpolator in chunks
    of 1e5 positions (describealign.py:232-245)."""
    chunk = 10 ** 5
    parts = []
    for k in range(0, len(samples), chunk):
        pos = samples[k:k + chunk]
        lo, hi = max(int(pos[0] - 2), 0), min(int(pos[-1] + 2), audio_desc_arr.shape[1])
        f = scipy.interpolate.interp1d(np.arange(lo, hi), audio_desc_arr[:, lo:hi], copy=False, bounds_error=False,
                                       fill_value=0, kind='quadratic', assume_sorted=True)
        parts.append(f(pos).astype(np.float16))
    return np.hstack(parts)


def replace_aligned_segments(video_arr, audio_desc_arr, audio_desc_times, video_times, no_pitch_correction, stretcher=None):
    """Drop-in for describealign.py:229-416: overwrite the aligned runs of `video_arr` (channels, samples, float16)
    with the description's audio brought onto the video's clock.  stretcher: replaces `stretch` (tests)."""
    stretcher = stretch if stretcher is None else stretcher
    x_samples = (audio_desc_times * AUDIO_SAMPLE_RATE).astype(int)
    y_samples = (video_times * AUDIO_SAMPLE_RATE).astype(int)
    dx, dy = np.diff(x_samples), np.diff(y_samples)
    slopes = dx / dy
    offsets = dy - dx
    y_mid = (y_samples[:-1] + y_samples[1:]) // 2
    progress_every = (video_arr.shape[1] // 100) + 1
    shown = -1
    for i in range(len(audio_desc_times) - 1):
        if dy[i] < (MIN_DURATION_TO_REPLACE_SECONDS * AUDIO_SAMPLE_RATE) or np.abs(1 - slopes[i]) > MAX_RATE_RATIO_DIFF_ALIGN:
            continue
        target = video_arr[:, slice(*y_samples[i:i + 2])]
        progress = int(y_mid[i] // progress_every)
        if progress > shown:
            shown = progress
            print(f"  stretching audio:{progress:3d}%                        \r", end='')
        # pitch correction only where the difference would be audible
        if no_pitch_correction or np.abs(1 - slopes[i]) <= JUST_NOTICEABLE_DIFF_IN_FREQ_RATIO or \
           abs(offsets[i]) < MIN_STRETCH_OFFSET:
            points = np.linspace(*x_samples[i:i + 2], num=dy[i], endpoint=False)
            target[:] = resample_quadratic(audio_desc_arr, points)
        else:
            stretcher(audio_desc_arr[:, slice(*x_samples[i:i + 2])], target)
