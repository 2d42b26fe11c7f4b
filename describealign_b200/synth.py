"""Synthetic PCM pairs (video soundtrack + audio-description track).

The reference ships no test media, so every parity test and the
benchmark run on generated PCM.  The generator follows the requirements collected in
SURVEY.md appendix C: three band-limited noise bands, each modulated by a fast and a
slow envelope whose peaks are bounded and whose dips go downward (the pass-2 energy gate
of the reference, describealign.py:934, rejects heavy-upper-tail audio); the description
is an independent intro followed by the video audio with segments inserted or removed,
with louder "narration" mixed over part of the time.

Everything here uses integer arithmetic or plain IEEE + - * / only (no exp/log/FFT and no
BLAS), so the same seed yields the same int16 samples on any x86-64 host.  That is what
lets golden fixtures made in the authoring container be replayed on the GPU box.

Layout of the returned arrays: int16, shape (S, ch), i.e. the interleaved s16le stream
ffmpeg would hand to the reference (describealign.py:152-156).  `as_reference_input`
turns it into the float16 (ch, S) array the reference functions take.
"""
from __future__ import annotations

import numpy as np

SAMPLE_RATE = 44100
FRAME = 210  # samples per feature frame (210 Hz frame rate)


def as_reference_input(pcm: np.ndarray) -> np.ndarray:
    """int16 (S, ch) -> float16 (ch, S) view, exactly as describealign.py:156 builds it."""
    ch = pcm.shape[1]
    return np.ascontiguousarray(pcm).reshape(-1).astype(np.float16).reshape((-1, ch)).T


def _white(rng: np.random.Generator, n: int) -> np.ndarray:
    w = rng.integers(-32768, 32768, size=n, dtype=np.int32)
    w += rng.integers(-32768, 32768, size=n, dtype=np.int32)
    return w


def _boxcar(x: np.ndarray, m: int) -> np.ndarray:
    """Centred moving SUM of length m (integer exact), same length, zero padded."""
    n = len(x)
    left = m // 2
    c = np.zeros(n + m + 1, dtype=np.int64)
    np.cumsum(x, dtype=np.int64, out=c[left + 1:left + 1 + n])
    c[left + 1 + n:] = c[left + n]
    return c[m:m + n] - c[:n]


def _smooth_unit_noise(rng: np.random.Generator, n: int, step: int, width: int) -> np.ndarray:
    """Smooth, roughly unit-variance Gaussian-like process sampled every sample.

    White integer noise every `step` samples, double boxcar of `width` knots, linear
    interpolation up to the sample rate.  Variance normalised analytically.
    """
    k = n // step + 3
    w = _white(rng, k)
    s = _boxcar(_boxcar(w, width), width).astype(np.float64)
    # var(w) = 2 * (65536^2 - 1) / 12 ; double boxcar = triangular kernel of 2*width-1 taps
    tri = np.convolve(np.ones(width), np.ones(width))
    var = (2.0 * (65536.0 ** 2 - 1.0) / 12.0) * float(np.sum(tri * tri))
    s /= np.sqrt(var)
    f = np.arange(step, dtype=np.float64) / float(step)
    up = s[:-1, None] * (1.0 - f)[None, :] + s[1:, None] * f[None, :]
    return up.reshape(-1)[:n]


def _envelope(g: np.ndarray) -> np.ndarray:
    """~exp(-1.5|g|) with plain arithmetic: bounded by 1 above, dips downward."""
    a = 1.5 * np.abs(g)
    return 1.0 / (1.0 + a + 0.5 * a * a)


def _programme(rng: np.random.Generator, n: int) -> np.ndarray:
    """One channel of 'programme' audio, float64, RMS about 3000."""
    out = np.zeros(n, dtype=np.float64)
    # (short boxcar, long boxcar, gain): difference of moving averages = crude band-pass
    bands = ((73, 551, 1.0), (18, 73, 0.8), (5, 18, 0.6))
    slow = _envelope(_smooth_unit_noise(rng, n, 4410, 5))       # ~0.5 Hz
    for short, long_, gain in bands:
        w = _white(rng, n)
        b = _boxcar(w, short).astype(np.float64) / short - _boxcar(w, long_).astype(np.float64) / long_
        # white variance 2*(65536^2-1)/12 ; band variance ~ var*(1/short - 1/long)
        sd = np.sqrt((2.0 * (65536.0 ** 2 - 1.0) / 12.0) * (1.0 / short - 1.0 / long_))
        fast = _envelope(_smooth_unit_noise(rng, n, 441, 6))     # ~8 Hz
        out += (gain * 5200.0 / sd) * b * fast * slow
    return out


def _to_int16(x: np.ndarray) -> np.ndarray:
    return np.clip(np.rint(x), -32768, 32767).astype(np.int16)


def _resample_linear(x: np.ndarray, ratio: float) -> np.ndarray:
    """Stretch x to round(len*ratio) samples by linear interpolation (plain arithmetic)."""
    m = int(round(len(x) * ratio))
    if m <= 1 or len(x) < 2:
        return x[:m].copy()
    t = np.arange(m, dtype=np.float64) * ((len(x) - 1) / float(m - 1))
    i0 = np.minimum(t.astype(np.int64), len(x) - 2)
    f = t - i0
    return x[i0] * (1.0 - f) + x[i0 + 1] * f


def make_pair(video_s: float, offset_s: float, skips=(), seed: int = 0, ch: int = 1,
              tail_s: float = 0.0, narration_frac: float = 0.3, warps=()):
    """Build one (video, description) PCM pair.

    video_s      length of the video soundtrack in seconds
    offset_s     length of the independent intro at the start of the description
    skips        iterable of (video_time_s, delta_s): delta > 0 inserts delta seconds of
                 unrelated audio into the description at that video time, delta < 0 removes
                 |delta| seconds of the video audio from the description
    warps        iterable of (video_t0_s, video_t1_s, ratio): that span of the video audio
                 plays `ratio` times slower in the description
    tail_s       unrelated audio appended to the description
    Returns (video int16 (Sv, ch), description int16 (Sa, ch)).
    """
    rng = np.random.default_rng(np.random.PCG64(seed))
    sr = SAMPLE_RATE
    nv = int(round(video_s * sr))
    vid = [_programme(rng, nv) for _ in range(ch)]
    if ch == 2:
        mid, side = vid
        vid = [mid + 0.3 * side, mid - 0.3 * side]

    # cut list over the video timeline, in order
    events = sorted([(float(t), "skip", float(d)) for t, d in skips] +
                    [(float(t0), "warp", (float(t1), float(r))) for t0, t1, r in warps])
    pieces = [[] for _ in range(ch)]

    def extra(n):
        e = [_programme(rng, n) for _ in range(ch)]
        if ch == 2:
            e = [e[0] + 0.3 * e[1], e[0] - 0.3 * e[1]]
        return e

    intro = extra(int(round(offset_s * sr)))
    for c in range(ch):
        pieces[c].append(intro[c])
    pos = 0
    for t, kind, arg in events:
        at = min(max(int(round(t * sr)), pos), nv)
        for c in range(ch):
            pieces[c].append(vid[c][pos:at])
        pos = at
        if kind == "skip":
            if arg > 0:
                ins = extra(int(round(arg * sr)))
                for c in range(ch):
                    pieces[c].append(ins[c])
            else:
                pos = min(nv, pos + int(round(-arg * sr)))
        else:
            t1, ratio = arg
            end = min(nv, max(pos, int(round(t1 * sr))))
            for c in range(ch):
                pieces[c].append(_resample_linear(vid[c][pos:end], ratio))
            pos = end
    for c in range(ch):
        pieces[c].append(vid[c][pos:])
    if tail_s > 0:
        tail = extra(int(round(tail_s * sr)))
        for c in range(ch):
            pieces[c].append(tail[c])
    desc = [np.concatenate(p) for p in pieces]
    na = len(desc[0])

    if narration_frac > 0:
        # narration: an independent louder signal gated on for ~narration_frac of the time
        gate_src = _smooth_unit_noise(rng, na, 44100, 3)
        # P(g > thr) ~ narration_frac for a unit Gaussian; fixed thresholds avoid erfinv
        thr = {0.3: 0.5244, 0.2: 0.8416, 0.5: 0.0}.get(round(narration_frac, 2), 0.5244)
        gate = np.clip((gate_src - thr) * 4.0, 0.0, 1.0)
        voice = _programme(rng, na) * 1.6
        for c in range(ch):
            desc[c] = desc[c] + gate * voice

    video_pcm = np.stack([_to_int16(v) for v in vid], axis=1)
    desc_pcm = np.stack([_to_int16(d) for d in desc], axis=1)
    return video_pcm, desc_pcm


# ---------------------------------------------------------------------------------------
# Named configurations.  Durations in seconds.
# ---------------------------------------------------------------------------------------

def config_pair(name: str, seed: int = 0, scale: float = 1.0):
    """Return (video_pcm, desc_pcm) for a named BASELINE configuration.

    scale < 1 shrinks every duration proportionally (for quick tests).
    """
    s = float(scale)
    if name == "C1":      # trimmed Ask Dad stand-in
        return make_pair(179 * s, 201.81 * s, skips=[(37 * s, 3 * s)], seed=seed)
    if name in ("C2", "C3"):   # Ask Dad full shape; C3 = stereo (--stretch_audio semantics)
        rng = np.random.default_rng(1000 + seed)
        ts = np.sort(rng.uniform(600, 1250, size=10))
        skips = [(float(t) * s, 3.0 * s) for t in ts]
        return make_pair(1320 * s, 202 * s, skips=skips, seed=seed, tail_s=68 * s,
                         ch=2 if name == "C3" else 1)
    if name == "C4":      # one 45-min episode of the batch config
        rng = np.random.default_rng(2000 + seed)
        k = int(rng.integers(4, 9))
        ts = np.sort(rng.uniform(120, 2600, size=k))
        ds = rng.uniform(1.5, 6.0, size=k) * rng.choice([-1.0, 1.0], size=k)
        off = float(rng.uniform(10, 90))
        return make_pair(2700 * s, off * s, skips=[(float(t) * s, float(d) * s) for t, d in zip(ts, ds)],
                         seed=seed)
    if name == "C5":      # long-form
        rng = np.random.default_rng(3000 + seed)
        ts = np.sort(rng.uniform(300, 8800, size=12))
        ds = rng.uniform(2.0, 8.0, size=12) * rng.choice([-1.0, 1.0], size=12)
        extra = 10800 - 9000 - 300 - float(np.sum(ds))
        return make_pair(9000 * s, 300 * s, skips=[(float(t) * s, float(d) * s) for t, d in zip(ts, ds)],
                         seed=seed, tail_s=max(0.0, extra) * s)
    raise ValueError(f"unknown config {name!r}")


# ---------------------------------------------------------------------------------------
# Long-form pair built from independently generated segments (C5 shape, quickly)
# ---------------------------------------------------------------------------------------

def _segment(arg):
    video_s, offset_s, skips, seed, tail_s = arg
    return make_pair(video_s, offset_s, skips=skips, seed=seed, tail_s=tail_s)


def long_pair(seed: int = 0, scale: float = 1.0, workers: int = 0, segments: int = 10):
    """The C5 shape (2.5-h film vs 3-h description, 300 s start offset, ~12 skips) as `segments` independently
    generated pieces laid end to end, so that the pieces can be made by a pool of worker processes (config_pair("C5")
    is one sequential pass, ~4 min at full scale).  Piece k is make_pair(video 9000/segments s, intro, one inner skip):
    the intro of piece 0 is the 300 s start offset, the intros of the later pieces act as inserted unrelated audio at
    the joins, the last piece carries the tail that brings the description to 10800 s.  Mono; deterministic in
    (seed, scale, segments)."""
    s = float(scale)
    rng = np.random.default_rng(4000 + seed)
    seg_v = 9000.0 / segments
    intros = [300.0] + [float(x) for x in rng.uniform(2.0, 8.0, size=segments - 1)]
    inner_t = rng.uniform(0.2 * seg_v, 0.8 * seg_v, size=segments)
    inner_d = rng.uniform(2.0, 8.0, size=segments) * rng.choice([-1.0, 1.0], size=segments)
    has_inner = rng.uniform(size=segments) < 0.3
    total = sum(intros) + 9000.0 + float(np.sum(inner_d[has_inner]))
    tail = max(0.0, 10800.0 - total)
    args = []
    for k in range(segments):
        skips = [(float(inner_t[k]) * s, float(inner_d[k]) * s)] if has_inner[k] else []
        args.append((seg_v * s, intros[k] * s, skips, 100003 * (seed + 1) + k, tail * s if k == segments - 1 else 0.0))
    if workers and workers > 1:
        from concurrent.futures import ProcessPoolExecutor
        with ProcessPoolExecutor(max_workers=min(workers, segments)) as ex:
            parts = list(ex.map(_segment, args))
    else:
        parts = [_segment(a) for a in args]
    video = np.concatenate([p[0] for p in parts], axis=0)
    desc = np.concatenate([p[1] for p in parts], axis=0)
    return video, desc
