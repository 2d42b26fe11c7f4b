"""TEST INFRASTRUCTURE ONLY - oracle for align() (reference describealign.py:595-1027).

stage_a   describealign.py:596-700   prep, digit codes, hash tables, candidate lookup,
                                     correlation scoring, frontier DP #1, traceback
stage_b   describealign.py:895-993   corridor scoring, frontier DP #2, traceback
align     the whole function; the host stage in the middle (describealign.py:702-893, the
          untimed "rate-change fit") is injected by the caller as a module-like object with
          the functions of describealign_b200/host_fit.py, so that the oracle and the CUDA
          path are compared on identical host-stage arithmetic.

The heavy loops live in oracle/c/oracle_align.c.  literal_* functions at the bottom follow
the reference's own data structures step by step (sorted frontier list, dict of back
pointers) in pure Python; they are slow and only used on small inputs to cross-check the
prefix-max restatements used by the C code and the CUDA kernels.
"""
from __future__ import annotations

import bisect
import ctypes

import numpy as np
import scipy.signal

from . import lib

FAILED_MSG = "Alignment failed, are the input files mismatched?"


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _pp(arrs):
    return (ctypes.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


def hann41():
    w = scipy.signal.windows.hann(43)[1:-1]
    return np.ascontiguousarray(w / np.sum(w))


def prep_track(features, is_video: bool):
    """Mean-subtracted features, sliding norms and digit codes of one track."""
    L = lib()
    h = hann41()
    ms, nrm, code, flags = [], [], [], []
    for f in features:
        f64 = np.ascontiguousarray(f, dtype=np.float64)
        n = len(f64)
        m = np.zeros(n)
        r = np.zeros(max(n - 40, 0))
        L.oracle_meansub_norm(_p(f64), n, _p(h), _p(m), _p(r))
        c = np.zeros(max(n - 40, 0), np.int32)
        fl = np.zeros(max(n - 40, 0), np.uint8)
        L.oracle_codes(_p(m), _p(r), n, 1 if is_video else 0, _p(c), _p(fl))
        ms.append(m); nrm.append(r); code.append(c); flags.append(fl)
    return ms, nrm, code, flags


def not_quiet(energy):
    """Indices t < len(energy) - 41 with energy[t] > 0.5 (describealign.py:629, 657)."""
    return np.flatnonzero(np.asarray(energy)[:-41] > .5).astype(np.int32)


def match_points(video_features, audio_features, video_energy, audio_energy, rows=None):
    """Match points (describealign.py:649-673) of the audio frames rows[0] <= i < rows[1] (all
    rows when None), sorted by (i, v).  Returns (i, v, q, Lv, debug dict)."""
    L = lib()
    v_ms, v_nrm, v_code, v_flags = prep_track(video_features, True)
    a_ms, a_nrm, a_code, _ = prep_track(audio_features, False)
    v_sel = np.ascontiguousarray(not_quiet(video_energy)[::4])
    a_nq = np.ascontiguousarray(not_quiet(audio_energy))
    if rows is not None:
        a_nq = np.ascontiguousarray(a_nq[(a_nq >= rows[0]) & (a_nq < rows[1])])
    Lv = max(len(c) for c in v_code)
    cap = max(1 << 20, 4 * len(a_nq))
    stats = np.zeros(2, np.int64)
    while True:
        pi = np.zeros(cap, np.int32); pv = np.zeros(cap, np.int32); pq = np.zeros(cap)
        n = L.oracle_match(_pp(a_ms), _pp(a_nrm), _pp(a_code), _p(a_nq), len(a_nq),
                           _pp(v_ms), _pp(v_nrm), _pp(v_code), _pp(v_flags), _p(v_sel), len(v_sel), Lv,
                           _p(pi), _p(pv), _p(pq), cap, _p(stats))
        if n <= cap:
            break
        cap = int(n)
    dbg = dict(v_ms=v_ms, v_nrm=v_nrm, v_code=v_code, v_flags=v_flags, a_ms=a_ms, a_nrm=a_nrm,
               a_code=a_code, v_sel=v_sel, a_nq=a_nq, touched=int(stats[0]), scored=int(stats[1]))
    return pi[:n].copy(), pv[:n].copy(), pq[:n].copy(), Lv, dbg


def dp1(pi, pv, pq, Lv):
    """Frontier DP #1 + traceback (describealign.py:674-697) over points sorted by (i, v)."""
    n = len(pi)
    pi = np.ascontiguousarray(pi, np.int32); pv = np.ascontiguousarray(pv, np.int32)
    pq = np.ascontiguousarray(pq, np.float64)
    path_i = np.zeros(max(n, 1), np.int32); path_v = np.zeros(max(n, 1), np.int32)
    cum = np.zeros(max(n, 1)); back = np.zeros(max(n, 1), np.int32)
    plen = lib().oracle_dp1(_p(pi), _p(pv), _p(pq), n, Lv, _p(path_i), _p(path_v), _p(cum), _p(back))
    return path_i[:plen].astype(np.int64), path_v[:plen].astype(np.int64), cum[:n], back[:n]


def stage_a(video_features, audio_features, video_energy, audio_energy, want_debug=False):
    pi, pv, pq, Lv, dbg = match_points(video_features, audio_features, video_energy, audio_energy)
    path_x, path_y, cum, back = dp1(pi, pv, pq, Lv)
    out = {
        "points_i": pi, "points_v": pv, "points_q": pq,
        "path_x": path_x, "path_y": path_y,
        "touched": dbg["touched"], "scored": dbg["scored"],
    }
    if want_debug:
        out.update(dbg, cum=cum, back=back)
    return out


def score_corridors(plans, audio_scaled, video_scaled):
    """Pass-2 points (describealign.py:931-944): for each planned corridor, every audio row
    gets one point on the cluster's line; the first cluster to claim (i, int(j)) keeps it.
    Returns arrays sorted by (i, j, cluster): i, j, cluster, qual."""
    a_max = np.max(audio_scaled[:, 0])
    v_max = np.max(video_scaled[:, 0])
    v64 = video_scaled.astype(np.float64)
    ii, jj, cc, qq = [], [], [], []
    for idx, lo, hi, slope, offset in plans:
        if hi <= lo:
            continue
        rows = np.arange(lo, hi)
        y = slope * rows + offset
        f = np.floor(y).astype(np.int64)
        t = (y - f)[:, None]
        v_m = v64[f] * (1.0 - t) + v64[f + 1] * t
        a_m = audio_scaled[lo:hi]
        q = np.sum(-.5 - np.log10(1e-4 + np.abs(a_m - v_m)), axis=1)
        q *= np.clip(v_m[:, 0] + 2.5 - v_max, 0, 1)
        q += np.clip(a_m[:, 0] + 2.5 - a_max, 0, 1) * .1
        ii.append(rows); jj.append(y); cc.append(np.full(len(rows), idx)); qq.append(q)
    if not ii:
        z = np.zeros(0)
        return z.astype(np.int32), z, z.astype(np.int32), z
    i = np.concatenate(ii); j = np.concatenate(jj); c = np.concatenate(cc); q = np.concatenate(qq)
    # first cluster (lowest plan order == lowest cluster index) to claim (i, int(j)) wins
    cell = j.astype(np.int64)
    order = np.lexsort((c, cell, i))
    i, j, c, q, cell = i[order], j[order], c[order], q[order], cell[order]
    first = np.ones(len(i), bool)
    first[1:] = (i[1:] != i[:-1]) | (cell[1:] != cell[:-1])
    i, j, c, q = i[first], j[first], c[first], q[first]
    order = np.lexsort((q, c, j, i))
    return i[order].astype(np.int32), j[order], c[order].astype(np.int32), q[order]


def dp2(pi, pj, pc, pq, n_clusters, n_video_scaled):
    uniq, rank = np.unique(pj, return_inverse=True)
    rank = (rank + 1).astype(np.int32)
    out = np.zeros((max(len(pi), 1), 5))
    n = lib().oracle_dp2(_p(np.ascontiguousarray(pi, np.int32)), _p(np.ascontiguousarray(pj, np.float64)),
                         _p(np.ascontiguousarray(pc, np.int32)), _p(np.ascontiguousarray(pq, np.float64)),
                         _p(rank), len(pi), len(uniq) + 1, n_clusters, n_video_scaled, _p(out))
    return out[:n].copy()


def stage_b(plans, n_clusters, audio_scaled, video_scaled):
    pi, pj, pc, pq = score_corridors(plans, audio_scaled, video_scaled)
    path = dp2(pi, pj, pc, pq, n_clusters, len(video_scaled))
    return {"points_i": pi, "points_j": pj, "points_c": pc, "points_q": pq, "path": path}


# ------------------------------------------------------------------------------------------
# corridor planning (describealign.py:895-932), on the reference's own numpy / scipy calls
# ------------------------------------------------------------------------------------------

def x_limits(x_first, x_last, offset, slope, n_audio, n_video, extend=(210 * 30), buffer_vert=4):
    """Half-open audio-row range of a cluster's corridor (describealign.py:895-900)."""
    lo = max(int(x_first) - extend, 0)
    hi = min(int(x_last) + extend, n_audio - 1)
    lo = max(lo, int(np.ceil((buffer_vert - offset) / slope)))
    hi = min(hi, int(np.floor((n_video - buffer_vert - offset) / slope)))
    return lo, hi


def lerp_video(video_scaled: np.ndarray, y: np.ndarray) -> np.ndarray:
    """Degree-1 spline through the video rows at fractional positions y, in the f64 form
    that is bit-identical to scipy's make_interp_spline(k=1)."""
    f = np.floor(y).astype(np.int64)
    t = (y - f)[:, None]
    v = video_scaled.astype(np.float64)
    return v[f] * (1.0 - t) + v[f + 1] * t


def plan_corridors(clusters, audio_scaled, video_scaled):
    """Per cluster: the audio-row range to score and the (possibly refined) line.

    Implements the control flow of describealign.py:912-932 up to the point where scoring
    starts, including the sub-frame offset refinement (describealign.py:916-930) and the
    quirk that the refinement branch shortens the corridor by one row.
    Returns a list of (cluster_index, lo, hi, slope, offset) for clusters that are scored.
    """
    n_audio, n_video = len(audio_scaled), len(video_scaled)
    plans = []
    for idx, (cx, offset, slope) in enumerate(clusters):
        lo, hi = x_limits(cx[0], cx[-1], offset, slope, n_audio, n_video, extend=0)
        if hi < lo + 5:
            continue
        x_first, x_last = cx[0], cx[-1]
        if hi > lo + 100:
            rows = np.arange(lo, hi)
            y = slope * rows + offset
            a_m = audio_scaled[lo:hi]
            v_m = lerp_video(video_scaled, y)
            err = a_m[1:-1] - v_m[1:-1]
            ok = np.mean(err, axis=-1) < 0.1
            if np.count_nonzero(ok) > 50:
                dv = ((v_m[2:] - v_m[:-2]) / 2.)[ok]
                err = err[ok]
                coef, resid, _, _ = np.linalg.lstsq(dv.reshape(-1, 1), err.flat, rcond=None)
                explained = 1 - (resid / np.sum(err ** 2))
                sigmas = np.sqrt(explained * np.prod(err.shape)) - 1.
                if sigmas > 8 and abs(coef[0]) < 2:
                    offset = offset + coef[0]
            x_first, x_last = rows[0], rows[-1]
        lo2, hi2 = x_limits(x_first, x_last, offset, slope, n_audio, n_video)
        plans.append((idx, lo2, hi2, float(slope), float(offset)))
    return plans



def align(video_features, audio_features, video_energy, audio_energy, host, details=None):
    """Whole align() with the host stage supplied by `host` (see module docstring).
    Returns (audio_times, video_times, similarity_percent, path, median_slope)."""
    a = stage_a(video_features, audio_features, video_energy, audio_energy)
    x, y = a["path_x"], a["path_y"]
    if len(x) < host.min_path_length(len(video_energy), len(audio_energy)):
        raise RuntimeError(FAILED_MSG)
    keep = host.continuity_error(x, y) < 3
    x, y = x[keep], y[keep]
    audio_scaled, video_scaled = host.scale_features(video_features, audio_features, x, y)
    fx, fy = host.compress_path(x, y)
    fit = host.rate_change_fit(fx, fy)
    clusters = host.line_clusters(fit)
    plans = plan_corridors(clusters, audio_scaled, video_scaled)
    b = stage_b(plans, len(clusters), audio_scaled, video_scaled)
    path = b["path"]
    if len(path) < host.min_path_length(len(video_energy), len(audio_energy)):
        raise RuntimeError(FAILED_MSG)
    if details is not None:
        details.update(stage_a=a, stage_b=b, fit=fit, clusters=clusters, plans=plans,
                       audio_scaled=audio_scaled, video_scaled=video_scaled, kept_x=x, kept_y=y)
    nx, ny, sim = host.build_nodes(path, len(audio_energy), len(video_energy), len(audio_scaled), len(video_scaled))
    return nx, ny, sim, path, fit.median_slope


# ------------------------------------------------------------------------------------------
# Literal pure-Python versions (small inputs only)
# ------------------------------------------------------------------------------------------

def literal_dp1(points):
    """points: iterable of (i, v, qual) sorted by (i, v).  Sorted frontier list with
    strictly increasing cum, as described in SURVEY.md A.5.  Returns [(v, i), ...]."""
    keys = [-1]
    front = [(-1, -1, 0.0)]
    back = {}
    for i, v, q in points:
        k = bisect.bisect_right(keys, v)
        pv, pi_, pc = front[k - 1]
        cum = pc + q
        while k < len(front) and front[k][2] <= cum:
            del front[k]; del keys[k]
        pos = bisect.bisect_right(keys, v)
        keys.insert(pos, v); front.insert(pos, (v, i, cum))
        back[(v, i)] = (pv, pi_)
    path = [front[-1][:2]]
    while path[-1] in back:
        path.append(back[path[-1]])
    path.pop()
    path.reverse()
    return path


def literal_dp2(points_by_row, n_clusters, n_video):
    """points_by_row: list over audio rows of sorted [(j, cluster, qual)].  Follows the
    data structures of SURVEY.md A.7 (sorted frontier, per-cluster best, prev_cache)."""
    keys = [0]
    front = [(0, 0, -1, 0, 0)]
    cbest = [(0, 0, 0, -1000) for _ in range(n_clusters)]
    back = {}
    cache = np.full((n_video, 5), -np.inf)
    cache[0] = (0, 0, -1, 0, 0)
    for i, row in enumerate(points_by_row):
        for j, c, q in row:
            k = bisect.bisect_right(keys, j)
            pj, pi_, pc, pq, best = front[k - 1]
            cl = cbest[c]
            if cl[3] >= best:
                pj, pi_, pq, best = cl
                pc = c
            for cell in range(max(0, int(j) - 2), int(j) + 1):
                nd = cache[cell].tolist()
                if c != nd[2]:
                    nd[4] -= 100 + 100 * ((j - nd[0]) - (i - nd[1])) ** 2
                if nd[1] >= (i - 2) and nd[0] <= j and nd[4] >= best:
                    pj, pi_, pc, pq, best = nd
            cum = best + q
            cache[int(j)] = (j, i, c, q, cum)
            jump = cum - 1000
            if front[k - 1][4] < jump:
                while k < len(front) and front[k][4] <= jump:
                    del front[k]; del keys[k]
                pos = bisect.bisect_right(keys, j)
                keys.insert(pos, j); front.insert(pos, (j, i, c, q, jump))
            if cl[3] < cum - 50:
                cbest[c] = (j, i, q, cum - 50)
            back[(j, i)] = (pj, pi_, pc, pq, best)
    path = [front[-1]]
    while tuple(path[-1][:2]) in back:
        path.append(back[tuple(path[-1][:2])])
    path.pop()
    path.reverse()
    return np.array(path, dtype=np.float64).reshape(-1, 5)
