#!/usr/bin/env python
"""Executable model of dp2_block_kernel (describealign_b200/csrc/stage_b.cu): the pass-2 frontier DP
(reference describealign.py:946-983) evaluated in blocks of up to 32 points.

The CUDA kernel cannot be run in the authoring container (no GPU), so its control flow - which
corridors are evaluated without the frontier ("leaders"), which with it ("followers"), what is
verified before a block is committed, when a point falls back to the one-point-at-a-time rules -
is stated here in plain Python, checked against the plain sequential rules on real stage-B inputs,
and used to count how many points each path handles.  Test/tool code only.

    python tools/dp2_block_model.py /tmp/c2_stageb_0.pkl
"""
from __future__ import annotations

import math
import pickle
import sys

import numpy as np

NEG = -math.inf
P2_NEAR, P2_VIS1, P2_VIS2, P2_GAP, P2_MAYQ = 1, 2, 4, 8, 16
LEAD_MARGIN = 500.0


def line_at(slope, offset, i):
    return slope * float(i) + offset


def point_flags(plans, pi, pj, pc):
    """Static per-point facts, as corridor_kernel<true> works them out."""
    n = len(pi)
    clus2k = {p[0]: k for k, p in enumerate(plans)}
    pk = np.array([clus2k[c] for c in pc], np.int32)
    cell = pj.astype(np.int64)
    where = {(int(pk[p]), int(pi[p])): int(cell[p]) for p in range(n)}
    flags = np.zeros(n, np.int32)
    ro = np.zeros(n, np.int32)
    for p in range(n):
        k, i, j, c = int(pk[p]), int(pi[p]), float(pj[p]), int(cell[p])
        lo = plans[k][1]
        ro[p] = i - lo
        a = where.get((k, i - 1))
        b = where.get((k, i - 2))
        has1, has2 = a is not None, b is not None
        vis1 = (a >= c - 2) if has1 else (has2 and b >= c - 2)
        vis2 = has1 and has2 and b >= c - 2 and b != a
        gap = (not has1) and ro[p] > 0
        near = mayq = False
        for k2, (_, lo2, hi2, s2, o2) in enumerate(plans):
            if k2 == k:
                continue
            for r in (i - 2, i - 1, i):
                if lo2 <= r < hi2:
                    oc = int(line_at(s2, o2, r))
                    near = near or (c - 2 <= oc <= c)
            if hi2 > lo2 and lo2 <= i - 1:
                r = min(i - 1, hi2 - 1)
                mayq = mayq or line_at(s2, o2, r) > j
        flags[p] = (P2_NEAR * near) | (P2_VIS1 * bool(vis1)) | (P2_VIS2 * bool(vis2)) | (P2_GAP * gap) | (P2_MAYQ * mayq)
    return pk, cell, ro, flags


class State:
    def __init__(self, plans, pj):
        self.plans = plans
        self.pj = pj
        nc = len(plans)
        self.c = [[NEG, NEG, NEG] for _ in range(nc)]
        self.id = [[-2, -2, -2] for _ in range(nc)]
        self.cl = [(-1000.0, -1) for _ in range(nc)]
        self.pmh = [(NEG, -2) for _ in range(nc)]
        self.filled = [-1] * nc
        self.pm = [[None] * max(0, p[2] - p[1]) for p in plans]
        self.top = (0.0, 0.0, -1)       # value, j, id
        self.cache = {}                 # prev_cache: cell -> (j, i, cluster, cum, id)

    def rows_le(self, k2, j, i):
        """number of processed rows of corridor k2 (rows <= i) whose coordinate is <= j"""
        _, lo, hi, s, o = self.plans[k2]
        rows = max(0, hi - lo)
        if rows == 0 or lo > i:
            return 0
        kk = int(min(max(math.floor((j - o) / s) - lo + 1, 0), rows))
        while kk < rows and line_at(s, o, lo + kk) <= j:
            kk += 1
        while kk > 0 and line_at(s, o, lo + kk - 1) > j:
            kk -= 1
        done = min(i + 1, lo + rows) - lo
        return min(kk, done)

    def query(self, k, i, j, cap=None, extra=()):
        """F(j): best frontier entry with j' <= j among the other corridors and the seed, arg-max on
        (value desc, j' asc, id asc).  cap: per-corridor last row that may be read (block mode: rows
        written before the block; the running maximum makes PM[min(x, cap)] the best entry among
        them).  extra: further (value, j', id) candidates (points of the current block)."""
        best = (0.0, 0.0, -1)           # the seed (value, j', id)
        cands = []
        for k2 in range(len(self.plans)):
            if k2 == k:
                continue
            idx = self.rows_le(k2, j, i)
            f = self.filled[k2] if cap is None else cap[k2]
            if idx <= 0 or f < 0:
                continue
            x = min(idx - 1, f)
            e = self.pmh[k2] if (cap is None and x >= self.filled[k2]) else self.pm[k2][x]
            if e[0] == NEG:
                continue
            cands.append((e[0], 0.0 if e[1] < 0 else float(self.pj[e[1]]), e[1]))
        for cand in list(cands) + list(extra):
            if cand[0] > best[0] or (cand[0] == best[0] and (cand[1], cand[2]) < (best[1], best[2])):
                best = cand
        return best[0], best[2]


def local_best(st, k, fl):
    """cluster best, then the corridor's points two and one rows back (later candidates win ties)"""
    m, mi = st.cl[k]
    if (fl & P2_VIS2) and st.c[k][1] >= m:
        m, mi = st.c[k][1], st.id[k][1]
    if (fl & P2_VIS1) and st.c[k][0] >= m:
        m, mi = st.c[k][0], st.id[k][0]
    return m, mi


def commit(st, k, p, j, ro, fl, best, pred, q, back):
    cum = best + q
    st.c[k] = [cum, st.c[k][0], st.c[k][1]]
    st.id[k] = [p, st.id[k][0], st.id[k][1]]
    if st.cl[k][0] < cum - 50.0:
        st.cl[k] = (cum - 50.0, p)
    jump = cum - 1000.0
    if fl & P2_GAP:
        for r in range(st.filled[k] + 1, ro):
            st.pm[k][r] = st.pmh[k]
    if jump > st.pmh[k][0]:
        st.pmh[k] = (jump, p)
    st.filled[k] = ro
    st.pm[k][ro] = st.pmh[k]
    back[p] = (best, pred)
    st.cache[int(j)] = (j, ro + st.plans[k][1], st.plans[k][0], cum, p)
    if jump > st.top[0] or (jump == st.top[0] and j < st.top[1]):
        st.top = (jump, j, p)


def scalar_point(st, p, k, i, j, q, ro, fl, back, counters):
    if fl & P2_NEAR:
        # generic rules (describealign.py:960-973): frontier, cluster best, prev_cache cells
        best, pred = (st.top[0], st.top[2]) if st.top[1] <= j else st.query(k, i, j)
        if st.cl[k][0] >= best:
            best, pred = st.cl[k]
        for cell in range(int(j) - 2, int(j) + 1):
            e = st.cache.get(cell)
            if e is None:
                continue
            ej, ei, ec, ecum, eid = e
            if ec != st.plans[k][0]:
                ecum = ecum - (100.0 + 100.0 * (((j - ej) - (i - ei)) * ((j - ej) - (i - ei))))
            if ei >= i - 2 and ej <= j and ecum >= best:
                best, pred = ecum, eid
        commit(st, k, p, j, ro, fl, best, pred, q, back)
        return
    m, mi = local_best(st, k, fl)
    left = st.top[1] <= j
    best, pred = (st.top[0], st.top[2]) if (left and st.top[0] > m) else (m, mi)
    if (fl & P2_MAYQ) and not left and m < st.top[0]:
        counters["scalar_query"] += 1
        fv, fi = st.query(k, i, j)
        if fv > m:
            best, pred = fv, fi
    commit(st, k, p, j, ro, fl, best, pred, q, back)


def run_scalar(plans, pi, pj, pq, pk, ro, flags):
    st = State(plans, pj)
    back = [None] * len(pi)
    counters = {"scalar_query": 0}
    for p in range(len(pi)):
        scalar_point(st, p, int(pk[p]), int(pi[p]), float(pj[p]), float(pq[p]), int(ro[p]), int(flags[p]), back, counters)
    return back, st, counters


def run_blocks(plans, pi, pj, pq, pk, ro, flags, min_block=2):
    """The block algorithm.  Groups of 32 consecutive points; inside a group, maximal ranges free of
    NEAR / GAP points are evaluated as blocks."""
    n = len(pi)
    st = State(plans, pj)
    back = [None] * n
    cnt = {"scalar_query": 0, "scalar_points": 0, "block_points": 0, "blocks": 0, "failed_blocks": 0,
           "fail_leader_top": 0, "fail_follower_top": 0, "fail_query": 0, "block_queries": 0, "wasted_points": 0}
    nc = len(plans)
    for base in range(0, n, 32):
        gcnt = min(32, n - base)
        t = 0
        if gcnt == 32 and fast_block(st, base, pi, pj, pq, pk, ro, flags, back, cnt):
            continue
        while t < gcnt:
            e = t
            while e < gcnt and not (int(flags[base + e]) & (P2_NEAR | P2_GAP)):
                e += 1
            if e - t >= min_block:
                glen = block(st, base, t, e, pi, pj, pq, pk, ro, flags, back, cnt, nc)
                t += glen
                if t >= gcnt:
                    break
                # the point at t failed verification (or is NEAR / GAP): one point by the scalar rules
            p = base + t
            scalar_point(st, p, int(pk[p]), int(pi[p]), float(pj[p]), float(pq[p]), int(ro[p]), int(flags[p]), back, cnt)
            cnt["scalar_points"] += 1
            t += 1
    return back, st, cnt


def fast_block(st, base, pi, pj, pq, pk, ro, flags, back, cnt):
    """A full group of 32 points that all belong to ONE leader corridor and all see the corridor's
    previous two points: under the assumption that the cluster best is never chosen, cum is the
    chain max(c0, c1) + q (the only serial part); everything else is decided per point afterwards
    and the assumption checked.  All or nothing: returns False (nothing committed) when a check fails."""
    pts = list(range(base, base + 32))
    k = int(pk[base])
    both = P2_VIS1 | P2_VIS2
    if any(int(pk[p]) != k or (int(flags[p]) & both) != both or (int(flags[p]) & (P2_NEAR | P2_GAP)) for p in pts):
        return False
    if not (st.c[k][0] >= st.top[0] + LEAD_MARGIN):
        return False
    cnt["fast_tried"] = cnt.get("fast_tried", 0) + 1
    a, b = st.c[k][0], st.c[k][1]
    cums = []
    for p in pts:
        m = a if a >= b else b
        cum = m + float(pq[p])
        cums.append(cum)
        b, a = a, cum
    # per point, in parallel
    pre_c = [st.c[k][1], st.c[k][0]]          # cums two and one before the block
    pre_id = [st.id[k][1], st.id[k][0]]
    allc = pre_c + cums
    allid = pre_id + pts
    clv = st.cl[k][0]
    top = st.top
    outs = []
    for u, p in enumerate(pts):
        c0, c1 = allc[u + 1], allc[u]
        m, pred = (c0, allid[u + 1]) if c0 >= c1 else (c1, allid[u])
        if not (m >= clv):
            return False                      # the cluster best would have been chosen
        if not (top[0] <= m):
            return False                      # the frontier could have been chosen
        outs.append((m, pred))
        cum = cums[u]
        clv = max(clv, cum - 50.0)
        jump = cum - 1000.0
        if jump > top[0] or (jump == top[0] and float(pj[p]) < top[1]):
            top = (jump, float(pj[p]), p)
    for u, p in enumerate(pts):
        commit(st, k, p, float(pj[p]), int(ro[p]), int(flags[p]), outs[u][0], outs[u][1], float(pq[p]), back)
        assert st.c[k][0] == cums[u]
    assert st.top == top
    cnt["fast_points"] = cnt.get("fast_points", 0) + 32
    return True


def block(st, base, t0, t1, pi, pj, pq, pk, ro, flags, back, cnt, nc):
    """Evaluate points base+t0 .. base+t1-1 as one block; commit the verified prefix; return its length."""
    cnt["blocks"] += 1
    top0 = st.top
    pts = list(range(base + t0, base + t1))
    leader = [st.c[k][0] >= top0[0] + LEAD_MARGIN for k in range(nc)]
    filled0 = list(st.filled)
    # working copies of the per-corridor state (the kernel: the owner lane's registers)
    wc = [list(x) for x in st.c]
    wid = [list(x) for x in st.id]
    wcl = list(st.cl)
    wpm = list(st.pmh)
    out = {}

    def own_step(p, frontier):
        k, fl, q = int(pk[p]), int(flags[p]), float(pq[p])
        c_before, cl_before = (wc[k][0], wc[k][1]), wcl[k][0]
        m, mi = wcl[k]
        if (fl & P2_VIS2) and wc[k][1] >= m:
            m, mi = wc[k][1], wid[k][1]
        if (fl & P2_VIS1) and wc[k][0] >= m:
            m, mi = wc[k][0], wid[k][0]
        best, pred = m, mi
        if frontier is not None and frontier[0] > m:
            best, pred = frontier
        cum = best + q
        wc[k] = [cum, wc[k][0], wc[k][1]]
        wid[k] = [p, wid[k][0], wid[k][1]]
        if wcl[k][0] < cum - 50.0:
            wcl[k] = (cum - 50.0, p)
        jump = cum - 1000.0
        if jump > wpm[k][0]:
            wpm[k] = (jump, p)
        # the kernel's LITE walk computes cum as max(c0, c1, frontier candidate) + q; that is what the
        # exact rules give unless the cluster best is the choice (the kernel also falls back to the exact
        # walk when two successive running maxima of cum round to the same cum - 50 or cum - 1000: a pure
        # id tie, which the exact rules resolve in favour of the earlier point)
        lite_ok = (fl & (P2_VIS1 | P2_VIS2)) == (P2_VIS1 | P2_VIS2) and \
            (max(c_before[0], c_before[1]) >= cl_before or (frontier is not None and frontier[0] > cl_before))
        out[p] = {"best": best, "pred": pred, "m": m, "cum": cum, "jump": jump, "pm": wpm[k], "lite_ok": lite_ok}

    # phase A: leaders, frontier ignored
    for p in pts:
        if leader[int(pk[p])]:
            own_step(p, None)
    # phase B: running top over the block order from the leaders' points (exclusive prefix)
    topat = {}
    cur = top0
    for p in pts:
        topat[p] = cur
        if leader[int(pk[p])]:
            jv, j = out[p]["jump"], float(pj[p])
            if jv > cur[0] or (jv == cur[0] and j < cur[1]):
                cur = (jv, j, p)
    # phase Q: frontier queries of follower points whose top lies to their right: entries written
    # before the block (per corridor, capped at its last final row) and the leaders' points of this block
    fq = {}
    for n_before, p in enumerate(pts):
        k, fl = int(pk[p]), int(flags[p])
        if leader[k] or not (fl & P2_MAYQ):
            continue
        if topat[p][1] <= float(pj[p]):
            continue
        cnt["block_queries"] += 1
        extra = [(out[u]["jump"], float(pj[u]), u) for u in pts[:n_before]
                 if leader[int(pk[u])] and int(pk[u]) != k and float(pj[u]) <= float(pj[p])]
        fq[p] = st.query(k, int(pi[p]), float(pj[p]), cap=filled0, extra=extra)
    # phase C: followers
    qfail = set()
    for p in pts:
        k, fl = int(pk[p]), int(flags[p])
        if leader[k]:
            continue
        tv, tj, ti = topat[p]
        if tj <= float(pj[p]):
            own_step(p, (tv, ti))
        elif fl & P2_MAYQ:
            own_step(p, fq[p])
        else:
            own_step(p, None)
    # a queried follower point must not be able to use a FOLLOWER's point of this block (those were
    # not among its candidates): every such point's jump value has to be <= what the point chose from
    for n_before, p in enumerate(pts):
        if p not in fq:
            continue
        k = int(pk[p])
        for u in pts[:n_before]:
            ku = int(pk[u])
            if not leader[ku] and ku != k and float(pj[u]) <= float(pj[p]) and not (out[u]["jump"] <= out[p]["m"]):
                qfail.add(p)
    # (kernel: if every point qualified for the lite walk the block was evaluated that way, otherwise -
    # or when a point turns out to restart from its cluster best - with the exact walk; same results)
    if all(out[p]["lite_ok"] for p in pts if p in out):
        cnt["lite_blocks"] = cnt.get("lite_blocks", 0) + 1
    else:
        cnt["exact_blocks"] = cnt.get("exact_blocks", 0) + 1
    # phase D: verification, first failure
    glen = 0
    for p in pts:
        k = int(pk[p])
        if p in qfail:
            cnt["fail_query"] += 1
            break
        if leader[k]:
            if not (topat[p][0] <= out[p]["m"]):
                cnt["fail_leader_top"] += 1
                break
        else:
            if not (out[p]["jump"] < topat[p][0]):
                cnt["fail_follower_top"] += 1
                break
        glen += 1
    if glen < len(pts):
        cnt["failed_blocks"] += 1
        cnt["wasted_points"] += len(pts) - glen
    # commit the verified prefix through the ordinary state update (the kernel re-runs the owner
    # loops up to the failing point from the saved registers; the values are the same)
    for n_done, p in enumerate(pts[:glen]):
        o = out[p]
        commit(st, int(pk[p]), p, float(pj[p]), int(ro[p]), int(flags[p]), o["best"], o["pred"], float(pq[p]), back)
        assert st.pmh[int(pk[p])] == o["pm"]
        if n_done + 1 < len(pts):
            assert st.top == topat[pts[n_done + 1]]
    cnt["block_points"] += glen
    return glen


def main():
    d = pickle.load(open(sys.argv[1], "rb"))
    plans, pi, pj, pc, pq = d["plans"], d["pi"], d["pj"], d["pc"], d["pq"]
    if len(sys.argv) > 2:
        nmax = int(sys.argv[2])
        pi, pj, pc, pq = pi[:nmax], pj[:nmax], pc[:nmax], pq[:nmax]
    pk, cell, ro, flags = point_flags(plans, pi, pj, pc)
    print("points", len(pi), "NEAR", int(np.sum(flags & P2_NEAR > 0)), "GAP", int(np.sum(flags & P2_GAP > 0)),
          "MAYQ", int(np.sum(flags & P2_MAYQ > 0)), "noVIS1", int(np.sum(flags & P2_VIS1 == 0)))
    b0, s0, c0 = run_scalar(plans, pi, pj, pq, pk, ro, flags)
    b1, s1, c1 = run_blocks(plans, pi, pj, pq, pk, ro, flags)
    bad = [p for p in range(len(pi)) if b0[p] != b1[p]]
    print("scalar", c0)
    print("blocks", c1)
    print("mismatching back records:", len(bad), bad[:5])
    print("top equal:", s0.top == s1.top)


if __name__ == "__main__":
    main()
