#!/usr/bin/env python
"""Executable model of dp2_scan_kernel (describealign_b200/csrc/stage_b.cu): the pass-2 frontier DP
(reference describealign.py:946-983) evaluated in blocks of up to NB points by an exact integer
max-plus scan followed by a verification in float64.

Idea.  Along one corridor the recurrence is cum_k = max(cum_{k-1}, cum_{k-2}, E_k) + q_k, E_k being
whatever enters from outside the chain (frontier jump, cluster best).  While every value of a chain
stays inside one binade [2^e, 2^(e+1)), an IEEE round-to-nearest add is  fl(x + q) = x + rn_u(q)
with u = 2^(e-52) and rn_u = rounding to a multiple of u (ties aside), i.e. exact INTEGER arithmetic
in units of u - and integer max-plus recurrences are associative, so a block of points is one
parallel scan of 2x2 max-plus matrices with an affine term instead of a serial f64 chain.  Whatever
the scan produced is then checked point by point with the reference's own float64 rules (each point
recomputes its predecessor choice and its cum from the block's values): if every point reproduces
its value bit for bit, the block is the reference's result by induction; otherwise E is refreshed
from the block's own values and the scan repeated (followers of a leader corridor, restarts from a
cluster best reached inside the block), and finally the verified prefix is committed and the first
failing point evaluated by the one-point rules (binade crossings, rounding ties, NEAR / GAP points).

Test/tool code only.   python tools/dp2_scan_model.py /tmp/c2_stageb_0.pkl
"""
from __future__ import annotations

import math
import pickle
import sys

import numpy as np

# ---------------------------------------------------------------------------------------------
# the one-point rules (what dp2_scan.cuh's sc_one_point evaluates) and the per-point flags
# corridor_kernel<true> works out; `run_scalar` is the plain sequential DP the scan is compared with
# ---------------------------------------------------------------------------------------------
NEG = -math.inf
P2_NEAR, P2_VIS1, P2_VIS2, P2_GAP, P2_MAYQ = 1, 2, 4, 8, 16


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


INEG = -(1 << 61)


def _clamp(x):
    return x if x > INEG else INEG


def to_int(x, scale):
    if x == NEG:
        return INEG
    return int(np.rint(x * scale))


def scan_chain(c0, c1, E, Q, v1, v2):
    """Integer max-plus evaluation of one chain (what the parallel scan computes; sequential here -
    the operation is associative so the order of evaluation cannot matter)."""
    out = []
    a, b = c0, c1
    for k in range(len(Q)):
        m = E[k]
        if v2[k] and b > m:
            m = b
        if v1[k] and a > m:
            m = a
        c = _clamp(m + Q[k]) if m > INEG else INEG
        out.append(c)
        b, a = a, c
    return out


def better(a, b):
    """frontier order: value desc, j' asc, id asc; entries are (val, j, id)"""
    return a[0] > b[0] or (a[0] == b[0] and (a[1], a[2]) < (b[1], b[2]))


def run_scan(plans, pi, pj, pq, pk, ro, flags, NB=1024, max_pass=3, verbose=False):
    n = len(pi)
    st = State(plans, pj)
    back = [None] * n
    cnt = {"blocks": 0, "passes": 0, "scalar_points": 0, "hard_points": 0, "block_points": 0, "wasted_points": 0,
           "full_ok": 0, "prefix_commits": 0, "scalar_query": 0}
    hard = (flags & (P2_NEAR | P2_GAP)) != 0
    nc = len(plans)
    p0 = 0
    while p0 < n:
        if hard[p0]:
            scalar_point(st, p0, int(pk[p0]), int(pi[p0]), float(pj[p0]), float(pq[p0]), int(ro[p0]), int(flags[p0]), back, cnt)
            cnt["hard_points"] += 1
            p0 += 1
            continue
        p1 = min(n, p0 + NB)
        hh = np.nonzero(hard[p0:p1])[0]
        if len(hh):
            p1 = p0 + int(hh[0])
        ids = list(range(p0, p1))
        chains = {}
        for p in ids:
            chains.setdefault(int(pk[p]), []).append(p)
        filled0 = list(st.filled)
        cnt["blocks"] += 1
        # ---- external inputs from the committed state only ----
        E = {}
        for p in ids:
            k = int(pk[p])
            fv, fi = st.query(k, int(pi[p]), float(pj[p]), cap=filled0)
            e = (fv, fi)
            if st.cl[k][0] >= e[0]:
                e = st.cl[k]
            E[p] = e
        good_upto = None
        res = None
        for npass in range(max_pass):
            cnt["passes"] += 1
            # ---- S1: integer scan per chain ----
            cum = {}
            for k, pts in chains.items():
                ref = st.c[k][0] if st.c[k][0] != NEG else E[pts[0]][0]
                if ref == 0.0 or not math.isfinite(ref):
                    e2 = 0
                else:
                    e2 = math.frexp(abs(ref))[1] - 1
                scale = math.ldexp(1.0, 52 - e2)
                u = math.ldexp(1.0, e2 - 52)
                Ei = [to_int(E[p][0], scale) for p in pts]
                Qi = [int(np.rint(float(pq[p]) * scale)) for p in pts]
                v1 = [bool(flags[p] & P2_VIS1) for p in pts]
                v2 = [bool(flags[p] & P2_VIS2) for p in pts]
                ci = scan_chain(to_int(st.c[k][0], scale), to_int(st.c[k][1], scale), Ei, Qi, v1, v2)
                for p, c in zip(pts, ci):
                    cum[p] = float(c) * u if c > INEG // 2 else NEG
            # ---- S2: per chain, cluster best before and running-max head after every point ----
            clb, pma = {}, {}
            for k, pts in chains.items():
                cl = st.cl[k]
                pm = st.pmh[k]
                for p in pts:
                    clb[p] = cl
                    if cl[0] < cum[p] - 50.0:
                        cl = (cum[p] - 50.0, p)
                    if cum[p] - 1000.0 > pm[0]:
                        pm = (cum[p] - 1000.0, p)
                    pma[p] = pm
            # ---- S3: F(p) from committed rows and the block's own entries ----
            first_ro = {k: int(ro[pts[0]]) for k, pts in chains.items()}
            def frontier(p):
                k, i, j = int(pk[p]), int(pi[p]), float(pj[p])
                best = (0.0, 0.0, -1)
                for k2 in range(nc):
                    if k2 == k:
                        continue
                    idx = st.rows_le(k2, j, i)
                    if idx <= 0:
                        continue
                    x = idx - 1
                    if k2 in chains and x >= first_ro[k2]:
                        pts2 = chains[k2]
                        x = min(x, first_ro[k2] + len(pts2) - 1)
                        e = pma[pts2[x - first_ro[k2]]]
                    else:
                        f = filled0[k2]
                        if f < 0:
                            continue
                        e = st.pm[k2][min(x, f)]
                    if e[0] == NEG:
                        continue
                    cand = (e[0], 0.0 if e[1] < 0 else float(pj[e[1]]), e[1])
                    if better(cand, best):
                        best = cand
                return best[0], best[2]
            # ---- S4: every point recomputes its choice and value with the float64 rules ----
            res = {}
            newE = {}
            bad = None
            for k, pts in chains.items():
                pc = [st.c[k][1], st.c[k][0]] + [cum[p] for p in pts]
                pid = [st.id[k][1], st.id[k][0]] + pts
                for u_, p in enumerate(pts):
                    fl = int(flags[p])
                    fv, fi = frontier(p)
                    e = (fv, fi)
                    if clb[p][0] >= e[0]:
                        e = clb[p]
                    newE[p] = e
                    best, pred = e
                    if (fl & P2_VIS2) and pc[u_] >= best:
                        best, pred = pc[u_], pid[u_]
                    if (fl & P2_VIS1) and pc[u_ + 1] >= best:
                        best, pred = pc[u_ + 1], pid[u_ + 1]
                    res[p] = (best, pred)
                    if best + float(pq[p]) != cum[p]:
                        bad = p if bad is None else min(bad, p)
            if bad is None:
                good_upto = p1
                break
            good_upto = bad
            E = newE
        if good_upto == p1:
            cnt["full_ok"] += 1
        else:
            cnt["prefix_commits"] += 1
            cnt["wasted_points"] += p1 - good_upto
            if verbose:
                print("block", p0, p1, "fails at", good_upto)
        for p in range(p0, good_upto):
            k = int(pk[p])
            commit(st, k, p, float(pj[p]), int(ro[p]), int(flags[p]), res[p][0], res[p][1], float(pq[p]), back)
        cnt["block_points"] += good_upto - p0
        p0 = good_upto
        if good_upto < p1:
            scalar_point(st, p0, int(pk[p0]), int(pi[p0]), float(pj[p0]), float(pq[p0]), int(ro[p0]), int(flags[p0]), back, cnt)
            cnt["scalar_points"] += 1
            p0 += 1
    return back, st, cnt


def main():
    d = pickle.load(open(sys.argv[1], "rb"))
    plans, pi, pj, pc, pq = d["plans"], d["pi"], d["pj"], d["pc"], d["pq"]
    nmax = int(sys.argv[2]) if len(sys.argv) > 2 else len(pi)
    NB = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    pi, pj, pc, pq = pi[:nmax], pj[:nmax], pc[:nmax], pq[:nmax]
    pk, cell, ro, flags = point_flags(plans, pi, pj, pc)
    b0, s0, c0 = run_scalar(plans, pi, pj, pq, pk, ro, flags)
    b1, s1, c1 = run_scan(plans, pi, pj, pq, pk, ro, flags, NB=NB, verbose=True)
    bad = [p for p in range(len(pi)) if b0[p] != b1[p]]
    print("scan", c1)
    print("mismatching back records:", len(bad), bad[:5])
    print("top equal:", s0.top == s1.top)


if __name__ == "__main__":
    main()
