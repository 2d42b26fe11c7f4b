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

import dp2_block_model as M
from dp2_block_model import NEG, P2_GAP, P2_NEAR, P2_VIS1, P2_VIS2

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
    st = M.State(plans, pj)
    back = [None] * n
    cnt = {"blocks": 0, "passes": 0, "scalar_points": 0, "hard_points": 0, "block_points": 0, "wasted_points": 0,
           "full_ok": 0, "prefix_commits": 0, "scalar_query": 0}
    hard = (flags & (P2_NEAR | P2_GAP)) != 0
    nc = len(plans)
    p0 = 0
    while p0 < n:
        if hard[p0]:
            M.scalar_point(st, p0, int(pk[p0]), int(pi[p0]), float(pj[p0]), float(pq[p0]), int(ro[p0]), int(flags[p0]), back, cnt)
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
            M.commit(st, k, p, float(pj[p]), int(ro[p]), int(flags[p]), res[p][0], res[p][1], float(pq[p]), back)
        cnt["block_points"] += good_upto - p0
        p0 = good_upto
        if good_upto < p1:
            M.scalar_point(st, p0, int(pk[p0]), int(pi[p0]), float(pj[p0]), float(pq[p0]), int(ro[p0]), int(flags[p0]), back, cnt)
            cnt["scalar_points"] += 1
            p0 += 1
    return back, st, cnt


def main():
    d = pickle.load(open(sys.argv[1], "rb"))
    plans, pi, pj, pc, pq = d["plans"], d["pi"], d["pj"], d["pc"], d["pq"]
    nmax = int(sys.argv[2]) if len(sys.argv) > 2 else len(pi)
    NB = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    pi, pj, pc, pq = pi[:nmax], pj[:nmax], pc[:nmax], pq[:nmax]
    pk, cell, ro, flags = M.point_flags(plans, pi, pj, pc)
    b0, s0, c0 = M.run_scalar(plans, pi, pj, pq, pk, ro, flags)
    b1, s1, c1 = run_scan(plans, pi, pj, pq, pk, ro, flags, NB=NB, verbose=True)
    bad = [p for p in range(len(pi)) if b0[p] != b1[p]]
    print("scan", c1)
    print("mismatching back records:", len(bad), bad[:5])
    print("top equal:", s0.top == s1.top)


if __name__ == "__main__":
    main()
