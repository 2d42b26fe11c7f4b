"""TEST INFRASTRUCTURE ONLY - CPU restatement of the jump search inside the reference's `stretch`
(describealign.py:252-304 `get_pearson_corrs_generator` + the two lines of :330-333 that reduce every 512-sample
window to its best jump location and loss), in plain numpy.  Checker of the CUDA kernel behind
describealign_b200.stretch.best_jumps; never imported by the product.

For an input (channels, n) float16 and a list of jump distances it returns, per 512-sample window w and jump k,
    loc[w, k]   the start position inside the window whose 512-sample Pearson correlation with the window `jump`
                samples further on (earlier, if `negative`) is largest (first maximum),
    best[w, k]  that correlation,
with the reference's own piece structure: the input is handled in pieces of 51 windows that overlap by two, each with
its own epsilon (1e-4 of the piece's largest window energy) and its own float64 running sums, exactly as the recursive
generator yields them (pieces advance by 49 windows; the first yields its windows 0-49, the others 1-49, the last all
that remain)."""
from __future__ import annotations

import numpy as np

WINDOW = 512
CACHED = 50                       # describealign.py:259
CUT = CACHED * WINDOW             # 25 600
LIMIT = (CACHED + 2) * 1.1 * WINDOW   # pieces longer than this are split (describealign.py:261)


def pieces(n: int):
    """[(first sample, length, first yielded local window, number of yielded windows)] in order."""
    out = []
    start, first = 0, True
    while True:
        left = n - start
        last = not (left > LIMIT)
        length = left if last else CUT + WINDOW
        lo = 0 if first else 1
        hi = (length // WINDOW) if last else CACHED
        out.append((start, length, lo, max(0, hi - lo)))
        if last:
            return out
        start += CUT - WINDOW
        first = False


def _piece_corrs(x: np.ndarray, negative: bool, jumps):
    """(positions, jumps) Pearson matrix of one piece (describealign.py:273-302)."""
    n = x.shape[1]
    if n < 3 * WINDOW - 1:
        raise RuntimeError("Invalid state in Pearson generator.")
    w = n - WINDOW + 1
    corrs = np.full((len(jumps), w), -np.inf)
    energy = np.sum(x.astype(np.float32) ** 2, axis=0)
    run = np.cumsum(energy, dtype=np.float64)
    run[WINDOW:] = run[WINDOW:] - run[:-WINDOW]
    rms = run[WINDOW - 1:]
    eps = 1e-4 * max(1, np.max(rms))
    rms = np.sqrt(rms + eps)
    for k, jump in enumerate(jumps):
        prod = np.sum(x[:, jump:].astype(np.float32) * x[:, :n - jump], axis=0)
        run = np.cumsum(prod, dtype=np.float64)
        run[WINDOW:] = run[WINDOW:] - run[:-WINDOW]
        num = run[WINDOW - 1:] + eps
        if negative:
            corrs[k, jump:] = num / rms[:len(rms) - jump]
        else:
            corrs[k, :w - jump] = num / rms[jump:]
    corrs = corrs / rms[None, :]
    return corrs.T


def best_jumps(x: np.ndarray, negative: bool, jumps):
    x = np.asarray(x)
    jumps = [int(j) for j in jumps]
    n = x.shape[1]
    nw = n // WINDOW
    loc = np.zeros((nw, len(jumps)), dtype=np.int16)
    best = np.full((nw, len(jumps)), -np.inf)
    g = 0
    for start, length, lo, count in pieces(n):
        c = _piece_corrs(x[:, start:start + length], negative, jumps)
        for local in range(lo, lo + count):
            rows = c[local * WINDOW:(local + 1) * WINDOW]
            if g >= nw:
                break
            at = np.argmax(rows, axis=0)
            loc[g] = at
            best[g] = rows[at, np.arange(len(jumps))]
            g += 1
    assert g == nw, (g, nw)
    return loc, best
