"""Which log10 the feature kernel reproduces.

The reference computes its energy and band features with np.log10 on float32 arrays
(describealign.py:554, 590).  On hosts without AVX-512 that is glibc's log10f, which the CUDA feature
kernel restates in IEEE operations ("portable" mode, the default, and what the oracle and the golden
fixtures use).  With AVX-512, numpy dispatches to its bundled SIMD library, which differs by up to 2 ulp
on about half of the inputs - enough to move ~2e-4 of the final path points.  "native" mode makes the
device follow THIS host's numpy instead: np.log10 is evaluated once over every float32 the features can
feed it (x in [1, 2^34), 34 binades of 2^23 values), compared with the device's glibc result, and the
differences are uploaded as a 4-bit-per-input table (143 MB of HBM).  "auto" does that only if the host's
np.log10 actually differs from glibc's.
"""
from __future__ import annotations

import numpy as np

BINADES = 34            # 1 + x with x a mean square of int16 samples: below 2^31; a few binades of head-room

_state = {"mode": "portable"}


def host_differs_from_glibc(ctx, samples: int = 200000, seed: int = 1) -> bool:
    """Does np.log10 (float32) on this host differ from the feature kernel's glibc restatement?"""
    rng = np.random.default_rng(seed)
    bits = rng.integers(0x3f800000, 0x3f800000 + (BINADES << 23), size=samples, dtype=np.int64).astype(np.uint32)
    x = bits.view(np.float32)
    return bool(np.any(np.log10(x).view(np.int32) != ctx.eval_log10f(x).view(np.int32)))


def build_correction(ctx) -> tuple[np.ndarray, int, dict]:
    """4-bit differences between the host's np.log10 and the device's glibc formula over [1, 2^BINADES)."""
    ctx.set_log10f_correction(None)
    count = BINADES << 23
    table = np.empty(count // 2, np.uint8)
    worst, changed = 0, 0
    for b in range(BINADES):
        bits = np.arange(0x3f800000 + (b << 23), 0x3f800000 + ((b + 1) << 23), dtype=np.uint32)
        x = bits.view(np.float32)
        d = np.log10(x).view(np.int32).astype(np.int64) - ctx.eval_log10f(x).view(np.int32).astype(np.int64)
        worst = max(worst, int(np.max(np.abs(d))))
        changed += int(np.count_nonzero(d))
        if worst > 7:
            raise RuntimeError(f"np.log10 differs from glibc's log10f by {worst} ulp: not a numpy SIMD log10 this table can describe")
        nib = (d + 8).astype(np.uint8)
        table[(b << 22):((b + 1) << 22)] = nib[0::2] | (nib[1::2] << 4)
    return table, count, {"max_ulp": worst, "inputs_that_differ": changed, "inputs": count}


def set_mode(mode: str, ctx=None) -> dict:
    """portable: glibc's log10f (default).  native: this host's np.log10.  auto: native only if it differs."""
    from . import api
    if mode not in ("portable", "native", "auto"):
        raise ValueError("log10 mode must be 'portable', 'native' or 'auto'")
    ctx = api.context() if ctx is None else ctx
    info = {"mode": mode}
    if mode == "portable":
        ctx.set_log10f_correction(None)
    else:
        ctx.set_log10f_correction(None)
        if mode == "auto" and not host_differs_from_glibc(ctx):
            info["mode"] = "portable"
            info["note"] = "np.log10 on this host is glibc's log10f: nothing to correct"
        else:
            table, count, stats = build_correction(ctx)
            ctx.set_log10f_correction(table, count)
            info.update(stats)
            info["mode"] = "native"
    _state.update(info)
    return info


def mode() -> str:
    return _state["mode"]
