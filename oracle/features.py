"""TEST INFRASTRUCTURE ONLY - oracle for the feature functions.

Python front for oracle/c/oracle_features.c.  Mirrors the reference's call signatures
(describealign.py:545-593) but takes the interleaved int16 PCM the decoder produced
(describealign.py:156 turns it into float16; the C code applies the same rounding).
"""
from __future__ import annotations

import ctypes

import numpy as np
import scipy.signal

from . import lib


def hann_f32(taps: int) -> np.ndarray:
    """Normalised float32 Hann window with the end zeros removed, built with the very
    numpy/scipy expression the reference uses (describealign.py:551-552, 569-570)."""
    w = scipy.signal.windows.hann(taps + 2)[1:-1].astype(np.float32)
    w = w / np.sum(w)
    return np.ascontiguousarray(w)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _pcm(pcm: np.ndarray):
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    if pcm.ndim == 1:
        pcm = pcm[:, None]
    return pcm, pcm.shape[0], pcm.shape[1]


def get_energy(pcm: np.ndarray) -> np.ndarray:
    pcm, S, ch = _pcm(pcm)
    out = np.zeros(S // 210 + 2, np.float32)
    n = lib().oracle_energy(_p(pcm), S, ch, _p(hann_f32(13)), _p(out))
    return out[:n].copy()


def get_zero_crossings(pcm: np.ndarray) -> np.ndarray:
    pcm, S, ch = _pcm(pcm)
    out = np.zeros(S // 210 + 1, np.float32)
    n = lib().oracle_zero_crossings(_p(pcm), S, ch, _p(hann_f32(13)), _p(out))
    return out[:n].copy()


def get_freq_bands(pcm: np.ndarray):
    pcm, S, ch = _pcm(pcm)
    L = S // 210
    b0 = np.zeros(L, np.float32)
    b1 = np.zeros(L, np.float32)
    b2 = np.zeros(L, np.float64)
    lib().oracle_freq_bands(_p(pcm), S, ch, _p(hann_f32(15)), _p(hann_f32(21)), _p(hann_f32(630)),
                            _p(hann_f32(90)), _p(b0), _p(b1), _p(b2))
    return [b0, b1, b2]


def all_features(pcm: np.ndarray):
    """[energy, zero_crossings, band0, band1, band2] in the order describealign.py:1104 uses."""
    return [get_energy(pcm), get_zero_crossings(pcm), *get_freq_bands(pcm)]
