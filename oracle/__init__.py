"""TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy + plain C) of the alignment hot path of julbean/describealign
(reference describealign.py:545-1027).  It exists to check the CUDA path and to provide
the CPU baseline in bench.py.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it; the product package
(describealign_b200) never does and fails loudly without its CUDA library.

Parity status: the reference has no golden vectors or tests for this path (SURVEY.md
section 4).  The oracle is pinned against outputs of the reference itself, produced in the
authoring container by tools/make_golden.py and committed under tests/golden/.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    """Compile oracle/c/*.c into oracle/liboracle.so with gcc (needs only libc/libm)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = sorted(os.path.join(_HERE, "c", f) for f in os.listdir(os.path.join(_HERE, "c")) if f.endswith(".c"))
    stale = force or not os.path.isfile(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        i64, p, i32 = ctypes.c_int64, ctypes.c_void_p, ctypes.c_int
        _LIB.oracle_energy.restype = i64
        _LIB.oracle_energy.argtypes = [p, i64, i32, p, p]
        _LIB.oracle_zero_crossings.restype = i64
        _LIB.oracle_zero_crossings.argtypes = [p, i64, i32, p, p]
        _LIB.oracle_freq_bands.restype = i64
        _LIB.oracle_freq_bands.argtypes = [p, i64, i32, p, p, p, p, p, p, p]
        _LIB.oracle_log10f.restype = ctypes.c_float
        _LIB.oracle_log10f.argtypes = [ctypes.c_float]
        _LIB.oracle_set_log10f_hook.restype = None
        _LIB.oracle_set_log10f_hook.argtypes = [p]
        _LIB.oracle_ddot.restype = ctypes.c_double
        _LIB.oracle_ddot.argtypes = [p, p, i64]
        _LIB.oracle_meansub_norm.restype = None
        _LIB.oracle_meansub_norm.argtypes = [p, i64, p, p, p]
        _LIB.oracle_codes.restype = None
        _LIB.oracle_codes.argtypes = [p, p, i64, i32, p, p]
        _LIB.oracle_match.restype = i64
        _LIB.oracle_match.argtypes = [p, p, p, p, i64, p, p, p, p, p, i64, i64, p, p, p, i64, p]
        _LIB.oracle_dp1.restype = i64
        _LIB.oracle_dp1.argtypes = [p, p, p, i64, i64, p, p, p, p]
        _LIB.oracle_dp2.restype = i64
        _LIB.oracle_dp2.argtypes = [p, p, p, p, p, i64, i64, i64, i64, p]
    return _LIB
