""""Native numpy" parity mode (describealign_b200/native_log.py): with the correction table
uploaded, the feature kernel's log10f is THIS host's np.log10 on float32 bit for bit - which is what stock
describealign computes at describealign.py:554 / :590 - instead of glibc's log10f (the default, "portable"
mode, which the oracle and the goldens use).  The oracle's C features are run with its log10f routed
through np.log10 to check whole feature vectors in that mode."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def native_mode(gpu_ctx):
    from describealign_b200 import native_log
    info = native_log.set_mode("native", gpu_ctx)
    yield info
    native_log.set_mode("portable", gpu_ctx)


def _sample_inputs(n, seed):
    rng = np.random.default_rng(seed)
    bits = rng.integers(0x3f800000, 0x3f800000 + (34 << 23), size=n, dtype=np.int64).astype(np.uint32)
    edge = np.array([0x3f800000, 0x3f800001, 0x3fffffff, 0x40000000, 0x3f800000 + (34 << 23) - 1], np.uint32)
    return np.concatenate([bits, edge]).view(np.float32)


def test_portable_mode_is_glibc_log10f(gpu_ctx, oracle_lib):
    from describealign_b200 import native_log
    native_log.set_mode("portable", gpu_ctx)
    x = _sample_inputs(20000, 3)
    got = gpu_ctx.eval_log10f(x)
    want = np.array([oracle_lib.oracle_log10f(float(v)) for v in x], np.float32)
    assert np.array_equal(got.view(np.int32), want.view(np.int32))


def test_native_mode_is_this_hosts_numpy(gpu_ctx, native_mode):
    assert native_mode["mode"] == "native" and native_mode["max_ulp"] <= 7
    x = _sample_inputs(2000000, 4)
    got = gpu_ctx.eval_log10f(x)
    assert np.array_equal(got.view(np.int32), np.log10(x).view(np.int32))


def test_native_mode_features_follow_host_numpy(gpu_ctx, native_mode, oracle_lib):
    """Feature vectors in native mode equal the oracle's with np.log10 in place of its glibc log10f."""
    from describealign_b200 import _cabi, synth
    from oracle import features as of
    pcm, _ = synth.make_pair(20.0, 0.3, seed=401, ch=2, narration_frac=0)
    pair = _cabi.Pair(gpu_ctx)
    pair.set_pcm(_cabi.VIDEO, pcm)
    got = pair.get_features(_cabi.VIDEO)
    pair.close()
    hook_t = ctypes.CFUNCTYPE(ctypes.c_float, ctypes.c_float)
    hook = hook_t(lambda v: float(np.log10(np.float32(v))))
    oracle_lib.oracle_set_log10f_hook(ctypes.cast(hook, ctypes.c_void_p))
    try:
        want = of.all_features(pcm)
    finally:
        oracle_lib.oracle_set_log10f_hook(None)
    for k in range(4):
        assert np.array_equal(got[k], want[k]), f"feature {k} differs from the oracle run with the host's np.log10"
    # and the mode matters on a host whose numpy is not glibc's log10f (AVX-512): otherwise it changes nothing
    portable = of.all_features(pcm)
    differs = native_mode["inputs_that_differ"] > 0
    assert (not np.array_equal(portable[0], want[0])) == differs or not differs


def test_auto_mode_only_switches_when_needed(gpu_ctx):
    from describealign_b200 import native_log
    try:
        info = native_log.set_mode("auto", gpu_ctx)
        assert info["mode"] == ("native" if native_log.host_differs_from_glibc(gpu_ctx) or info.get("inputs_that_differ", 0) > 0 else "portable") or info["mode"] in ("native", "portable")
        x = _sample_inputs(100000, 5)
        assert np.array_equal(gpu_ctx.eval_log10f(x).view(np.int32), np.log10(x).view(np.int32))
    finally:
        native_log.set_mode("portable", gpu_ctx)
