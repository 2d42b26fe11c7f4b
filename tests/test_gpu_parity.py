"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle and the
golden fixtures made from the reference (tests/golden/, tools/make_golden.py).

Bars: float32 features bit-exact; band 2 (float64) within 4 ulp; integer outputs (match
points, pass-1 path, pass-2 (i, int(j), cluster)) identical; float64 quals rel 1e-12 (pow /
log10 are the device's libm, not glibc's); node times within 1e-9 s.
"""
import numpy as np
import pytest

from conftest import golden_pair_pcm

pytestmark = pytest.mark.gpu


def _features_gpu(ctx, pcm):
    from describealign_b200 import _cabi
    pair = _cabi.Pair(ctx)
    pair.set_pcm(_cabi.VIDEO, pcm)
    f = pair.get_features(_cabi.VIDEO)
    pair.close()
    return f


def _check_features(got, want, what):
    names = ["energy", "zero_crossings", "band0", "band1", "band2"]
    for k in range(5):
        assert got[k].shape == want[k].shape, f"{what}: {names[k]} length {got[k].shape} != {want[k].shape}"
    for k in range(4):
        bad = np.flatnonzero(got[k] != want[k])
        assert bad.size == 0, f"{what}: {names[k]} differs at {bad[:5]} ({bad.size} of {got[k].size})"
    err = np.abs(got[4] - want[4])
    assert np.all(err <= 4 * np.spacing(np.abs(want[4]))), f"{what}: band2 max err {err.max()}"


def test_features_golden(gpu_ctx, golden_features):
    """CUDA features == the reference's own outputs (portable mode) on the golden clips."""
    from describealign_b200 import synth
    data, meta = golden_features
    for case in meta["cases"]:
        pcm, _ = synth.make_pair(case["seconds"], 0.5, seed=case["seed"], ch=case["ch"], narration_frac=0)
        got = _features_gpu(gpu_ctx, pcm)
        want = [data[f"{case['name']}.{k}"] for k in ("energy", "zc", "b0", "b1", "b2")]
        _check_features(got, want, case["name"])


def test_features_extremes(gpu_ctx, golden_features):
    data, _ = golden_features
    S = 210 * 60 + 17
    ext = {"silence": np.zeros((S, 1), np.int16),
           "square": (np.where((np.arange(S) // 3) % 2 == 0, 32767, -32768).astype(np.int16))[:, None],
           "impulse": np.zeros((S, 1), np.int16)}
    ext["impulse"][S // 2, 0] = 12345
    for name, pcm in ext.items():
        got = _features_gpu(gpu_ctx, pcm)
        want = [data[f"{name}.{k}"] for k in ("energy", "zc", "b0", "b1", "b2")]
        _check_features(got, want, name)


@pytest.mark.parametrize("ch,seconds,seed", [(1, 33.37, 101), (2, 21.013, 102), (1, 1.0, 103), (2, 0.5, 104)])
def test_features_vs_oracle(gpu_ctx, ch, seconds, seed):
    from describealign_b200 import synth
    from oracle import features as of
    pcm, _ = synth.make_pair(seconds, 0.2, seed=seed, ch=ch, narration_frac=0)
    _check_features(_features_gpu(gpu_ctx, pcm), of.all_features(pcm), f"ch{ch}")
    # float16 input (the reference's own array type) gives the same result as int16 input
    f16 = np.ascontiguousarray(pcm).astype(np.float16)
    _check_features(_features_gpu(gpu_ctx, f16), of.all_features(pcm), f"ch{ch} f16")


def test_features_tiny_and_empty(gpu_ctx):
    from oracle import features as of
    rng = np.random.default_rng(5)
    for S in (0, 1, 104, 105, 209, 210, 211, 419, 420, 1000):
        pcm = rng.integers(-3000, 3000, size=(S, 1)).astype(np.int16)
        got = _features_gpu(gpu_ctx, pcm)
        want = of.all_features(pcm)
        _check_features(got, want, f"S={S}")


def _stage_a_gpu(ctx, V, A):
    from describealign_b200 import _cabi
    pair = _cabi.Pair(ctx)
    pair.set_features(_cabi.VIDEO, V)
    pair.set_features(_cabi.AUDIO, A)
    pair.stage_a()
    return pair


def test_stage_a_golden(gpu_ctx, golden_align):
    """Match points and the pass-1 path equal the reference's on the golden pairs."""
    from oracle import features as of
    data, meta = golden_align
    for name in meta["cases"]:
        g = data[name]
        v, a = golden_pair_pcm(meta, name)
        V, A = of.all_features(v), of.all_features(a)
        pair = _stage_a_gpu(gpu_ctx, V, A)
        pi, pv, pq = pair.points1()
        assert np.array_equal(pi, g["points1_i"]) and np.array_equal(pv, g["points1_v"]), name
        np.testing.assert_allclose(pq, g["points1_q"], rtol=1e-12, atol=0)
        x, y = pair.path1()
        assert np.array_equal(x, g["path1_x"]) and np.array_equal(y, g["path1_y"]), name
        pair.close()


def test_end_to_end_golden(gpu_ctx, golden_align):
    """PCM in -> nodes out through the public API equals the reference's result."""
    from describealign_b200 import api
    data, meta = golden_align
    for name in meta["cases"]:
        g = data[name]
        v, a = golden_pair_pcm(meta, name)
        det = {}
        nx, ny, sim, path, med = api.align_pcm(v, a, details=det)
        gp = g["path2"]
        assert path.shape == gp.shape, name
        assert np.array_equal(path[:, 1] * 210, gp[:, 1]) or np.allclose(path[:, 1] * 210, gp[:, 1], atol=1e-6)
        assert np.array_equal(path[:, 2], gp[:, 2]), name
        np.testing.assert_allclose(path[:, 0] * 210, gp[:, 0], rtol=0, atol=1e-7)
        np.testing.assert_allclose(path[:, 3], gp[:, 3], rtol=0, atol=1e-9)
        np.testing.assert_allclose(nx, g["nodes_x"], rtol=0, atol=1e-9)
        np.testing.assert_allclose(ny, g["nodes_y"], rtol=0, atol=1e-9)
        assert abs(sim - float(g["similarity"])) < 1e-9


def test_end_to_end_vs_oracle(gpu_ctx):
    """A pair that is not in the fixtures: CUDA path vs oracle on identical PCM."""
    from describealign_b200 import api, host_fit, synth
    from oracle import align_oracle as ao, features as of
    v, a = synth.make_pair(140.0, 9.0, skips=[(50.0, 2.0), (90.0, -3.0)], seed=77, ch=2)
    det = {}
    nx, ny, sim, path, med = api.align_pcm(v, a, details=det)
    V, A = of.all_features(v), of.all_features(a)
    odet = {}
    ox, oy, osim, opath, omed = ao.align(V, A, V[0], A[0], host_fit, details=odet)
    assert path.shape == opath.shape
    assert np.array_equal(path[:, 1], opath[:, 1]) and np.array_equal(path[:, 2], opath[:, 2])
    np.testing.assert_allclose(path[:, 0], opath[:, 0], rtol=0, atol=1e-9)
    np.testing.assert_allclose(path[:, 3:], opath[:, 3:], rtol=0, atol=1e-8)
    np.testing.assert_allclose(nx, ox, rtol=0, atol=1e-9)
    np.testing.assert_allclose(ny, oy, rtol=0, atol=1e-9)
    assert abs(sim - osim) < 1e-9 and med == omed


def test_mismatched_inputs_fail_like_the_reference(gpu_ctx):
    from describealign_b200 import api, synth
    v, _ = synth.make_pair(60.0, 1.0, seed=201)
    _, a = synth.make_pair(60.0, 1.0, seed=202)
    with pytest.raises(RuntimeError, match="Alignment failed, are the input files mismatched"):
        api.align_pcm(v, a)
