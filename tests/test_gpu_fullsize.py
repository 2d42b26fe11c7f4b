"""The named BASELINE.json configurations at (or near) full size on the GPU, against the oracle on the
same PCM: C3 = the 22/27-min pair in stereo (--stretch_audio feature semantics), C4 = one 45-min
episode of the batch config, C5 = the long-form pair at quarter scale (37.5 min / 45 min; the full 2.5 h
pair needs more host memory to generate than a test should take).  C2 at full size is compared inside
every bench.py run ("parity").  Slow: each case generates its PCM (~1 min of CPU) and runs the CPU
oracle (seconds)."""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.slow]


def _compare(v, a):
    from describealign_b200 import api, host_fit
    from oracle import align_oracle as ao, features as of
    det = {}
    nx, ny, sim, path, med = api.align_pcm(v, a, details=det)
    V, A = of.all_features(v), of.all_features(a)
    for k in range(4):
        assert np.array_equal(det["video_features"][k], V[k]) and np.array_equal(det["audio_features"][k], A[k]), f"feature {k}"
    ox, oy, osim, opath, omed = ao.align(V, A, V[0], A[0], host_fit)
    assert path.shape == opath.shape, (path.shape, opath.shape)
    assert np.array_equal(path[:, 1], opath[:, 1]) and np.array_equal(path[:, 2], opath[:, 2])
    np.testing.assert_allclose(path[:, 0], opath[:, 0], rtol=0, atol=1e-9)
    np.testing.assert_allclose(nx, ox, rtol=0, atol=1e-9)
    np.testing.assert_allclose(ny, oy, rtol=0, atol=1e-9)
    assert abs(sim - osim) < 1e-9 and med == omed
    return det


@pytest.mark.parametrize("config,scale,seed", [("C3", 1.0, 0), ("C4", 1.0, 3), ("C5", 0.25, 0)])
def test_named_config_vs_oracle(gpu_ctx, config, scale, seed):
    from describealign_b200 import synth
    v, a = synth.config_pair(config, seed, scale)
    if config == "C3":
        assert v.shape[1] == 2
    det = _compare(v, a)
    assert det["stats"]["n_path2"] > 0.3 * min(len(v), len(a)) / 210


def test_c4_pairs_through_the_batch_engine(gpu_ctx):
    """Four quarter-scale C4 episodes through batch.align_batch (the engine path batch mode uses) against
    the synchronous API."""
    from describealign_b200 import api, batch, synth
    pairs = [synth.config_pair("C4", 10 + k, 0.25) for k in range(4)]
    got = batch.align_batch(pairs, in_flight=4)
    for p, g in zip(pairs, got):
        want = api.align_pcm(*p)
        assert not isinstance(g, Exception), g
        np.testing.assert_array_equal(g[3], want[3])
        np.testing.assert_array_equal(g[0], want[0])
        np.testing.assert_array_equal(g[1], want[1])
