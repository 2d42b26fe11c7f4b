"""--stretch_audio resynthesis (describealign.py:229-416).

CPU: the host logic of describealign_b200.stretch (segment loop, jump distances, drift DP, traceback, cross-fades) with
the numpy oracle of the jump search reproduces the arrays the unmodified reference produced (golden SHA-256,
tools/make_golden_stretch.py), and - where /root/reference is mounted - the reference run live.
GPU: the CUDA jump search equals the oracle bit for bit, and the whole function equals the goldens."""
import contextlib
import hashlib
import io
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stretch_ref.json")


def _cases():
    with open(GOLD) as f:
        return json.load(f)["cases"]


def _arrays(case):
    from describealign_b200 import synth
    v, a = synth.make_pair(**case["make_pair"])
    return synth.as_reference_input(v).copy(), synth.as_reference_input(a).copy()


def _oracle_stretcher():
    from describealign_b200 import stretch as st
    from oracle import stretch_oracle as so

    def run(segment, output):
        n_in, n_out = segment.shape[1], output.shape[1]
        jumps = st.jump_distances(n_out - n_in)
        loc, best = so.best_jumps(segment, n_out > n_in, jumps)
        orig = st.best_jumps
        st.best_jumps = lambda seg, neg, j: (loc, best)
        try:
            st.stretch(segment, output)
        finally:
            st.best_jumps = orig
    return run


def _run(case, stretcher):
    from describealign_b200 import stretch as st
    va, aa = _arrays(case)
    with contextlib.redirect_stdout(io.StringIO()):
        st.replace_aligned_segments(va, aa, np.array(case["audio_times"]), np.array(case["video_times"]),
                                    case["no_pitch_correction"], stretcher=stretcher)
    return va


@pytest.mark.parametrize("name", sorted(_cases()))
def test_host_logic_with_oracle_jump_search_equals_reference_golden(name):
    case = _cases()[name]
    got = _run(case, _oracle_stretcher())
    assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).hexdigest() == case["sha256"]


def test_jump_distance_sets():
    from describealign_b200 import stretch as st
    assert st.jump_distances(20000) == list(st.BASE_JUMPS)
    assert st.jump_distances(-5000) == list(st.BASE_JUMPS) + [30, 31, 33, 37, 45, 61, 93, 157]
    assert st.jump_distances(700) == list(range(30, 512))


def test_oracle_pieces_cover_every_window_once():
    from oracle import stretch_oracle as so
    for n in (1535, 29286, 29287, 52224, 88200, 1000003):
        ps = so.pieces(n)
        assert sum(c for *_, c in ps) == n // 512
        assert all(length >= 3 * 512 - 1 for _, length, _, _ in ps)


@pytest.mark.gpu
@pytest.mark.parametrize("ch,n,negative,jumps,seed", [
    (1, 88200, False, "base", 1), (1, 88200, True, "base", 2), (2, 61000, True, "extended", 3),
    (2, 29287, False, "extended", 4), (1, 29286, True, "all", 5), (1, 1535, False, "base", 6), (1, 3000, True, "all", 7),
    (1, 400000, False, "base", 8),
])
def test_cuda_jump_search_equals_oracle(gpu_ctx, ch, n, negative, jumps, seed):
    from describealign_b200 import stretch as st
    from oracle import stretch_oracle as so
    rng = np.random.default_rng(seed)
    # programme-like signal with silent stretches and float16 scaling as combine() leaves it (describealign.py:1134-1148)
    x = (rng.normal(0, 3000, size=(ch, n)) * (1 + np.sin(np.arange(n) / 900.0))).astype(np.float32)
    x[:, n // 3:n // 3 + 2000] = 0
    x = (x / 1.37).astype(np.float16)
    jl = {"base": st.jump_distances(20000), "extended": st.jump_distances(5000), "all": st.jump_distances(500)}[jumps]
    loc, best = st.best_jumps(x, negative, jl)
    oloc, obest = so.best_jumps(x, negative, jl)
    assert loc.shape == oloc.shape and np.array_equal(loc, oloc)
    assert np.array_equal(best.view(np.int64), obest.view(np.int64))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(_cases()))
def test_replace_aligned_segments_on_the_gpu_equals_reference_golden(gpu_ctx, name):
    case = _cases()[name]
    got = _run(case, None)
    assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).hexdigest() == case["sha256"]


@pytest.mark.parametrize("seed,nw,total,jumps", [(1, 300, 4000, "extended"), (2, 300, -4000, "extended"), (3, 900, 30000, "base"),
                                                  (4, 900, -41000, "base"), (5, 180, 700, "all"), (6, 180, -333, "all"),
                                                  (7, 2, 40, "all")])
def test_native_drift_dp_equals_numpy(seed, nw, total, jumps):
    """dab_host_stretch_plan (C++) against the numpy statement of describealign.py:320-371 on random correlations."""
    from describealign_b200 import stretch as st
    rng = np.random.default_rng(seed)
    jl = {"base": st.jump_distances(20000), "extended": st.jump_distances(5000), "all": st.jump_distances(500)}[jumps]
    n_in = nw * 512 + int(rng.integers(0, 512))
    loc = rng.integers(0, 512, size=(nw, len(jl))).astype(np.int16)
    best = rng.uniform(-0.2, 1.0, size=(nw, len(jl)))
    best[rng.uniform(size=best.shape) < 0.02] = -np.inf          # windows without a valid position for a jump
    try:
        want = st.plan_jumps_numpy(n_in, n_in + total, jl, loc, best)
    except IndexError:
        with pytest.raises(IndexError):
            st.plan_jumps(n_in, n_in + total, jl, loc, best)
        return
    got = st.plan_jumps(n_in, n_in + total, jl, loc, best)
    assert got.shape == want.shape and np.array_equal(got, want)
