"""CPU tests that pin the oracle (oracle/) against the golden fixtures produced by running
the UNMODIFIED reference in the authoring container (tools/make_golden.py, portable mode).

The reference has no tests or golden vectors of its own for this path,
so these fixtures are the pin: every stage boundary of describealign.py:545-1027 that the
oracle restates is compared with what the reference itself produced on the same PCM.
"""
import numpy as np
import pytest

from conftest import golden_pair_pcm

FEATURE_KEYS = ("energy", "zc", "b0", "b1", "b2")


def _check_features(got, want, what):
    for k in range(5):
        assert got[k].shape == want[k].shape, f"{what}: feature {k} length"
        assert got[k].dtype == want[k].dtype, f"{what}: feature {k} dtype"
    for k in range(4):   # float32 features: bit-exact
        bad = np.flatnonzero(got[k] != want[k])
        assert bad.size == 0, f"{what}: feature {k} differs at {bad[:5]} ({bad.size} of {got[k].size})"
    err = np.abs(got[4] - want[4])   # float64 band: OpenBLAS kernel choice moves the last bit
    assert np.all(err <= 4 * np.spacing(np.abs(want[4]))), f"{what}: band2 max err {err.max()}"


def test_oracle_features_match_reference(oracle_lib, golden_features):
    """describealign.py:545-593 on five synthetic clips (mono/stereo, S mod 210 on both sides of 105)."""
    from describealign_b200 import synth
    from oracle import features as of
    data, meta = golden_features
    assert "portable" in meta["env"]["mode"]
    for case in meta["cases"]:
        pcm, _ = synth.make_pair(case["seconds"], 0.5, seed=case["seed"], ch=case["ch"], narration_frac=0)
        assert pcm.shape[0] == case["samples"]
        want = [data[f"{case['name']}.{k}"] for k in FEATURE_KEYS]
        _check_features(of.all_features(pcm), want, case["name"])


def test_oracle_features_extremes(oracle_lib, golden_features):
    """silence, a full-scale 7.35 kHz square wave and a single impulse."""
    from oracle import features as of
    data, _ = golden_features
    S = 210 * 60 + 17
    ext = {"silence": np.zeros((S, 1), np.int16),
           "square": (np.where((np.arange(S) // 3) % 2 == 0, 32767, -32768).astype(np.int16))[:, None],
           "impulse": np.zeros((S, 1), np.int16)}
    ext["impulse"][S // 2, 0] = 12345
    for name, pcm in ext.items():
        want = [data[f"{name}.{k}"] for k in FEATURE_KEYS]
        _check_features(of.all_features(pcm), want, name)


def test_oracle_feature_lengths_ragged(oracle_lib):
    """len(energy) = ceil((S // 105) / 2), the others S // 210 (describealign.py:548-555, 560-566)."""
    from oracle import features as of
    rng = np.random.default_rng(3)
    for S in (0, 1, 104, 105, 209, 210, 211, 314, 315, 419, 420, 2100 + 104, 2100 + 105):
        pcm = rng.integers(-2000, 2000, size=(S, 1)).astype(np.int16)
        f = of.all_features(pcm)
        assert len(f[0]) == (S // 105 + 1) // 2
        assert all(len(x) == S // 210 for x in f[1:])


def test_oracle_log10f_is_glibc(oracle_lib):
    """The restated glibc log10f against the host libm on 200 000 inputs >= 1."""
    import ctypes
    import ctypes.util
    libm = ctypes.CDLL(ctypes.util.find_library("m"))
    libm.log10f.restype = ctypes.c_float
    libm.log10f.argtypes = [ctypes.c_float]
    rng = np.random.default_rng(0)
    xs = np.concatenate([1.0 + rng.random(50000, dtype=np.float32),
                         np.exp(rng.uniform(0, 23, 150000)).astype(np.float32)])
    for x in xs[::40]:
        assert oracle_lib.oracle_log10f(float(x)) == libm.log10f(float(x)), x


@pytest.mark.parametrize("name", ["pair_a", "pair_warp"])
def test_oracle_stage_a_matches_reference(oracle_lib, golden_align, name):
    """Match points (i, v, qual), back pointers' result (the pass-1 path) vs the reference
    (describealign.py:596-700)."""
    from oracle import align_oracle as ao, features as of
    data, meta = golden_align
    g = data[name]
    v, a = golden_pair_pcm(meta, name)
    V, A = of.all_features(v), of.all_features(a)
    sa = ao.stage_a(V, A, V[0], A[0])
    assert np.array_equal(sa["points_i"], g["points1_i"])
    assert np.array_equal(sa["points_v"], g["points1_v"])
    np.testing.assert_allclose(sa["points_q"], g["points1_q"], rtol=1e-12, atol=0)
    assert np.array_equal(sa["path_x"], g["path1_x"])
    assert np.array_equal(sa["path_y"], g["path1_y"])


@pytest.mark.parametrize("name", ["pair_a", "pair_warp"])
def test_oracle_align_matches_reference(oracle_lib, golden_align, name):
    """Whole align(): filter, fit points, LP solution, clusters, pass-2 points, final path,
    nodes and similarity vs the reference (describealign.py:702-1027)."""
    from describealign_b200 import host_fit
    from oracle import align_oracle as ao, features as of
    data, meta = golden_align
    g = data[name]
    v, a = golden_pair_pcm(meta, name)
    V, A = of.all_features(v), of.all_features(a)
    det = {}
    nx, ny, sim, path, med = ao.align(V, A, V[0], A[0], host_fit, details=det)
    assert np.array_equal(det["kept_x"], g["kept_x"]) and np.array_equal(det["kept_y"], g["kept_y"])
    np.testing.assert_array_equal(det["fit"].x, g["fit_x"])
    np.testing.assert_array_equal(det["fit"].y, g["fit_y"])
    np.testing.assert_allclose(det["fit"].slopes, g["slopes"], rtol=0, atol=1e-12)
    assert len(det["clusters"]) == int(g["n_clusters"])
    for k, (cx, off, sl) in enumerate(det["clusters"]):
        np.testing.assert_array_equal(cx, g[f"cluster{k}_x"])
        np.testing.assert_allclose([off, sl], g[f"cluster{k}_line"], rtol=1e-12, atol=1e-9)
    np.testing.assert_array_equal(det["audio_scaled"][:64], g["scaled_audio_head"])
    np.testing.assert_array_equal(det["video_scaled"][:64], g["scaled_video_head"])
    b = det["stage_b"]
    p2 = g["points2"]
    assert len(b["points_i"]) == len(p2)
    assert np.array_equal(b["points_i"], p2[:, 0]) and np.array_equal(b["points_c"], p2[:, 2])
    np.testing.assert_allclose(b["points_j"], p2[:, 1], rtol=0, atol=1e-9)
    np.testing.assert_allclose(b["points_q"], p2[:, 3], rtol=0, atol=1e-9)
    gp = g["path2"]          # captured before the /210 scaling
    assert path.shape == gp.shape
    np.testing.assert_array_equal(np.rint(path[:, 1] * 210), gp[:, 1])
    np.testing.assert_array_equal(path[:, 2], gp[:, 2])
    np.testing.assert_allclose(path[:, 0] * 210, gp[:, 0], rtol=0, atol=1e-7)
    np.testing.assert_allclose(path[:, 3], gp[:, 3], rtol=0, atol=1e-9)
    np.testing.assert_allclose(nx, g["nodes_x"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(ny, g["nodes_y"], rtol=0, atol=1e-9)
    assert abs(sim - float(g["similarity"])) < 1e-9
    assert abs(med - float(g["median_slope"])) < 1e-12


def test_host_lp_matches_reference(golden_align):
    """The LP assembled by host_fit (describealign.py:769-836) equals the matrix the reference
    handed to linprog, entry for entry."""
    import scipy.sparse
    from describealign_b200 import host_fit
    data, meta = golden_align
    for name in meta["cases"]:
        g = data[name]
        cost, a_eq, b_eq, bounds = host_fit._lp_problem(g["fit_x"], g["fit_y"])
        m = scipy.sparse.csc_matrix(a_eq)
        m.sum_duplicates()
        m.sort_indices()
        np.testing.assert_array_equal(m.indptr, g["lp_indptr"])
        np.testing.assert_array_equal(m.indices, g["lp_indices"])
        np.testing.assert_array_equal(m.data, g["lp_data"])
        np.testing.assert_array_equal(cost, g["lp_c"])
        np.testing.assert_array_equal(b_eq, g["lp_b"])
        assert len(bounds) == len(cost)


def test_literal_dp1_equals_prefix_max(oracle_lib):
    """The sorted-frontier-list DP written the way the reference does it (describealign.py:
    654-682) and the prefix-max restatement in C agree on tie-heavy random instances."""
    from oracle import align_oracle as ao
    import ctypes
    rng = np.random.default_rng(11)
    for trial in range(60):
        n_rows, n_cols = int(rng.integers(1, 40)), int(rng.integers(1, 25))
        pts = set()
        for _ in range(int(rng.integers(1, 200))):
            pts.add((int(rng.integers(0, n_rows)), int(rng.integers(0, n_cols))))
        pts = sorted(pts)
        q = rng.integers(1, 4, size=len(pts)).astype(np.float64)   # small integers: many exact ties
        lit = ao.literal_dp1([(i, v, float(w)) for (i, v), w in zip(pts, q)])
        pi = np.array([p[0] for p in pts], np.int32)
        pv = np.array([p[1] for p in pts], np.int32)
        n = len(pts)
        path_i = np.zeros(n, np.int32); path_v = np.zeros(n, np.int32)
        cum = np.zeros(n); back = np.zeros(n, np.int32)
        _p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        plen = oracle_lib.oracle_dp1(_p(pi), _p(pv), _p(q), n, n_cols, _p(path_i), _p(path_v), _p(cum), _p(back))
        assert [(int(v), int(i)) for v, i in zip(path_v[:plen], path_i[:plen])] == [(int(v), int(i)) for v, i in lit], trial


def test_literal_dp2_equals_restatement(oracle_lib):
    """Same for DP #2 (describealign.py:946-983) on small random corridors."""
    from oracle import align_oracle as ao
    rng = np.random.default_rng(12)
    for trial in range(25):
        n_rows, n_video, n_clusters = int(rng.integers(5, 60)), 80, int(rng.integers(1, 4))
        rows = []
        for i in range(n_rows):
            row, cells = [], set()
            for c in range(n_clusters):
                if rng.random() < 0.8:
                    j = float(np.round(rng.uniform(4, n_video - 5), 3)) if rng.random() < 0.3 else 4.0 + 0.9 * i + 3 * c
                    if j < n_video - 4 and int(j) not in cells:
                        cells.add(int(j))
                        row.append((j, c, float(rng.integers(0, 5))))
            rows.append(sorted(row))
        lit = ao.literal_dp2(rows, n_clusters, n_video)
        pi = np.array([i for i, r in enumerate(rows) for _ in r], np.int32)
        pj = np.array([p[0] for r in rows for p in r], np.float64)
        pc = np.array([p[1] for r in rows for p in r], np.int32)
        pq = np.array([p[2] for r in rows for p in r], np.float64)
        if len(pi) == 0:
            continue
        got = ao.dp2(pi, pj, pc, pq, n_clusters, n_video)
        assert got.shape == lit.shape, trial
        np.testing.assert_array_equal(got, lit)
