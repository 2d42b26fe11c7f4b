"""Native host stage (csrc/host_stage.cpp) against its numpy statement, bit for bit, and against
the reference's own intermediate values in tests/golden/ (kept path, fit points, LP arrays)."""
import numpy as np
import pytest
import scipy.sparse

from describealign_b200 import host_fit


def _random_path(rng, n, jitter=True):
    """A pass-1-like path: mostly unit steps with a few jumps, repeated audio indices and false matches."""
    x = np.cumsum(rng.integers(0, 3, size=n)).astype(np.int64) + 40
    y = (x * 0.97).astype(np.int64) + rng.integers(-2, 3, size=n)
    for at in rng.integers(n // 5, max(n // 5 + 1, n - n // 5), size=4):
        y[at:] += rng.integers(50, 400)
    if jitter:
        bad = rng.integers(0, n, size=n // 50)
        y[bad] += rng.integers(-3000, 3000, size=len(bad))
    return x, np.maximum(y, 0)


def _same(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.int64), b.view(np.int64))   # NaN == NaN, -0 != +0


@pytest.mark.parametrize("seed,n", [(1, 700), (2, 5000), (3, 120), (4, 51), (5, 20011)])
def test_continuity_error_native_equals_numpy(seed, n):
    rng = np.random.default_rng(seed)
    x, y = _random_path(rng, n)
    for deriv in (False, True):
        assert _same(host_fit.continuity_error(x, y, deriv=deriv), host_fit.continuity_error_numpy(x, y, deriv=deriv))
    # float input (the fit points of the LP)
    xf, yf = x.astype(np.float64) + rng.uniform(0, 1, n).round(3), y.astype(np.float64) + 0.25
    assert _same(host_fit.continuity_error(xf, yf, deriv=True), host_fit.continuity_error_numpy(xf, yf, deriv=True))


def test_continuity_error_flat_stretches_give_the_same_nans():
    x = np.arange(400, dtype=np.int64)
    x[100:160] = x[100]                  # zero denominators: 0/0 and x/0 as numpy produces them
    y = np.arange(400, dtype=np.int64) * 2
    with np.errstate(all="ignore"):
        want = host_fit.continuity_error_numpy(x, y)
    got = host_fit.continuity_error(x, y)
    assert _same(got, want) and np.isnan(want).any()


@pytest.mark.parametrize("seed,n", [(11, 300), (12, 4321), (13, 91), (14, 161), (15, 30000)])
def test_compress_path_native_equals_numpy(seed, n):
    rng = np.random.default_rng(seed)
    x, y = _random_path(rng, n, jitter=seed % 2 == 0)
    with np.errstate(all="ignore"):
        wx, wy = host_fit.compress_path_numpy(x, y)
    gx, gy = host_fit.compress_path(x, y)
    assert _same(gx, wx) and _same(gy, wy)


def test_compress_path_too_short_fails_like_the_reference():
    x = np.arange(90, dtype=np.int64)
    with pytest.raises(RuntimeError, match="Alignment failed"):
        host_fit.compress_path(x, x)
    with pytest.raises(RuntimeError, match="Alignment failed"):
        host_fit.compress_path_numpy(x, x)


@pytest.mark.parametrize("seed,n", [(21, 60), (22, 323), (23, 2194)])
def test_lp_assembly_native_equals_scipy(seed, n):
    rng = np.random.default_rng(seed)
    x = np.cumsum(rng.uniform(1, 80, size=n))
    y = x * 1.01 + np.cumsum(rng.normal(0, 0.3, size=n))
    c0, a0, b0, bd0 = host_fit._lp_problem_numpy(x, y)
    c1, a1, b1, bd1 = host_fit._lp_problem(x, y)
    assert _same(c1, c0) and _same(b1, b0) and bd1 == bd0
    a0 = scipy.sparse.csc_matrix(a0)
    a0.sort_indices()
    assert a1.shape == a0.shape
    assert np.array_equal(a1.indptr, a0.indptr) and np.array_equal(a1.indices, a0.indices) and _same(a1.data, a0.data)


@pytest.mark.parametrize("name", ["pair_a", "pair_warp"])
def test_native_host_stage_reproduces_the_references_intermediates(golden_align, name):
    """kept path -> fit points -> LP arrays, as recorded from the unmodified reference (tools/make_golden.py)."""
    data, _ = golden_align
    g = data[name]
    x, y = g["path1_x"].astype(np.int64), g["path1_y"].astype(np.int64)
    keep = host_fit.continuity_error(x, y) < 3
    assert np.array_equal(x[keep], g["kept_x"]) and np.array_equal(y[keep], g["kept_y"])
    fx, fy = host_fit.compress_path(x[keep], y[keep])
    assert _same(fx, g["fit_x"]) and _same(fy, g["fit_y"])
    cost, a_eq, b_eq, _ = host_fit._lp_problem(fx, fy)
    assert _same(cost, g["lp_c"]) and _same(b_eq, g["lp_b"])
    assert np.array_equal(a_eq.indptr, g["lp_indptr"]) and np.array_equal(a_eq.indices, g["lp_indices"])
    assert _same(a_eq.data, g["lp_data"])


def _random_fit(rng, n, segments):
    """A RateFit like the LP's output: piece-wise linear path with a few rate changes and jumps, small fit errors."""
    x = np.cumsum(rng.choice([1.0, 35.0, 70.0, 70.0, 12.5], size=n)) + 100.0
    bounds = np.sort(rng.choice(np.arange(5, n - 5), size=segments - 1, replace=False))
    slope_of = np.ones(n - 1)
    y = np.empty(n)
    y[0] = 40.0
    seg = 0
    cur = 1.0
    for k in range(1, n):
        if seg < len(bounds) and k == bounds[seg]:
            seg += 1
            cur = float(rng.choice([1.0, 1.000001, 0.98, 1.04, 0.05, 12.0]))
            y[k] = y[k - 1] + cur * (x[k] - x[k - 1]) + float(rng.choice([0.0, 400.0, -90.0, 2.4]))
        else:
            y[k] = y[k - 1] + cur * (x[k] - x[k - 1])
        slope_of[k - 1] = cur
    slopes = slope_of + rng.normal(0, 2e-7, size=n - 1) * (rng.uniform(size=n - 1) < 0.3)
    fit_err = rng.normal(0, 0.3, size=n) * (rng.uniform(size=n) < 0.2)
    return host_fit.RateFit(x=x, y=y + fit_err, fit_err=fit_err, slopes=slopes, median_slope=1.0)


@pytest.mark.parametrize("seed,n,segments", [(31, 60, 2), (32, 400, 5), (33, 2194, 9), (34, 2194, 40), (35, 7, 1)])
def test_line_clusters_native_equals_python(seed, n, segments):
    rng = np.random.default_rng(seed)
    fit = _random_fit(rng, n, segments)
    want = host_fit.line_clusters_numpy(fit)
    got = host_fit.line_clusters(fit)
    assert len(got) == len(want)
    for (gx, go, gs), (wx, wo, ws) in zip(got, want):
        assert _same(gx, wx) and _same(go, wo) and _same(gs, ws)


@pytest.mark.parametrize("name", ["pair_a", "pair_warp"])
def test_line_clusters_native_reproduces_the_references_clusters(golden_align, name):
    """The LP solution recorded from the reference -> the clusters the reference formed from it."""
    data, _ = golden_align
    g = data[name]
    n = len(g["fit_x"])
    sol = g["lp_x"]
    fit_err = sol[:n] - sol[n:2 * n]
    jumps = sol[8 * n - 4:9 * n - 5] - sol[9 * n - 5:10 * n - 6]
    fit = host_fit.RateFit(x=g["fit_x"], y=g["fit_y"], fit_err=fit_err, slopes=sol[-1] + jumps / np.diff(g["fit_x"]),
                           median_slope=sol[-1])
    assert _same(fit.slopes, g["slopes"])
    clusters = host_fit.line_clusters(fit)
    assert len(clusters) == int(g["n_clusters"])
    for k, (cx, offset, slope) in enumerate(clusters[:3]):
        if f"cluster{k}_x" in g:
            assert _same(cx, g[f"cluster{k}_x"]) and _same(np.array([offset, slope]), g[f"cluster{k}_line"])
