"""The oracle against the UNMODIFIED reference, live, on inputs that are not in the committed fixtures.

Runs only where the reference is mounted (/root/reference, the authoring container) - on the GPU box the
committed fixtures under tests/golden/ (made the same way by tools/make_golden.py) stand in.  The comparison
runs in a subprocess because numpy's AVX-512 log10 dispatch has to be switched off before numpy is imported
("portable" oracle mode): that is the arithmetic the oracle and the CUDA kernels restate.
"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import json, sys
import numpy as np
sys.path.insert(0, %(root)r)
from describealign_b200 import host_fit, synth
from oracle import align_oracle as ao, features as of
from oracle.ref_loader import load_reference
import oracle
oracle.build()
da = load_reference()
seed, ch = %(seed)d, %(ch)d
if %(config)r:
    v, a = synth.config_pair(%(config)r, seed, 1.0)
else:
    v, a = synth.make_pair(%(video_s)f, %(offset_s)f, skips=%(skips)r, seed=seed, ch=ch)
out = {}
feats = {}
for name, pcm in (("video", v), ("audio", a)):
    arr = synth.as_reference_input(pcm)                      # float16 (ch, S), describealign.py:156
    ref = [da.get_energy(arr), da.get_zero_crossings(arr)] + list(da.get_freq_bands(arr))
    mine = of.all_features(pcm)
    feats[name] = ref
    out[name + "_f32_identical"] = bool(all(np.array_equal(np.asarray(r), m) for r, m in zip(ref[:4], mine[:4])))
    out[name + "_band2_rel"] = float(np.max(np.abs(np.asarray(ref[4]) - mine[4]) / np.maximum(np.abs(mine[4]), 1e-300)))
import io, contextlib
with contextlib.redirect_stdout(io.StringIO()):
    rx, ry, rsim, rpath, rmed = da.align(feats["video"], feats["audio"], feats["video"][0], feats["audio"][0])
V, A = of.all_features(v), of.all_features(a)
ox, oy, osim, opath, omed = ao.align(V, A, V[0], A[0], host_fit)
rpath = np.asarray(rpath, dtype=np.float64)
out["path_rows"] = [int(len(rpath)), int(len(opath))]
same = rpath.shape == opath.shape
out["path_int_identical"] = bool(same and np.array_equal(rpath[:, 1], opath[:, 1]) and np.array_equal(rpath[:, 2], opath[:, 2]))
out["path_float_max"] = float(np.max(np.abs(rpath[:, [0, 3, 4]] - opath[:, [0, 3, 4]]))) if same else None
out["nodes_max"] = float(max(np.max(np.abs(np.asarray(rx) - ox)), np.max(np.abs(np.asarray(ry) - oy)))) if len(rx) == len(ox) else None
out["similarity_diff"] = float(abs(rsim - osim))
out["median_slope_equal"] = bool(rmed == omed)
print("RESULT " + json.dumps(out))
'''


def _reference_present():
    return os.path.isfile("/root/reference/describealign.py")


@pytest.mark.skipif(not _reference_present(), reason="the reference is only mounted in the authoring container")
@pytest.mark.parametrize("seed,ch,video_s,offset_s,skips,config", [
    (301, 1, 70.0, 5.0, [(25.0, 2.0), (50.0, -1.0)], ""),
    (302, 2, 64.0, 3.0, [(30.0, 1.5)], ""),
    (303, 1, 0.0, 0.0, [], "C1"),        # BASELINE.json config 1 stand-in at full size: 179 s video, 378 s description
    (0, 1, 0.0, 0.0, [], "C2"),          # the bench workload itself at full size (22-min video, 27-min description)
])
def test_oracle_equals_reference_on_fresh_pairs(seed, ch, video_s, offset_s, skips, config):
    env = dict(os.environ)
    env["NPY_DISABLE_CPU_FEATURES"] = "AVX512F AVX512CD AVX512_SKX AVX512_CLX AVX512_CNL AVX512_ICL AVX512_SPR"
    env["OMP_NUM_THREADS"] = "1"
    code = CHILD % dict(root=ROOT, seed=seed, ch=ch, video_s=video_s, offset_s=offset_s, skips=skips, config=config)
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=900)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-3000:]
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    out = json.loads(line[len("RESULT "):])
    assert out["video_f32_identical"] and out["audio_f32_identical"], out       # energy, zero crossings, bands 0/1
    assert out["video_band2_rel"] < 1e-13 and out["audio_band2_rel"] < 1e-13, out   # f64 band: BLAS kernel order only
    assert out["path_int_identical"], out                                     # (audio i, cluster) of every path row
    assert out["path_float_max"] < 1e-8 and out["nodes_max"] < 1e-9, out
    assert out["similarity_diff"] < 1e-9 and out["median_slope_equal"], out


@pytest.mark.skipif(not _reference_present(), reason="the reference is only mounted in the authoring container")
@pytest.mark.parametrize("seed,ch,nodes", [
    (411, 1, ([1.5, 9.0, 17.2, 24.0], [0.0, 7.45, 16.0, 22.3])),
    (412, 2, ([2.0, 12.6, 20.0], [0.0, 10.0, 17.9])),
])
def test_stretch_host_logic_and_oracle_equal_the_reference_live(seed, ch, nodes):
    """replace_aligned_segments (describealign.py:229-416) of the unmodified reference against the product's host logic
    (native drift DP, numpy cross-fades) fed by the numpy oracle of the jump search, on pairs that are not in the goldens."""
    import contextlib
    import io
    import numpy as np
    sys.path.insert(0, ROOT)
    from describealign_b200 import stretch as st, synth
    from oracle import ref_loader, stretch_oracle as so
    da = ref_loader.load_reference()
    v, a = synth.make_pair(28.0, nodes[0][0], seed=seed, ch=ch)
    va, aa = synth.as_reference_input(v).copy(), synth.as_reference_input(a).copy()
    xt, yt = np.array(nodes[0]), np.array(nodes[1])

    def oracle_stretcher(segment, output):
        n_in, n_out = segment.shape[1], output.shape[1]
        jumps = st.jump_distances(n_out - n_in)
        loc, best = so.best_jumps(segment, n_out > n_in, jumps)
        orig = st.best_jumps
        st.best_jumps = lambda seg, neg, j: (loc, best)
        try:
            st.stretch(segment, output)
        finally:
            st.best_jumps = orig

    want, got = va.copy(), va.copy()
    with contextlib.redirect_stdout(io.StringIO()):
        da.replace_aligned_segments(want, aa, xt, yt, False)
        st.replace_aligned_segments(got, aa, xt, yt, False, stretcher=oracle_stretcher)
    assert int(np.sum(want != va)) > 100000                     # something was replaced
    assert np.array_equal(want.view(np.uint16), got.view(np.uint16))
