import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: generates full-size PCM and runs the CPU oracle on it (about a minute each)")


@pytest.fixture(scope="session")
def golden_features():
    data = np.load(os.path.join(GOLDEN, "features_ref.npz"))
    with open(os.path.join(GOLDEN, "features_ref.json")) as f:
        meta = json.load(f)
    return data, meta


@pytest.fixture(scope="session")
def golden_align():
    with open(os.path.join(GOLDEN, "align_ref.json")) as f:
        meta = json.load(f)
    data = {name: np.load(os.path.join(GOLDEN, f"align_{name}.npz")) for name in meta["cases"]}
    return data, meta


_pair_cache = {}


def golden_pair_pcm(meta, name):
    """Regenerate the PCM of a golden align case (deterministic generator) and check its hash."""
    import hashlib
    from describealign_b200 import synth
    if name not in _pair_cache:
        kw = dict(meta["cases"][name]["make_pair"])
        kw["skips"] = [tuple(s) for s in kw.get("skips", [])]
        kw["warps"] = [tuple(w) for w in kw.get("warps", [])]
        v, a = synth.make_pair(**kw)
        sha = lambda x: hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest()
        assert sha(v) == meta["cases"][name]["video_pcm_sha256"], "synthetic generator is not reproducing the golden PCM"
        assert sha(a) == meta["cases"][name]["audio_pcm_sha256"]
        _pair_cache[name] = (v, a)
    return _pair_cache[name]


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle
    oracle.build()
    return oracle.lib()


@pytest.fixture(scope="session")
def gpu_ctx():
    from describealign_b200 import _cabi, build
    build.build()
    return _cabi.Context(-1)
