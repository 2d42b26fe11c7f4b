"""Host-side mirror of the reference's hot-path interface.

Same names, arguments, return values, progress output and errors as the functions of
julbean/describealign that `combine()` calls between decoding the PCM and writing the
output (reference describealign.py:1098-1122):

    get_energy(arr)            describealign.py:545
    get_zero_crossings(arr)    describealign.py:557
    get_freq_bands(arr)        describealign.py:575
    align(video_features, audio_desc_features, video_energy, audio_desc_energy)   :595

plus the coarser seam `align_pcm` (PCM in, alignment out; features never leave the GPU
except for the small host stage).  All arithmetic of the device stages runs in
libdescribealign_b200.so; without it, or without a CUDA device, these functions raise.
"""
from __future__ import annotations

import threading
import time
import weakref

import numpy as np

from . import _cabi, host_fit
from ._cabi import AUDIO, VIDEO

FAILED_MSG = host_fit.FAILED_MSG

_tls = threading.local()


_device = -1     # GPU index of this process (-1: whatever device is current in the calling thread)


def set_device(index: int):
    """The GPU this process works on (one process per GPU).  Needed when host threads other than the
    main one call into the library - CUDA's "current device" is per thread and starts at 0 - e.g. the
    worker threads of batch.run_local under torchrun."""
    global _device
    _device = int(index)


_log10_request = None     # a mode asked for before CUDA was initialised (launcher.patch)


def request_log10_mode(mode: str):
    """Remember which log10 the feature kernel should follow (native_log.set_mode); applied when the
    first context is created, so that importing / patching never initialises CUDA."""
    global _log10_request
    if mode not in ("portable", "native", "auto"):
        raise ValueError("log10 mode must be 'portable', 'native' or 'auto'")
    _log10_request = mode


def _current_device_of_this_thread() -> int:
    import sys
    torch = sys.modules.get("torch")
    try:
        if torch is not None and torch.cuda.is_available() and torch.cuda.is_initialized():
            return int(torch.cuda.current_device())
    except Exception:
        pass
    return -1


def context() -> _cabi.Context:
    """Per-thread CUDA context handle, created on first use (never at import time, so that
    a GUI parent process that only imports this module does not initialise CUDA before it
    forks its worker; reference describealign.py:1432)."""
    global _log10_request, _device
    ctx = getattr(_tls, "ctx", None)
    if ctx is None:
        if _device < 0:
            # No set_device(): the whole process stays on the GPU that is current in the FIRST thread that
            # needs a context (what a caller who did torch.cuda.set_device(local_rank) expects) - a fresh
            # worker thread's own "current device" would be 0 whatever the rank.
            _device = _current_device_of_this_thread()
        ctx = _cabi.Context(_device)
        _tls.ctx = ctx
        _tls.pairs = []
        if _log10_request is not None:
            from . import native_log
            req, _log10_request = _log10_request, None
            native_log.set_mode(req, ctx)
    return ctx


def acquire_pair() -> _cabi.Pair:
    ctx = context()
    return _tls.pairs.pop() if _tls.pairs else _cabi.Pair(ctx)


def release_pair(pair: _cabi.Pair):
    _tls.pairs.append(pair)


# ---------------------------------------------------------------------------------------------
# feature functions
# ---------------------------------------------------------------------------------------------
_feature_cache = []   # [(weakref(arr), data_ptr, shape, features)] newest last, at most 4 entries


def _interleaved(arr: np.ndarray) -> np.ndarray:
    """(ch, S) array as handed out by describealign.py:156 -> contiguous (S, ch) samples."""
    arr = np.asarray(arr)
    if arr.ndim != 2 or arr.shape[0] not in (1, 2):
        raise ValueError("expected a (channels, samples) array with 1 or 2 channels")
    if arr.dtype not in (np.float16, np.int16):
        arr = arr.astype(np.float16)
    return np.ascontiguousarray(arr.T)


def track_features(arr: np.ndarray):
    """All five feature vectors of one track from one pass over the PCM on the GPU.
    The reference calls three functions on the same array; the results are cached per
    array so the PCM is uploaded and scanned once."""
    if hasattr(arr, "features") and hasattr(arr, "device"):
        return arr.features()          # a decode.DevicePcm: the samples are already in HBM
    arr = np.asarray(arr)
    ptr = arr.__array_interface__["data"][0]
    for ref, p, shape, feats in _feature_cache:
        if ref() is arr and p == ptr and shape == arr.shape:
            return feats
    pair = acquire_pair()
    try:
        pair.set_pcm(VIDEO, _interleaved(arr))
        feats = pair.get_features(VIDEO)
    finally:
        release_pair(pair)
    try:
        _feature_cache.append((weakref.ref(arr), ptr, arr.shape, feats))
        del _feature_cache[:-4]
    except TypeError:
        pass
    return feats


def remember_features(arr: np.ndarray, feats):
    """Attach already computed features to a host array (launcher: a stereo track decoded straight to the device whose
    samples --stretch_audio needs on the host as well), so that the feature functions do not upload it again."""
    try:
        _feature_cache.append((weakref.ref(arr), arr.__array_interface__["data"][0], arr.shape, feats))
        del _feature_cache[:-4]
    except TypeError:
        pass


def get_energy(arr):
    return track_features(arr)[0]


def get_zero_crossings(arr):
    return track_features(arr)[1]


def get_freq_bands(arr):
    return list(track_features(arr)[2:5])


# ---------------------------------------------------------------------------------------------
# align
# ---------------------------------------------------------------------------------------------

class AlignJob:
    """One pair walking through the path: device stage A -> host fit -> device stage B.

    The three steps are separate methods so that a batch driver can keep several pairs in
    flight (one CUDA stream per pair) and so that the benchmark can time the device stages
    without the host-side rate-change fit, which BASELINE.json excludes from the metric.
    """

    def __init__(self, pair: _cabi.Pair | None = None, detached: bool = False):
        """detached: no dab_pair of its own - the device stages run in an engine slot (batch.run_local) and
        this object only carries the host-side steps (length rules, host fit, nodes)."""
        self.pair = None if detached else (pair if pair is not None else acquire_pair())
        self._own = pair is None and not detached
        self.video_features = self.audio_features = None
        self.want_all_features = False
        self.device_scaling = True   # stage B scales the pair's device-resident features itself (6 floats up)
        self.device_planning = True  # ... and plans the corridors there (describealign.py:895-932, csrc/refine.cuh)
        self.test_planner = None     # tests: a host-side corridor planner to compare the device's planning with
        self.gains = None
        self.h2d_bytes = self.d2h_bytes = 0
        self.host_ms = {}        # wall time of each host-side call of this job (diagnostics)

    def _timed(self, name, fn, *args):
        t0 = time.perf_counter()
        out = fn(*args)
        self.host_ms[name] = self.host_ms.get(name, 0.0) + 1e3 * (time.perf_counter() - t0)
        return out

    # -- inputs ---------------------------------------------------------------------------
    def load_pcm(self, video_pcm, audio_pcm):
        """Host PCM: interleaved int16 (S, ch), or the reference's float16 (ch, S) arrays."""
        def prep(p):
            p = np.asarray(p)
            if p.dtype == np.float16 and p.ndim == 2 and p.shape[0] in (1, 2) and p.shape[1] > 2:
                return _interleaved(p)
            return p
        v, a = prep(video_pcm), prep(audio_pcm)
        self._timed("set_pcm", self.pair.set_pcm, VIDEO, v)
        self._timed("set_pcm", self.pair.set_pcm, AUDIO, a)
        self.h2d_bytes += v.nbytes + a.nbytes
        self._features_on_device = True

    def load_pcm_device(self, video, audio):
        """Device-resident PCM: (device pointer, samples per channel, channels) per track."""
        self.pair.set_pcm_device(VIDEO, *video)
        self.pair.set_pcm_device(AUDIO, *audio)
        self._features_on_device = True

    def load_features(self, video_features, audio_features, video_energy=None, audio_energy=None):
        """video_energy / audio_energy: align()'s separate energy arguments when they are not features[0]."""
        self.pair.set_features(VIDEO, video_features)
        self.pair.set_features(AUDIO, audio_features)
        if video_energy is not None:
            self.pair.set_gate_energy(VIDEO, video_energy)
        if audio_energy is not None:
            self.pair.set_gate_energy(AUDIO, audio_energy)
        self.video_features, self.audio_features = video_features, audio_features
        self.h2d_bytes += sum(np.asarray(f).nbytes for f in list(video_features) + list(audio_features))
        self._features_on_device = False

    # -- stages ---------------------------------------------------------------------------
    def device_stage_a(self):
        """Features (if PCM was given) + stage A on the device; brings back what the host fit
        needs: the integer pass-1 path and the feature vectors."""
        self._timed("stage_a", self.pair.stage_a)
        return self.after_stage_a()

    def after_stage_a(self):
        """Length rule of describealign.py:698-699 and the copies the host fit needs."""
        n_path = self.pair.n_path1
        if self.video_features is None:
            # the host fit uses the first three features only (describealign.py:735)
            count = 5 if self.want_all_features else 3
            # views of the pair's pinned staging unless the caller keeps them (details)
            keep = self.want_all_features
            self.video_features = self._timed("get_features", self.pair.get_features, VIDEO, count, keep)
            self.audio_features = self._timed("get_features", self.pair.get_features, AUDIO, count, keep)
            self.d2h_bytes += sum(f.nbytes for f in self.video_features + self.audio_features)
        self.check_path1_length(n_path)
        self.x, self.y = self._timed("path1", self.pair.path1)
        self.d2h_bytes += 8 * n_path
        return self.x, self.y

    def check_path1_length(self, n_path):
        """Length rule of describealign.py:698-699 (needs the feature vectors for the track lengths)."""
        self.n_video_energy = len(self.video_features[0])
        self.n_audio_energy = len(self.audio_features[0])
        self.min_len = host_fit.min_path_length(self.n_video_energy, self.n_audio_energy)
        if n_path < self.min_len:
            raise RuntimeError(FAILED_MSG)

    def stage_b_input(self):
        """What dab_engine_submit_b / dab_pair_stage_b_gains need from the host fit."""
        return dict(gains=self.gains[0], audio_stds=self.gains[1], n_audio=self.n_audio_scaled,
                    n_video=self.n_video_scaled, lines=self.lines)

    def host_stage(self):
        """The untimed "rate-change fit" on the host (describealign.py:702-893)."""
        keep = host_fit.continuity_error(self.x, self.y) < 3
        self.kept_x, self.kept_y = self.x[keep], self.y[keep]
        g, sd, self.n_audio_scaled, self.n_video_scaled = host_fit.feature_gains(
            self.video_features, self.audio_features, self.kept_x, self.kept_y)
        self.gains = (g, sd)
        fit_x, fit_y = host_fit.compress_path(self.kept_x, self.kept_y)
        self.fit = host_fit.rate_change_fit(fit_x, fit_y)
        self.clusters = host_fit.line_clusters(self.fit)
        self.lines = host_fit.cluster_lines(self.clusters)

    def device_stage_b(self):
        """Scaling (describealign.py:737-741), corridor planning (:895-932), scoring, DP 2 and traceback on
        the device, from the host fit's six scalars and line clusters.  (device_planning / device_scaling
        off: the oracle's numpy planning / uploaded scaled arrays instead - test switches.)"""
        if self.device_planning and self.device_scaling:
            self._timed("stage_b", self.pair.stage_b_clusters, self.gains[0], self.gains[1], self.n_audio_scaled,
                        self.n_video_scaled, self.lines)
            self.h2d_bytes += 24 + 40 * len(self.lines)
        else:
            audio_scaled, video_scaled = host_fit.scale_features(self.video_features, self.audio_features, self.kept_x, self.kept_y)
            if self.test_planner is None:
                raise RuntimeError("device_planning / device_scaling are test switches: they need job.test_planner")
            plans = self.test_planner(self.clusters, audio_scaled, video_scaled)
            if self.device_scaling:
                self._timed("stage_b", self.pair.stage_b_gains, self.gains[0], self.gains[1], audio_scaled,
                            video_scaled, plans, len(self.clusters))
                self.h2d_bytes += 24
            else:
                self._timed("stage_b", self.pair.stage_b, audio_scaled, video_scaled, plans, len(self.clusters))
                self.h2d_bytes += audio_scaled.nbytes + video_scaled.nbytes
        self.path = self._timed("path2", self.pair.path2)
        self.d2h_bytes += self.path.nbytes
        if len(self.path) < self.min_len:
            raise RuntimeError(FAILED_MSG)
        return self.path

    def finish(self, details=None):
        if details is not None:
            details.update(kept_x=self.kept_x, kept_y=self.kept_y, fit=self.fit, clusters=self.clusters,
                           plans=self.pair.corridors(), gains=self.gains,
                           video_features=self.video_features, audio_features=self.audio_features,
                           path1=(self.x, self.y), stats=self.pair.stats(), timings=self.pair.timings(),
                           h2d_bytes=self.h2d_bytes, d2h_bytes=self.d2h_bytes)
        nx, ny, sim = host_fit.build_nodes(self.path, self.n_audio_energy, self.n_video_energy,
                                           self.n_audio_scaled, self.n_video_scaled)
        return nx, ny, sim, self.path, self.fit.median_slope

    def close(self):
        if self._own and self.pair is not None:
            release_pair(self.pair)
        self.pair = None

    def run(self, details=None):
        print("  matching audio...  \r", end='')
        self.device_stage_a()
        print("  refining match: pass 1 of 2...\r", end='')
        self.host_stage()
        print("  refining match: pass 2 of 2...\r", end='')
        self.device_stage_b()
        return self.finish(details)


def align(video_features, audio_desc_features, video_energy, audio_desc_energy, details=None):
    """Drop-in for describealign.align (describealign.py:595-1027).

    Returns (audio_desc_times, video_times, similarity_percent, path, median_slope) with the
    reference's meaning; raises RuntimeError("Alignment failed, ...") under the reference's
    length rule (:698-699, :991-992)."""
    gates = []
    for feats, energy in ((video_features, video_energy), (audio_desc_features, audio_desc_energy)):
        # the reference's only caller passes features[0] (describealign.py:1121); anything else decides the
        # not-quiet frames in its place (:629, :657) and the path-length rule (:698)
        same = energy is feats[0] or (np.shape(energy) == np.shape(feats[0]) and np.array_equal(energy, feats[0]))
        if not same and len(energy) != len(feats[0]):
            raise ValueError("the energy arguments must have the length of features[0]")
        gates.append(None if same else np.asarray(energy))
    print("  memorizing video...        \r", end='')
    job = AlignJob()
    try:
        job.load_features(list(video_features), list(audio_desc_features), gates[0], gates[1])
        return job.run(details)
    finally:
        job.close()


def align_pcm(video_pcm, audio_desc_pcm, details=None):
    """PCM in, alignment out (the block describealign.py:1096-1125 in one call): features
    never leave the GPU except the copies the host-side fit needs."""
    print("  memorizing video...        \r", end='')
    job = AlignJob()
    try:
        job.want_all_features = details is not None
        job.load_pcm(video_pcm, audio_desc_pcm)
        return job.run(details)
    finally:
        job.close()


def align_streams(video_stream, audio_desc_stream, details=None):
    """Two decode.DevicePcm tracks in, alignment out: align_pcm for samples that the decode hand-off
    (describealign_b200.decode, describealign.py:149-157) already put into HBM."""
    print("  memorizing video...        \r", end='')
    job = AlignJob()
    try:
        job.want_all_features = details is not None
        job.load_pcm_device(video_stream.device(), audio_desc_stream.device())
        return job.run(details)
    finally:
        job.close()
