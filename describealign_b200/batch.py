"""Batch mode and very long pairs over the GPUs of one box.

The reference's batch mode is a sequential `for` over file pairs in one process
(describealign.py:1077); pairs share no state.  Here:

* `align_batch`      one process per GPU (torchrun), pairs assigned longest-first to ranks, several
                     pairs in flight per GPU (one CUDA stream each, so the serial frontier DPs of
                     different pairs overlap), results gathered on rank 0.  No data-path collective.
* `align_long_pair`  one pair on all ranks: the audio query rows of the match stage
                     (describealign.py:649-673) are sharded over the ranks, the scored match points
                     are all-gathered (NCCL over NVLink; gloo in the CPU tests) and every rank runs
                     the sequential frontier DP on the full list.  The DP does not shard.

The collective plumbing is `torch.distributed`; everything that computes is behind the C ABI.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor
from typing import Callable, Sequence

import numpy as np


# ---------------------------------------------------------------------------------------------
# partitioning (pure host logic)
# ---------------------------------------------------------------------------------------------

def assign_pairs(durations: Sequence[float], world: int) -> list[list[int]]:
    """Longest-processing-time-first assignment of pairs to `world` ranks.

    Returns, per rank, the indices of its pairs in the order they should be started (longest
    first).  Deterministic: ties go to the lower pair index and the lower rank."""
    if world <= 0:
        raise ValueError("world must be positive")
    order = sorted(range(len(durations)), key=lambda k: (-float(durations[k]), k))
    load = [0.0] * world
    out: list[list[int]] = [[] for _ in range(world)]
    for k in order:
        r = min(range(world), key=lambda x: (load[x], x))
        out[r].append(k)
        load[r] += float(durations[k])
    return out


def row_shards(n_rows: int, world: int) -> list[tuple[int, int]]:
    """Contiguous, ordered audio-row ranges [lo, hi) covering [0, n_rows), one per rank."""
    base, rem = divmod(max(int(n_rows), 0), world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def pair_duration(pair) -> float:
    """Seconds of audio in a (video_pcm, description_pcm) pair of (S, ch) / (ch, S) arrays."""
    def samples(p):
        p = np.asarray(p)
        return p.shape[0] if p.ndim == 1 or p.shape[0] > 2 else p.shape[1]
    v, a = pair
    return (samples(v) + samples(a)) / 44100.0


# ---------------------------------------------------------------------------------------------
# distributed helpers
# ---------------------------------------------------------------------------------------------

def _dist():
    import torch.distributed as dist
    return dist


def _rank_world(group=None):
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def exchange_varlen(tensors, group=None):
    """All-gather 1-D tensors whose length differs per rank (every rank passes the same number of
    tensors, element k having the same length on a rank): returns, per tensor, the concatenation over
    ranks in rank order.  One all_gather of the lengths, one padded all_gather per tensor."""
    import torch
    dist = _dist()
    rank, world = _rank_world(group)
    if world == 1:
        return list(tensors)
    dev = tensors[0].device
    n = torch.tensor([tensors[0].numel()], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    out = []
    for x in tensors:
        pad = torch.zeros(cap, dtype=x.dtype, device=dev)
        pad[:x.numel()] = x
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        out.append(torch.cat([parts[r][:counts[r]] for r in range(world)]))
    return out


def exchange_points(i, v, q, group=None):
    """All-gather variable-length match-point shards.

    i, v: int32 tensors, q: float64 tensor, all of length n_rank, on the CPU (gloo) or on the
    rank's GPU (nccl).  Returns the concatenation over ranks in rank order, which is sorted by
    (audio frame, video frame) when the shards are ordered row ranges."""
    gi, gv, gq = exchange_varlen([i, v, q], group)
    return gi, gv, gq


# ---------------------------------------------------------------------------------------------
# batch mode
# ---------------------------------------------------------------------------------------------

def gpu_runner(pair):
    """One pair through the CUDA path, synchronously: PCM in, the reference's align() tuple out."""
    from . import api
    job = api.AlignJob()
    try:
        job.load_pcm(*pair)
        job.device_stage_a()
        job.host_stage()
        job.device_stage_b()
        return job.finish()
    finally:
        job.close()


_engines = {}


def engine(slots: int = 32):
    """This process's batch engine (one per GPU and slot count), created on first use."""
    from . import _cabi, api
    key = (api._device, int(slots))
    eng = _engines.get(key)
    if eng is None:
        eng = _cabi.Engine(api.context(), slots)
        _engines[key] = eng
    return eng


def _interleaved_pcm(p):
    from . import api
    if isinstance(p, tuple):
        return p                          # (device pointer, samples, channels): PCM already on the GPU
    p = np.asarray(p)
    if p.dtype == np.float16 and p.ndim == 2 and p.shape[0] in (1, 2) and p.shape[1] > 2:
        return api._interleaved(p)        # the reference's (ch, S) float16 arrays
    return p


def run_engine(pairs, in_flight: int = 32, host_workers: int = 0, host_stage: Callable | None = None, finish: bool = True):
    """Run `pairs` on this process's GPU through the batch engine with up to `in_flight` pairs on the
    device at once.  One Python thread serves the engine's event queue; the host fit of each pair
    (describealign.py:702-893, ~1 s of scipy per 22-min pair) runs in a pool of `host_workers` threads
    (default: one per host core, at most 16) so that it overlaps the device stages of other pairs.
    host_stage(job): replaces AlignJob.host_stage (the benchmark passes a cache lookup).
    Returns one result per pair in input order: the tuple align() returns (or the job when
    finish=False), or the exception the pair raised."""
    from concurrent.futures import ThreadPoolExecutor
    from . import _cabi, api
    eng = engine(in_flight)
    n = len(pairs)
    results: list = [None] * n
    jobs: dict = {}
    if host_workers <= 0:
        host_workers = max(1, min(16, os.cpu_count() or 1))
    pool = ThreadPoolExecutor(max_workers=host_workers)
    fits = []            # (future, tag, slot)
    submitted = done = 0
    # PCM is read by asynchronous copies: keep no more pairs loaded than the engine can start soon
    window = in_flight + 4

    def fail(tag, slot, exc):
        nonlocal done
        results[tag] = exc
        if slot is not None:
            eng.release(slot)
        jobs.pop(tag, None)
        done += 1

    def fit(job):
        (host_stage or api.AlignJob.host_stage)(job)
        return job

    try:
        while done < n:
            while submitted < n and submitted - done < window:
                p = pairs[submitted]
                try:
                    v, a = p() if callable(p) else p
                    jobs[submitted] = api.AlignJob(detached=True)
                    jobs[submitted].tag = submitted
                    eng.submit(submitted, _interleaved_pcm(v), _interleaved_pcm(a))
                except Exception as e:      # reported per pair; the batch goes on
                    fail(submitted, None, e)
                submitted += 1
            # host fits that finished: hand their pairs back to the device
            still = []
            for fut, tag, slot in fits:
                if not fut.done():
                    still.append((fut, tag, slot))
                    continue
                try:
                    job = fut.result()
                    eng.submit_b(slot, **job.stage_b_input())
                except Exception as e:
                    fail(tag, slot, e)
            fits = still
            evt = eng.next(2 if fits else 50)
            if evt is None:
                continue
            tag = int(evt.tag)
            job = jobs.get(tag)
            if evt.status != 0:
                fail(tag, evt.slot, _cabi.DabError(f"describealign_b200 error {evt.status}: {eng.error(evt)}"))
                continue
            try:
                if evt.kind == _cabi.EVENT_STAGE_A:
                    job.video_features = [f.copy() for f in eng.features(evt, _cabi.VIDEO)]
                    job.audio_features = [f.copy() for f in eng.features(evt, _cabi.AUDIO)]
                    job.check_path1_length(int(evt.n_path1))
                    x, y = eng.path1(evt)
                    job.x, job.y = x.astype(np.int64), y.astype(np.int64)
                    fits.append((pool.submit(fit, job), tag, evt.slot))
                else:
                    job.path = eng.rows(evt).copy()
                    job.stats = evt.stats.as_dict()
                    job.timings = eng.timings(evt)
                    eng.release(evt.slot)
                    if len(job.path) < job.min_len:
                        raise RuntimeError(api.FAILED_MSG)
                    results[tag] = job.finish() if finish else job
                    jobs.pop(tag, None)
                    done += 1
            except Exception as e:
                fail(tag, evt.slot if evt.kind == _cabi.EVENT_STAGE_A else None, e)
    finally:
        pool.shutdown(wait=True)
    return results


def run_local(pairs, in_flight: int = 32, runner: Callable | None = None, **engine_kw):
    """Run `pairs` on this process's GPU with up to `in_flight` pairs in progress at once.
    A pair that fails (e.g. RuntimeError("Alignment failed, ...")) yields its exception.
    With the default runner the pairs go through the batch engine (run_engine, which takes engine_kw:
    host_workers, host_stage, finish); a custom runner (CPU tests, instrumentation) is called once per
    pair from a thread pool."""
    if runner is None or runner is gpu_runner:
        return run_engine(pairs, in_flight, **engine_kw)

    def one(p):
        try:
            return runner(p() if callable(p) else p)
        except Exception as e:   # reported per pair; the batch goes on
            return e

    if in_flight <= 1 or len(pairs) <= 1:
        return [one(p) for p in pairs]
    with ThreadPoolExecutor(max_workers=in_flight) as ex:
        return list(ex.map(one, pairs))


def align_batch(pairs, durations: Sequence[float] | None = None, in_flight: int = 32, group=None,
                runner: Callable | None = None, gather: bool = True, **engine_kw):
    """Align a list of (video_pcm, description_pcm) pairs (or zero-argument loaders returning
    such a pair) over all ranks of the current process group.

    Every rank must call this with the same list (loaders are only invoked on the owning rank; entries a
    rank does not own are never touched and may be None when durations are given).
    Returns, on rank 0, one result per pair in input order - the tuple `align()` returns, or the
    exception the pair raised; other ranks return None.  gather=False: no collective at the end, every
    rank gets [(pair index, result), ...] for its own pairs."""
    rank, world = _rank_world(group)
    if durations is None:
        if any(callable(p) for p in pairs):
            raise ValueError("durations are required when pairs are given as loaders")
        durations = [pair_duration(p) for p in pairs]
    mine = assign_pairs(durations, world)[rank]
    local = run_local([pairs[k] for k in mine], in_flight, runner, **engine_kw)
    if not gather:
        # every rank keeps its own results: [(pair index, result), ...]
        return list(zip(mine, local))
    if world == 1:
        out = [None] * len(pairs)
        for k, res in zip(mine, local):
            out[k] = res
        return out
    dist = _dist()
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(list(zip(mine, local)), gathered, dst=0, group=group)
    if rank != 0:
        return None
    out = [None] * len(pairs)
    for part in gathered:
        for k, res in part:
            out[k] = res
    return out


# ---------------------------------------------------------------------------------------------
# one very long pair on all ranks
# ---------------------------------------------------------------------------------------------

HOST_PHASES = ("host_fit_and_broadcast", "nodes_on_host_and_broadcast")


def align_long_pair(video_pcm, audio_desc_pcm, group=None, details=None, root: int = 0, host_stage: Callable | None = None):
    """One very long pair on all ranks of the group.  Every rank passes the same PCM and
    gets the same result.

    * features and codes: computed redundantly per rank (one pass over the PCM - cheaper than broadcasting
      them);
    * match stage (describealign.py:649-673): audio query rows sharded over the ranks, ONE all-gather of the
      scored match points (NCCL over NVLink);
    * DP #1, traceback and the host fit (:674-893): on `root` only - they do not shard - and the fit's
      six scalars and line clusters are broadcast;
    * corridor scoring (:931-944): audio rows sharded again, ONE all-gather of the quals (the point list
      itself only depends on the corridors and is built on every rank);
    * DP #2, traceback, nodes (:946-1027): on `root`, result broadcast.

    host_stage(job): replaces job.host_stage() on `root` (bench.py caches the rate-change fit there).
    details["phases_ms"]: time between CUDA events recorded after a device synchronisation at the end of each phase on
    this rank; the phases named in HOST_PHASES are host work on `root` (and the other ranks' wait for it), the rest is
    device work.  details["host_fit_s"]: wall time of the host fit on `root`."""
    import time
    import torch
    from . import _cabi, api
    dist = _dist()
    rank, world = _rank_world(group)
    job = api.AlignJob()
    info = {}
    marks = []

    def mark(name):
        e = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e.record()
        marks.append((name, e))

    try:
        mark("start")
        if isinstance(video_pcm, tuple):      # device-resident PCM: (device pointer, samples per channel, channels)
            job.load_pcm_device(video_pcm, audio_desc_pcm)
        else:
            job.load_pcm(video_pcm, audio_desc_pcm)
        pair = job.pair
        dev = torch.device("cuda", torch.cuda.current_device())
        mark("features")
        # ---- stage A ----
        n_rows = int(pair.feature_lens(_cabi.AUDIO)[0])
        lo, hi = row_shards(n_rows, world)[rank]
        n = pair.stage_a_match(lo, hi)
        ti = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        tv = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        tq = torch.empty(max(n, 1), dtype=torch.float64, device=dev)
        pair.export_points1_device(ti.data_ptr(), tv.data_ptr(), tq.data_ptr())
        mark("match_shard")
        gi, gv, gq = exchange_points(ti[:n], tv[:n], tq[:n], group)
        mark("all_gather_points")
        info["shard_a"] = (lo, hi, n, int(gi.numel()))
        msg = [None]
        host_fit_s = 0.0
        if rank == root:
            try:
                gi, gv, gq = gi.contiguous(), gv.contiguous(), gq.contiguous()
                torch.cuda.current_stream().synchronize()
                pair.import_points1_device(gi.data_ptr(), gv.data_ptr(), gq.data_ptr(), gi.numel())
                pair.dp1()
                job.after_stage_a()
                mark("dp1_trace")
                t_fit = time.perf_counter()
                if host_stage is None:
                    job.host_stage()
                else:
                    host_stage(job)
                host_fit_s = time.perf_counter() - t_fit
                msg[0] = ("ok", job.stage_b_input())
            except Exception as e:       # every rank must learn that the pair failed
                msg[0] = ("error", e)
        if world > 1:
            dist.broadcast_object_list(msg, src=root, group=group)
        if msg[0][0] == "error":
            raise msg[0][1]
        b_in = msg[0][1]
        mark("host_fit_and_broadcast")       # not device time: the root's host fit and the wait for it (excluded by BASELINE.json)
        # ---- stage B ----
        lo2, hi2 = row_shards(int(b_in["n_audio"]), world)[rank]
        n2, first, mine = pair.stage_b_score(b_in["gains"], b_in["audio_stds"], b_in["n_audio"], b_in["n_video"],
                                             b_in["lines"], lo2, hi2)
        tq2 = torch.empty(max(mine, 1), dtype=torch.float64, device=dev)
        pair.export_quals2_device(tq2.data_ptr(), first, mine)
        mark("corridor_shard")
        (q_all,) = exchange_varlen([tq2[:mine]], group)
        mark("all_gather_quals")
        info["shard_b"] = (lo2, hi2, mine, int(q_all.numel()))
        out = [None]
        if rank == root:
            try:
                q_all = q_all.contiguous()
                torch.cuda.current_stream().synchronize()
                if int(q_all.numel()) != n2:
                    raise RuntimeError("long pair: the ranks' corridor shards do not add up to the point list")
                pair.import_quals2_device(q_all.data_ptr(), n2)
                pair.dp2()
                mark("dp2_trace")
                job.path = pair.path2()
                mark("path_to_host")
                if len(job.path) < job.min_len:
                    raise RuntimeError(api.FAILED_MSG)
                out[0] = ("ok", job.finish(details))
            except Exception as e:
                out[0] = ("error", e)
        if world > 1:
            dist.broadcast_object_list(out, src=root, group=group)
        if out[0][0] == "error":
            raise out[0][1]
        mark("nodes_on_host_and_broadcast")   # not device time: similarity / node list in Python on root (:993-1026), result broadcast
        if details is not None:
            details["shards"] = info
            details["phases_ms"] = {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(marks[:-1], marks[1:])}
            details["host_fit_s"] = host_fit_s
        return out[0][1]
    finally:
        job.close()


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Pin this process (and therefore the page-locked buffers it allocates from now on) to the CPUs of
    the NUMA node the GPU hangs off, read from sysfs.  Uploads from the far socket cross the inter-socket
    link and share it between all ranks of a box.  Returns what was found; does nothing on a single-node
    host or when the topology cannot be read."""
    info = {"device": int(device_index), "numa_node": None, "cpus": None, "bound": False}
    try:
        import subprocess
        out = subprocess.run(["nvidia-smi", "-i", str(int(device_index)), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if not out:
            return info
        bdf = out[-12:] if len(out) >= 12 else out          # 00000000:1B:00.0 -> 0000:1b:00.0
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        info["numa_node"] = node
        if node < 0:
            # sysfs does not say (virtualised PCI topology).  If the host still shows several memory nodes, assume the
            # usual HGX layout - the first half of the GPUs on the first socket, the second half on the second.
            import glob
            nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
            n_gpus = len([ln for ln in subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=20).stdout.splitlines()
                          if ln.startswith("GPU ")])
            if len(nodes) < 2 or n_gpus < 2:
                return info
            node = nodes[min(len(nodes) - 1, int(device_index) * len(nodes) // n_gpus)]
            info["numa_node_assumed"] = node
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        info["cpus"] = len(allowed)
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["bound"] = True
    except Exception as e:       # best effort: an unreadable topology must never stop a run
        info["error"] = repr(e)
    return info


def init_from_env(backend: str = "nccl"):
    """torchrun plumbing: RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* from the environment."""
    import torch
    dist = _dist()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if backend == "nccl":
        torch.cuda.set_device(local)
        from . import api
        api.set_device(local)      # the library's worker threads must not fall back to device 0
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        kw = {"device_id": torch.device("cuda", local)} if backend == "nccl" else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world, local
