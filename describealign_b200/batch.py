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


def exchange_points(i, v, q, group=None):
    """All-gather variable-length match-point shards.

    i, v: int32 tensors, q: float64 tensor, all of length n_rank, on the CPU (gloo) or on the
    rank's GPU (nccl).  Returns the concatenation over ranks in rank order, which is sorted by
    (audio frame, video frame) when the shards are ordered row ranges."""
    import torch
    dist = _dist()
    rank, world = _rank_world(group)
    if world == 1:
        return i, v, q
    dev = i.device
    n = torch.tensor([i.numel()], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)

    def gather(x):
        pad = torch.zeros(cap, dtype=x.dtype, device=dev)
        pad[:x.numel()] = x
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        return torch.cat([parts[r][:counts[r]] for r in range(world)])

    return gather(i), gather(v), gather(q)


# ---------------------------------------------------------------------------------------------
# batch mode
# ---------------------------------------------------------------------------------------------

def gpu_runner(pair):
    """One pair through the CUDA path: PCM in, the reference's align() tuple out."""
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


def run_local(pairs, in_flight: int = 64, runner: Callable | None = None):
    """Run `pairs` on this process's GPU with up to `in_flight` pairs in progress at once.
    A pair that fails (e.g. RuntimeError("Alignment failed, ...")) yields its exception."""
    runner = gpu_runner if runner is None else runner

    def one(p):
        try:
            return runner(p() if callable(p) else p)
        except Exception as e:   # reported per pair; the batch goes on
            return e

    if in_flight <= 1 or len(pairs) <= 1:
        return [one(p) for p in pairs]
    if runner is gpu_runner:
        # one host thread per pair in flight: they must sleep, not spin, while their streams drain
        from . import _cabi, api
        _cabi.set_host_wait(api._device, 2)
    with ThreadPoolExecutor(max_workers=in_flight) as ex:
        return list(ex.map(one, pairs))


def align_batch(pairs, durations: Sequence[float] | None = None, in_flight: int = 64, group=None,
                runner: Callable | None = None):
    """Align a list of (video_pcm, description_pcm) pairs (or zero-argument loaders returning
    such a pair) over all ranks of the current process group.

    Every rank must call this with the same list (loaders are only invoked on the owning rank).
    Returns, on rank 0, one result per pair in input order - the tuple `align()` returns, or the
    exception the pair raised; other ranks return None."""
    rank, world = _rank_world(group)
    if durations is None:
        if any(callable(p) for p in pairs):
            raise ValueError("durations are required when pairs are given as loaders")
        durations = [pair_duration(p) for p in pairs]
    mine = assign_pairs(durations, world)[rank]
    local = run_local([pairs[k] for k in mine], in_flight, runner)
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

def align_long_pair(video_pcm, audio_desc_pcm, group=None, details=None):
    """Every rank passes the same PCM and gets the same result.  Features and codes are computed
    redundantly per rank (one pass over the PCM, cheaper than broadcasting them); the match stage
    is sharded by audio rows; one all-gather of the scored match points precedes DP #1."""
    import torch
    from . import _cabi, api
    rank, world = _rank_world(group)
    job = api.AlignJob()
    try:
        job.load_pcm(video_pcm, audio_desc_pcm)
        pair = job.pair
        n_rows = int(pair.feature_lens(_cabi.AUDIO)[0])
        lo, hi = row_shards(n_rows, world)[rank]
        n = pair.stage_a_match(lo, hi)
        dev = torch.device("cuda", torch.cuda.current_device())
        ti = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        tv = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        tq = torch.empty(max(n, 1), dtype=torch.float64, device=dev)
        pair.export_points1_device(ti.data_ptr(), tv.data_ptr(), tq.data_ptr())
        gi, gv, gq = exchange_points(ti[:n], tv[:n], tq[:n], group)
        gi, gv, gq = gi.contiguous(), gv.contiguous(), gq.contiguous()
        torch.cuda.current_stream().synchronize()
        pair.import_points1_device(gi.data_ptr(), gv.data_ptr(), gq.data_ptr(), gi.numel())
        pair.dp1()
        job.after_stage_a()
        job.host_stage()
        job.device_stage_b()
        if details is not None:
            details["shard"] = (lo, hi, n, int(gi.numel()))
        return job.finish(details)
    finally:
        job.close()


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
