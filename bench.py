#!/usr/bin/env python
"""Benchmark of the alignment hot path (BASELINE.json: audio-hours aligned per second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs B] [--impl ours|reference]

A step = one pass of the hot path over one batch of B synthetic pairs per GPU (W in flight) of the C2 shape
(22-min video audio vs 27-min description, 202 s start offset, injected skips):
features of both tracks, device stage A (describealign.py:596-700) and device stage B
(:895-993).  The host-side rate-change fit between the two device stages (:702-893) is outside
the timed regions, as BASELINE.json prescribes; it is solved once per pair and reused for as long
as stage A returns the identical pass-1 path.  W pairs are in flight at once (one CUDA stream
each): the frontier DPs are one warp per pair, so throughput comes from overlapping pairs.

value   device-resident: PCM already in HBM when the timed region starts; CUDA events
        bracketing each device stage on the streams the kernels are launched on.
e2e     the same passes through the public API (AlignJob.load_pcm + stages) from pinned HOST
        buffers, H2D/D2H copies inside the timed region.
One rank per GPU (torchrun for N > 1); ranks process disjoint pairs (weak scaling), no
data-path collective; timings are max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# one CUDA stream per pair in flight: ask for the maximum number of hardware queues before CUDA starts
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "audio_hours_aligned_per_second"
UNIT = "audio-hours/s"
WORKLOAD = "C2: synthetic 22-min video audio vs 27-min description, 202 s offset, 10 inserted skips, mono 44.1 kHz s16"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=0, help="pairs per GPU per step (default: 4 x workers)")
    ap.add_argument("--workers", type=int, default=0,
                    help="pairs in flight per GPU, one CUDA stream and one host thread each (default: 128, 96 with 4 ranks "
                         "and 64 with 8 ranks sharing the host's cores)")
    ap.add_argument("--distinct", type=int, default=0,
                    help="distinct synthetic pairs per GPU, a step cycles over them (default: 8, or 4 per GPU when more than two ranks "
                         "share the host's cores for generating them; 4 pairs are 1 GB of PCM, still 8x the L2)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the C2 durations (debugging only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--timeline", default=None, help="write the device timeline of the last step of each arm (per pair and stage) to <name>_dev.json / <name>_e2e.json")
    ap.add_argument("--dp-reserve-kb", type=int, default=0, help="see dab_set_option: keeps other pairs' big CTAs off the DP's SM")
    ap.add_argument("--switch-interval", type=float, default=0.005, help="Python thread switch interval (s)")
    ap.add_argument("--host-wait", type=int, default=2, help="dab_set_host_wait mode: 0 spin (CUDA default), 1 blocking sync, 2 query + sleep")
    return ap.parse_args()


def make_pairs(n, first_seed, scale, world=1):
    """n C2 pairs with distinct seeds, generated in parallel worker processes (the host's cores are
    shared by the `world` ranks of the node)."""
    from concurrent.futures import ProcessPoolExecutor
    from describealign_b200 import synth
    seeds = [first_seed + k for k in range(n)]
    if n == 1:
        return [synth.config_pair("C2", seeds[0], scale)]
    with ProcessPoolExecutor(max_workers=max(1, min(n, (os.cpu_count() or 1) // max(1, world)))) as ex:
        return list(ex.map(_make_one, [(s, scale) for s in seeds]))


def _make_one(arg):
    from describealign_b200 import synth
    return synth.config_pair("C2", arg[0], arg[1])


def audio_hours(pairs):
    return sum(v.shape[0] + a.shape[0] for v, a in pairs) / 44100.0 / 3600.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle port; the reference is Python and not present on the GPU box)
# ------------------------------------------------------------------------------------------------

def cpu_pass(pair, keep=False):
    """One pass of the hot path on the CPU through the oracle; returns (seconds without the host
    fit, seconds of the host fit, outputs or path length)."""
    from describealign_b200 import host_fit
    from oracle import align_oracle as ao, features as of
    v, a = pair
    t0 = time.perf_counter()
    V, A = of.all_features(v), of.all_features(a)
    sa = ao.stage_a(V, A, V[0], A[0])
    t1 = time.perf_counter()
    x, y = sa["path_x"], sa["path_y"]
    kp = host_fit.continuity_error(x, y) < 3
    kx, ky = x[kp], y[kp]
    a_s, v_s = host_fit.scale_features(V, A, kx, ky)
    fx, fy = host_fit.compress_path(kx, ky)
    fit = host_fit.rate_change_fit(fx, fy)
    clusters = host_fit.line_clusters(fit)
    plans = host_fit.plan_corridors(clusters, a_s, v_s)
    t2 = time.perf_counter()
    sb = ao.stage_b(plans, len(clusters), a_s, v_s)
    t3 = time.perf_counter()
    out = len(sb["path"])
    if keep:
        path = sb["path"]
        nx, ny, sim = host_fit.build_nodes(path, len(A[0]), len(V[0]), len(a_s), len(v_s))
        out = {"V": V, "A": A, "path1": (x, y), "path": path, "nodes": (nx, ny), "similarity": sim}
    return (t1 - t0) + (t3 - t2), t2 - t1, out


def _cpu_worker(pair):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    return cpu_pass(pair)


def run_reference(args, rank, world):
    """--impl reference: the oracle port of the reference's CPU path on all host cores (one pair
    per process, which is how the single-threaded reference would be run in parallel)."""
    if rank != 0:
        return
    import oracle
    oracle.build()
    # bounded sample: one pair per host core (at most 16), whatever --pairs / --gpus say
    pairs = make_pairs(max(1, min(16, os.cpu_count() or 1)), 0, args.scale)
    hours = audio_hours(pairs)
    cores = min(len(pairs), os.cpu_count() or 1)
    from concurrent.futures import ProcessPoolExecutor
    times = []
    with ProcessPoolExecutor(max_workers=cores) as ex:
        for step in range(args.warmup + args.steps):
            res = list(ex.map(_cpu_worker, pairs))
            # each worker times its own pass (host fit excluded, as in the GPU arm); with one
            # process per pair running concurrently the step takes as long as the slowest one
            dev = [r[0] for r in res]
            step_s = max(dev) if cores >= len(pairs) else sum(dev) / cores
            if step >= args.warmup:
                times.append(max(step_s, 1e-9))
    ms = 1e3 * float(np.mean(times))
    value = hours / (ms / 1e3)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "pairs_per_step": len(pairs), "audio_hours_per_step": hours,
                       "scale": args.scale},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{len(pairs)} full C2 pairs per step, one oracle process per pair"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------

def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from describealign_b200 import api, build
    build.build()
    # W host threads wait on W pair streams: they must sleep, not spin (16 host cores, W >> 16)
    sched = api._cabi.set_host_wait(local_rank, args.host_wait)
    api.set_device(local_rank)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # B pairs per step, cycling over a few distinct synthetic pairs (generating one takes ~35 s of CPU)
    if args.workers <= 0:
        args.workers = 128 if world <= 2 else (96 if world <= 4 else 64)
    if args.pairs <= 0:
        args.pairs = 4 * args.workers
    B = args.pairs
    distinct = max(1, min(B, args.distinct if args.distinct > 0 else (8 if world <= 2 else 4)))
    base_pairs = make_pairs(distinct, rank * distinct, args.scale, world)
    pairs = [base_pairs[k % distinct] for k in range(B)]
    hours_rank = audio_hours(pairs)

    # device-resident copies (int16 interleaved) and pinned host copies of the distinct pairs
    dev = [(torch.from_numpy(np.ascontiguousarray(v)).cuda(), torch.from_numpy(np.ascontiguousarray(a)).cuda())
           for v, a in base_pairs]
    pinned = [(torch.from_numpy(np.ascontiguousarray(v)).pin_memory(), torch.from_numpy(np.ascontiguousarray(a)).pin_memory())
              for v, a in base_pairs]
    # what the PCIe link gives a single large pinned copy (the bound of the end-to-end number)
    h2d_gbs = None
    try:
        src = pinned[0][1]
        dst = torch.empty_like(src, device="cuda")
        best = 0.0
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); dst.copy_(src, non_blocking=True); e1.record(); e1.synchronize()
            best = max(best, src.numel() * src.element_size() / (e0.elapsed_time(e1) * 1e-3) / 1e9)
        h2d_gbs = best
        del dst
    except Exception:
        pass
    ctx = api.context()
    if args.dp_reserve_kb:
        ctx.set_option("dp_reserve_kb", args.dp_reserve_kb)
    sys.setswitchinterval(args.switch_interval)
    # W pairs in flight: one dab_pair (device buffers + CUDA stream) per worker slot.  The frontier DPs
    # are one warp per pair and tens of milliseconds long, so throughput comes from keeping W of them
    # running while the data-parallel kernels of other pairs fill the SMs.  A step ends with a tail in
    # which only the last DPs run, so a step is several rounds of W pairs (B = 4 W by default).
    W = max(1, min(B, args.workers))
    slot_pairs = [api._cabi.Pair(ctx) for _ in range(W)]
    streams = [torch.cuda.ExternalStream(p.stream) for p in slot_pairs]
    cur = torch.cuda.current_stream()

    import queue
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=W)
    host_cache = {}

    def host_stage(k, job):
        """The rate-change fit is outside the metric (BASELINE.json); its result only depends on the
        pass-1 path, so it is solved once per distinct pair (during warm-up) and reused for as long as
        stage A keeps returning that same path.  Inside the timed region it is a dictionary lookup."""
        c = host_cache.get(k % distinct)
        if c is not None and np.array_equal(c["x"], job.x) and np.array_equal(c["y"], job.y):
            for name in ("kept_x", "kept_y", "audio_scaled", "video_scaled", "fit", "clusters", "plans", "gains"):
                setattr(job, name, c[name])
            return
        job.host_stage()
        host_cache[k % distinct] = {"x": job.x, "y": job.y, **{name: getattr(job, name) for name in
                                    ("kept_x", "kept_y", "audio_scaled", "video_scaled", "fit", "clusters", "plans", "gains")}}

    def one_step(host_input: bool):
        """All B pairs through stage A -> (cached) host fit -> stage B, W at a time.  Returns the
        device time between a start event every pair stream waits on and a stop event that waits on
        every pair stream."""
        jobs = [None] * B
        slots = queue.SimpleQueue()
        for pr in slot_pairs:
            slots.put(pr)

        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def work(k):
            pr = slots.get()
            t_in = time.perf_counter()
            try:
                job = api.AlignJob(pr)
                jobs[k] = job
                if host_input:
                    hv, ha = pinned[k % distinct]
                    job.load_pcm(hv.numpy(), ha.numpy())
                else:
                    v, a = dev[k % distinct]
                    job.load_pcm_device((v.data_ptr(), v.shape[0], v.shape[1]), (a.data_ptr(), a.shape[0], a.shape[1]))
                job.device_stage_a()
                host_stage(k, job)
                job.device_stage_b()
                job.kernel_ms = pr.timings()
                job.work = pr.stats()
                if args.timeline:
                    job.timeline = pr.timeline(start.cuda_event)
                    job.host_span = (t_in, time.perf_counter())
                job.host_ms["whole_pair"] = 1e3 * (time.perf_counter() - t_in)
            finally:
                slots.put(pr)

        start.record(cur)
        t_step = time.perf_counter()
        for st in streams:
            st.wait_event(start)
        list(pool.map(work, range(B)))
        for st in streams:
            e = torch.cuda.Event()
            e.record(st)
            cur.wait_event(e)
        stop.record(cur)
        stop.synchronize()
        if args.timeline:
            tl = [{"pair": k, "slot_host_ms": [1e3 * (j.host_span[0] - t_step), 1e3 * (j.host_span[1] - t_step)], **j.timeline}
                  for k, j in enumerate(jobs)]
            name = args.timeline.replace(".json", "") + ("_e2e.json" if host_input else "_dev.json")
            with open(name, "w") as f:
                json.dump({"step_ms": start.elapsed_time(stop), "host_input": host_input, "pairs": tl}, f)
        return start.elapsed_time(stop), jobs

    alloc0 = {}

    def run_steps(host_input):
        for _ in range(args.warmup):
            one_step(host_input)
        barrier()
        l0 = ctx.launches()
        alloc0.update(api._cabi.alloc_stats())
        total, jobs = 0.0, None
        for _ in range(args.steps):
            ms, jobs = one_step(host_input)
            total += ms
        barrier()
        return total / args.steps, ctx.launches() - l0, jobs

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, launches, jobs = run_steps(False)
    alloc1 = api._cabi.alloc_stats()
    alloc_timed = {k: alloc1[k] - alloc0[k] for k in alloc1}
    clocks = sampler.stop() if rank == 0 else None
    timings = [j.kernel_ms for j in jobs]
    host_calls = {k: {"mean": float(np.mean([j.host_ms.get(k, 0.0) for j in jobs])),
                      "max": float(np.max([j.host_ms.get(k, 0.0) for j in jobs]))}
                  for k in sorted(set().union(*[j.host_ms.keys() for j in jobs]))}
    stats = [j.work for j in jobs]
    ms_e2e, _, jobs_e = run_steps(True)
    h2d = sum(j.h2d_bytes for j in jobs_e)
    d2h = sum(j.d2h_bytes for j in jobs_e)
    host_calls_e2e = {k: {"mean": float(np.mean([j.host_ms.get(k, 0.0) for j in jobs_e])),
                          "max": float(np.max([j.host_ms.get(k, 0.0) for j in jobs_e]))}
                      for k in sorted(set().union(*[j.host_ms.keys() for j in jobs_e]))}

    # one pair alone on the GPU (device-resident PCM, three passes, the last one is kept): kernel
    # durations without the other 63 pairs' work in the way
    solo_ms = solo_work = solo_wall_ms = None
    if rank == 0:
        pr0 = slot_pairs[0]
        for _ in range(3):
            job = api.AlignJob(pr0)
            v0, a0 = dev[0]
            job.load_pcm_device((v0.data_ptr(), v0.shape[0], v0.shape[1]), (a0.data_ptr(), a0.shape[0], a0.shape[1]))
            job.device_stage_a()
            host_stage(0, job)
            job.device_stage_b()
            solo_ms = {k: v for k, v in pr0.timings().items() if not k.startswith("host_in_")}
            solo_work = pr0.stats()
            # BASELINE.json's second figure, "ms per 22-min pair": host wall time of the device stages and
            # of the result copies for this one pair (the host fit between the stages is not in it)
            solo_wall_ms = sum(job.host_ms.get(k, 0.0) for k in ("stage_a", "stage_b", "get_features", "path1", "path2"))

    # parity inside the run: rank 0 checks its first pair against the oracle
    parity = None
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        import oracle
        oracle.build()
        cpu_s, fit_s, o = cpu_pass(pairs[0], keep=True)
        # one more (untimed) pass of pair 0 through the public API, keeping its intermediates
        det = {}
        import contextlib
        with contextlib.redirect_stdout(sys.stderr):     # progress labels (describealign.py:604,635,726,860) stay off the JSON line
            nx, ny, sim, path, med = api.align_pcm(pairs[0][0], pairs[0][1], details=det)

        class _J:
            video_features, audio_features = det["video_features"], det["audio_features"]
            x, y = det["path1"]
        j0 = _J
        ox, oy = o["nodes"]
        opath = o["path"]
        same_shape = path.shape == opath.shape
        parity = {
            "features_f32_identical": bool(all(np.array_equal(j0.video_features[k], o["V"][k]) and
                                               np.array_equal(j0.audio_features[k], o["A"][k])
                                               for k in range(min(4, len(j0.video_features))))),
            "path1_identical": bool(np.array_equal(j0.x, o["path1"][0]) and np.array_equal(j0.y, o["path1"][1])),
            "path2_rows": int(len(path)),
            "path2_int_identical": bool(same_shape and np.array_equal(path[:, 1], opath[:, 1]) and
                                        np.array_equal(path[:, 2], opath[:, 2]) and
                                        np.allclose(path[:, 0], opath[:, 0], rtol=0, atol=1e-9)),
            "nodes_max_abs_diff_s": float(max(np.max(np.abs(nx - ox)), np.max(np.abs(ny - oy)))) if len(nx) == len(ox) else None,
            "similarity_diff": float(abs(sim - o["similarity"])),
        }
        h1 = audio_hours(pairs[:1])
        cpu = {"value": h1 / cpu_s, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "1 full C2 pair through the oracle (C features, C match + DPs, numpy corridors); host fit (%.2f s) excluded as in the GPU arm" % fit_s,
               "seconds": cpu_s, "host_cpu": _cpu_model(), "host_cores": os.cpu_count()}

    # max over ranks
    t = torch.tensor([ms_dev, ms_e2e], device="cuda", dtype=torch.float64)
    tot = torch.tensor([hours_rank, float(launches), float(h2d), float(d2h)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_dev, ms_e2e = float(t[0]), float(t[1])
    hours, launches_all, h2d_all, d2h_all = (float(x) for x in tot)

    if rank == 0:
        peaks, peak_src = measured_peaks()
        # per-kernel device times of the last timed step, summed over this rank's pairs (64 in flight: the
        # events around a kernel then also count the time it waited for SMs), and of one pair alone
        agg = {k: sum(tm[k] for tm in timings) for k in timings[0] if not k.startswith("host_in_")}
        lib_host = {k: float(np.mean([tm[k] for tm in timings])) for k in timings[0] if k.startswith("host_in_")}
        clk_hz = 1e6 * float(peaks.get("sm_max_mhz", 1965.0))

        def rooflines(kernel_ms, work, pair_list):
            feat_ms = kernel_ms["features_video"] + kernel_ms["features_audio"]
            # feature kernel: algorithmic bytes = PCM read once + 24 B per output frame
            feat_bytes = sum(2 * (v.shape[0] * v.shape[1] + a.shape[0] * a.shape[1]) + 24 * (v.shape[0] // 210 + a.shape[0] // 210)
                             for v, a in pair_list)
            out = {"features": {"bound": "hbm", "achieved": feat_bytes / (feat_ms * 1e-3) / 1e9 if feat_ms > 0 else None,
                                "peak": peaks["hbm_gbs"], "unit": "GB/s", "ms": feat_ms, "launches": 2 * len(pair_list),
                                "algorithmic_bytes": feat_bytes}}
            # DP kernels: 24 B per point (i, v, qual in; back pointer out); what bounds them is
            # the dependent chain: one f64 add + select per point, measured at 52 SM cycles on this chip
            # (profiles/r1_v18_dp2_block_hot_loops.txt)
            for key, npts in (("dp1_trace", work["n_points1"]), ("dp2_trace", work["n_points2"])):
                ms = kernel_ms[key]
                out[key] = {"bound": "hbm", "achieved": 24 * npts / (ms * 1e-3) / 1e9 if ms > 0 else None,
                            "peak": peaks["hbm_gbs"], "unit": "GB/s", "ms": ms, "launches": len(pair_list), "points": npts,
                            "points_per_s": npts / (ms * 1e-3) if ms > 0 else None,
                            "note": "latency-bound serial dependency, not bandwidth-bound",
                            "serial_chain_floor_ms": 1e3 * 52.0 * npts / clk_hz,
                            "frac_of_serial_chain_floor": (1e3 * 52.0 * npts / clk_hz) / ms if ms > 0 else None}
            for v in out.values():
                v["frac"] = v["achieved"] / v["peak"] if v.get("achieved") is not None else None
            return out

        work_all = {k: sum(st[k] for st in stats) for k in stats[0]}
        roof_load = rooflines(agg, work_all, pairs)
        roof = rooflines(solo_ms, solo_work, pairs[:1])
        dominant = max((k for k in solo_ms if k in roof or k.startswith("features")), key=lambda k: solo_ms[k])
        dom_key = dominant if dominant in roof else "features"
        main_roof = dict(roof[dom_key])
        # DRAM bytes per launch of that kernel from the committed ncu --set full captures of the same
        # workload (seed 0 pair): profiles/r1_v18_dp2_block_full.txt, profiles/r1_v8_features_full.txt
        ncu_traffic = {"dp2_trace": 12.918784e6 + 2048, "features": (116.465152e6 + 5.258240e6 + 143.008e6 + 9.717248e6) / 2}
        main_roof.update({"kernel": dom_key, "peak_source": peak_src,
                          "measured": "CUDA events on the pair's stream, one C2 pair alone on the GPU right after the timed steps",
                          "traffic": ncu_traffic.get(dom_key) if args.scale == 1.0 else None,
                          "traffic_note": "bytes per launch (dram read + write), ncu --set full of one C2 pair, see profiles/"})
        line = {
            "metric": METRIC, "value": hours / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32+f64 (u32 packed codes)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_gpu_per_step": B, "pairs_in_flight_per_gpu": W,
                       "distinct_pairs_per_gpu": distinct, "audio_hours_per_step": hours,
                       "l2_policy": "inputs larger than L2 (each pair streams 259 MB of PCM, %d distinct pairs per step; 126 MB L2)" % distinct,
                       "timed": "one region per step: every pair through device stage A (features, prep, tables, gate, score, DP1, traceback) and device stage B (corridors, DP2, traceback), W pairs in flight; the host rate-change fit between them is solved during warm-up and is a cache lookup inside the region (untimed by BASELINE.json)",
                       "scale": args.scale},
            "ms_per_pair": ms_dev / B,
            "pairs_in_flight": W,
            "e2e": {"value": hours / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d_all, "d2h_bytes_per_step": d2h_all,
                    "h2d_gbs_achieved": h2d_all / world / (ms_e2e * 1e-3) / 1e9,
                    "h2d_gbs_single_copy": h2d_gbs,
                    "note": "PCM is 2 bytes per sample per channel; the end-to-end rate is bounded by the host-to-device link"},
            "host_wait": {"cuda_schedule_flags": sched, "mode": ["spin (CUDA default)", "blocking sync", "query + sleep polling"][args.host_wait]},
            "gpu_launches": int(launches_all),
            "roofline": main_roof,
            "roofline_by_kernel": roof,
            "roofline_by_kernel_under_load": roof_load,
            "kernel_ms_one_pair_alone": solo_ms,
            "ms_per_pair_alone": {"device_stages_and_result_copies_host_wall": solo_wall_ms,
                                  "kernels_only": sum(v for k, v in solo_ms.items() if k != "dp2"),
                                  "cpu_port_same_pair": 1e3 * cpu["seconds"] if cpu else None,
                                  "note": "one C2 pair (22-min video, 27-min description) alone on the GPU, PCM device-resident"},
            "kernel_ms_last_step": agg,
            "host_call_ms_last_step": host_calls,
            "host_call_ms_last_step_e2e": host_calls_e2e,
            "host_ms_inside_library_per_pair": lib_host,
            "allocator_activity_in_timed_steps": alloc_timed,
            "work": work_all,
            "clocks": clocks,
            "cpu_baseline": cpu,
            "parity": parity,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
