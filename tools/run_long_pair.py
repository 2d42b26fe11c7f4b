#!/usr/bin/env python
"""One pair on all ranks (torchrun, NCCL): match stage and corridor scoring sharded by audio rows, the DPs
and the host fit on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        tools/run_long_pair.py [--seconds S | --config C5] [--check]

--check compares rank 0's result with the single-GPU path on the same PCM (identical outputs).
Prints one JSON line with device timings of the sharded stage.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=240.0)
    ap.add_argument("--config", default=None, help="a synth.config_pair name, e.g. C5")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from describealign_b200 import api, batch, build, synth
    build.build()
    rank, world, local = batch.init_from_env("nccl")
    if args.config:
        v, a = synth.config_pair(args.config, 0, args.scale)
    else:
        v, a = synth.make_pair(args.seconds, 7.0, skips=[(args.seconds * 0.4, 2.0), (args.seconds * 0.7, -1.5)], seed=31)
    api.context()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    det = {}
    t0 = time.perf_counter()
    out = batch.align_long_pair(v, a, details=det)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    ok = True
    if args.check:
        ref = api.align_pcm(v, a)
        ok = all(np.array_equal(np.asarray(x), np.asarray(y)) for x, y in zip(out[:2], ref[:2])) and \
            np.array_equal(out[3], ref[3]) and out[2] == ref[2]
    flags = torch.tensor([1 if ok else 0], device="cuda")
    if world > 1:
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "wall_s": t1 - t0, "shards": det.get("shards"), "timings": det.get("timings"),
                          "stats": det.get("stats"), "audio_hours": (len(v) + len(a)) / 44100 / 3600}))
        if int(flags.item()) == 1:
            print("long pair ok")
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if int(flags.item()) == 1 else 1)


if __name__ == "__main__":
    main()
