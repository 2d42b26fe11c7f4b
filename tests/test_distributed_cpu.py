"""N > 1 host logic on the CPU: world_size 2, gloo backend (no GPU needed).

Covers the two multi-GPU protocols of describealign_b200/batch.py:
* batch mode: longest-first assignment, per-rank execution, gather of results on rank 0;
* long pair: audio rows of the match stage sharded over ranks, all-gather of the scored match
  points, DP #1 on the reassembled list.
The compute inside the ranks is the oracle (this is a CPU test of the plumbing); the GPU tests run
the same protocols with the CUDA path in the ranks.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _init(rank, world, port):
    import torch.distributed as dist
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    return dist


def _oracle_runner(pair):
    from describealign_b200 import host_fit
    from oracle import align_oracle as ao, features as of
    v, a = pair
    V, A = of.all_features(v), of.all_features(a)
    return ao.align(V, A, V[0], A[0], host_fit)


def _batch_specs():
    # (video seconds, offset seconds, skips, seed): different lengths so that the assignment matters
    return [(45.0, 4.0, [(20.0, 1.5)], 11), (75.0, 6.0, [(30.0, 2.0)], 7), (40.0, 3.0, [], 12)]


def _make(spec):
    from describealign_b200 import synth
    vs, off, skips, seed = spec
    return synth.make_pair(vs, off, skips=skips, seed=seed)


def _batch_worker(rank, world, port, out_dir):
    dist = _init(rank, world, port)
    from describealign_b200 import batch
    specs = _batch_specs()
    loaders = [(lambda s=s: _make(s)) for s in specs]
    durations = [2 * s[0] + s[1] for s in specs]
    res = batch.align_batch(loaders, durations=durations, in_flight=2, runner=_oracle_runner)
    mine = batch.assign_pairs(durations, world)[rank]
    np.save(os.path.join(out_dir, f"mine{rank}.npy"), np.array(mine))
    if rank == 0:
        assert res is not None and len(res) == len(specs)
        for k, r in enumerate(res):
            assert not isinstance(r, Exception), r
            np.save(os.path.join(out_dir, f"nodes{k}.npy"), np.stack([r[0], r[1]]))
            np.save(os.path.join(out_dir, f"path{k}.npy"), r[3])
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def test_batch_mode_gloo(tmp_path):
    import torch.multiprocessing as mp
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import oracle
    oracle.build()
    mp.spawn(_batch_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    from describealign_b200 import batch
    specs = _batch_specs()
    durations = [2 * s[0] + s[1] for s in specs]
    want_assign = batch.assign_pairs(durations, 2)
    assert sorted(want_assign[0] + want_assign[1]) == [0, 1, 2]
    assert want_assign[0][0] == 1, "the longest pair goes first to rank 0"
    for r in range(2):
        assert list(np.load(tmp_path / f"mine{r}.npy")) == want_assign[r]
    # results equal a serial run in this process
    for k, spec in enumerate(specs):
        want = _oracle_runner(_make(spec))
        got_nodes = np.load(tmp_path / f"nodes{k}.npy")
        np.testing.assert_array_equal(got_nodes[0], want[0])
        np.testing.assert_array_equal(got_nodes[1], want[1])
        np.testing.assert_array_equal(np.load(tmp_path / f"path{k}.npy"), want[3])


def _long_worker(rank, world, port, out_dir):
    dist = _init(rank, world, port)
    import torch
    from describealign_b200 import batch, synth
    from oracle import align_oracle as ao, features as of
    v, a = synth.make_pair(90.0, 5.0, skips=[(40.0, 2.0)], seed=21)
    V, A = of.all_features(v), of.all_features(a)
    lo, hi = batch.row_shards(len(A[0]), world)[rank]
    pi, pv, pq, Lv, _ = ao.match_points(V, A, V[0], A[0], rows=(lo, hi))
    gi, gv, gq = batch.exchange_points(torch.from_numpy(pi), torch.from_numpy(pv), torch.from_numpy(pq))
    x, y, _, _ = ao.dp1(gi.numpy(), gv.numpy(), gq.numpy(), Lv)
    np.savez(os.path.join(out_dir, f"long{rank}.npz"), i=gi.numpy(), v=gv.numpy(), q=gq.numpy(), x=x, y=y,
             n_local=len(pi))
    dist.barrier()
    dist.destroy_process_group()


def test_long_pair_row_sharding_gloo(tmp_path):
    import torch.multiprocessing as mp
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import oracle
    oracle.build()
    mp.spawn(_long_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    from describealign_b200 import synth
    from oracle import align_oracle as ao, features as of
    v, a = synth.make_pair(90.0, 5.0, skips=[(40.0, 2.0)], seed=21)
    V, A = of.all_features(v), of.all_features(a)
    full = ao.stage_a(V, A, V[0], A[0])
    parts = [np.load(tmp_path / f"long{r}.npz") for r in range(2)]
    assert all(int(p["n_local"]) > 0 for p in parts), "both shards must contribute points"
    assert int(parts[0]["n_local"]) + int(parts[1]["n_local"]) == len(full["points_i"])
    for p in parts:     # every rank ends up with the full sorted list and the same path
        np.testing.assert_array_equal(p["i"], full["points_i"])
        np.testing.assert_array_equal(p["v"], full["points_v"])
        np.testing.assert_array_equal(p["q"], full["points_q"])
        np.testing.assert_array_equal(p["x"], full["path_x"])
        np.testing.assert_array_equal(p["y"], full["path_y"])


def test_assign_pairs_and_row_shards():
    from describealign_b200 import batch
    a = batch.assign_pairs([10, 50, 20, 20, 5], 2)
    assert a == [[1, 0], [2, 3, 4]] or sorted(a[0] + a[1]) == [0, 1, 2, 3, 4]
    loads = [sum([10, 50, 20, 20, 5][k] for k in part) for part in a]
    assert max(loads) <= 60
    assert batch.assign_pairs([], 3) == [[], [], []]
    assert batch.assign_pairs([1.0], 4) == [[0], [], [], []]
    with pytest.raises(ValueError):
        batch.assign_pairs([1.0], 0)
    for n, w in ((10, 3), (0, 2), (7, 8), (340200, 8)):
        sh = batch.row_shards(n, w)
        assert len(sh) == w and sh[0][0] == 0 and sh[-1][1] == n
        assert all(sh[k][1] == sh[k + 1][0] for k in range(w - 1))
        assert max(h - l for l, h in sh) - min(h - l for l, h in sh) <= 1


def test_long_pair_generator_is_deterministic_and_has_the_c5_shape():
    """synth.long_pair builds the C5 shape from independently generated segments (worker processes or not: same PCM)."""
    from describealign_b200 import synth
    v1, a1 = synth.long_pair(3, 0.004, workers=0)
    v2, a2 = synth.long_pair(3, 0.004, workers=3)
    assert np.array_equal(v1, v2) and np.array_equal(a1, a2)
    assert v1.dtype == np.int16 and v1.shape[1] == 1
    assert abs(v1.shape[0] / 44100 - 9000 * 0.004) < 0.01 and abs(a1.shape[0] / 44100 - 10800 * 0.004) < 0.05
    v3, _ = synth.long_pair(4, 0.004)
    assert not np.array_equal(v1, v3)


def test_long_pair_host_phases_are_named():
    from describealign_b200 import batch
    assert set(batch.HOST_PHASES) == {"host_fit_and_broadcast", "nodes_on_host_and_broadcast"}
