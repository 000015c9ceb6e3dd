"""Frame sharding + the statistics all-reduce (the only multi-rank exchange), exercised with gloo on CPU at
world_size 2.  The per-shard compute here is the CPU oracle: tests may use it, the product may not."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition_the_frames():
    wifi = importlib.import_module("80211parallelestimation_b200")
    for n in (0, 1, 7, 8, 1000003, 8 * 2 ** 20):
        for g in (1, 2, 3, 4, 8):
            edges = [wifi.shard_range(n, r, g) for r in range(g)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(g - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        wifi.shard_range(10, 2, 2)


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import synth
    from oracle.pyoracle import Oracle
    wifi = importlib.import_module("80211parallelestimation_b200")
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    fr = synth.make_frames(n, seed=42)                       # every rank can regenerate the global sequence
    est = wifi.ShardedEstimator(n, rank, world)
    o = Oracle()

    def fn(lo, cnt):
        H = o.lt_ls(fr["tx_pre"][lo:lo + cnt], fr["rx_pre"][lo:lo + cnt])
        ref = fr["H_true"][lo:lo + cnt].copy(); ref[:, 26] = 0
        d = np.abs(H - ref)
        return torch.tensor([np.sum(d ** 2), np.sum(np.abs(ref) ** 2), d.size, d.max()], dtype=torch.float64)

    res = est.reduce_stats(est.run(fn))
    q.put((rank, est.lo, est.hi, res))
    dist.destroy_process_group()


def test_two_rank_stats_allreduce_matches_single_rank():
    import torch.multiprocessing as mp
    n = 101
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    [p.start() for p in procs]
    got = sorted(q.get(timeout=120) for _ in procs)
    [p.join(60) for p in procs]
    assert [g[1:3] for g in got] == [(0, 50), (50, 101)]
    assert got[0][3] == got[1][3]                            # both ranks hold the same reduced numbers
    # single-rank answer
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import synth
    from oracle.pyoracle import Oracle
    fr = synth.make_frames(n, seed=42)
    H = Oracle().lt_ls(fr["tx_pre"], fr["rx_pre"]); ref = fr["H_true"].copy(); ref[:, 26] = 0
    d = np.abs(H - ref)
    res = got[0][3]
    assert res["count"] == d.size
    assert abs(res["sum_sq_err"] - np.sum(d ** 2)) < 1e-12 * np.sum(d ** 2) + 1e-30
    assert res["max_abs_err"] == d.max()
    assert 0 < res["nmse"] < 1e-2
