"""Multi-GPU plumbing on CPU: world_size-2 ``gloo``.  Episodes shard contiguously with no data-path
collective; only the final metric gather + rank-ordered aggregate combination cross ranks.  The per-rank result
blocks here are synthetic (the kernels need a GPU); what is tested is that ``gather_results`` reassembles ragged
shards into the exact whole and that the aggregate combination is deterministic and rank-ordered."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bgr_pathplanning_control_b200 import _abi as abi
from bgr_pathplanning_control_b200 import sharding, sweeps


def test_shard_ranges_cover_exactly():
    for E in (0, 1, 7, 1000, 1 << 20, (1 << 22) + 3):
        for G in (1, 2, 3, 4, 8):
            rs = [sharding.shard_range(E, r, G) for r in range(G)]
            assert rs[0][0] == 0 and rs[-1][1] == E
            assert all(rs[i][1] == rs[i + 1][0] for i in range(G - 1))
            sizes = sharding.shard_sizes(E, G)
            assert sum(sizes) == E and max(sizes) - min(sizes) <= 1


def test_sweep_rows_are_shard_invariant():
    """Row i of every parameter generator depends only on (seed, i): any shard reproduces its rows."""
    whole = sweeps.stanley_lqg_params(200000, start=0)
    lo, hi = sharding.shard_range(200000, 1, 3)
    assert np.array_equal(sweeps.stanley_lqg_params(hi - lo, start=lo), whole[lo:hi])
    whole = sweeps.pp_sweep_params(5000, start=100)
    assert np.array_equal(sweeps.pp_sweep_params(1000, start=2100), whole[2000:3000])
    whole = sweeps.monte_carlo_params(70000)
    assert np.array_equal(sweeps.monte_carlo_params(100, start=65500), whole[65500:65600])


def _fake_block(lo, hi):
    """Deterministic stand-in for a rank's per-episode results: a function of the GLOBAL episode id only."""
    e = np.arange(lo, hi, dtype=np.float64)
    m = np.stack([np.sin(e * (k + 1) * 1e-3) for k in range(abi.N_METRICS)], axis=1).astype(np.float32)
    s = np.stack([(e * (k + 3)).astype(np.int64) % 1000 for k in range(abi.N_STATUS)], axis=1).astype(np.int32)
    agg = np.zeros(abi.N_AGG)
    agg[abi.A_EPISODES] = hi - lo
    agg[abi.A_STEPS] = s[:, abi.S_STEPS].sum()
    agg[abi.A_SUM_LAT_MEAN] = m[:, abi.M_LAT_MEAN].astype(np.float64).sum()
    agg[abi.A_MIN_LAP_TIME] = 10.0 + lo
    agg[abi.A_CONE_HITS] = s[:, abi.S_CONE_HITS].sum()
    return m, s, agg


def _worker(rank, world, port, E, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = sharding.shard_range(E, rank, world)
        m, s, agg = _fake_block(lo, hi)
        tm, ts, ta = torch.from_numpy(m), torch.from_numpy(s), torch.from_numpy(agg)
        M, S, A = sharding.gather_results(tm, ts, ta, E)                       # gather to rank 0 + reduce
        Ma, Sa, Aa = sharding.gather_results(tm, ts, ta, E, dst=None)          # all-gather variant
        _, _, Ar = sharding.gather_results(tm, ts, ta, E, rows=False)          # reduce only
        assert (M is None) == (rank != 0) and (S is None) == (rank != 0)
        assert torch.equal(A, Aa) and torch.equal(A, Ar)
        if rank == 0:
            assert torch.equal(M, Ma) and torch.equal(S, Sa)
        q.put((rank, Ma.numpy(), Sa.numpy(), A.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("E", [1001, 64])
def test_gather_results_world_size_2_gloo(E):
    world = 2
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, E, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m, s, _ = _fake_block(0, E)
    parts = [_fake_block(*sharding.shard_range(E, r, world))[2] for r in range(world)]
    want = sharding.combine_aggregates(parts)
    for rank, M, S, A in got:                      # identical on every rank, equal to the unsharded whole
        assert np.array_equal(M, m) and np.array_equal(S, s)
        assert np.array_equal(A, want)
        assert A[abi.A_EPISODES] == E and A[abi.A_MIN_LAP_TIME] == 10.0
