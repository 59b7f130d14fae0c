"""CPU, world_size = 2 over gloo: the N > 1 candidate-sharding path (same host logic bench.py runs over NCCL).
The per-rank backend here is the CPU oracle - allowed in tests - so the check is: sharded == single-process."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import get_model, quadruped_inputs


class _OracleBackend:
    def __init__(self, m):
        from mujoco_mpc_b200.blob import to_blob
        from oracle import pyoracle
        self.o = pyoracle.Oracle(to_blob(m), m, 64)

    def rollout_spline(self, state, time, mocap, knots, kt, interp, H):
        r = self.o.rollout_spline(state, time, mocap, knots, kt, interp, H, nthreads=2, full=False)
        return r["returns"], r["failure"], np.argsort(r["returns"], kind="stable")


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mujoco_mpc_b200.planner import SamplingPlanner
    from mujoco_mpc_b200.sharding import ShardedRollouts
    m = get_model("quadruped")
    sh = ShardedRollouts(_OracleBackend(m), dist)
    state, mocap, knots, kt = quadruped_inputs(m, N=11, H=12)     # 11 candidates: ragged shards (6 + 5)
    ret, fail, order = sh.rollout_spline(state, 0.0, mocap, knots, kt, 2, 12)
    # two planner iterations on top of the sharded backend: every rank must install the same policy
    pl = SamplingPlanner(m, sh, num_trajectory=11, horizon=12)
    pl.reset(); pl.set_state(state, 0.0, mocap)
    pl.optimize_policy(); pl.optimize_policy()
    q.put((rank, ret, fail, order, pl.values, pl.winner))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds():
    from mujoco_mpc_b200.sharding import shard_bounds
    assert shard_bounds(256, 8) == [(32 * r, 32 * (r + 1)) for r in range(8)]
    assert shard_bounds(11, 2) == [(0, 6), (6, 11)]
    assert shard_bounds(3, 4) == [(0, 1), (1, 2), (2, 3), (3, 3)]     # empty shard on the last rank
    for N, G in ((1024, 8), (7, 3), (1, 1)):
        b = shard_bounds(N, G)
        assert b[0][0] == 0 and b[-1][1] == N and all(b[i][1] == b[i + 1][0] for i in range(G - 1))


def test_two_rank_gloo_matches_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(2)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m = get_model("quadruped")
    state, mocap, knots, kt = quadruped_inputs(m, N=11, H=12)
    ref = _OracleBackend(m).rollout_spline(state, 0.0, mocap, knots, kt, 2, 12)
    for rank, ret, fail, order, values, winner in res:
        np.testing.assert_allclose(ret, ref[0].astype(np.float32), rtol=1e-6)     # bit-identical up to the fp32 transport
        assert list(order) == list(np.argsort(ref[0].astype(np.float32), kind="stable"))
        assert not fail.any()
    np.testing.assert_array_equal(res[0][4], res[1][4])      # same installed policy on both ranks
    assert res[0][5] == res[1][5]
