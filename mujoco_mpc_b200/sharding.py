"""Multi-GPU sharding of the candidate batch (one process per GPU, torch.distributed for the plumbing).

Candidates are independent units given (state, time, mocap, task snapshot, nominal knots), so the path shards
with NO data-path collective: rank g owns the contiguous candidate range [g*N/G, (g+1)*N/G) (candidate 0, the
un-noised nominal, lives on rank 0; mjpc/planners/sampling/planner.cc:374).  The single exchange step per planning
iteration is an all-gather of the per-candidate returns (N floats: pure latency over NVLink/NVSwitch).  With an Engine
that owns an NCCL communicator (Engine.comm_init) all of it - sharding, ncclAllGather on the engine stream, ranking -
runs inside libmjpc_b200.so (mjpc_b200_rollout_spline_sharded) and this class only forwards; the torch.distributed
path below is the host-logic twin used by the CPU tests (gloo, oracle backend).  Every rank computes the identical
ranking and installs the same policy.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(N: int, world: int):
    """Contiguous, balanced ranges; the first N % world ranks get one extra candidate."""
    base, rem = divmod(N, world)
    lo = [r * base + min(r, rem) for r in range(world)]
    return [(lo[r], lo[r] + base + (1 if r < rem else 0)) for r in range(world)]


def rank_order(returns: np.ndarray) -> np.ndarray:
    """Ascending return, ties to the lower index (what rank_kernel computes on the device)."""
    return np.argsort(returns, kind="stable")


class ShardedRollouts:
    """Wraps this rank's backend (Engine, or anything with rollout_spline) behind the single-process signature."""

    def __init__(self, backend, dist=None, device="cpu"):
        self.backend = backend
        self.dist = dist
        self.device = device
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1
        self.last_local = None

    def rollout_spline(self, state, time, mocap, knots, knot_times, interp, H):
        # product path: the engine owns an NCCL communicator (Engine.comm_init) - sharding, the all-gather of returns on
        # the engine stream and the ranking all happen inside libmjpc_b200.so (mjpc_b200_rollout_spline_sharded)
        if getattr(self.backend, "nranks", 1) > 1 and hasattr(self.backend, "rollout_spline_sharded"):
            N = knots.shape[0]
            self.last_local = shard_bounds(N, self.world)[self.rank]
            return self.backend.rollout_spline_sharded(state, time, mocap, knots, knot_times, interp, H)
        import torch
        N = knots.shape[0]
        bounds = shard_bounds(N, self.world)
        lo, hi = bounds[self.rank]
        ret, fail, _ = self.backend.rollout_spline(state, time, mocap, knots[lo:hi], knot_times, interp, H)
        self.last_local = (lo, hi)
        if self.world == 1:
            return np.asarray(ret), np.asarray(fail), rank_order(np.asarray(ret))
        width = max(b[1] - b[0] for b in bounds)
        buf = torch.full((width, 2), np.inf, dtype=torch.float32, device=self.device)
        buf[: hi - lo, 0] = torch.as_tensor(np.asarray(ret, np.float32), device=self.device)
        buf[: hi - lo, 1] = torch.as_tensor(np.asarray(fail, np.float32), device=self.device)
        out = torch.empty((self.world, width, 2), dtype=torch.float32, device=self.device)
        self.dist.all_gather_into_tensor(out.view(-1), buf.view(-1))   # the one collective per planning iteration
        out = out.cpu().numpy()
        full_ret = np.concatenate([out[r, : b[1] - b[0], 0] for r, b in enumerate(bounds)])
        full_fail = np.concatenate([out[r, : b[1] - b[0], 1] for r, b in enumerate(bounds)]).astype(np.uint8)
        return full_ret, full_fail, rank_order(full_ret)

    def owner_of(self, candidate: int, N: int) -> int:
        for r, (lo, hi) in enumerate(shard_bounds(N, self.world)):
            if lo <= candidate < hi:
                return r
        raise IndexError(candidate)

    def fetch_winner(self, candidate: int, N: int):
        """Trajectory of a (global) candidate index from the rank that rolled it out; None on the other ranks."""
        lo, hi = self.last_local
        if lo <= candidate < hi and hasattr(self.backend, "fetch_trajectory"):
            return self.backend.fetch_trajectory(candidate - lo)
        return None
