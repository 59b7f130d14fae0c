"""Gradient-descent planner and iLQS planner around the device sweeps (Python mirrors of
mjpc/planners/gradient/planner.cc:159-330, gradient.cc:44-107, spline_mapping.cc, and mjpc/planners/ilqs/planner.cc:87-215).

  spline_mapping      linear operator "spline parameters -> actions at the trajectory times" (Zero / Linear / Cubic;
                      spline_mapping.cc:35-208).  It is block diagonal over the action dimension, so only the scalar
                      [num_output][num_input] weights are stored (the reference stores their Kronecker product with I_nu)
  gradient_sweep      Gradient::Compute (gradient.cc:44-107): Vx backward recursion, k_t = -Qu_t, dV[0] = sum k.Qu
  GradientPlanner     ResamplePolicy, nominal rollout, {model derivatives, cost derivatives, gradient sweep, total
                      derivative through the mapping, K line-search rollouts} x max_rollout, winner, policy update
  ILQSPlanner         alternates Predictive Sampling and one iLQG iteration; converts the iLQG trajectory policy to spline
                      parameters through the least-squares inverse of the mapping when sampling follows iLQG

GradientPolicy::Action (gradient/policy.cc:80-103) is FindInterval + Zero/Linear/CubicInterpolation (utilities.cc:303-422)
+ Clamp: for >= 3 spline points the same function as TimeSpline::Sample (spline.cc:103-156, 250-287), which is what the
rollout kernel evaluates, so the line-search rollouts are ONE mjpc_b200_rollout_spline launch of K candidates.
"""
from __future__ import annotations

import numpy as np

from .ilqg import ILQGPlanner, log_scale
from .planner import SamplingPlanner, clamp, sample_spline


def find_interval(seq, value):
    """FindInterval (utilities.h:125-144): upper_bound, then clamp."""
    length = len(seq)
    upper = int(np.searchsorted(np.asarray(seq, float), value, side="right"))
    lower = upper - 1
    if lower < 0:
        return 0, 0
    if lower > length - 1:
        return length - 1, length - 1
    return max(lower, 0), min(upper, length - 1)


def spline_mapping(input_times, output_times, representation):
    """W [num_output][num_input] with action(t_i) = sum_p W[i][p] * parameter_p (before clamping)."""
    ti = np.asarray(input_times, float); to = np.asarray(output_times, float)
    P, T = len(ti), len(to)
    W = np.zeros((T, P))
    if representation == 0:
        for i in range(T):
            W[i, find_interval(ti, to[i])[0]] = 1.0
        return W
    if representation == 1:
        for i in range(T):
            b0, b1 = find_interval(ti, to[i])
            if b0 == b1:
                W[i, b0] = 1.0
            else:
                a = (to[i] - ti[b0]) / (ti[b1] - ti[b0])
                W[i, b0] = 1.0 - a; W[i, b1] = a
        return W
    # cubic: points + finite-difference slopes (spline_mapping.cc:118-205)
    S = np.zeros((2 * P, P))
    S[:P] = np.eye(P)
    for i in range(P):
        dt1 = 1.0 / (ti[i] - ti[i - 1]) if i > 0 else 0.0
        dt2 = 1.0 / (ti[i + 1] - ti[i]) if i < P - 1 else 0.0
        if 0 < i < P - 1:
            dt1 *= 0.5; dt2 *= 0.5
        if i - 1 >= 0:
            S[P + i, i - 1] = -dt1
        S[P + i, i] = dt1 - dt2
        if i + 1 <= P - 1:
            S[P + i, i + 1] = dt2
    O = np.zeros((T, 2 * P))
    for i in range(T):
        b0, b1 = find_interval(ti, to[i])
        if b0 == b1:
            c = (1.0, 0.0, 0.0, 0.0)
        else:
            t = (to[i] - ti[b0]) / (ti[b1] - ti[b0]); d = ti[b1] - ti[b0]
            c = (2 * t ** 3 - 3 * t ** 2 + 1, (t ** 3 - 2 * t ** 2 + t) * d, -2 * t ** 3 + 3 * t ** 2, (t ** 3 - t ** 2) * d)
        O[i, b0] = c[0]; O[i, P + b0] = c[1]
        if b0 != b1:
            O[i, b1] = c[2]; O[i, P + b1] = c[3]
    return O @ S


def gradient_sweep(A, B, cx, cu):
    """Gradient::Compute: k [T][m] (last row repeats the previous one), dV0."""
    T = cx.shape[0]
    m = cu.shape[1]
    k = np.zeros((T, m))
    Vx = cx[T - 1].astype(float).copy()
    dV0 = 0.0
    for t in range(T - 1, 0, -1):
        Qx = cx[t - 1] + A[t - 1].T @ Vx
        Qu = cu[t - 1] + B[t - 1].T @ Vx
        k[t - 1] = -Qu
        Vx = Qx
        dV0 += float(k[t - 1] @ Qu)
    if T >= 2:
        k[T - 1] = k[T - 2]
    return k, dV0


class GradientSettings:                   # gradient/settings.h:21-27
    max_rollout = 1
    min_linesearch_step = 1.0e-8
    fd_tolerance = 1.0e-5
    fd_mode = 0
    action_limits = 1
    derivative_skip = 0
    differentiable = 1                    # agent.cc:158-164: gradient-based planners plan on the differentiable model


class GradientPlanner:
    def __init__(self, model, backend, horizon, num_trajectory=None, num_spline_points=None, representation=None,
                 fd_tolerance=None, fd_mode=None):
        m = self.model = model
        self.backend = backend
        self.settings = GradientSettings()
        if fd_tolerance is not None:
            self.settings.fd_tolerance = fd_tolerance
        if fd_mode is not None:
            self.settings.fd_mode = int(fd_mode)
        num = m.numeric
        self.H = int(horizon)
        self.K = int(num_trajectory or num.get("gradient_num_trajectory", [32])[0])
        self.P = int(num_spline_points or num.get("gradient_spline_points", [min(self.H, 25)])[0])
        self.representation = int(representation if representation is not None else num.get("gradient_representation", [1])[0])
        self.timestep = float(m.opt_timestep)
        self.ctrlrange = np.asarray(m.actuator_ctrlrange, float).reshape(-1, 2)
        self.nu = m.nu
        self.reset()

    def reset(self, initial_repeated_action=None):
        a = np.zeros(self.nu) if initial_repeated_action is None else np.asarray(initial_repeated_action, float)
        self.parameters = np.tile(a, (self.P, 1)); self.times = np.zeros(self.P)
        self.previous = (self.parameters.copy(), self.times.copy())
        self.winner = -1
        self.action_step = self.expected = self.improvement = self.surprise = 0.0
        self.best = None

    def set_state(self, state, time, mocap):
        self.state, self.time, self.mocap = np.asarray(state, float), float(time), np.asarray(mocap, float)

    def _action(self, params, times, t):
        return clamp(sample_spline(times, params, self.representation, t), self.ctrlrange)

    def _steps(self):
        s = log_scale(1.0, self.settings.min_linesearch_step, self.K - 1) if self.K > 1 else np.zeros(0)
        return np.concatenate([s, [0.0]])

    def _differentiable(self, on):
        if self.settings.differentiable and hasattr(self.backend, "set_differentiable"):
            self.backend.set_differentiable(on)

    # ResamplePolicy (planner.cc:356-383)
    def resample(self, params, times):
        shift = max((self.H - 1) * self.timestep / (self.P - 1), 1.0e-5) if self.P > 1 else 1.0e-5
        new_t = self.time + shift * np.arange(self.P)
        new_p = np.stack([self._action(params, times, tt) for tt in new_t])
        return new_p, new_t

    def optimize_policy(self):
        self._differentiable(True)
        try:
            return self._optimize_policy()
        finally:
            self._differentiable(False)

    def _optimize_policy(self):
        s, be, H = self.settings, self.backend, self.H
        params, times = self.resample(self.parameters, self.times)
        # nominal rollout (candidate 0)
        ret, fail, _ = be.rollout_spline(self.state, self.time, self.mocap, params[None], times, self.representation, H)
        tr = be.fetch_trajectory(0)
        c_prev = float(ret[0]) if not fail[0] else 1.0e6
        c_best = c_prev
        steps = self._steps()
        cand = np.repeat(params[None], self.K, 0)
        self.winner = self.K - 1
        for _ in range(s.max_rollout):
            A, B, C, D = be.model_derivatives(tr["states"], tr["actions"], tr["times"], self.mocap, s.fd_tolerance,
                                              skip=s.derivative_skip, mode=int(s.fd_mode))
            cx, cu, cxx, cuu, cxu = be.cost_derivatives(tr["residual"], C, D)
            k, dV0 = gradient_sweep(np.asarray(A, float), np.asarray(B, float), np.asarray(cx, float), np.asarray(cu, float))
            W = spline_mapping(times, np.asarray(tr["times"], float)[: H - 1], self.representation)
            update = W.T @ k[: H - 1]                                          # mapping^T k (planner.cc:240-245)
            cand = params[None] + steps[:, None, None] * update[None]          # Rollouts: parameters += step * update
            ret, fail, _ = be.rollout_spline(self.state, self.time, self.mocap, cand, times, self.representation, H)
            self.winner = self.K - 1
            for j in range(self.K - 1, -1, -1):
                c = float(ret[j]) if not fail[j] else 1.0e6
                if c < c_best:
                    c_best, self.winner = c, j
            params = cand[self.winner].astype(float)
            tr = be.fetch_trajectory(self.winner)
            self.action_step = float(steps[self.winner])
            self.expected = -self.action_step * dV0 - 1.0e-16
            self.improvement = c_prev - c_best
            self.surprise = min(max(0.0, self.improvement / self.expected), 2.0)
        if c_best >= c_prev:
            self.winner = self.K - 1
            params = cand[self.winner].astype(float)                            # step 0: the resampled nominal
        self.previous = (self.parameters, self.times)
        self.parameters, self.times = params, times
        self.best = tr
        self.total_return = c_best
        return c_best < c_prev

    def action_from_policy(self, time, use_previous=False):
        p, t = self.previous if use_previous else (self.parameters, self.times)
        return self._action(p, t, time)


class ILQSPlanner:
    """iLQSPlanner (ilqs/planner.cc): a SamplingPlanner and an ILQGPlanner on the same state; `active_policy` says whose
    policy acts."""
    K_SAMPLING, K_ILQG = 0, 1

    def __init__(self, model, sampling_backend, ilqg_backend, horizon, num_trajectory=None, num_rollouts=None,
                 fd_tolerance=None, seed=0x5EED, fd_mode=None):
        self.model = model
        self.sampling = SamplingPlanner(model, sampling_backend, num_trajectory=num_trajectory, horizon=horizon, seed=seed)
        self.ilqg = ILQGPlanner(model, ilqg_backend, horizon=horizon, num_rollouts=num_rollouts, fd_tolerance=fd_tolerance, fd_mode=fd_mode)
        self.H = int(horizon)
        self.reset()

    def reset(self, initial_repeated_action=None):
        self.sampling.reset(initial_repeated_action); self.ilqg.reset(initial_repeated_action)
        self.active_policy = self.previous_active_policy = self.K_SAMPLING

    def set_state(self, state, time, mocap):
        self.sampling.set_state(state, time, mocap); self.ilqg.set_state(state, time, mocap)

    def optimize_policy(self):
        sp, il, H = self.sampling, self.ilqg, self.H
        self.previous_active_policy = self.active_policy
        if self.previous_active_policy == self.K_ILQG:
            # spline parameters from the iLQG trajectory policy: least-squares inverse of the mapping (planner.cc:98-172)
            il.nominal_trajectory()
            P = sp.P
            shift = max((H - 1) * sp.timestep / (P - 1), 1.0e-5)
            new_t = sp.time + shift * np.arange(P)
            W = spline_mapping(new_t, np.asarray(il.cand["times"], float)[: H - 1], sp.interp)
            M = W.T @ W
            L = np.linalg.cholesky(M)
            inv = np.linalg.solve(L.T, np.linalg.solve(L, W.T))               # (A'A)^-1 A'
            params = inv @ np.asarray(il.cand["actions"], float)[: H - 1]
            # sampling.policy.plan = these nodes (:160-172).  Restated literally: with the default non-sliding plan the
            # very next step, SamplingPlanner::UpdateNominalPolicy, resamples candidate_policy[winner] - NOT policy - into
            # policy.plan (sampling/planner.cc:296-321), so the converted spline never reaches a rollout; it is kept for
            # inspection only (the C++ class writes it into sampling.policy.plan exactly as the reference does).
            self.converted = (new_t, clamp(params, sp.ctrlrange))
        ret, fail = sp.optimize_policy()
        ref = float(ret[0]) if self.previous_active_policy == self.K_SAMPLING else float(il.cand["total_return"])
        if sp.winner > 0 and float(ret[sp.winner]) < ref:
            self.active_policy = self.K_SAMPLING
            return True
        if self.previous_active_policy == self.K_SAMPLING:
            # ilqg.candidate_policy[0].trajectory = sampling.trajectory[0] (the un-noised nominal rollout)
            tr = sp.backend.fetch_trajectory(0)
            il.cand = dict(states=np.asarray(tr["states"], float), actions=np.asarray(tr["actions"], float),
                           times=np.asarray(tr["times"], float), residual=np.asarray(tr["residual"], float),
                           gains=il.gains.copy(), du=il.du.copy(), total_return=float(ret[0]))
        ok = il.iteration()
        if ok:
            new = il.total_return
            # ilqg.trajectory[0] is the action rollout with the SMALLEST non-zero step (LogScale ascends): restated literally
            old = float(ret[sp.winner]) if self.previous_active_policy == self.K_SAMPLING else float(il.last_returns[0])
            if new < old:
                self.active_policy = self.K_ILQG
        return ok

    def action_from_policy(self, time, state=None):
        if self.active_policy == self.K_SAMPLING:
            return self.sampling.action_from_policy(time)
        return self.ilqg.action_from_policy(time, state)
