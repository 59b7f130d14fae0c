"""Host-side iLQG driver around the device sweeps (Python mirror of mjpc/planners/ilqg/planner.cc).

  NominalTrajectory   :167-223  feedback-scaling line search over K time-indexed rollouts, BestRollout
  Iteration           :377-627  model derivatives -> cost derivatives -> backward pass (with the regularisation
                                retry loop :429-520) -> K action rollouts -> winner, expected / improvement /
                                surprise, UpdateRegularization (backward_pass.cc:327-356)
  BestRollout         :727-740  scans from the last rollout to the first with a strict '<' (ties -> larger index)
  LogScale            mjpc/utilities.cc:819-825; step sizes 1 -> min_linesearch_step, last forced to 0

The backend supplies the five hot-path calls of the C ABI (rollout_feedback, fetch_trajectory, model_derivatives,
cost_derivatives, backward_pass); mujoco_mpc_b200.engine.Engine is the product backend, tests also run the same
driver on the CPU oracle.
"""
from __future__ import annotations

import math

import numpy as np


def log_scale(max_value, min_value, steps):
    step = (math.log(max_value) - math.log(min_value)) / max(steps - 1, 1)
    return np.array([math.exp(math.log(min_value) + i * step) for i in range(steps)])


class ILQGSettings:                      # mjpc/planners/ilqg/settings.h:21-36
    min_linesearch_step = 1.0e-3
    fd_tolerance = 1.0e-6
    min_regularization = 1.0e-6
    max_regularization = 1.0e6
    regularization_type = 0              # 0 control, 1 feedback, 2 value, 3 none
    max_regularization_iterations = 5
    action_limits = 1
    nominal_feedback_scaling = 1
    fd_mode = 0                          # 0 one-sided, 1 centred (settings.h:24)
    derivative_skip = 0                  # iLQGPlanner::derivative_skip_
    differentiable = 1                   # Agent::PlanIteration MakeDifferentiable default for gradient planners (agent.cc:158-164)


class ILQGPlanner:
    def __init__(self, model, backend, horizon, num_rollouts=None, fd_tolerance=None, representation=None, fd_mode=None):
        m = self.model = model
        self.backend = backend
        self.settings = ILQGSettings()
        if fd_tolerance is not None:
            self.settings.fd_tolerance = fd_tolerance
        if fd_mode is not None:
            self.settings.fd_mode = int(fd_mode)   # the reference default (0, with 1e-6) is an fp64 setting; the engine wants centred 3e-4
        self.H = int(horizon)
        self.K = int(num_rollouts or m.numeric.get("ilqg_num_rollouts", [10])[0])
        self.representation = int(representation if representation is not None else m.numeric.get("ilqg_representation", [1])[0])
        self.nu, self.ds, self.n = m.nu, m.nq + m.nv, 2 * m.nv
        self.reset()

    # -- iLQGPolicy::Reset / iLQGBackwardPass::Reset
    def reset(self, initial_repeated_action=None):
        H = self.H
        a = np.zeros(self.nu) if initial_repeated_action is None else np.asarray(initial_repeated_action, float)
        self.states = np.zeros((H, self.ds)); self.actions = np.tile(a, (H, 1)); self.times = np.zeros(H)
        self.gains = np.zeros((H, self.nu, self.n)); self.du = np.zeros((H, self.nu))
        self.total_return = 0.0
        self.residual = None
        self.regularization, self.regularization_rate, self.regularization_factor = 1.0, 1.0, 2.0
        self.feedback_scaling = 1.0
        self.winner = 0
        self.improvement = self.expected = self.surprise = 0.0

    def set_state(self, state, time, mocap):
        self.state, self.time, self.mocap = np.asarray(state, float), float(time), np.asarray(mocap, float)

    def _steps(self):
        s = log_scale(1.0, self.settings.min_linesearch_step, self.K - 1) if self.K > 1 else np.zeros(0)
        # LogScale ascends from min to max in the reference (values[i] = exp(log(min) + i*step))
        return np.concatenate([s, [0.0]])

    @staticmethod
    def _best(returns, failure):
        best, best_ret = -1, 0.0
        for j in range(len(returns) - 1, -1, -1):
            if failure[j]:
                continue
            if best == -1 or returns[j] < best_ret:
                best_ret, best = returns[j], j
        return best

    def _fetch_candidate(self, tr, ret):
        """candidate_policy[0].trajectory = trajectory[i] (planner.cc:214,560): the working copy, not the live policy"""
        c = self.cand
        c["states"] = np.asarray(tr["states"], float); c["actions"] = np.asarray(tr["actions"], float)
        c["times"] = np.asarray(tr["times"], float); c["residual"] = np.asarray(tr["residual"], float)
        c["total_return"] = float(ret)

    def _differentiable(self, on=True):
        """MakeDifferentiable while planning, restored afterwards (agent.cc:296-309,346-356)"""
        if self.settings.differentiable and hasattr(self.backend, "set_differentiable"):
            self.backend.set_differentiable(on)

    # -- iLQGPlanner::NominalTrajectory (planner.cc:167-223): works on a copy; the live policy is not touched
    def nominal_trajectory(self):
        steps = self._steps()
        self.cand = dict(states=self.states.copy(), actions=self.actions.copy(), times=self.times.copy(),
                         residual=None if self.residual is None else self.residual.copy(), gains=self.gains.copy(),
                         du=self.du.copy(), total_return=self.total_return)
        c = self.cand
        self._differentiable(True)
        try:
            ret, fail, _ = self.backend.rollout_feedback(self.state, self.time, self.mocap, c["actions"], c["states"],
                                                         c["times"], c["gains"], None, steps, self.representation)
        finally:
            self._differentiable(False)
        best = self._best(ret, fail)
        if best == -1:
            self.feedback_scaling = 0.0
            return False
        self._fetch_candidate(self.backend.fetch_trajectory(best), ret[best])
        self.feedback_scaling = float(steps[best])       # planner diagnostic (planner.cc:217); the live policy keeps 1
        return True

    def _scale_regularization(self, factor):
        s = self.settings
        if factor > 1:
            self.regularization_rate = max(self.regularization_rate * factor, factor)
        else:
            self.regularization_rate = min(self.regularization_rate * factor, factor)
        self.regularization = min(max(self.regularization * self.regularization_rate, s.min_regularization), s.max_regularization)

    def _update_regularization(self, z, s_):
        f = self.regularization_factor
        if not (math.isfinite(z) and math.isfinite(s_)):
            self._scale_regularization(f * f)
        elif z > 0.5 or s_ > 0.3:
            self._scale_regularization(1.0 / f)
        elif z < 0.1 or s_ < 0.06:
            self._scale_regularization(f)

    # -- iLQGPlanner::Iteration (planner.cc:377-627)
    def iteration(self):
        self._differentiable(True)
        try:
            return self._iteration()
        finally:
            self._differentiable(False)

    def _iteration(self):
        s = self.settings
        c = self.cand
        previous_return = c["total_return"]
        steps = self._steps()
        if c["residual"] is None:
            return False                 # no nominal rollout succeeded yet: nothing to differentiate
        A, B, C, D = self.backend.model_derivatives(c["states"], c["actions"], c["times"], self.mocap, s.fd_tolerance,
                                                    skip=s.derivative_skip, mode=s.fd_mode)
        cx, cu, cxx, cuu, cxu = self.backend.cost_derivatives(c["residual"], C, D)
        status, reg_iter, bp = 0, 0, None
        while reg_iter < s.max_regularization_iterations and status == 0:
            bp = self.backend.backward_pass(A, B, cx, cu, cxx, cxu, cuu, c["actions"], mu=self.regularization,
                                            reg_type=s.regularization_type, limits=s.action_limits)
            status = int(bp["status"])
            if status == 0 and self.regularization <= s.max_regularization:
                self._scale_regularization(self.regularization_factor)
                reg_iter += 1
        if status == 0:
            return False                 # backward-pass failure: the live policy is untouched (:523-531)
        c["gains"], c["du"], self.dV = np.asarray(bp["K"], float), np.asarray(bp["du"], float), np.asarray(bp["dV"], float)
        ret, fail, _ = self.backend.rollout_feedback(self.state, self.time, self.mocap, c["actions"], c["states"],
                                                     c["times"], c["gains"], c["du"], steps, 3)
        self.last_returns = np.asarray(ret, float)      # trajectory[j].total_return of the K action rollouts
        best = self._best(ret, fail)
        if best == -1:
            return False                 # nothing is published (:548-550)
        self.winner = best
        action_step = float(steps[best])
        old = dict(states=c["states"].copy(), actions=c["actions"].copy(), times=c["times"].copy(), residual=c["residual"].copy())
        self._fetch_candidate(self.backend.fetch_trajectory(best), ret[best])
        self.expected = -1.0 * action_step * (self.dV[0] + action_step * self.dV[1]) + 1.0e-16
        self.improvement = previous_return - c["total_return"]
        self.surprise = min(max(0.0, self.improvement / self.expected), 2.0)
        self._update_regularization(self.surprise, action_step)
        # policy.CopyFrom(candidate_policy[winner]) (:597-605): candidate j keeps the OLD nominal states with actions
        # old + step_j * du (:639-643); only candidate 0 received the winning rollout (:560)
        if best == 0:
            self.states, self.actions, self.times, self.residual = c["states"], c["actions"], c["times"], c["residual"]
        else:
            self.states, self.times, self.residual = old["states"], old["times"], old["residual"]
            self.actions = np.asarray(np.asarray(old["actions"], np.float32) + np.float32(action_step) * np.asarray(c["du"], np.float32), float)
        self.gains, self.du = c["gains"], c["du"]
        self.total_return = c["total_return"]
        return True

    # -- iLQGPolicy::Action on the live policy (ilqg/policy.cc:82-161) through the host library
    def action_from_policy(self, time, state=None):
        from .engine import host_ilqg_policy_action
        return host_ilqg_policy_action(self.model, self.actions, self.states, self.times, self.gains, self.representation,
                                       1.0, state, time)

    # -- iLQGPlanner::OptimizePolicy
    def optimize_policy(self):
        self.nominal_trajectory()
        return self.iteration()
