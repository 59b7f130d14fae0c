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


class ILQGPlanner:
    def __init__(self, model, backend, horizon, num_rollouts=None, fd_tolerance=None, representation=None):
        m = self.model = model
        self.backend = backend
        self.settings = ILQGSettings()
        if fd_tolerance is not None:
            self.settings.fd_tolerance = fd_tolerance
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

    def _install(self, tr, ret):
        self.states = np.asarray(tr["states"], float); self.actions = np.asarray(tr["actions"], float)
        self.times = np.asarray(tr["times"], float); self.residual = np.asarray(tr["residual"], float)
        self.total_return = float(ret)

    # -- iLQGPlanner::NominalTrajectory
    def nominal_trajectory(self):
        steps = self._steps()
        ret, fail, _ = self.backend.rollout_feedback(self.state, self.time, self.mocap, self.actions, self.states,
                                                     self.times, self.gains, None, steps, self.representation)
        best = self._best(ret, fail)
        if best == -1:
            self.feedback_scaling = 0.0
            return False
        self._install(self.backend.fetch_trajectory(best), ret[best])
        self.feedback_scaling = float(steps[best])
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

    # -- iLQGPlanner::Iteration
    def iteration(self):
        s = self.settings
        previous_return = self.total_return
        steps = self._steps()
        A, B, C, D = self.backend.model_derivatives(self.states, self.actions, self.times, self.mocap, s.fd_tolerance)
        cx, cu, cxx, cuu, cxu = self.backend.cost_derivatives(self.residual, C, D)
        status, reg_iter, bp = 0, 0, None
        while reg_iter < s.max_regularization_iterations and status == 0:
            bp = self.backend.backward_pass(A, B, cx, cu, cxx, cxu, cuu, self.actions, mu=self.regularization,
                                            reg_type=s.regularization_type, limits=s.action_limits)
            status = int(bp["status"])
            if status == 0 and self.regularization <= s.max_regularization:
                self._scale_regularization(self.regularization_factor)
                reg_iter += 1
        if status == 0:
            return False
        self.gains, self.du, self.dV = np.asarray(bp["K"], float), np.asarray(bp["du"], float), np.asarray(bp["dV"], float)
        ret, fail, _ = self.backend.rollout_feedback(self.state, self.time, self.mocap, self.actions, self.states,
                                                     self.times, self.gains, self.du, steps, 3)
        best = self._best(ret, fail)
        if best == -1:
            return False
        self.winner = best
        self._install(self.backend.fetch_trajectory(best), ret[best])
        action_step = float(steps[best])
        self.expected = -1.0 * action_step * (self.dV[0] + action_step * self.dV[1]) + 1.0e-16
        self.improvement = previous_return - self.total_return
        self.surprise = min(max(0.0, self.improvement / self.expected), 2.0)
        self._update_regularization(self.surprise, action_step)
        self.feedback_scaling = 1.0
        return True

    # -- iLQGPlanner::OptimizePolicy
    def optimize_policy(self):
        self.nominal_trajectory()
        return self.iteration()
