"""Timing probe for BASELINE config 4 (Quadruped, iLQG, H=64): per-sweep device+transfer time through the C ABI."""
import os, sys, time
import numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
from conftest import get_model, mocap_of
from mujoco_mpc_b200.engine import Engine
from mujoco_mpc_b200.ilqg import ILQGPlanner
m = get_model("quadruped")
e = Engine(m, 64, 64)
pl = ILQGPlanner(m, e, horizon=64, num_rollouts=10, fd_tolerance=1e-3)
pl.set_state(np.concatenate([m.key_qpos[0], np.zeros(m.nv)]), 0.0, mocap_of(m))
for _ in range(3):
    pl.optimize_policy()
H, n, nu = 64, 36, 12
def tm(f, reps=5):
    f(); t0 = time.perf_counter()
    for _ in range(reps): out = f()
    return (time.perf_counter() - t0) / reps * 1e3, out
t_fd, (A, B, C, D) = tm(lambda: e.model_derivatives(pl.states, pl.actions, pl.times, pl.mocap, 1e-3))
t_cd, cd = tm(lambda: e.cost_derivatives(pl.residual, C, D))
t_bp, bp = tm(lambda: e.backward_pass(A, B, cd[0], cd[1], cd[2], cd[4], cd[3], pl.actions, mu=pl.regularization))
t_ro, _ = tm(lambda: e.rollout_feedback(pl.state, 0.0, pl.mocap, pl.actions, pl.states, pl.times, bp["K"], bp["du"], pl._steps(), 3))
t_it, _ = tm(lambda: pl.optimize_policy(), reps=5)
fd_steps = H * (1 + nu + 2 * m.nv)
print("quadruped iLQG H=64 (ms, host call incl. H2D/D2H): model_derivatives %.2f (%d mj_step-equivalents, %.2e steps/s) | cost_derivatives %.2f | "
      "backward_pass %.2f (status %d) | 10 line-search rollouts %.2f | full OptimizePolicy %.2f | return %.4f" % (
          t_fd, fd_steps, fd_steps / (t_fd * 1e-3), t_cd, t_bp, bp["status"], t_ro, t_it, pl.total_return))
