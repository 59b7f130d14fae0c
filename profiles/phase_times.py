"""Per-phase cycle split of the rollout kernel (needs the -DMJPC_PHASE_TIMING build: MJPC_B200_SO=profiles/var_phase.so)."""
import os, sys
import numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
from conftest import get_model, quadruped_inputs
from mujoco_mpc_b200.engine import Engine
m = get_model("quadruped")
e = Engine(m, 256, 64)
d = np.load(os.path.join(R, "profiles", "inputs_quadruped_256x64.npz"))
state, mocap, knots, kt = d["state"], d["mocap"], d["knots"], d["kt"]   # the profiled inputs (prof_rollout.py)
for _ in range(2):
    e.rollout_spline(state, 0.0, mocap, knots, kt, 2, 64)
st = e.fetch_stats().astype(float)
tot = st[:, 0]
names = ["kinematics+com+crb", "collision", "make_constraint", "vel+smooth+reference", "solve (rest)", "solve: Hessian assembly", "solve: Cholesky factor+solve", "policy+residual+cost+euler+output"]
print("kernel %.2f ms; per-candidate cycles median %.3g" % (e.last_kernel_ms, np.median(tot)))
for k, n in enumerate(names):
    if n != "-":
        print("  %-36s %5.1f%% of cycles  (%.0f cycles/step)" % (n, 100 * st[:, 4 + k].sum() / tot.sum(), st[:, 4 + k].mean() / 64))
print("  newton iterations/step %.2f  -> cycles per Newton iteration: all of solve %.0f, Hessian %.0f, Cholesky %.0f" % (
    st[:, 1].mean() / 64, st[:, 8:11].sum() / st[:, 1].sum(), st[:, 9].sum() / st[:, 1].sum(), st[:, 10].sum() / st[:, 1].sum()))
