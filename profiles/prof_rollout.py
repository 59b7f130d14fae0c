import sys, numpy as np
import os; R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
from conftest import get_model, mocap_of, quadruped_inputs
from mujoco_mpc_b200.engine import Engine
m = get_model("quadruped")
e = Engine(m, 256, 64)
state, mocap, knots, kt = quadruped_inputs(m, N=256, H=64)
for i in range(2):
    ret, fail, order = e.rollout_spline(state, 0.0, mocap, knots, kt, 2, 64)
print("kernel ms", e.last_kernel_ms)
st = e.fetch_stats()
cyc = st[:, 0] / 1.965e6
print("per-candidate ms: min %.2f median %.2f max %.2f ; newton iters/step mean %.2f max-cand %.2f ; ncon/step %.2f nefc/step %.2f" % (
    cyc.min(), np.median(cyc), cyc.max(), st[:, 1].mean() / 64, st[:, 1].max() / 64, st[:, 2].mean() / 64, st[:, 3].mean() / 64))
print("slowest candidates", np.argsort(-cyc)[:5], np.sort(-cyc)[:5])
tr = e.fetch_trajectory(5)
# solver iteration counts along a trajectory
its = []
for t in range(0, 64, 4):
    s = tr["states"][t]
    g = e.step_debug(s[:19], s[19:], tr["actions"][t], mocap)
    its.append((t, g["ncon"], g["nefc"], g["niter"]))
print(its)
