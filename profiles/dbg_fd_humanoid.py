import os, sys
import numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
from conftest import get_model, mocap_of
from mujoco_mpc_b200.engine import Engine
from mujoco_mpc_b200.blob import to_blob
from mujoco_mpc_b200.planner import candidate_knots
from oracle import pyoracle
m = get_model("humanoid")
e = Engine(m, 64, 64)
o = pyoracle.Oracle(to_blob(m), m, 64); o32 = pyoracle.Oracle(to_blob(m), m, 32)
H, P = 8, 3
cr = np.asarray(m.actuator_ctrlrange).reshape(-1, 2)
state = np.concatenate([m.qpos0, np.zeros(m.nv)])
kt = np.arange(P) * (H - 1) * m.opt_timestep / (P - 1)
knots = candidate_knots(np.zeros((P, m.nu)), 0.3, cr, 0, 2)[1:2]
r = o.rollout_spline(state, 0.0, mocap_of(m), knots, kt, 2, H)
xs, us, ts = r["states"][0], r["actions"][0], r["times"][0]
A, B, C, D = e.model_derivatives(xs, us, ts, mocap_of(m), 1e-3)
Ao, Bo, Co, Do = o.model_derivatives(xs, us, ts, mocap_of(m), tol=1e-3)
A3, B3, C3, D3 = o32.model_derivatives(xs, us, ts, mocap_of(m), tol=1e-3)
for nm, G, Rr, R3 in (("A", A, Ao, A3), ("B", B, Bo, B3)):
    err = np.abs(G - Rr); e3 = np.abs(R3 - Rr)
    idx = np.dstack(np.unravel_index(np.argsort(-err.ravel())[:6], err.shape))[0]
    print(nm, "device-vs-fp64 max %.3g ; fp32oracle-vs-fp64 max %.3g" % (err.max(), e3.max()))
    for t, i, j in idx:
        print("   t=%d row=%d col=%d  dev %.4f  o64 %.4f  o32 %.4f" % (t, i, j, G[t, i, j], Rr[t, i, j], R3[t, i, j]))
# how many columns are affected at the worst t
errB = np.abs(B - Bo); t = np.unravel_index(errB.argmax(), errB.shape)[0]
print("worst t", t, "per-column max err", np.round(errB[t].max(0), 3))
for tt in range(H):
    g = e.step_debug(xs[tt, :m.nq], xs[tt, m.nq:], us[tt], mocap_of(m))
    rr = o.forward_debug(xs[tt, :m.nq], xs[tt, m.nq:], us[tt], mocap_of(m))
    print("t", tt, "ncon/nefc dev", g["ncon"], g["nefc"], "oracle", rr["ncon"], rr["nefc"], "niter", g["niter"], rr["niter"])
