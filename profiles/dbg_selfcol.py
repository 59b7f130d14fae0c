"""Debug: one quadruped step, device vs oracle, with / without self-collision pairs, static vs generic kernel."""
import os, sys, numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
from conftest import mocap_of
from mujoco_mpc_b200 import models
from mujoco_mpc_b200.blob import to_blob
from mujoco_mpc_b200.engine import Engine
from oracle import pyoracle
for sc in (False, True):
    m = models.load("quadruped", self_collision=sc)
    o = pyoracle.Oracle(to_blob(m), m, 64)
    e = Engine(m, 4, 8)
    rng = np.random.default_rng(0)
    for trial in range(3):
        q = m.key_qpos[0].copy(); q[2] = [0.245, 0.25, 0.26][trial]; q[7:] += rng.normal(size=12) * 0.05
        v = rng.normal(size=m.nv) * (0.0 if trial == 0 else 0.3); u = rng.uniform(-1, 1, m.nu)
        r = o.forward_debug(q, v, u, mocap_of(m))
        g = e.step_debug(q, v, u, mocap_of(m))
        b = e.step_batch(q[None], v[None], u[None], mocap_of(m), [0.0])
        print("selfcol", sc, "trial", trial, "npair", m.npair, "| oracle ncon %d nefc %d niter %d | debug(generic) ncon %d nefc %d niter %d warn %d err %.2e | batch(static=%s) niter %d warn %d err %.2e"
              % (r["ncon"], r["nefc"], r["niter"], g["ncon"], g["nefc"], g["niter"], g["warning"], np.abs(g["qacc"] - r["qacc"]).max(),
                 e.last_kernel_static, b["niter"][0], b["warning"][0], np.abs(b["qacc"][0] - r["qacc"]).max()))
    e.close()
