"""Kernel-time probe for BASELINE config 3 (Humanoid Track PS, H=128, P=16 cubic, dt 0.005): per-GPU share 128
candidates (the 8-GPU configuration) and the whole 1024 on one GPU."""
import os, sys
import numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
from conftest import get_model
from mujoco_mpc_b200.engine import Engine
m = get_model("humanoid_track")
H, P = 128, 16
mocap = np.concatenate([m.key_mpos[0].reshape(-1, 3), np.tile([1.0, 0, 0, 0], (m.nmocap, 1))], 1).reshape(-1)
state = np.concatenate([m.key_qpos[0], np.zeros(m.nv)])
kt = np.arange(P) * (H - 1) * 0.005 / (P - 1)
rng = np.random.default_rng(0)
for N in (128, 256, 1024):
    e = Engine(m, N, H)
    knots = np.clip(0.15 * rng.standard_normal((N, P, m.nu)), -1, 1); knots[0] = 0
    for i in range(3):
        ret, fail, order = e.rollout_spline(state, 0.0, mocap, knots, kt, 2, H)
    st = e.fetch_stats()
    print("humanoid track N=%d H=%d static=%d: kernel %.2f ms -> %.3e env-steps/s | newton/step %.2f contacts/step %.2f rows/step %.1f | failures %d" % (
        N, H, e.last_kernel_static, e.last_kernel_ms, N * H / (e.last_kernel_ms * 1e-3), st[:, 1].mean() / H, st[:, 2].mean() / H,
        st[:, 3].mean() / H, int(fail.sum())))
    e.close()
