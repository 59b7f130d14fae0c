"""Interleaved A/B of the two static kernel shapes at a given N (same process, same box): kernel ms of alternating launches."""
import os, sys, numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
from conftest import get_model
from mujoco_mpc_b200.engine import Engine
m = get_model("quadruped")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
e = Engine(m, N, 64)
d = np.load(os.path.join(R, "profiles", "inputs_quadruped_256x64.npz"))
kn = np.concatenate([d["knots"]] * ((N + 255) // 256))[:N]
ms = {"wide": [], "plain": []}
for rep in range(12):
    for shape in ("wide", "plain"):
        os.environ["MJPC_B200_SHAPE"] = shape
        e.rollout_spline(d["state"], 0.0, d["mocap"], kn, d["kt"], 2, 64)
        if rep >= 2: ms[shape].append(e.last_kernel_ms)
for k, v in ms.items():
    print("N=%d %s: kernel ms min %.3f median %.3f max %.3f" % (N, k, min(v), np.median(v), max(v)))
