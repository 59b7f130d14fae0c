"""Dump every recorded array of the profiled 256x64 launch (for bitwise comparison of experiment builds):
python profiles/dump_all.py out.npz ; python profiles/dump_all.py --cmp a.npz b.npz"""
import sys, os, numpy as np
if sys.argv[1] == "--cmp":
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    for k in a.files:
        x, y = a[k], b[k]
        if np.array_equal(x, y): print("%-10s bitwise" % k); continue
        d = np.argwhere(x != y)
        first_t = d[:, 1].min() if d.ndim == 2 and d.shape[1] > 1 else -1
        print("%-10s differs in %d entries (%d candidates), first step %d, max abs %.3g; first idx %s" % (k, len(d), len(set(d[:, 0])), first_t, np.abs(x - y).max(), d[0]))
    sys.exit(0)
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
from conftest import get_model
from mujoco_mpc_b200.engine import Engine
m = get_model("quadruped")
e = Engine(m, 256, 64)
d = np.load(os.path.join(R, "profiles", "inputs_quadruped_256x64.npz"))
ret, fail, order = e.rollout_spline(d["state"], 0.0, d["mocap"], d["knots"], d["kt"], 2, 64)
o = e.fetch_all()
np.savez(sys.argv[1], returns=ret, **o)
print("kernel ms", e.last_kernel_ms)
