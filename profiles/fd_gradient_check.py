"""How good is the iLQG gradient on the device?  dJ/du of the nominal return from the FD model derivatives + cost
derivatives + the backward recursion (gradient_sweep), against brute-force central differences of the return itself
through device rollouts, for several FD step sizes and one-sided / centred differences (Quadruped, MakeDifferentiable on).
The fp64 oracle with eps 1e-6 (the reference's setting) is the yardstick.  usage: python profiles/fd_gradient_check.py [H]"""
import os, sys
import numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
from conftest import get_model, mocap_of, OracleBackend
from mujoco_mpc_b200.gradient import gradient_sweep
from mujoco_mpc_b200.ilqg import ILQGPlanner
m = get_model("quadruped")
H = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 64
state = np.concatenate([m.key_qpos[0], np.zeros(m.nv)])
probes = [(0, 1), (0, 2), (5, 4), (10, 7), (20, 10), (H // 2, 0), (H - 8, 5), (H - 4, 8)]


def run(be, label, combos, brute_h):
    be.set_differentiable(True)
    pl = ILQGPlanner(m, be, horizon=H, num_rollouts=4, fd_tolerance=1e-3)
    pl.settings.differentiable = 0
    pl.set_state(state, 0.0, mocap_of(m))
    pl.nominal_trajectory(); c = pl.cand
    gains = np.zeros((H, m.nu, 2 * m.nv))

    def J(du, a):
        ret, fail, _ = be.rollout_feedback(state, 0.0, pl.mocap, c["actions"], c["states"], c["times"], gains, du, np.array([a, -a]), 3)
        return (ret[0] - ret[1]) / (2 * a)
    brute = []
    for (t, j) in probes:
        du = np.zeros((H, m.nu)); du[t, j] = 1.0
        brute.append(J(du, brute_h))
    brute = np.array(brute)
    print("%s: brute-force dJ/du (h = %g): %s" % (label, brute_h, np.array2string(brute, precision=6)))
    for eps, mode in combos:
        A, B, C, D = be.model_derivatives(c["states"], c["actions"], c["times"], pl.mocap, eps, skip=0, mode=mode)
        cx, cu, cxx, cuu, cxu = be.cost_derivatives(c["residual"], C, D)
        k, _ = gradient_sweep(np.asarray(A, float), np.asarray(B, float), np.asarray(cx, float), np.asarray(cu, float))
        g = np.array([-k[t, j] for (t, j) in probes])
        print("  eps %-6g %-9s analytic %s | rel err vs brute: median %.2f max %.2f" % (
            eps, "centred" if mode else "one-sided", np.array2string(g, precision=6),
            np.median(np.abs(g - brute) / (np.abs(brute) + 1e-6)), np.max(np.abs(g - brute) / (np.abs(brute) + 1e-6))))
    be.set_differentiable(False)
    return brute


if "--oracle" in sys.argv:
    run(OracleBackend(m, threads=8), "fp64 oracle", [(1e-6, 0), (1e-6, 1), (1e-3, 0), (1e-4, 1)], 1e-4)
else:
    from mujoco_mpc_b200.engine import Engine
    e = Engine(m, 64, H)
    run(e, "device fp32", [(1e-3, 0), (3e-4, 0), (1e-4, 0), (1e-3, 1), (3e-4, 1), (1e-4, 1), (3e-5, 1)], 3e-3)
    run(OracleBackend(m, threads=8), "fp64 oracle", [(1e-6, 0), (1e-6, 1)], 1e-4)
    # what the planner makes of it: 8 iLQG iterations from the home keyframe (zero nominal), return after each
    for eps, mode in [(1e-3, 0), (3e-4, 0), (1e-3, 1), (3e-4, 1), (1e-4, 1), (3e-5, 1)]:
        pl = ILQGPlanner(m, e, horizon=H, num_rollouts=10, fd_tolerance=eps)
        pl.settings.fd_mode = mode
        pl.set_state(state, 0.0, mocap_of(m))
        pl.nominal_trajectory(); first = pl.cand["total_return"]
        rets = []
        for _ in range(8):
            pl.optimize_policy(); rets.append(pl.total_return)
        print("iLQG device eps %-6g %-9s first %.6f -> %s" % (eps, "centred" if mode else "one-sided", first, np.array2string(np.array(rets), precision=6)))
    ob = OracleBackend(m, threads=8)
    pl = ILQGPlanner(m, ob, horizon=H, num_rollouts=10, fd_tolerance=1e-6)
    pl.set_state(state, 0.0, mocap_of(m))
    pl.nominal_trajectory(); first = pl.cand["total_return"]
    rets = []
    for _ in range(8):
        pl.optimize_policy(); rets.append(pl.total_return)
    print("iLQG fp64 oracle eps 1e-6 one-sided first %.6f -> %s" % (first, np.array2string(np.array(rets), precision=6)))
