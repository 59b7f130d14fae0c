"""GPU: BASELINE config 5 (Shadow Hand cube reorientation, PS 512 x 48) on the documented stand-in hand, generic
(DynSpec) kernels.  Many small contacts (finger capsules and palm box against the cube box), fixed tendons, a ball joint
and a free joint in one tree - the regime config 5 names.  Same two checks as configs 2 / 3: teacher-forced single steps
from the fp64 oracle's own trajectory states, and returns of a 128 x 48 share of the batch (one quarter; the candidates
are independent, so the share is the same arithmetic)."""
import numpy as np
import pytest

from conftest import get_model, mocap_of

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def case(oracle_lib):
    from mujoco_mpc_b200.blob import to_blob
    from mujoco_mpc_b200.engine import Engine
    from mujoco_mpc_b200.planner import candidate_knots
    m = get_model("shadow_reorient")
    N, H, P = 128, 48, 5
    o = oracle_lib.Oracle(to_blob(m), m, 64)
    state = np.concatenate([m.key_qpos[0], np.zeros(m.nv)])
    # nominal = hold the keyframe (position actuators: targets = current lengths); candidates = nominal + N(0, 0.1) clamped
    from test_shadow_cpu import _hold_ctrl
    nominal = np.tile(_hold_ctrl(m, m.key_qpos[0]), (P, 1))
    ctrlrange = np.asarray(m.actuator_ctrlrange, float)
    knots = candidate_knots(nominal, 0.1, ctrlrange, 0, N, seed=7).astype(np.float32)
    kt = np.linspace(0.0, (H - 1) * m.opt_timestep, P)
    r64 = o.rollout_spline(state, 0.0, mocap_of(m), knots, kt, 2, H, nthreads=8, full=True)
    e = Engine(m, N, H)
    yield dict(m=m, N=N, H=H, o=o, state=state, knots=knots, kt=kt, r64=r64, e=e)
    e.close()


def test_shadow_teacher_forced_steps(case):
    m, r64, e, o = case["m"], case["r64"], case["e"], case["o"]
    N, H = case["N"], case["H"]
    S = r64["states"][:, : H - 1].reshape(-1, m.nq + m.nv)
    U = r64["actions"][:, : H - 1].reshape(-1, m.nu)
    T = np.tile(np.arange(H - 1) * m.opt_timestep, N)
    ref = o.step_batch(S[:, : m.nq], S[:, m.nq:], U, mocap_of(m), T, nthreads=8)
    dev = e.step_batch(S[:, : m.nq], S[:, m.nq:], U, mocap_of(m), T)
    assert (ref["ncon"] > 0).mean() > 0.7                      # the cube really rests in the hand
    same = (dev["ncon"] == ref["ncon"]) & (dev["nefc"] == ref["nefc"])
    assert same.mean() > 0.995, same.mean()
    ok = same & (ref["warning"] == 0) & (dev["warning"] == 0)
    assert ok.mean() > 0.99
    err = np.abs(dev["next_qvel"] - ref["next_qvel"]).max(1)[ok]
    med, p99, mx = float(np.median(err)), float(np.percentile(err, 99)), float(err.max())
    print("shadow teacher-forced: %d steps, ncon up to %d / nefc up to %d, qvel err median %.2e p99 %.2e max %.2e"
          % (ok.sum(), ref["ncon"].max(), ref["nefc"].max(), med, p99, mx))
    assert med <= 2e-5 and p99 <= 1e-3 and mx <= 2e-2
    rerr = np.abs(dev["residual"][ok][:, : m.task_num_residual] - ref["residual"][ok][:, : m.task_num_residual]).max()
    cerr = (np.abs(dev["cost"][ok] - ref["cost"][ok]) / np.maximum(np.abs(ref["cost"][ok]), 1e-3)).max()
    assert rerr <= 5e-4 and cerr <= 5e-5, (rerr, cerr)


def test_shadow_returns(case):
    m, r64, e = case["m"], case["r64"], case["e"]
    ret, fail, order = e.rollout_spline(case["state"], 0.0, mocap_of(m), case["knots"], case["kt"], 2, case["H"])
    assert not fail.any() and not r64["failure"].any()
    rel = np.abs(ret - r64["returns"]) / np.abs(r64["returns"])
    print("shadow 128 x 48 returns: rel err median %.2e p90 %.2e max %.2e, > 1e-4: %d" %
          (np.median(rel), np.percentile(rel, 90), rel.max(), (rel > 1e-4).sum()))
    assert np.median(rel) < 2e-5
    # contact-rich: candidates whose fp64 return is itself ill-conditioned may miss 1e-4 (same classifier as config 2 / 3)
    from test_gpu_teacher_forced import _stable_mask
    stable = _stable_mask(case["o"], m, case["state"], mocap_of(m), case["knots"], case["kt"], case["H"], r64["returns"])
    print("  well-conditioned: %d / %d, max rel on them %.2e" % (stable.sum(), len(rel), rel[stable].max()))
    # (the velocity-perturbation classifier is harsh on a cube resting on fingertips: fp64 itself flags 3 of 4 candidates;
    #  the oracle's own fp32 instantiation misses 1e-4 on 2 of 128 with max 1.05e-4 - the device gets the same allowance)
    assert stable.mean() >= 0.2
    assert (rel[stable] > 1e-4).sum() == 0
    assert (rel > 1e-4).sum() <= 4 and rel.max() < 1e-3
    best64 = int(np.argmin(r64["returns"]))
    assert abs(ret[int(order[0])] - r64["returns"][best64]) / abs(r64["returns"][best64]) < 1e-4
