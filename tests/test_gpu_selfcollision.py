"""GPU: contacts between two MOVING bodies of the robot (capsule-capsule, sphere-capsule, sphere-sphere; verdict item 7).
The models keep every pair MuJoCo's own filters keep (models.load(..., self_collision=True) is the default); here random
joint configurations high above the floor produce self-contacts only, and single device steps are compared with the fp64
oracle from the same state (teacher-forced: no trajectory divergence to hide behind)."""
import numpy as np
import pytest

from conftest import get_model, mocap_of

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["humanoid", "quadruped"])
def test_self_collision_steps_match_oracle(name, oracle_lib):
    from mujoco_mpc_b200.blob import to_blob
    from mujoco_mpc_b200.engine import Engine
    m = get_model(name)
    o = oracle_lib.Oracle(to_blob(m), m, 64)
    rng = np.random.default_rng(3)
    B = 1500
    q0 = np.asarray(m.key_qpos[0] if m.nkey else m.qpos0, float)
    q = np.tile(q0, (B, 1)); q[:, 2] = 3.0                              # far above the floor and the props
    lo = np.array([m.jnt_range[j][0] for j in range(1, m.njnt)]); hi = np.array([m.jnt_range[j][1] for j in range(1, m.njnt)])
    q[:, 7:] = lo + (hi - lo) * rng.uniform(0.0, 1.0, (B, m.nq - 7))     # anywhere inside the joint ranges
    v = np.zeros((B, m.nv)); v[:, 6:] = 0.5 * rng.standard_normal((B, m.nv - 6))
    u = rng.uniform(-0.3, 0.3, (B, m.nu))
    t = np.zeros(B)
    ref = o.step_batch(q[:, : m.nq], v, u, mocap_of(m), t, nthreads=8)
    touching = (ref["ncon"] > 0) & (ref["warning"] == 0)
    assert touching.sum() >= 50, touching.sum()                          # the sweep really produces self-contacts
    e = Engine(m, 4, 8)
    try:
        dev = e.step_batch(q[:, : m.nq], v, u, mocap_of(m), t)
    finally:
        e.close()
    same = (dev["ncon"] == ref["ncon"]) & (dev["nefc"] == ref["nefc"])
    # a pair exactly at its margin may be seen by one arithmetic only: allow a handful
    assert (~same[touching]).sum() <= 0.01 * touching.sum() + 2, (~same[touching]).sum()
    ok = touching & same & (dev["warning"] == 0)
    err = np.abs(dev["next_qvel"] - ref["next_qvel"]).max(1)[ok]
    scale = np.abs(ref["qacc"]).max(1)[ok] * m.opt_timestep + 1.0
    rel = err / scale
    print("%s: %d self-contact states (up to %d contacts), qvel err / scale median %.2e p99 %.2e max %.2e"
          % (name, ok.sum(), ref["ncon"][ok].max(), np.median(rel), np.percentile(rel, 99), rel.max()))
    assert np.median(rel) < 5e-5 and np.percentile(rel, 99) < 5e-3 and rel.max() < 5e-2
    # overflow semantics: a state with more contacts than the buffer holds raises a warning on BOTH sides (-> failure)
    both_warn = (ref["warning"] != 0)
    assert ((dev["warning"] != 0) == both_warn).mean() > 0.99
