"""CPU: the humanoid model (mjpc/tasks/humanoid/humanoid.xml.patch + stand/task.xml) through the MJCF compiler and
the oracle - the physics features the A1 does not exercise: pyramidal cones, fixed-tendon limits, joint springs,
several hinges per body."""
import numpy as np

from conftest import get_model


def _oracle(m, precision=64):
    from mujoco_mpc_b200.blob import to_blob
    from oracle import pyoracle
    return pyoracle.Oracle(to_blob(m), m, precision)


def test_humanoid_model_dimensions():
    m = get_model("humanoid")
    # nq/nv/nu = 28/27/21, 1 + 20 bodies; stand task: 46 residuals in 5 terms, dt 0.015
    assert (m.nq, m.nv, m.nu, m.nbody) == (28, 27, 21, 21)
    assert m.task_num_residual == 46 and m.task_num_term == 5 and abs(m.opt_timestep - 0.015) < 1e-12
    assert m.opt_cone == 0 and m.ntendon == 2
    assert list(m.tendon_num) == [2, 2] and np.allclose(m.wrap_coef, [0.5, -0.5, 0.5, -0.5])
    assert np.allclose(m.tendon_range, [[-0.3, 2], [-0.3, 2]]) and (m.tendon_invweight0 > 0).all()
    # gears and class inheritance from the patch: hip_y 120, knee 100; hip_y stiffness 10 (joint_big), knee 1, elbow 0
    names = m.actuator_names
    assert np.ravel(m.actuator_gear[names.index("hip_y_right")])[0] == 120 and np.ravel(m.actuator_gear[names.index("knee_left")])[0] == 100
    jn = m.jnt_names
    assert m.jnt_stiffness[jn.index("hip_y_left")] == 10 and m.jnt_stiffness[jn.index("knee_right")] == 1
    assert m.jnt_stiffness[jn.index("elbow_left")] == 0 and m.jnt_stiffness[jn.index("abdomen_z")] == 20
    assert np.allclose(np.degrees(m.jnt_range[jn.index("knee_left")]), [-160, 2])


def test_humanoid_mass_matrix_matches_independent_numpy():
    from mujoco_mpc_b200.refmath import mass_matrix_and_jacobians
    m = get_model("humanoid")
    o = _oracle(m)
    rng = np.random.default_rng(1)
    for _ in range(3):
        q = m.qpos0.copy()
        q[7:] += 0.4 * rng.standard_normal(m.nq - 7)
        q[3:7] = rng.standard_normal(4); q[3:7] /= np.linalg.norm(q[3:7]); q[2] = 3.0
        r = o.forward_debug(q, 0.3 * rng.standard_normal(m.nv), np.zeros(m.nu), np.zeros(0))
        M, _ = mass_matrix_and_jacobians(m, q)
        assert np.abs(r["qM"] - M).max() < 1e-10 * np.abs(M).max()


def test_humanoid_pyramidal_contacts_and_tendon_limits():
    m = get_model("humanoid")
    o = _oracle(m)
    q, v, u = m.qpos0.copy(), np.zeros(m.nv), np.zeros(m.nu)
    r = o.forward_debug(q, v, u, np.zeros(0))
    # standing pose: both ends of the four foot capsules touch; condim max(1, 3) = 3 -> 4 pyramid edges per contact
    assert r["ncon"] == 8 and r["nefc"] == 4 * r["ncon"] and not r["warning"]
    assert (r["efc_force"] >= -1e-12).all() and r["efc_force"].sum() > 0          # one-sided rows push only
    # in the air no rows; bending the right hip back with a straight knee shortens the hamstring tendon below its
    # lower limit (length = 0.5 hip_y - 0.5 knee, range [-0.3, 2]) -> exactly one extra (tendon) row beside the
    # hip_y joint-limit rows that this pose may also activate
    q2 = q.copy(); q2[2] = 3.0
    base = o.forward_debug(q2, v, u, np.zeros(0))
    assert base["ncon"] == 0 and base["nefc"] == 0
    jn = m.jnt_names
    hip, knee = m.jnt_qposadr[jn.index("hip_y_right")], m.jnt_qposadr[jn.index("knee_right")]
    q3 = q2.copy(); q3[hip] = -1.0; q3[knee] = 0.0         # length -0.5 < -0.3, hip_y still inside [-150, 20] deg
    r3 = o.forward_debug(q3, v, u, np.zeros(0))
    assert r3["ncon"] == 0 and r3["nefc"] == 1
    dof_hip, dof_knee = m.jnt_dofadr[jn.index("hip_y_right")], m.jnt_dofadr[jn.index("knee_right")]
    fc = r3["qfrc_constraint"]
    # J = -side * coef with side = -1 (lower limit): force on hip_y positive, on the knee negative, same magnitude
    assert fc[dof_hip] > 0 and abs(fc[dof_hip] + fc[dof_knee]) < 1e-9 * abs(fc[dof_hip])
    assert np.abs(np.delete(fc, [dof_hip, dof_knee])).max() < 1e-12


def test_humanoid_stand_residual_and_rollout():
    m = get_model("humanoid")
    o = _oracle(m)
    q, v = m.qpos0.copy(), np.zeros(m.nv)
    u = np.linspace(-0.5, 0.5, m.nu)
    r = o.forward_debug(q, v, u, np.zeros(0))
    res = r["residual"][:46]
    assert np.allclose(res[4:25], 0) and np.allclose(res[25:46], u)              # joint velocities, controls
    assert abs(res[0] - (1.282 + 0.19 - 0.0 - 1.4 - (1.282 - 0.04 - 0.4 - 0.39 - 0.4 + 0.0))) < 0.5  # head-feet height - goal
    H = int(0.35 / 0.015 + 1)
    knots = np.zeros((4, 3, m.nu)); knots[1:] = 0.1 * np.random.default_rng(0).standard_normal((3, 3, m.nu))
    out = o.rollout_spline(np.concatenate([q, v]), 0.0, np.zeros(0), knots, np.array([0.0, 0.17, 0.34]), 2, H, nthreads=2)
    assert not out["failure"].any() and np.isfinite(out["returns"]).all() and (out["returns"] > 0).all()
    assert abs(out["states"][0, -1, 2] - 1.282) < 0.05        # zero control: the springs hold the pose for 0.35 s
    assert (out["states"][:, -1, 2] > 0.5).all()


def test_humanoid_track_residual_against_independent_numpy():
    """Tracking residual (tracking.cc:94-216) of the oracle vs a numpy restatement built on refmath (independent
    kinematics + Jacobians): joint velocities, controls, mean-centred marker errors, marker velocity errors."""
    from mujoco_mpc_b200 import task as T
    from mujoco_mpc_b200.refmath import body_jacobian, kinematics
    m = get_model("humanoid_track")
    assert (m.nbody, m.nmocap, m.nkey, m.task_num_residual, m.task_num_term) == (37, 16, 1889, 141, 21)
    assert sum(T.TRACK_MOTION_LENGTHS) == m.nkey and abs(m.opt_timestep - 0.005) < 1e-12
    o = _oracle(m)
    rng = np.random.default_rng(2)
    mocap = np.concatenate([np.concatenate([m.key_mpos[0].reshape(-1, 3), np.tile([1.0, 0, 0, 0], (16, 1))], 1).reshape(-1)])
    for mode, t in ((0, 0.0), (0, 1.2345), (3, 0.51), (9, 100.0)):      # last: clamps to the clip's final frame
        o.set_task(task_state=np.array([float(mode), 0.25]))
        q = m.key_qpos[sum(T.TRACK_MOTION_LENGTHS[:mode]) + 5].copy()
        q[7:] += 0.05 * rng.standard_normal(m.nq - 7); q[2] += 1.0      # airborne: no contact forces needed here
        v = 0.5 * rng.standard_normal(m.nv)
        u = rng.uniform(-1, 1, m.nu)
        r = o.forward_debug(q, v, u, mocap, time=t)["residual"][:141]
        start = sum(T.TRACK_MOTION_LENGTHS[:mode]); last = start + T.TRACK_MOTION_LENGTHS[mode] - 1
        idx = min(max((t - 0.25) * 30.0 + start, 0.0), float(last))
        k0 = int(np.floor(idx)); k1 = min(k0 + 1, last); w1 = idx - k0
        kin = kinematics(m, q)
        mp = (m.key_mpos[k0] * (1 - w1) + m.key_mpos[k1] * w1).reshape(16, 3)
        sp, sv = np.zeros((16, 3)), np.zeros((16, 3))
        for b, name in enumerate(T.TRACK_BODIES):
            s = m.site_names.index("tracking[%s]" % name); body = m.site_bodyid[s]
            sp[b] = kin["xpos"][body] + kin["xmat"][body] @ m.site_pos[s]
            sv[b] = body_jacobian(m, kin, body, sp[b])[:3] @ v
        expect = np.concatenate([v[6:], u, mp.mean(0) - sp.mean(0), ((mp - mp.mean(0)) - (sp - sp.mean(0))).reshape(-1),
                                 ((m.key_mpos[k1] - m.key_mpos[k0]).reshape(16, 3) * 30.0 - sv).reshape(-1)])
        assert np.abs(r - expect).max() < 1e-9, (mode, t, np.abs(r - expect).argmax())


def test_humanoid_track_transition():
    """Tracking::TransitionLocked (tracking.cc:218-267): clip switch restarts the clock and resets the plant to the
    clip's first keyframe; markers follow the linearly interpolated keyframes and clamp at the clip end."""
    from mujoco_mpc_b200 import task as T
    from mujoco_mpc_b200.transition import HumanoidTrackTransition
    m = get_model("humanoid_track")
    tr = HumanoidTrackTransition(m)
    q, v, mocap = tr.transition(0.0, m.qpos0.copy(), np.ones(m.nv))
    assert tr.current_mode == 0 and tr.reference_time == 0.0
    np.testing.assert_allclose(q, m.key_qpos[0]); np.testing.assert_allclose(v, m.key_qvel[0])   # tracking.cc:236-238
    np.testing.assert_allclose(mocap.reshape(16, 7)[:, :3].reshape(-1), m.key_mpos[0])
    q2, v2, mocap = tr.transition(0.05, q + 0.01, v)                 # 1.5 frames into the clip: state untouched
    np.testing.assert_allclose(q2, q + 0.01)
    np.testing.assert_allclose(mocap.reshape(16, 7)[:, :3].reshape(-1), 0.5 * (m.key_mpos[1] + m.key_mpos[2]), atol=1e-12)
    tr.mode = 3                                                        # GUI switches the clip at t = 1.0
    q3, v3, mocap = tr.transition(1.0, q2, v2)
    start = sum(T.TRACK_MOTION_LENGTHS[:3])
    assert tr.current_mode == 3 and tr.reference_time == 1.0
    np.testing.assert_allclose(q3, m.key_qpos[start]); np.testing.assert_allclose(tr.task_state(), [3.0, 1.0])
    _, _, mocap = tr.transition(1000.0, q3, v3)                        # far past the end: last frame of clip 3
    np.testing.assert_allclose(mocap.reshape(16, 7)[:, :3].reshape(-1), m.key_mpos[start + T.TRACK_MOTION_LENGTHS[3] - 1])
