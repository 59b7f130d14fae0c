"""CPU: BASELINE config 5 task (Shadow Hand cube reorientation) on the documented primitive-geom stand-in hand.
What is pinned here is what the reference tree itself defines: dimensions (35 / 33 / 20, 81 residuals in 6 terms), the
residual of hand.cc:37-84 with its literal qpos + 7 / qvel + 6 offsets (independent numpy restatement on refmath
kinematics), the transition of hand.cc:90-119 (Python vs C++), and fixed-tendon actuator transmission (closed form)."""
import ctypes as C

import numpy as np

from conftest import get_model, mocap_of


def _hold_ctrl(m, q):
    """position targets that hold configuration q (joint actuators: the joint angle, tendon actuators: the tendon length)"""
    u = np.zeros(m.nu)
    for i in range(m.nu):
        if m.actuator_trntype[i] == 0:
            u[i] = q[m.jnt_qposadr[m.actuator_trnid[i]]]
        else:
            t = m.actuator_trnid[i]
            u[i] = sum(m.wrap_coef[w] * q[m.wrap_qposadr[w]] for w in range(m.tendon_adr[t], m.tendon_adr[t] + m.tendon_num[t]))
    return u


def test_shadow_dimensions_match_the_reference_task():
    m = get_model("shadow_reorient")
    assert (m.nq, m.nv, m.nu) == (35, 33, 20)
    assert m.task_num_residual == 81 and list(m.task_dim_norm_residual) == [3, 3, 3, 20, 26, 26]
    assert m.jnt_type[0] == 1 and m.jnt_type[1] == 0               # goal ball joint first, cube free joint second
    assert len(m.key_qpos[0]) == 35 and abs(m.key_qpos[0][4] - 0.33326) < 1e-9
    assert int(m.actuator_trntype.sum()) == 4 and m.ntendon == 4   # the coupled distal joints (J0 = J2 + J1)
    assert int(m.numeric["agent_planner"][0]) == 5 and int(m.numeric["sampling_spline_points"][0]) == 5
    from mujoco_mpc_b200.engine import load_library
    assert load_library().mjpc_b200_agent_steps(C.c_double(0.4701), C.c_double(0.01)) == 48


def test_shadow_residual_matches_numpy_restatement(oracle_lib):
    from mujoco_mpc_b200.blob import to_blob
    from mujoco_mpc_b200.refmath import body_jacobian, kinematics
    m = get_model("shadow_reorient")
    o = oracle_lib.Oracle(to_blob(m), m, 64)
    rng = np.random.default_rng(1)
    cube, goal = m.body_names.index("cube"), m.body_names.index("goal")
    site = m.site_names.index("grasp_site")
    for trial in range(4):
        q = m.key_qpos[0].copy(); q[11:] += 0.1 * rng.standard_normal(24)
        gq = rng.standard_normal(4); q[0:4] = gq / np.linalg.norm(gq)
        cq = q[7:11] + 0.2 * rng.standard_normal(4); q[7:11] = cq / np.linalg.norm(cq)
        v = 0.3 * rng.standard_normal(m.nv)
        u = _hold_ctrl(m, q) + 0.05 * rng.standard_normal(m.nu)
        r = o.forward_debug(q, v, u, mocap_of(m))
        res = r["residual"][:81]
        kin = kinematics(m, q)
        sb = m.site_bodyid[site]
        palm = kin["xpos"][sb] + kin["xmat"][sb] @ m.site_pos[site]
        np.testing.assert_allclose(res[0:3], kin["xpos"][cube] - palm, atol=1e-12)

        def qmul(a, b):
            return np.array([a[0] * b[0] - a[1:] @ b[1:], *(a[0] * b[1:] + b[0] * a[1:] + np.cross(a[1:], b[1:]))])
        qd = qmul(np.array([q[7], -q[8], -q[9], -q[10]]), q[0:4])              # cube^-1 * goal (mju_subQuat)
        s = np.linalg.norm(qd[1:]); ang = 2 * np.arctan2(s, qd[0])
        ang = ang - 2 * np.pi if ang > np.pi else ang
        np.testing.assert_allclose(res[3:6], qd[1:] / s * ang, atol=1e-10)
        np.testing.assert_allclose(res[6:9], body_jacobian(m, kin, cube, kin["xpos"][cube])[:3] @ v, atol=1e-10)
        np.testing.assert_allclose(res[29:55], q[7:33] - m.key_qpos[0][7:33], atol=1e-12)   # qpos + 7: starts at the cube quaternion
        np.testing.assert_allclose(res[55:81], v[6:32], atol=1e-12)                          # qvel + 6: starts at the cube's angular velocity
        # Residual (3) = actuator_force of position actuators, joint and tendon transmission (kp (ctrl - length), clamped)
        for i in range(m.nu):
            if m.actuator_trntype[i] == 0:
                length = q[m.jnt_qposadr[m.actuator_trnid[i]]]
            else:
                t = m.actuator_trnid[i]
                length = sum(m.wrap_coef[w] * q[m.wrap_qposadr[w]] for w in range(m.tendon_adr[t], m.tendon_adr[t] + m.tendon_num[t]))
            kp = m.actuator_gainprm[i][0]
            vel = 0.0                                                            # kv = 0
            uc = np.clip(u[i], m.actuator_ctrlrange[i][0], m.actuator_ctrlrange[i][1])   # ctrl is clamped first (ctrllimited)
            f = np.clip(kp * (uc - length) + vel, m.actuator_forcerange[i][0], m.actuator_forcerange[i][1])
            assert abs(res[9 + i] - f) < 1e-10, (i, res[9 + i], f)


def test_tendon_actuator_moment_closed_form(oracle_lib):
    """A force on a fixed tendon acts on every wrapped dof with its coefficient: both distal joints of a finger get the
    same generalized force from the J0 actuator, nothing else changes."""
    from mujoco_mpc_b200.blob import to_blob
    m = get_model("shadow_reorient")
    o = oracle_lib.Oracle(to_blob(m), m, 64)
    q = m.key_qpos[0].copy()
    u0 = _hold_ctrl(m, q)
    i = int(np.nonzero(m.actuator_trntype)[0][0]); t = m.actuator_trnid[i]
    dofs = [int(m.wrap_dof[w]) for w in range(m.tendon_adr[t], m.tendon_adr[t] + m.tendon_num[t])]
    u0[i] = m.actuator_ctrlrange[i][0] + 0.1; u1 = u0.copy(); u1[i] += 0.3
    a = o.forward_debug(q, np.zeros(m.nv), u0, mocap_of(m)); b = o.forward_debug(q, np.zeros(m.nv), u1, mocap_of(m))
    d = b["qfrc_smooth"] - a["qfrc_smooth"]
    np.testing.assert_allclose(d[dofs], 0.3 * m.actuator_gainprm[i][0], atol=1e-12)
    mask = np.ones(m.nv, bool); mask[dofs] = False
    assert np.abs(d[mask]).max() < 1e-12


def test_shadow_transition_python_matches_cpp():
    from mujoco_mpc_b200.engine import load_library
    from mujoco_mpc_b200.transition import ShadowReorientTransition
    m = get_model("shadow_reorient")
    tr = ShadowReorientTransition(m)
    lib = load_library()
    q0c = np.ascontiguousarray(m.qpos0[tr.qadr:tr.qadr + 7], float)
    h = C.c_void_p(lib.mjpc_b200_shadow_transition_create(tr.qadr, tr.dadr, q0c.ctypes.data_as(C.POINTER(C.c_double))))
    rng = np.random.default_rng(0)
    dp = C.POINTER(C.c_double)
    for on_floor, speed in ((True, 1e-4), (True, 0.5), (False, 1e-5)):
        q = rng.standard_normal(m.nq); v = rng.standard_normal(m.nv); lv = np.array([speed, 0, 0.0])
        q1, v1, reset = tr.transition(q, v, on_floor, lv)
        q2, v2 = q.copy(), v.copy()
        rc = lib.mjpc_b200_shadow_transition_step(h, q2.ctypes.data_as(dp), v2.ctypes.data_as(dp), int(on_floor), lv.ctypes.data_as(dp))
        assert bool(rc) == reset == (on_floor and speed < 1e-3)
        np.testing.assert_array_equal(q1, q2); np.testing.assert_array_equal(v1, v2)
        if reset:
            np.testing.assert_array_equal(q1[tr.qadr:tr.qadr + 7], m.qpos0[tr.qadr:tr.qadr + 7]); assert not v1[tr.dadr:tr.dadr + 6].any()
    assert tr.on_floor([(tr.floor_geom, tr.cube_geom)]) and not tr.on_floor([(tr.cube_geom, 5)])
    lib.mjpc_b200_shadow_transition_destroy(h)
