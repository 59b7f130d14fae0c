"""CPU: QuadrupedFlat::TransitionLocked restated on the host (mujoco_mpc_b200/transition.py) - the state machine that
produces the task-state block the rollout kernel consumes (quadruped.cc:228-395)."""
import numpy as np

from conftest import get_model


def _view(time, vel=(0, 0, 0), yaw=0.0, pos=(0, 0, 0.26)):
    c, s = np.cos(yaw), np.sin(yaw)
    return dict(time=time, torso_subtreelinvel=np.array(vel, float), torso_xmat=np.array([c, -s, 0, s, c, 0, 0, 0, 1.0]),
                torso_xpos=np.array(pos, float), torso_xquat=np.array([np.cos(yaw / 2), 0, 0, np.sin(yaw / 2)]),
                head_site_xpos=np.array([pos[0] + 0.3, pos[1], pos[2]]), torso_subtreecom=np.array(pos, float),
                ground=lambda p: 0.0)


def test_first_transition_and_gait_tables():
    from mujoco_mpc_b200 import task as T
    from mujoco_mpc_b200.transition import K_GAIT_PARAM, QuadrupedFlatTransition
    m = get_model("quadruped")
    tr = QuadrupedFlatTransition(m)
    tr.transition(_view(0.0))
    st = tr.task_state()
    # cadence 2 Hz from the task XML -> phase velocity 4 pi; Stand is the current gait so nothing is overwritten yet
    assert abs(st[T.QS_PHASE_VELOCITY] - 4 * np.pi) < 1e-12 and st[T.QS_MODE] == 0 and st[T.QS_GAIT] == 0
    # manual switch to Trot: parameters and weights take the gait table's values (quadruped.h:88-97)
    tr.parameters[tr.p["select_Gait switch"]] = 0
    tr.parameters[tr.p["select_Gait"]] = 2
    tr.transition(_view(0.01))
    gp = K_GAIT_PARAM[2]
    assert tr.parameters[tr.p["Duty ratio"]] == gp[0] and tr.parameters[tr.p["Cadence"]] == gp[1]
    assert tr.parameters[tr.p["Amplitude"]] == gp[2] and tr.weight[tr.w["Balance"]] == gp[3]
    # the cadence change is picked up on the NEXT transition and keeps the phase continuous
    ph_before = tr.get_phase(0.02)
    tr.transition(_view(0.02))
    assert abs(tr.phase_velocity - 2 * np.pi * gp[1]) < 1e-12 and abs(tr.get_phase(0.02) - ph_before) < 1e-12


def test_auto_gait_follows_filtered_com_speed():
    from mujoco_mpc_b200.transition import QuadrupedFlatTransition
    m = get_model("quadruped")
    tr = QuadrupedFlatTransition(m)
    assert tr.parameters[tr.p["select_Gait switch"]] == 1            # auto switching is the task default
    t = 0.0
    for _ in range(400):                                              # 4 s at 0.3 m/s: 0.02 < v <= 0.6 -> trot
        tr.transition(_view(t, vel=(0.3, 0, 0))); t += 0.01
    assert int(tr.current_gait) == 2
    for _ in range(400):                                              # 1.0 m/s: 0.6 < v <= 2 -> canter
        tr.transition(_view(t, vel=(1.0, 0, 0))); t += 0.01
    assert int(tr.current_gait) == 3
    for _ in range(400):                                              # standing still again -> stand
        tr.transition(_view(t, vel=(0, 0, 0))); t += 0.01
    assert int(tr.current_gait) == 0


def test_walk_mode_moves_goal_and_forbidden_transitions():
    from mujoco_mpc_b200 import task as T
    from mujoco_mpc_b200.transition import K_MODE_BIPED, K_MODE_FLIP, K_MODE_QUADRUPED, K_MODE_WALK, QuadrupedFlatTransition
    m = get_model("quadruped")
    tr = QuadrupedFlatTransition(m)
    tr.transition(_view(0.0))
    tr.parameters[tr.p["Walk speed"]] = 0.5
    tr.mode = K_MODE_WALK
    g0 = tr.goal_pos.copy()
    tr.transition(_view(1.0))
    st = tr.task_state()
    assert st[T.QS_MODE] == K_MODE_WALK and st[T.QS_MODE_START_TIME] == 1.0 and st[T.QS_SPEED] == 0.5
    np.testing.assert_allclose(tr.goal_pos, g0, atol=1e-12)          # t - mode_start = 0: goal where it was
    tr.transition(_view(3.0))
    heading = g0[:2] - np.zeros(2)
    np.testing.assert_allclose(tr.goal_pos[:2], g0[:2] + 2.0 * 0.5 * heading / np.linalg.norm(heading), atol=1e-12)
    # turning: the goal moves on a circle around the saved axis
    tr.parameters[tr.p["Walk turn"]] = 1.0
    tr.transition(_view(3.0))
    r = np.linalg.norm(tr.goal_pos[:2] - tr.state[T.QS_POSITION:T.QS_POSITION + 2])
    tr.transition(_view(4.0))
    assert abs(np.linalg.norm(tr.goal_pos[:2] - tr.state[T.QS_POSITION:T.QS_POSITION + 2]) - r) < 1e-12
    # Biped -> Flip is forbidden (stateful modes only from Quadruped): reverts to Quadruped
    tr.mode = K_MODE_BIPED; tr.transition(_view(4.01))
    tr.mode = K_MODE_FLIP; tr.transition(_view(4.02))
    assert tr.current_mode == K_MODE_QUADRUPED


def test_flip_saves_and_restores_weights():
    from mujoco_mpc_b200 import task as T
    from mujoco_mpc_b200.transition import K_MODE_FLIP, K_MODE_QUADRUPED, QuadrupedFlatTransition
    m = get_model("quadruped")
    tr = QuadrupedFlatTransition(m)
    tr.transition(_view(0.0))
    w0 = tr.weight.copy()
    tr.mode = K_MODE_FLIP
    tr.transition(_view(0.5, yaw=0.3))
    st = tr.task_state()
    assert st[T.QS_MODE] == K_MODE_FLIP and st[T.QS_MODE_START_TIME] == 0.5
    np.testing.assert_allclose(st[T.QS_ORIENTATION:T.QS_ORIENTATION + 4], [np.cos(0.15), 0, 0, np.sin(0.15)])
    assert tr.weight[tr.w["Height"]] == 5 and tr.weight[tr.w["Gait"]] == 0 and tr.parameters[tr.p["select_Gait switch"]] == 0
    total = st[T.QS_JUMP_TIME] + st[T.QS_FLIGHT_TIME] + st[T.QS_LAND_TIME]
    tr.transition(_view(0.5 + total - 1e-3))
    assert tr.current_mode == K_MODE_FLIP
    tr.transition(_view(0.5 + total + 1e-3, pos=(1.0, 2.0, 0.26)))
    assert tr.current_mode == K_MODE_QUADRUPED
    np.testing.assert_allclose(tr.weight, w0)
    np.testing.assert_allclose(tr.goal_pos[:2], [1.3, 2.0]) and tr.parameters[tr.p["select_Gait switch"]] == 1


def test_cpp_transitions_match_python(tmp_path):
    """csrc/host/task_transition.cc (C++) vs transition.py on the same random plant-view sequence, including mode
    switches (Walk, Flip, forbidden ones), auto gait switching and clip changes of the tracking task."""
    import ctypes as C

    from mujoco_mpc_b200 import task as T
    from mujoco_mpc_b200.engine import load_library
    from mujoco_mpc_b200.transition import HumanoidTrackTransition, QuadrupedFlatTransition
    lib = load_library()
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    m = get_model("quadruped")
    py = QuadrupedFlatTransition(m)
    ids = np.array([py.p[n] for n in ("select_Gait", "select_Gait switch", "Cadence", "Amplitude", "Duty ratio", "Walk speed",
                                      "Walk turn")] +
                   [py.w[n] for n in ("Upright", "Height", "Position", "Gait", "Balance", "Effort", "Posture")], np.int32)
    P0, W0, S0, G0 = py.parameters.copy(), py.weight.copy(), py.state.copy(), py.goal_pos.copy()
    h = C.c_void_p(lib.mjpc_b200_quadruped_transition_create(ids.ctypes.data_as(C.POINTER(C.c_int)), dp(P0), len(P0), dp(W0),
                                                             len(W0), dp(S0), len(S0), dp(G0)))
    assert h.value
    rng = np.random.default_rng(0)
    t, mode = 0.0, 0
    P, W, S, G = np.zeros_like(P0), np.zeros_like(W0), np.zeros(T.QS_SIZE), np.zeros(3)
    requests = {100: 2, 400: 0, 700: 1, 705: 4, 900: 0, 905: 4, 1300: 3}   # Walk, Quadruped, Biped, Flip (forbidden), Quadruped, Flip, Scramble
    seen = set()
    for k in range(1500):
        t += 0.01
        if k in requests:
            mode = requests[k]
        if k == 100:                                              # a GUI edit of two parameters, mirrored on both sides
            py.parameters[py.p["Walk speed"]] = 0.4; py.parameters[py.p["Walk turn"]] = 0.5
            lib.mjpc_b200_quadruped_transition_set(h, dp(py.parameters), None)
        view = _view(t, vel=(1.5 * abs(np.sin(0.004 * k)), 0.1 * rng.standard_normal(), 0), yaw=0.3 * np.sin(0.01 * k),
                     pos=(0.01 * k, 0.2 * np.sin(0.01 * k), 0.26))
        flat = np.concatenate([[view["time"]], view["torso_subtreelinvel"], view["torso_xmat"], view["torso_xpos"],
                               view["torso_xquat"], view["head_site_xpos"], [0.0]])
        py.mode = mode
        py.transition(view)
        cm = C.c_int(mode)
        lib.mjpc_b200_quadruped_transition_step(h, C.byref(cm), dp(flat), dp(P), dp(W), dp(S), dp(G))
        np.testing.assert_allclose(P, py.parameters, atol=1e-12, err_msg="parameters at step %d" % k)
        np.testing.assert_allclose(W, py.weight, atol=1e-12, err_msg="weights at step %d" % k)
        np.testing.assert_allclose(S, py.task_state(), atol=1e-9, err_msg="task state at step %d" % k)
        np.testing.assert_allclose(G, py.goal_pos, atol=1e-9, err_msg="goal at step %d" % k)
        assert cm.value == py.mode
        mode = py.mode                                            # Transition may rewrite Task::mode
        seen.add((py.current_mode, int(py.current_gait)))
    assert {mm for mm, _ in seen} == {0, 1, 2, 3, 4} and len({g for _, g in seen}) >= 3     # all modes, several gaits
    lib.mjpc_b200_quadruped_transition_destroy(h)
    # ---- tracking
    mt = get_model("humanoid_track")
    tp = HumanoidTrackTransition(mt)
    kq, kv, km = (np.ascontiguousarray(x, np.float64) for x in (mt.key_qpos, mt.key_qvel, mt.key_mpos))
    th = C.c_void_p(lib.mjpc_b200_track_transition_create(mt.nq, mt.nv, mt.nmocap, mt.nkey, dp(kq), dp(kv), dp(km)))
    assert th.value
    q, v = mt.qpos0.copy(), np.ones(mt.nv)
    qc, vc = q.copy(), v.copy()
    mp, ts = np.zeros(3 * mt.nmocap), np.zeros(2)
    for k, (mode, time) in enumerate([(0, 0.0), (0, 0.033), (0, 0.5), (3, 0.7), (3, 0.71), (3, 9.0), (9, 9.5), (9, 30.0)]):
        tp.mode = mode
        q, v, mocap = tp.transition(time, q, v)
        lib.mjpc_b200_track_transition_step(th, mode, C.c_double(time), dp(qc), dp(vc), dp(mp), dp(ts))
        np.testing.assert_allclose(qc, q, atol=1e-12); np.testing.assert_allclose(vc, v, atol=1e-12)
        np.testing.assert_allclose(mp, mocap.reshape(-1, 7)[:, :3].reshape(-1), atol=1e-12)
        np.testing.assert_allclose(ts, tp.task_state(), atol=1e-12)
        q = q + 0.01; v = v * 0.9; qc = qc + 0.01; vc = vc * 0.9
    lib.mjpc_b200_track_transition_destroy(th)
