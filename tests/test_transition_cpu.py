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
