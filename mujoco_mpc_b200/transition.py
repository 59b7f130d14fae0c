"""Host-side Task::Transition for the tracking task (runs once per real step on the plant side of the loop; never on
the device).  Mirrors mjpc/tasks/humanoid/tracking/tracking.cc:218-267: on a mode switch (or at time 0) the clip
clock restarts and the plant is reset to the clip's first keyframe; every step the 16 mocap markers are moved to the
interpolated keyframe.  The (mode, reference_time) pair it maintains is the task_state block the rollout kernel
consumes (mjpc_b200_set_task), i.e. the residual_fn_ snapshot of Agent::PlanIteration."""
from __future__ import annotations

import numpy as np

from . import task as T


class HumanoidTrackTransition:
    def __init__(self, model):
        self.m = model
        self.mode = 0                    # GUI-selected clip (Task::mode)
        self.current_mode = -1           # residual_.current_mode_
        self.reference_time = 0.0

    @staticmethod
    def motion_start(mode):
        return int(sum(T.TRACK_MOTION_LENGTHS[:mode]))

    def task_state(self):
        return np.array([float(max(self.current_mode, 0)), self.reference_time])

    def transition(self, time, qpos, qvel):
        """Returns (qpos, qvel, mocap[7*nmocap]) after TransitionLocked; qpos/qvel are replaced on a clip switch."""
        m = self.m
        start, length = self.motion_start(self.mode), T.TRACK_MOTION_LENGTHS[self.mode]
        if self.current_mode != self.mode or time == 0.0:
            self.current_mode = self.mode
            self.reference_time = time
            qpos, qvel = m.key_qpos[start].copy(), m.key_qvel[start].copy()
        last = start + length - 1
        idx = min(max((time - self.reference_time) * T.TRACK_FPS + start, 0.0), float(last))
        k0 = int(np.floor(idx)); k1 = min(k0 + 1, last); w1 = idx - k0
        mpos = (m.key_mpos[k0] * (1.0 - w1) + m.key_mpos[k1] * w1).reshape(m.nmocap, 3)
        mocap = np.concatenate([mpos, np.tile([1.0, 0, 0, 0], (m.nmocap, 1))], 1).reshape(-1)
        return qpos, qvel, mocap


# ---------------------------------------------------------------------------------------------- quadruped
K_MODE_QUADRUPED, K_MODE_BIPED, K_MODE_WALK, K_MODE_SCRAMBLE, K_MODE_FLIP = range(5)
K_GAIT_STAND, K_GAIT_WALK, K_GAIT_TROT, K_GAIT_CANTER, K_GAIT_GALLOP = range(5)
# duty ratio, cadence, amplitude, balance, upright, height (quadruped.h:88-97)
K_GAIT_PARAM = ((1, 1, 0, 0, 1, 1), (0.75, 1, 0.03, 0, 1, 1), (0.45, 2, 0.03, 0.2, 1, 1), (0.4, 4, 0.05, 0.03, 0.5, 0.2),
                (0.3, 3.5, 0.10, 0.03, 0.2, 0.1))
K_GAIT_AUTO = (0, 0.02, 0.02, 0.6, 2)          # quadruped.h:100-107
K_AUTO_GAIT_FILTER, K_AUTO_GAIT_MIN_TIME, K_MIN_ANGVEL = 0.2, 1.0, 0.01


class QuadrupedFlatTransition:
    """QuadrupedFlat::TransitionLocked (mjpc/tasks/quadruped/quadruped.cc:228-395) on the host.

    Owns what the reference's Task owns between planning iterations - `mode`, `parameters`, `weight`, the goal mocap
    position and the ResidualFn state block (the QS_* layout of task.py that the rollout kernel consumes) - and
    updates them once per real step from plant quantities passed as a dict:
      time, torso_subtreelinvel[3], torso_xmat[9], torso_xpos[3], torso_xquat[4], head_site_xpos[3], torso_subtreecom[3]
    and, for the Flip mode only, `ground(pos) -> height` (mjpc::Ground, a downward ray cast against group-0 geoms).
    The per-iteration snapshot for the engine is (weight, parameters, task_state())."""

    def __init__(self, model):
        from . import mjcf
        m = self.m = model
        self.mode = K_MODE_QUADRUPED
        self.parameters = np.asarray(m.task_parameters, float).copy()
        self.weight = np.asarray(m.task_weight, float).copy()
        self.state = np.asarray(m.task_state, float).copy()
        self.goal_pos = np.asarray(m.mocap_pos0, float)[int(m.task_ids[T.QI_GOAL_MOCAP])].copy()
        pn = [p[len("residual_"):] for p in m.task_parameter_names]
        self.p = {n: pn.index(n) for n in pn}
        terms = [n for n, t in zip(m.sensor_names, m.sensor_type) if t == mjcf.SENS_TYPES.get("user")]
        self.w = {n: terms.index(n) for n in terms}
        self.current_mode = K_MODE_QUADRUPED
        self.last_transition_time = -1.0
        self.com_vel = np.zeros(2)
        self.gait_switch_time = 0.0
        self.current_gait = float(K_GAIT_STAND)
        self.phase_velocity = 0.0
        self.save_weight, self.save_gait_switch = None, 0.0

    def task_state(self):
        s = self.state.copy()
        s[T.QS_MODE] = self.current_mode
        s[T.QS_GAIT] = self.current_gait
        s[T.QS_PHASE_VELOCITY] = self.phase_velocity
        return s

    def get_phase(self, time):
        s = self.state
        return s[T.QS_PHASE_START] + (time - s[T.QS_PHASE_START_TIME]) * self.phase_velocity

    def get_gait(self):
        return K_GAIT_TROT if self.current_mode == K_MODE_BIPED else int(self.current_gait)

    def walk(self, time):
        s = self.state
        heading, pos = s[T.QS_HEADING:T.QS_HEADING + 2], s[T.QS_POSITION:T.QS_POSITION + 2]
        speed, angvel = s[T.QS_SPEED], s[T.QS_ANGVEL]
        if abs(angvel) < K_MIN_ANGVEL:
            fwd = heading / max(np.linalg.norm(heading), 1e-15)
            return pos + heading + time * speed * fwd
        a = time * angvel
        return np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]) @ heading + pos

    def transition(self, d):
        s, P, W = self.state, self.parameters, self.weight
        time = float(d["time"])
        # ---- mjData reset
        if time < self.last_transition_time or self.last_transition_time == -1:
            if self.mode not in (K_MODE_QUADRUPED, K_MODE_BIPED):
                self.mode = K_MODE_QUADRUPED
            self.last_transition_time = s[T.QS_PHASE_START_TIME] = s[T.QS_PHASE_START] = time
        # ---- forbidden mode transitions: stateful modes only from Quadruped
        if self.mode != self.current_mode and self.current_mode != K_MODE_QUADRUPED:
            if self.mode in (K_MODE_WALK, K_MODE_FLIP):
                self.mode = K_MODE_QUADRUPED
        # ---- phase velocity change
        pv = 2 * np.pi * P[self.p["Cadence"]]
        if pv != self.phase_velocity:
            s[T.QS_PHASE_START] = self.get_phase(time)
            s[T.QS_PHASE_START_TIME] = time
            self.phase_velocity = pv
        # ---- automatic gait switching
        comvel = np.asarray(d["torso_subtreelinvel"], float)
        beta = np.exp(-(time - self.last_transition_time) / K_AUTO_GAIT_FILTER)
        self.com_vel = beta * self.com_vel + (1 - beta) * comvel[:2]
        auto = int(P[self.p["select_Gait switch"]])
        if self.mode == K_MODE_BIPED:
            P[self.p["select_Gait"]] = K_GAIT_TROT
        elif auto:
            speed = float(np.linalg.norm(self.com_vel))
            for g in range(5):
                if self.mode == K_MODE_SCRAMBLE and g == K_GAIT_STAND:
                    continue
                lower = speed > K_GAIT_AUTO[g]
                upper = g == K_GAIT_GALLOP or speed <= K_GAIT_AUTO[g + 1]
                wait = abs(self.gait_switch_time - time) > K_AUTO_GAIT_MIN_TIME
                if lower and upper and wait:
                    P[self.p["select_Gait"]] = g
                    self.gait_switch_time = time
        # ---- gait switch, manual or auto
        sel = P[self.p["select_Gait"]]
        if sel != self.current_gait:
            self.current_gait = sel
            gp = K_GAIT_PARAM[self.get_gait()]
            P[self.p["Duty ratio"]], P[self.p["Cadence"]], P[self.p["Amplitude"]] = gp[0], gp[1], gp[2]
            W[self.w["Balance"]], W[self.w["Upright"]], W[self.w["Height"]] = gp[3], gp[4], gp[5]
        # ---- Walk
        if self.mode == K_MODE_WALK:
            angvel, speed = P[self.p["Walk turn"]], P[self.p["Walk speed"]]
            xmat = np.asarray(d["torso_xmat"], float)
            fwd = np.array([xmat[0], xmat[3]]); fwd /= max(np.linalg.norm(fwd), 1e-15)
            left = np.array([-fwd[1], fwd[0]])
            if self.mode != self.current_mode or s[T.QS_ANGVEL] != angvel or s[T.QS_SPEED] != speed:
                s[T.QS_MODE_START_TIME] = time
                s[T.QS_SPEED], s[T.QS_ANGVEL] = speed, angvel
                axis = np.asarray(d["torso_xpos"], float)[:2].copy()
                if abs(angvel) > K_MIN_ANGVEL:
                    axis += speed / angvel * left
                s[T.QS_POSITION:T.QS_POSITION + 2] = axis
                s[T.QS_HEADING:T.QS_HEADING + 2] = self.goal_pos[:2] - axis
            self.goal_pos[:2] = self.walk(time - s[T.QS_MODE_START_TIME])
        # ---- Flip
        if self.mode == K_MODE_FLIP:
            if self.mode != self.current_mode:
                s[T.QS_MODE_START_TIME] = time
                s[T.QS_ORIENTATION:T.QS_ORIENTATION + 4] = np.asarray(d["torso_xquat"], float)
                s[T.QS_GROUND] = float(d["ground"](np.asarray(d["torso_subtreecom"], float)))
                self.save_weight = W.copy()
                self.save_gait_switch = P[self.p["select_Gait switch"]]
                for n, v in (("Upright", 0.2), ("Height", 5), ("Position", 0), ("Gait", 0), ("Balance", 0), ("Effort", 0.005),
                             ("Posture", 0.1)):
                    W[self.w[n]] = v
                P[self.p["select_Gait switch"]] = 0
            flip_time = time - s[T.QS_MODE_START_TIME]
            if flip_time >= s[T.QS_JUMP_TIME] + s[T.QS_FLIGHT_TIME] + s[T.QS_LAND_TIME]:
                self.mode = K_MODE_QUADRUPED
                W[:] = self.save_weight
                P[self.p["select_Gait switch"]] = self.save_gait_switch
                self.goal_pos[:2] = np.asarray(d["head_site_xpos"], float)[:2]
        self.current_mode = self.mode
        self.last_transition_time = time


class ShadowReorientTransition:
    """ShadowReorient::TransitionLocked (mjpc/tasks/shadow_reorient/hand.cc:90-119): when the cube lies on the floor
    (a cube-floor contact exists) and is at rest (|cube linear velocity| < 1e-3) it is put back into the hand: its 7 qpos
    take the model's qpos0 values, its 6 qvel are zeroed.  The contact list and the framelinvel sensor belong to the
    plant; the caller passes what the reference reads from mjData."""

    def __init__(self, model):
        self.m = model
        cube = model.body_names.index("cube")
        j = int(model.body_jntadr[cube])
        self.qadr, self.dadr = int(model.jnt_qposadr[j]), int(model.jnt_dofadr[j])
        self.cube_geom, self.floor_geom = model.geom_names.index("cube"), model.geom_names.index("floor")

    def on_floor(self, contact_geoms):
        """contact_geoms: iterable of (geom1, geom2) of the plant's active contacts (hand.cc:96-104)."""
        return any({int(a), int(b)} == {self.cube_geom, self.floor_geom} for a, b in contact_geoms)

    def transition(self, qpos, qvel, on_floor, cube_linvel):
        qpos, qvel = np.array(qpos, float), np.array(qvel, float)
        reset = bool(on_floor) and float(np.linalg.norm(cube_linvel)) < 1e-3
        if reset:
            qpos[self.qadr:self.qadr + 7] = self.m.qpos0[self.qadr:self.qadr + 7]
            qvel[self.dadr:self.dadr + 6] = 0.0
        return qpos, qvel, reset
