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
