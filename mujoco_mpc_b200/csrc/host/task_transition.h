// task_transition.h - C++ host restatement of the config tasks' Task::Transition (runs once per real step on the plant
// side of the loop, never on the device):
//   QuadrupedFlat::TransitionLocked  mjpc/tasks/quadruped/quadruped.cc:228-395 (+ constants quadruped.h:72-128)
//   Tracking::TransitionLocked       mjpc/tasks/humanoid/tracking/tracking.cc:218-267
// They own what the reference's Task owns between planning iterations (mode, parameters, weights, goal mocap, the
// ResidualFn state) and produce the task-state block the rollout kernels consume (layouts: mujoco_mpc_b200/task.py,
// csrc/dev_task.cuh QS_* / tracking [mode, reference_time]); mjpc_b200_set_task is the per-iteration snapshot.
#pragma once
#include <vector>

namespace mjpc_b200_host {

// indices into the quadruped task-state block (same numbering as task.py / dev_task.cuh)
enum QuadrupedState { QS_MODE = 0, QS_MODE_START_TIME = 1, QS_POSITION = 2, QS_HEADING = 5, QS_SPEED = 7, QS_ANGVEL = 8,
                      QS_GROUND = 9, QS_ORIENTATION = 10, QS_GAIT = 14, QS_PHASE_START = 15, QS_PHASE_START_TIME = 16,
                      QS_PHASE_VELOCITY = 17, QS_FLIGHT_TIME = 19, QS_JUMP_TIME = 23, QS_LAND_TIME = 25, QS_SIZE = 31 };
enum QuadrupedMode { kModeQuadruped = 0, kModeBiped, kModeWalk, kModeScramble, kModeFlip };
enum QuadrupedGait { kGaitStand = 0, kGaitWalk, kGaitTrot, kGaitCanter, kGaitGallop, kNumGait };

struct QuadrupedIds {   // positions inside `parameters` / `weight` (resolved by name once, quadruped.cc:540-606)
  int p_gait, p_gait_switch, p_cadence, p_amplitude, p_duty, p_walk_speed, p_walk_turn;
  int w_upright, w_height, w_position, w_gait, w_balance, w_effort, w_posture;
};

struct QuadrupedPlantView {   // what TransitionLocked reads from mjData
  double time;
  double torso_subtreelinvel[3], torso_xmat[9], torso_xpos[3], torso_xquat[4], head_site_xpos[3];
  double ground_under_com;    // mjpc::Ground(torso_subtreecom): only read when the Flip mode starts
};

class QuadrupedFlatTransition {
 public:
  QuadrupedFlatTransition(const QuadrupedIds& ids, const std::vector<double>& parameters,
                          const std::vector<double>& weight, const std::vector<double>& task_state, const double goal_pos[3]);
  void Transition(const QuadrupedPlantView& d);
  std::vector<double> TaskState() const;
  double GetPhase(double time) const;
  int GetGait() const { return current_mode_ == kModeBiped ? kGaitTrot : (int)current_gait_; }

  int mode = kModeQuadruped;            // Task::mode (GUI / caller writes it)
  std::vector<double> parameters, weight;
  double goal_pos[3];

 private:
  void Walk(double pos[2], double time) const;
  QuadrupedIds ids_;
  std::vector<double> state_;
  int current_mode_ = kModeQuadruped;
  double last_transition_time_ = -1, com_vel_[2] = {0, 0}, gait_switch_time_ = 0, current_gait_ = kGaitStand;
  double phase_velocity_ = 0, save_gait_switch_ = 0;
  std::vector<double> save_weight_;
};

class HumanoidTrackTransition {
 public:
  // key_qpos [nkey][nq], key_qvel [nkey][nv], key_mpos [nkey][3*nmocap]
  HumanoidTrackTransition(int nq, int nv, int nmocap, int nkey, const double* key_qpos, const double* key_qvel,
                          const double* key_mpos);
  // qpos / qvel are overwritten on a clip switch; mocap_pos [3*nmocap] always
  void Transition(double time, double* qpos, double* qvel, double* mocap_pos);
  int mode = 0;
  int current_mode() const { return current_mode_; }
  double reference_time() const { return reference_time_; }

 private:
  int nq_, nv_, nmocap_, nkey_, current_mode_ = -1;
  double reference_time_ = 0;
  std::vector<double> key_qpos_, key_qvel_, key_mpos_;
};

// ShadowReorient::TransitionLocked (mjpc/tasks/shadow_reorient/hand.cc:90-119): a cube lying at rest on the floor is put
// back into the hand (its free joint takes qpos0, its velocity is zeroed).  `on_floor` = a cube-floor contact exists in
// the plant's contact list; `cube_linvel` = the cube_linear_velocity sensor.  Returns true when the reset happened.
class ShadowReorientTransition {
 public:
  ShadowReorientTransition(int cube_qposadr, int cube_dofadr, const double* qpos0_cube /*[7]*/)
      : qadr_(cube_qposadr), dadr_(cube_dofadr) { for (int i = 0; i < 7; i++) qpos0_[i] = qpos0_cube[i]; }
  bool Transition(double* qpos, double* qvel, bool on_floor, const double cube_linvel[3]) const;

 private:
  int qadr_, dadr_;
  double qpos0_[7];
};

}  // namespace mjpc_b200_host
