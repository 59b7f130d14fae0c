// task_transition.cc - see task_transition.h.
#include "task_transition.h"

#include <algorithm>
#include <cmath>
#include <cstring>

#include "../../../include/mjpc_b200.h"

namespace mjpc_b200_host {

namespace {
// duty ratio, cadence, amplitude, balance, upright, height (quadruped.h:88-97)
constexpr double kGaitParam[kNumGait][6] = {{1, 1, 0, 0, 1, 1}, {0.75, 1, 0.03, 0, 1, 1}, {0.45, 2, 0.03, 0.2, 1, 1},
                                            {0.4, 4, 0.05, 0.03, 0.5, 0.2}, {0.3, 3.5, 0.10, 0.03, 0.2, 0.1}};
constexpr double kGaitAuto[kNumGait] = {0, 0.02, 0.02, 0.6, 2};   // quadruped.h:100-107
constexpr double kAutoGaitFilter = 0.2, kAutoGaitMinTime = 1, kMinAngvel = 0.01;
constexpr int kMotionLengths[10] = {121, 154, 115, 78, 145, 188, 260, 279, 39, 510};   // tracking.cc:43-54
constexpr double kFps = 30.0;
}  // namespace

QuadrupedFlatTransition::QuadrupedFlatTransition(const QuadrupedIds& ids, const std::vector<double>& p,
                                                 const std::vector<double>& w, const std::vector<double>& task_state,
                                                 const double goal[3])
    : parameters(p), weight(w), ids_(ids), state_(task_state) {
  state_.resize(QS_SIZE, 0.0);
  std::copy(goal, goal + 3, goal_pos);
}

double QuadrupedFlatTransition::GetPhase(double time) const {
  return state_[QS_PHASE_START] + (time - state_[QS_PHASE_START_TIME]) * phase_velocity_;
}

std::vector<double> QuadrupedFlatTransition::TaskState() const {
  std::vector<double> s = state_;
  s[QS_MODE] = current_mode_; s[QS_GAIT] = current_gait_; s[QS_PHASE_VELOCITY] = phase_velocity_;
  return s;
}

void QuadrupedFlatTransition::Walk(double pos[2], double time) const {   // quadruped.cc:633-649
  const double *heading = &state_[QS_HEADING], *position = &state_[QS_POSITION];
  const double speed = state_[QS_SPEED], angvel = state_[QS_ANGVEL];
  if (std::fabs(angvel) < kMinAngvel) {
    const double n = std::max(std::hypot(heading[0], heading[1]), 1e-15);
    pos[0] = position[0] + heading[0] + time * speed * heading[0] / n;
    pos[1] = position[1] + heading[1] + time * speed * heading[1] / n;
  } else {
    const double a = time * angvel, c = std::cos(a), s = std::sin(a);
    pos[0] = c * heading[0] - s * heading[1] + position[0];
    pos[1] = s * heading[0] + c * heading[1] + position[1];
  }
}

void QuadrupedFlatTransition::Transition(const QuadrupedPlantView& d) {
  std::vector<double>& s = state_;
  std::vector<double>& P = parameters;
  std::vector<double>& W = weight;
  const double time = d.time;
  // ---- mjData reset
  if (time < last_transition_time_ || last_transition_time_ == -1) {
    if (mode != kModeQuadruped && mode != kModeBiped) mode = kModeQuadruped;
    last_transition_time_ = s[QS_PHASE_START_TIME] = s[QS_PHASE_START] = time;
  }
  // ---- forbidden mode transitions: stateful modes only from Quadruped
  if (mode != current_mode_ && current_mode_ != kModeQuadruped)
    if (mode == kModeWalk || mode == kModeFlip) mode = kModeQuadruped;
  // ---- phase velocity change
  const double pv = 2 * M_PI * P[ids_.p_cadence];
  if (pv != phase_velocity_) {
    s[QS_PHASE_START] = GetPhase(time);
    s[QS_PHASE_START_TIME] = time;
    phase_velocity_ = pv;
  }
  // ---- automatic gait switching
  const double beta = std::exp(-(time - last_transition_time_) / kAutoGaitFilter);
  com_vel_[0] = beta * com_vel_[0] + (1 - beta) * d.torso_subtreelinvel[0];
  com_vel_[1] = beta * com_vel_[1] + (1 - beta) * d.torso_subtreelinvel[1];
  const int auto_switch = (int)P[ids_.p_gait_switch];
  if (mode == kModeBiped) {
    P[ids_.p_gait] = kGaitTrot;
  } else if (auto_switch) {
    const double com_speed = std::hypot(com_vel_[0], com_vel_[1]);
    for (int g = 0; g < kNumGait; g++) {
      if (mode == kModeScramble && g == kGaitStand) continue;
      const bool lower = com_speed > kGaitAuto[g];
      const bool upper = g == kGaitGallop || com_speed <= kGaitAuto[g + 1];
      const bool wait = std::fabs(gait_switch_time_ - time) > kAutoGaitMinTime;
      if (lower && upper && wait) { P[ids_.p_gait] = g; gait_switch_time_ = time; }
    }
  }
  // ---- gait switch, manual or auto
  if (P[ids_.p_gait] != current_gait_) {
    current_gait_ = P[ids_.p_gait];
    const double* gp = kGaitParam[current_mode_ == kModeBiped ? kGaitTrot : (int)current_gait_];
    P[ids_.p_duty] = gp[0]; P[ids_.p_cadence] = gp[1]; P[ids_.p_amplitude] = gp[2];
    W[ids_.w_balance] = gp[3]; W[ids_.w_upright] = gp[4]; W[ids_.w_height] = gp[5];
  }
  // ---- Walk
  if (mode == kModeWalk) {
    const double angvel = P[ids_.p_walk_turn], speed = P[ids_.p_walk_speed];
    double fwd[2] = {d.torso_xmat[0], d.torso_xmat[3]};
    const double n = std::max(std::hypot(fwd[0], fwd[1]), 1e-15);
    fwd[0] /= n; fwd[1] /= n;
    const double left[2] = {-fwd[1], fwd[0]};
    if (mode != current_mode_ || s[QS_ANGVEL] != angvel || s[QS_SPEED] != speed) {
      s[QS_MODE_START_TIME] = time;
      s[QS_SPEED] = speed; s[QS_ANGVEL] = angvel;
      double axis[2] = {d.torso_xpos[0], d.torso_xpos[1]};
      if (std::fabs(angvel) > kMinAngvel) { axis[0] += speed / angvel * left[0]; axis[1] += speed / angvel * left[1]; }
      s[QS_POSITION] = axis[0]; s[QS_POSITION + 1] = axis[1];
      s[QS_HEADING] = goal_pos[0] - axis[0]; s[QS_HEADING + 1] = goal_pos[1] - axis[1];
    }
    Walk(goal_pos, time - s[QS_MODE_START_TIME]);
  }
  // ---- Flip
  if (mode == kModeFlip) {
    if (mode != current_mode_) {
      s[QS_MODE_START_TIME] = time;
      for (int k = 0; k < 4; k++) s[QS_ORIENTATION + k] = d.torso_xquat[k];
      s[QS_GROUND] = d.ground_under_com;
      save_weight_ = W;
      save_gait_switch_ = P[ids_.p_gait_switch];
      W[ids_.w_upright] = 0.2; W[ids_.w_height] = 5; W[ids_.w_position] = 0; W[ids_.w_gait] = 0; W[ids_.w_balance] = 0;
      W[ids_.w_effort] = 0.005; W[ids_.w_posture] = 0.1;
      P[ids_.p_gait_switch] = 0;
    }
    const double flip_time = time - s[QS_MODE_START_TIME];
    if (flip_time >= s[QS_JUMP_TIME] + s[QS_FLIGHT_TIME] + s[QS_LAND_TIME]) {
      mode = kModeQuadruped;
      W = save_weight_;
      P[ids_.p_gait_switch] = save_gait_switch_;
      goal_pos[0] = d.head_site_xpos[0]; goal_pos[1] = d.head_site_xpos[1];
    }
  }
  current_mode_ = mode;
  last_transition_time_ = time;
}

HumanoidTrackTransition::HumanoidTrackTransition(int nq, int nv, int nmocap, int nkey, const double* key_qpos,
                                                 const double* key_qvel, const double* key_mpos)
    : nq_(nq), nv_(nv), nmocap_(nmocap), nkey_(nkey), key_qpos_(key_qpos, key_qpos + (size_t)nkey * nq),
      key_qvel_(key_qvel, key_qvel + (size_t)nkey * nv), key_mpos_(key_mpos, key_mpos + (size_t)nkey * 3 * nmocap) {}

void HumanoidTrackTransition::Transition(double time, double* qpos, double* qvel, double* mocap_pos) {
  int start = 0;
  for (int i = 0; i < mode; i++) start += kMotionLengths[i];
  const int length = kMotionLengths[mode];
  if (current_mode_ != mode || time == 0.0) {
    current_mode_ = mode;
    reference_time_ = time;
    std::copy(&key_qpos_[(size_t)start * nq_], &key_qpos_[(size_t)start * nq_] + nq_, qpos);
    std::copy(&key_qvel_[(size_t)start * nv_], &key_qvel_[(size_t)start * nv_] + nv_, qvel);
  }
  const int last = start + length - 1;
  const double idx = std::min(std::max((time - reference_time_) * kFps + start, 0.0), (double)last);
  const int k0 = (int)std::floor(idx), k1 = std::min(k0 + 1, last);
  const double w1 = idx - k0, w0 = 1.0 - w1;
  const int n3 = 3 * nmocap_;
  for (int i = 0; i < n3; i++) mocap_pos[i] = key_mpos_[(size_t)k0 * n3 + i] * w0 + key_mpos_[(size_t)k1 * n3 + i] * w1;
}

bool ShadowReorientTransition::Transition(double* qpos, double* qvel, bool on_floor, const double v[3]) const {
  if (!(on_floor && std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) < 0.001)) return false;
  for (int i = 0; i < 7; i++) qpos[qadr_ + i] = qpos0_[i];
  for (int i = 0; i < 6; i++) qvel[dadr_ + i] = 0.0;
  return true;
}

}  // namespace mjpc_b200_host

// ------------------------------------------------------------------------------------------ C entry points
using mjpc_b200_host::HumanoidTrackTransition;
using mjpc_b200_host::QuadrupedFlatTransition;
using mjpc_b200_host::QuadrupedIds;
using mjpc_b200_host::QuadrupedPlantView;

extern "C" {

// ids[14] = {p_gait, p_gait_switch, p_cadence, p_amplitude, p_duty, p_walk_speed, p_walk_turn,
//            w_upright, w_height, w_position, w_gait, w_balance, w_effort, w_posture}
void* mjpc_b200_quadruped_transition_create(const int* ids, const double* parameters, int nparam, const double* weight,
                                            int nweight, const double* task_state, int nstate, const double* goal_pos) {
  if (!ids || !parameters || !weight || !task_state || !goal_pos) return nullptr;
  QuadrupedIds q;
  std::memcpy(&q, ids, sizeof(q));
  return new QuadrupedFlatTransition(q, std::vector<double>(parameters, parameters + nparam),
                                     std::vector<double>(weight, weight + nweight),
                                     std::vector<double>(task_state, task_state + nstate), goal_pos);
}
void mjpc_b200_quadruped_transition_destroy(void* p) { delete (QuadrupedFlatTransition*)p; }
// GUI / caller edits of Task::parameters and Task::weight between transitions (either pointer may be NULL)
void mjpc_b200_quadruped_transition_set(void* pv, const double* parameters, const double* weight) {
  auto* p = (QuadrupedFlatTransition*)pv;
  if (parameters) std::copy(parameters, parameters + p->parameters.size(), p->parameters.begin());
  if (weight) std::copy(weight, weight + p->weight.size(), p->weight.begin());
}
// view[24] = {time, subtreelinvel[3], xmat[9], xpos[3], xquat[4], head_site_xpos[3], ground_under_com}.
// mode_inout: Task::mode before / after; outputs (any may be NULL): parameters, weight, task_state [31], goal_pos [3]
void mjpc_b200_quadruped_transition_step(void* pv, int* mode_inout, const double* view, double* parameters, double* weight,
                                         double* task_state, double* goal_pos) {
  auto* p = (QuadrupedFlatTransition*)pv;
  QuadrupedPlantView d;
  d.time = view[0];
  std::copy(view + 1, view + 4, d.torso_subtreelinvel); std::copy(view + 4, view + 13, d.torso_xmat);
  std::copy(view + 13, view + 16, d.torso_xpos); std::copy(view + 16, view + 20, d.torso_xquat);
  std::copy(view + 20, view + 23, d.head_site_xpos); d.ground_under_com = view[23];
  if (mode_inout) p->mode = *mode_inout;
  p->Transition(d);
  if (mode_inout) *mode_inout = p->mode;
  if (parameters) std::copy(p->parameters.begin(), p->parameters.end(), parameters);
  if (weight) std::copy(p->weight.begin(), p->weight.end(), weight);
  if (task_state) { auto s = p->TaskState(); std::copy(s.begin(), s.end(), task_state); }
  if (goal_pos) std::copy(p->goal_pos, p->goal_pos + 3, goal_pos);
}

void* mjpc_b200_track_transition_create(int nq, int nv, int nmocap, int nkey, const double* key_qpos,
                                        const double* key_qvel, const double* key_mpos) {
  if (!key_qpos || !key_qvel || !key_mpos || nkey < 1889) return nullptr;   // the ten clips of tracking.cc:43-54
  return new HumanoidTrackTransition(nq, nv, nmocap, nkey, key_qpos, key_qvel, key_mpos);
}
void mjpc_b200_track_transition_destroy(void* p) { delete (HumanoidTrackTransition*)p; }
// mode: the selected clip; qpos/qvel are overwritten on a clip switch; task_state [2] = {current_mode, reference_time}
void mjpc_b200_track_transition_step(void* pv, int mode, double time, double* qpos, double* qvel, double* mocap_pos,
                                     double* task_state) {
  auto* p = (HumanoidTrackTransition*)pv;
  p->mode = mode;
  p->Transition(time, qpos, qvel, mocap_pos);
  if (task_state) { task_state[0] = p->current_mode(); task_state[1] = p->reference_time(); }
}

void* mjpc_b200_shadow_transition_create(int cube_qposadr, int cube_dofadr, const double* qpos0_cube) {
  return new mjpc_b200_host::ShadowReorientTransition(cube_qposadr, cube_dofadr, qpos0_cube);
}
void mjpc_b200_shadow_transition_destroy(void* p) { delete (mjpc_b200_host::ShadowReorientTransition*)p; }
int mjpc_b200_shadow_transition_step(void* p, double* qpos, double* qvel, int on_floor, const double* cube_linvel) {
  return ((mjpc_b200_host::ShadowReorientTransition*)p)->Transition(qpos, qvel, on_floor != 0, cube_linvel) ? 1 : 0;
}

}  // extern "C"
