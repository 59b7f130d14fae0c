// TEST INFRASTRUCTURE - CPU oracle (see oracle/model.h header).
// residuals.h: task residual functions (the ResidualFn::Residual plug-ins) and the down-ray helper.
//   particle       <- mjpc/test/testdata/particle_residual.h:36-45
//   particle copy  <- mjpc/test/agent/rollout_test.cc:36-40
//   cartpole       <- mjpc/tasks/cartpole/cartpole.cc:36-49
//   quadruped flat <- mjpc/tasks/quadruped/quadruped.cc:33-226, 609-720; Ground: mjpc/utilities.cc:556-574
// Index layouts of task_ids / task_state: mujoco_mpc_b200/task.py (QI_*, QS_*), mirrored below.
#pragma once
#include <cmath>

#include "physics.h"

namespace oracle {

enum { RESIDUAL_PARTICLE = 0, RESIDUAL_PARTICLE_COPY = 1, RESIDUAL_CARTPOLE = 2, RESIDUAL_QUADRUPED_FLAT = 3,
       RESIDUAL_HUMANOID_STAND = 4, RESIDUAL_HUMANOID_TRACK = 5, RESIDUAL_SHADOW_REORIENT = 6 };
enum { SI_GRASP_SITE = 0, SI_CUBE_BODY = 1, SI_GOAL_BODY = 2, SI_KEY_GRASP = 3 };
enum { QS_MODE = 0, QS_MODE_START_TIME = 1, QS_POSITION = 2, QS_HEADING = 5, QS_SPEED = 7, QS_ANGVEL = 8, QS_GROUND = 9,
       QS_ORIENTATION = 10, QS_GAIT = 14, QS_PHASE_START = 15, QS_PHASE_START_TIME = 16, QS_PHASE_VELOCITY = 17,
       QS_JUMP_VEL = 18, QS_FLIGHT_TIME = 19, QS_JUMP_ACC = 20, QS_CROUCH_TIME = 21, QS_LEAP_TIME = 22,
       QS_JUMP_TIME = 23, QS_CROUCH_VEL = 24, QS_LAND_TIME = 25, QS_LAND_ACC = 26, QS_FLIGHT_ROT_VEL = 27,
       QS_JUMP_ROT_VEL = 28, QS_JUMP_ROT_ACC = 29, QS_LAND_ROT_ACC = 30 };
enum { QI_TORSO_BODY = 0, QI_HEAD_SITE = 1, QI_GOAL_MOCAP = 2, QI_FOOT_GEOM = 3, QI_PARAM_GAIT = 7,
       QI_PARAM_BIPED_TYPE = 8, QI_PARAM_CADENCE = 9, QI_PARAM_AMPLITUDE = 10, QI_PARAM_DUTY = 11,
       QI_PARAM_ARM_POSTURE = 12, QI_PARAM_HEADING = 13, QI_PARAM_FLIP_DIR = 14, QI_KEY_HOME = 15, QI_KEY_CROUCH = 16 };
enum { kModeQuadruped = 0, kModeBiped, kModeWalk, kModeScramble, kModeFlip };
enum { kFootFL = 0, kFootHL, kFootFR, kFootHR };

// nearest non-negative intersection distance of a ray with one geom, -1 if none ([EXT] mj_ray semantics)
template <class T>
T ray_geom(const T* gpos, const T* gmat, const T* size, int type, const T* pnt, const T* vec) {
  T dp[3] = {pnt[0] - gpos[0], pnt[1] - gpos[1], pnt[2] - gpos[2]};
  T lp[3], lv[3];
  rot_vec_T(lp, gmat, dp);
  rot_vec_T(lv, gmat, vec);
  if (type == GEOM_PLANE) {
    if (lv[2] > -kMinVal<T>()) return -1;
    T x = -lp[2] / lv[2];
    if (x < 0) return -1;
    T p0 = lp[0] + x * lv[0], p1 = lp[1] + x * lv[1];
    if ((size[0] <= 0 || mm::fabs(p0) <= size[0]) && (size[1] <= 0 || mm::fabs(p1) <= size[1])) return x;
    return -1;
  }
  if (type == GEOM_SPHERE) {
    T a = dot3(lv, lv), b = dot3(lv, lp), c = dot3(lp, lp) - size[0] * size[0];
    T det = b * b - a * c;
    if (det < 0 || a < kMinVal<T>()) return -1;
    T sq = mm::sqrt(det);
    T x0 = (-b - sq) / a, x1 = (-b + sq) / a;
    if (x0 >= 0) return x0;
    if (x1 >= 0) return x1;
    return -1;
  }
  if (type == GEOM_BOX) {
    T best = -1;
    for (int i = 0; i < 3; i++) {
      if (mm::fabs(lv[i]) <= kMinVal<T>()) continue;
      for (int s = -1; s <= 1; s += 2) {
        T x = ((T)s * size[i] - lp[i]) / lv[i];
        if (x < 0) continue;
        int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
        T q1 = lp[i1] + x * lv[i1], q2 = lp[i2] + x * lv[i2];
        if (mm::fabs(q1) <= size[i1] && mm::fabs(q2) <= size[i2] && (best < 0 || x < best)) best = x;
      }
    }
    return best;
  }
  return -1;
}

// mjpc::Ground: height of the nearest group-0 geom under pos; sets *ok=false if nothing is hit
template <class T>
T Ground(const Model<T>& m, const Data<T>& d, const T* pos, bool* ok) {
  T down[3] = {0, 0, -1};
  const T height_offset = (T)0.5;
  T query[3] = {pos[0], pos[1], pos[2] + height_offset};
  T best = -1;
  for (int g : m.ray_geoms) {
    T x = ray_geom(&d.geom_xpos[3 * g], &d.geom_xmat[9 * g], &m.geom_size[3 * g], m.geom_type[g], query, down);
    if (x >= 0 && (best < 0 || x < best)) best = x;
  }
  if (best < 0) { *ok = false; return 0; }
  return pos[2] + height_offset - best;
}

template <class T>
void sub_quat(T* res, const T* qa, const T* qb) {  // qb * quat(res) = qa
  T qneg[4] = {qb[0], -qb[1], -qb[2], -qb[3]}, qd[4];
  quat_mul(qd, qneg, qa);
  T axis[3] = {qd[1], qd[2], qd[3]};
  T s = norm3(axis);
  if (s < kMinVal<T>()) { axis[0] = 1; axis[1] = axis[2] = 0; } else { axis[0] /= s; axis[1] /= s; axis[2] /= s; }
  T speed = 2 * mm::atan2(s, qd[0]);
  if (speed > (T)M_PI) speed -= 2 * (T)M_PI;
  for (int c = 0; c < 3; c++) res[c] = axis[c] * speed;
}

template <class T>
void residual_particle(const Model<T>& m, Data<T>& d, T* r) {
  for (int i = 0; i < m.nq; i++) r[i] = d.qpos[i];
  r[0] -= d.mocap_pos[0];
  r[1] -= d.mocap_pos[1];
  for (int i = 0; i < m.nv; i++) r[2 + i] = d.qvel[i];
}
template <class T>
void residual_particle_copy(const Model<T>& m, Data<T>& d, T* r) {
  for (int i = 0; i < m.nq; i++) r[i] = d.qpos[i];
  for (int i = 0; i < m.nv; i++) r[m.nq + i] = d.qvel[i];
}
template <class T>
void residual_cartpole(const Model<T>& m, Data<T>& d, T* r) {
  r[0] = mm::cos(d.qpos[1]) - 1;
  r[1] = d.qpos[0] - m.parameters[0];
  r[2] = d.qvel[1];
  r[3] = d.ctrl[0];
}

// Humanoid Stand <- mjpc/tasks/humanoid/stand/stand.cc:30-97.  task_ids = {torso body, head body, site sp0..sp3};
// sensors restated: framepos(objtype=body) = xipos, framepos(site) = site_xpos, subtreecom / subtreelinvel of torso.
enum { HI_TORSO_BODY = 0, HI_HEAD_BODY, HI_SITE_SP0, HI_SIZE = 6 };
template <class T>
void residual_humanoid_stand(const Model<T>& m, Data<T>& d, T* r) {
  const int* I = m.task_ids.data();
  const T* f[4];
  for (int k = 0; k < 4; k++) f[k] = &d.site_xpos[3 * I[HI_SITE_SP0 + k]];
  const T* head = &d.xipos[3 * I[HI_HEAD_BODY]];
  int counter = 0;
  r[counter++] = head[2] - (T)0.25 * (f[0][2] + f[1][2] + f[2][2] + f[3][2]) - m.parameters[0];
  const T* com = &d.subtree_com[3 * I[HI_TORSO_BODY]];
  const T* vel = &d.subtree_linvel[3 * I[HI_TORSO_BODY]];
  const T kFallTime = (T)0.2;
  T dx[2];
  for (int c = 0; c < 2; c++)
    dx[c] = (T)0.25 * (f[0][c] + f[1][c] + f[2][c] + f[3][c]) - (com[c] + vel[c] * kFallTime);
  r[counter++] = mm::sqrt(dx[0] * dx[0] + dx[1] * dx[1]);
  r[counter++] = vel[0];
  r[counter++] = vel[1];
  for (int i = 6; i < m.nv; i++) r[counter++] = d.qvel[i];
  for (int i = 0; i < m.nu; i++) r[counter++] = d.ctrl[i];
}

// Humanoid Track <- mjpc/tasks/humanoid/tracking/tracking.cc:29-216.  task_ids = 16 tracking-site ids then 16 mocap
// ids (body order of tracking.cc:71-75); task_state = [current_mode, reference_time].  Sensors restated:
// tracking_pos[x] = site_xpos, tracking_linvel[x] = world-frame linear velocity of the site (mj_objectVelocity).
constexpr int kTrackMotionLengths[10] = {121, 154, 115, 78, 145, 188, 260, 279, 39, 510};
constexpr double kTrackFps = 30.0;
template <class T>
void residual_humanoid_track(const Model<T>& m, Data<T>& d, T* r) {
  const int* I = m.task_ids.data();
  const int mode = (int)m.task_state[0];
  const T reference_time = m.task_state[1];
  int start = 0;
  for (int i = 0; i < mode; i++) start += kTrackMotionLengths[i];
  const int length = kTrackMotionLengths[mode];
  const T current_index = (d.time - reference_time) * (T)kTrackFps + start;
  const int last_key_index = start + length - 1;
  // ComputeInterpolationValues (tracking.cc:29-39)
  const T clamped = mm::max((T)0, mm::min((T)last_key_index, current_index));
  const int k0 = (int)std::floor(mm::dbl(clamped));
  const int k1 = k0 + 1 < last_key_index ? k0 + 1 : last_key_index;
  const T w1 = clamped - (T)k0, w0 = (T)1 - w1;
  const int nm = m.nmocap;
  int counter = 0;
  for (int i = 6; i < m.nv; i++) r[counter++] = d.qvel[i];
  for (int i = 0; i < m.nu; i++) r[counter++] = d.ctrl[i];
  T mpos[16][3], spos[16][3], avg_m[3] = {0, 0, 0}, avg_s[3] = {0, 0, 0};
  for (int b = 0; b < 16; b++) {
    const T* p0 = &m.key_mpos[(size_t)nm * 3 * k0 + 3 * I[16 + b]];
    const T* p1 = &m.key_mpos[(size_t)nm * 3 * k1 + 3 * I[16 + b]];
    for (int c = 0; c < 3; c++) {
      mpos[b][c] = p0[c] * w0 + p1[c] * w1;
      spos[b][c] = d.site_xpos[3 * I[b] + c];
      avg_m[c] += mpos[b][c]; avg_s[c] += spos[b][c];
    }
  }
  for (int c = 0; c < 3; c++) { avg_m[c] /= 16; avg_s[c] /= 16; }
  for (int c = 0; c < 3; c++) r[counter++] = avg_m[c] - avg_s[c];
  for (int b = 0; b < 16; b++)
    for (int c = 0; c < 3; c++) r[counter++] = (mpos[b][c] - avg_m[c]) - (spos[b][c] - avg_s[c]);
  for (int b = 0; b < 16; b++) {
    const T* p0 = &m.key_mpos[(size_t)nm * 3 * k0 + 3 * I[16 + b]];
    const T* p1 = &m.key_mpos[(size_t)nm * 3 * k1 + 3 * I[16 + b]];
    const int body = m.site_bodyid[I[b]];
    const T* cv = &d.cvel[6 * body];
    T off[3], wx[3];
    for (int c = 0; c < 3; c++) off[c] = d.site_xpos[3 * I[b] + c] - d.subtree_com[3 * m.body_rootid[body] + c];
    cross3(wx, cv, off);
    for (int c = 0; c < 3; c++) r[counter++] = (p1[c] - p0[c]) * (T)kTrackFps - (cv[3 + c] + wx[c]);
  }
}

template <class T>
struct QuadrupedFn {
  const Model<T>& m;
  const T* S;  // task state
  const int* I;
  explicit QuadrupedFn(const Model<T>& model) : m(model), S(model.task_state.data()), I(model.task_ids.data()) {}
  T param(int qi) const { return m.parameters[I[qi]]; }
  int mode() const { return (int)S[QS_MODE]; }
  T GetPhase(T time) const { return S[QS_PHASE_START] + (time - S[QS_PHASE_START_TIME]) * S[QS_PHASE_VELOCITY]; }
  int GetGait() const { return mode() == kModeBiped ? 2 : (int)S[QS_GAIT]; }
  T StepHeight(T time, T footphase, T duty_ratio) const {
    const T pi = (T)M_PI;
    T angle = mm::fmod(time + pi - footphase, 2 * pi) - pi;
    T value = 0;
    if (duty_ratio < 1) {
      angle *= (T)0.5 / (1 - duty_ratio);
      value = mm::cos(mm::max(-pi / 2, mm::min(pi / 2, angle)));
    }
    return mm::fabs(value) < (T)1e-6 ? (T)0 : value;
  }
  void FootStep(T* step, T time, int gait) const {
    static const double kGaitPhase[5][4] = {{0, 0, 0, 0}, {0, 0.75, 0.5, 0.25}, {0, 0.5, 0.5, 0},
                                            {0, 0.33, 0.33, 0.66}, {0, 0.4, 0.05, 0.35}};
    T amplitude = param(QI_PARAM_AMPLITUDE), duty = param(QI_PARAM_DUTY);
    for (int f = 0; f < 4; f++) step[f] = amplitude * StepHeight(time, 2 * (T)M_PI * (T)kGaitPhase[gait][f], duty);
  }
  void Walk(T* pos, T time) const {
    const T* heading = S + QS_HEADING;
    const T* position = S + QS_POSITION;
    if (mm::fabs(S[QS_ANGVEL]) < (T)0.01) {
      T fw[2] = {heading[0], heading[1]};
      T n = mm::sqrt(fw[0] * fw[0] + fw[1] * fw[1]);
      if (n < kMinVal<T>()) { fw[0] = 1; fw[1] = 0; } else { fw[0] /= n; fw[1] /= n; }
      pos[0] = position[0] + heading[0] + time * S[QS_SPEED] * fw[0];
      pos[1] = position[1] + heading[1] + time * S[QS_SPEED] * fw[1];
    } else {
      T angle = time * S[QS_ANGVEL];
      T c = mm::cos(angle), s = mm::sin(angle);
      pos[0] = c * heading[0] - s * heading[1] + position[0];
      pos[1] = s * heading[0] + c * heading[1] + position[1];
    }
  }
  T FlipHeight(T time) const {
    const T kHeightQuadruped = (T)0.25, kLeapHeight = (T)0.5;
    T jump = S[QS_JUMP_TIME], flight = S[QS_FLIGHT_TIME], land = S[QS_LAND_TIME];
    if (time >= jump + flight + land) return kHeightQuadruped + S[QS_GROUND];
    T h = 0;
    if (time < jump) {
      h = kHeightQuadruped + time * S[QS_CROUCH_VEL] + (T)0.5 * time * time * S[QS_JUMP_ACC];
    } else if (time >= jump && time < jump + flight) {
      time -= jump;
      h = kLeapHeight + S[QS_JUMP_VEL] * time - (T)0.5 * (T)9.81 * time * time;
    } else if (time >= jump + flight) {
      time -= jump + flight;
      h = kLeapHeight - S[QS_JUMP_VEL] * time + (T)0.5 * S[QS_LAND_ACC] * time * time;
    }
    return h + S[QS_GROUND];
  }
  void FlipQuat(T* quat, T time) const {
    const T pi = (T)M_PI;
    T jump = S[QS_JUMP_TIME], flight = S[QS_FLIGHT_TIME], land = S[QS_LAND_TIME], crouch = S[QS_CROUCH_TIME];
    T angle = 0;
    if (time >= jump + flight + land) {
      angle = 2 * pi;
    } else if (time >= crouch && time < jump) {
      time -= crouch;
      angle = (T)0.5 * S[QS_JUMP_ROT_ACC] * time * time + S[QS_JUMP_ROT_VEL] * time;
    } else if (time >= jump && time < jump + flight) {
      time -= jump;
      angle = pi / 2 + S[QS_FLIGHT_ROT_VEL] * time;
    } else if (time >= jump + flight) {
      time -= jump + flight;
      angle = (T)1.75 * pi + S[QS_FLIGHT_ROT_VEL] * time - (T)0.5 * S[QS_LAND_ROT_ACC] * time * time;
    }
    int flip_dir = (int)param(QI_PARAM_FLIP_DIR);
    T axis[3] = {0, flip_dir ? (T)1 : (T)-1, 0};
    T q[4];
    axis_angle_quat(q, axis, angle);
    quat_mul(quat, S + QS_ORIENTATION, q);
  }

  void Residual(Data<T>& d, T* residual) const {
    const T kHeightQuadruped = (T)0.25, kHeightBiped = (T)0.6, kFootRadius = (T)0.02;
    const T kJointPostureGain[3] = {2, 1, 1};
    int counter = 0;
    int cur = mode();
    const T* foot_pos[4];
    for (int f = 0; f < 4; f++) foot_pos[f] = &d.geom_xpos[3 * I[QI_FOOT_GEOM + f]];
    int handstand_i = (int)param(QI_PARAM_BIPED_TYPE);
    // average foot position (quadruped.cc:609-627)
    T avg[3];
    if (cur == kModeBiped) {
      int a = handstand_i ? kFootFL : kFootHL, b = handstand_i ? kFootFR : kFootHR;
      for (int c = 0; c < 3; c++) avg[c] = (T)0.5 * (foot_pos[a][c] + foot_pos[b][c]);
    } else {
      for (int c = 0; c < 3; c++)
        avg[c] = (T)0.25 * (((foot_pos[kFootHL][c] + foot_pos[kFootHR][c]) + foot_pos[kFootFL][c]) + foot_pos[kFootFR][c]);
    }
    int torso = I[QI_TORSO_BODY];
    const T* torso_xmat = &d.xmat[9 * torso];
    const T* goal_pos = &d.mocap_pos[3 * I[QI_GOAL_MOCAP]];
    const T* compos = &d.subtree_com[3 * torso];
    // ---------- Upright
    if (cur != kModeFlip) {
      if (cur == kModeBiped) residual[counter++] = torso_xmat[6] - (handstand_i ? -1 : 1);
      else residual[counter++] = torso_xmat[8] - 1;
      residual[counter++] = 0;
      residual[counter++] = 0;
    } else {
      T quat[4];
      FlipQuat(quat, d.time - S[QS_MODE_START_TIME]);
      sub_quat(residual + counter, &d.xquat[4 * torso], quat);
      counter += 3;
    }
    // ---------- Height
    const T* torso_pos = &d.xipos[3 * torso];
    bool is_biped = cur == kModeBiped;
    T height_goal = is_biped ? kHeightBiped : kHeightQuadruped;
    if (cur == kModeScramble) residual[counter++] = 0;
    else if (cur == kModeFlip) residual[counter++] = torso_pos[2] - FlipHeight(d.time - S[QS_MODE_START_TIME]);
    else residual[counter++] = (torso_pos[2] - avg[2]) - height_goal;
    // ---------- Position
    const T* head = &d.site_xpos[3 * I[QI_HEAD_SITE]];
    T target[3] = {goal_pos[0], goal_pos[1], goal_pos[2]};
    if (cur == kModeWalk) Walk(target, d.time - S[QS_MODE_START_TIME]);
    residual[counter++] = head[0] - target[0];
    residual[counter++] = head[1] - target[1];
    residual[counter++] = cur == kModeScramble ? 2 * (head[2] - target[2]) : (T)0;
    // ---------- Gait
    int gait = GetGait();
    T step[4];
    FootStep(step, GetPhase(d.time), gait);
    for (int f = 0; f < 4; f++) {
      if (is_biped) {
        bool front_hand = !handstand_i && (f == kFootFL || f == kFootFR);
        bool back_hand = handstand_i && (f == kFootHL || f == kFootHR);
        if (front_hand || back_hand) { residual[counter++] = 0; continue; }
      }
      T query[3] = {foot_pos[f][0], foot_pos[f][1], foot_pos[f][2]};
      if (cur == kModeScramble) {
        T tg[3];
        for (int c = 0; c < 3; c++) tg[c] = goal_pos[c] - foot_pos[f][c];
        tg[2] = 0;
        normalize3(tg);
        for (int c = 0; c < 3; c++) query[c] += (T)0.15 * tg[c];
      }
      bool ok = true;
      T ground_height = Ground(m, d, query, &ok);
      if (!ok) d.warning = true;  // reference: mju_error("no group 0 geom detected by raycast")
      T height_target = ground_height + kFootRadius + step[f];
      T height_difference = foot_pos[f][2] - height_target;
      if (cur == kModeScramble) height_difference = mm::min((T)0, height_difference);
      residual[counter++] = step[f] ? height_difference : (T)0;
    }
    // ---------- Balance
    const T* comvel = &d.subtree_linvel[3 * torso];
    T fall_time = mm::sqrt(2 * height_goal / (T)9.81);
    residual[counter++] = compos[0] + comvel[0] * fall_time - avg[0];
    residual[counter++] = compos[1] + comvel[1] * fall_time - avg[1];
    // ---------- Effort
    for (int i = 0; i < m.nu; i++) residual[counter + i] = (T)2e-2 * d.actuator_force[i];
    counter += m.nu;
    // ---------- Posture
    const T* home = &m.key_qpos[m.nq * I[QI_KEY_HOME]];
    for (int i = 0; i < m.nu; i++) residual[counter + i] = d.qpos[7 + i] - home[7 + i];
    if (cur == kModeFlip) {
      T flip_time = d.time - S[QS_MODE_START_TIME];
      if (flip_time < S[QS_CROUCH_TIME]) {
        const T* crouch = &m.key_qpos[m.nq * I[QI_KEY_CROUCH]];
        for (int i = 0; i < m.nu; i++) residual[counter + i] = d.qpos[7 + i] - crouch[7 + i];
      } else if (flip_time >= S[QS_CROUCH_TIME] && flip_time < S[QS_JUMP_TIME] + S[QS_FLIGHT_TIME]) {
        for (int i = 0; i < m.nu; i++) residual[counter + i] = 0;
      }
    }
    for (int f = 0; f < 4; f++)
      for (int j = 0; j < 3; j++) residual[counter + 3 * f + j] *= kJointPostureGain[j];
    if (cur == kModeBiped) {
      T arm = param(QI_PARAM_ARM_POSTURE);
      int base = handstand_i ? 6 : 0;
      for (int i = 0; i < 6; i++) residual[counter + base + i] *= arm;
    }
    counter += m.nu;
    // ---------- Yaw
    T th[2] = {torso_xmat[0], torso_xmat[3]};
    if (cur == kModeBiped) {
      int hs = handstand_i ? 1 : -1;
      th[0] = hs * torso_xmat[2];
      th[1] = hs * torso_xmat[5];
    }
    T n = mm::sqrt(th[0] * th[0] + th[1] * th[1]);
    if (n < kMinVal<T>()) { th[0] = 1; th[1] = 0; } else { th[0] /= n; th[1] /= n; }
    T heading_goal = param(QI_PARAM_HEADING);
    residual[counter++] = th[0] - mm::cos(heading_goal);
    residual[counter++] = th[1] - mm::sin(heading_goal);
    // ---------- Angular momentum (sensor "torso_angmom" is declared subtreelinvel: task_flat.xml:144)
    for (int c = 0; c < 3; c++) residual[counter++] = comvel[c];
  }
};

template <class T>
void residual_quadruped(const Model<T>& m, Data<T>& d, T* r) {
  QuadrupedFn<T>(m).Residual(d, r);
}

// Shadow Hand cube reorientation <- mjpc/tasks/shadow_reorient/hand.cc:37-84 (81 residuals).  Sensors restated:
// palm_position = grasp_site xpos, cube_position / cube_orientation = cube body frame, cube_goal_orientation = goal body
// frame, cube_linear_velocity = world-frame velocity of the cube body origin (framelinvel -> mj_objectVelocity).
// The posture / velocity terms read qpos + 7 and qvel + 6 for 26 values LITERALLY: with the goal ball joint first
// (4 qpos / 3 qvel) and the cube free joint second they start at the cube QUATERNION and at the cube's ANGULAR
// velocity, not at the first hand joint.
template <class T>
void residual_shadow_reorient(const Model<T>& m, Data<T>& d, T* r) {
  const int* I = m.task_ids.data();
  const int cube = I[SI_CUBE_BODY], goal = I[SI_GOAL_BODY];
  int counter = 0;
  const T* palm = &d.site_xpos[3 * I[SI_GRASP_SITE]];
  const T* pos = &d.xpos[3 * cube];
  for (int c = 0; c < 3; c++) r[counter++] = pos[c] - palm[c];
  T gq[4] = {d.xquat[4 * goal], d.xquat[4 * goal + 1], d.xquat[4 * goal + 2], d.xquat[4 * goal + 3]};
  quat_normalize(gq);
  sub_quat(r + counter, gq, &d.xquat[4 * cube]);
  counter += 3;
  {
    const T* cv = &d.cvel[6 * cube];
    T off[3], wx[3];
    for (int c = 0; c < 3; c++) off[c] = d.xpos[3 * cube + c] - d.subtree_com[3 * m.body_rootid[cube] + c];
    cross3(wx, cv, off);
    for (int c = 0; c < 3; c++) r[counter++] = cv[3 + c] + wx[c];
  }
  for (int i = 0; i < m.nu; i++) r[counter++] = d.actuator_force[i];
  const T* key = &m.key_qpos[m.nq * I[SI_KEY_GRASP]];
  for (int i = 0; i < 26; i++) r[counter++] = d.qpos[7 + i] - key[7 + i];
  for (int i = 0; i < 26; i++) r[counter++] = d.qvel[6 + i];
}

template <class T>
ResidualCallback<T> residual_by_id(int id) {
  switch (id) {
    case RESIDUAL_PARTICLE: return residual_particle<T>;
    case RESIDUAL_PARTICLE_COPY: return residual_particle_copy<T>;
    case RESIDUAL_CARTPOLE: return residual_cartpole<T>;
    case RESIDUAL_QUADRUPED_FLAT: return residual_quadruped<T>;
    case RESIDUAL_HUMANOID_STAND: return residual_humanoid_stand<T>;
    case RESIDUAL_HUMANOID_TRACK: return residual_humanoid_track<T>;
    case RESIDUAL_SHADOW_REORIENT: return residual_shadow_reorient<T>;
  }
  return nullptr;
}

}  // namespace oracle
