// TEST INFRASTRUCTURE - CPU oracle. Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference leg may use anything under oracle/.  The product path
// (mujoco_mpc_b200/csrc) never includes or links this directory.
//
// PARITY PINNING: spline, norm, cost, Riccati backward pass and the particle rollout are pinned
// against the reference's own known-answer tests (tests/test_oracle_golden.py).  Contact physics
// is "parity unpinned": MuJoCo (google-deepmind/mujoco @ 088079eff0450e32b98ee743141780ed68307506,
// CMakeLists.txt:55-58) is not in the reference tree and not installable offline, so its published
// algorithm is restated here and anchored on MJPC's call sites (mjpc/trajectory.cc:158,198).
//
// model.h: flat model view over the blob written by mujoco_mpc_b200/blob.py.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace oracle {

enum { GEOM_PLANE = 0, GEOM_HFIELD, GEOM_SPHERE, GEOM_CAPSULE, GEOM_ELLIPSOID, GEOM_CYLINDER, GEOM_BOX };
enum { JNT_FREE = 0, JNT_BALL, JNT_SLIDE, JNT_HINGE };
enum { OBJ_BODY = 0, OBJ_XBODY, OBJ_GEOM, OBJ_SITE };

struct BlobReader {
  const uint8_t* p;
  size_t n;
  int nent;
  BlobReader(const void* data, size_t nbytes) : p((const uint8_t*)data), n(nbytes) {
    if (nbytes < 16 || std::memcmp(p, "MJPCB200", 8) != 0) throw std::runtime_error("bad model blob");
    std::memcpy(&nent, p + 12, 4);
  }
  bool find(const char* name, int* dtype, int* count, int64_t* off) const {
    for (int i = 0; i < nent; i++) {
      const uint8_t* e = p + 16 + 56 * (size_t)i;
      if (std::strncmp((const char*)e, name, 40) == 0) {
        std::memcpy(dtype, e + 40, 4);
        std::memcpy(count, e + 44, 4);
        std::memcpy(off, e + 48, 8);
        return true;
      }
    }
    return false;
  }
  std::vector<int> ints(const char* name) const {
    int dt, c; int64_t off;
    if (!find(name, &dt, &c, &off) || dt != 0) throw std::runtime_error(std::string("blob: missing int ") + name);
    std::vector<int> v(c);
    if (c) std::memcpy(v.data(), p + off, 4 * (size_t)c);
    return v;
  }
  std::vector<double> reals(const char* name) const {
    int dt, c; int64_t off;
    if (!find(name, &dt, &c, &off) || dt != 1) throw std::runtime_error(std::string("blob: missing real ") + name);
    std::vector<double> v(c);
    if (c) std::memcpy(v.data(), p + off, 8 * (size_t)c);
    return v;
  }
  int i(const char* name) const { return ints(name).at(0); }
  double r(const char* name) const { return reals(name).at(0); }
};

template <class T>
struct Model {
  // sizes
  int nq, nv, nu, na, nbody, njnt, ngeom, nsite, nmocap, nkey, nuserdata, nsensordata, npair, ntendon;
  // options
  T timestep, impratio, tolerance, ls_tolerance, meaninertia;
  T gravity[3];
  int cone, iterations, ls_iterations, integrator;
  int disable_contact, disable_eulerdamp, disable_frictionloss, disable_limit, disable_refsafe, disable_warmstart;
  // bodies
  std::vector<int> body_parentid, body_rootid, body_weldid, body_jntnum, body_jntadr, body_dofnum, body_dofadr,
      body_mocapid;
  std::vector<T> body_pos, body_quat, body_ipos, body_iquat, body_mass, body_inertia, body_subtreemass,
      body_invweight0;
  // joints / dofs
  std::vector<int> jnt_type, jnt_qposadr, jnt_dofadr, jnt_bodyid, jnt_limited;
  std::vector<T> jnt_pos, jnt_axis, jnt_range, jnt_stiffness, jnt_margin, jnt_solref, jnt_solimp, qpos0, qpos_spring;
  std::vector<int> dof_bodyid, dof_jntid, dof_parentid;
  std::vector<T> dof_damping, dof_armature, dof_frictionloss, dof_solref, dof_solimp, dof_invweight0;
  // geoms / sites
  std::vector<int> geom_type, geom_bodyid, geom_condim, geom_priority, geom_group, site_bodyid;
  std::vector<T> geom_size, geom_pos, geom_quat, geom_friction, geom_solmix, geom_solref, geom_solimp, geom_margin,
      geom_gap, geom_rbound, site_pos, site_quat;
  // actuators
  std::vector<int> actuator_trnid, actuator_trntype, actuator_biastype, actuator_ctrllimited, actuator_forcelimited;
  std::vector<T> actuator_gear, actuator_gainprm, actuator_biasprm, actuator_ctrlrange, actuator_forcerange;
  // collision candidates, ray-cast set
  std::vector<int> pair_geom1, pair_geom2, ray_geoms;
  // fixed tendons (length = sum coef * qpos)
  std::vector<int> tendon_adr, tendon_num, tendon_limited, wrap_dof, wrap_qposadr;
  std::vector<T> wrap_coef, tendon_range, tendon_margin, tendon_solref, tendon_solimp, tendon_invweight0;
  // keyframes
  std::vector<T> key_qpos, key_qvel, key_ctrl, key_mpos, key_mquat;
  // task (mjpc/task.cc:147-248 parse result + residual registry id + per-task state block)
  int num_term, num_residual, num_trace, residual_id;
  T risk;
  std::vector<int> dim_norm_residual, norm, num_norm_parameter, trace_objtype, trace_objid, task_ids;
  std::vector<T> weight, norm_parameter, parameters, task_state;

  static std::vector<T> cv(const std::vector<double>& v) { return std::vector<T>(v.begin(), v.end()); }

  explicit Model(const void* blob, size_t nbytes) {
    BlobReader b(blob, nbytes);
    nq = b.i("nq"); nv = b.i("nv"); nu = b.i("nu"); na = b.i("na"); nbody = b.i("nbody"); njnt = b.i("njnt");
    ngeom = b.i("ngeom"); nsite = b.i("nsite"); nmocap = b.i("nmocap"); nkey = b.i("nkey");
    nuserdata = b.i("nuserdata"); nsensordata = b.i("nsensordata"); npair = b.i("npair"); ntendon = b.i("ntendon");
    timestep = (T)b.r("opt_timestep"); impratio = (T)b.r("opt_impratio"); tolerance = (T)b.r("opt_tolerance");
    ls_tolerance = (T)b.r("opt_ls_tolerance"); meaninertia = (T)b.r("stat_meaninertia");
    auto g = b.reals("opt_gravity");
    for (int k = 0; k < 3; k++) gravity[k] = (T)g[k];
    cone = b.i("opt_cone"); iterations = b.i("opt_iterations"); ls_iterations = b.i("opt_ls_iterations");
    integrator = b.i("opt_integrator");
    disable_contact = b.i("opt_disable_contact"); disable_eulerdamp = b.i("opt_disable_eulerdamp");
    disable_frictionloss = b.i("opt_disable_frictionloss"); disable_limit = b.i("opt_disable_limit");
    disable_refsafe = b.i("opt_disable_refsafe"); disable_warmstart = b.i("opt_disable_warmstart");
#define LI(x) x = b.ints(#x)
#define LR(x) x = cv(b.reals(#x))
    LI(body_parentid); LI(body_rootid); LI(body_weldid); LI(body_jntnum); LI(body_jntadr); LI(body_dofnum);
    LI(body_dofadr); LI(body_mocapid);
    LR(body_pos); LR(body_quat); LR(body_ipos); LR(body_iquat); LR(body_mass); LR(body_inertia);
    LR(body_subtreemass); LR(body_invweight0);
    LI(jnt_type); LI(jnt_qposadr); LI(jnt_dofadr); LI(jnt_bodyid); LI(jnt_limited);
    LR(jnt_pos); LR(jnt_axis); LR(jnt_range); LR(jnt_stiffness); LR(jnt_margin); LR(jnt_solref); LR(jnt_solimp);
    LR(qpos0); LR(qpos_spring);
    LI(dof_bodyid); LI(dof_jntid); LI(dof_parentid);
    LR(dof_damping); LR(dof_armature); LR(dof_frictionloss); LR(dof_solref); LR(dof_solimp); LR(dof_invweight0);
    LI(geom_type); LI(geom_bodyid); LI(geom_condim); LI(geom_priority); LI(geom_group); LI(site_bodyid);
    LR(geom_size); LR(geom_pos); LR(geom_quat); LR(geom_friction); LR(geom_solmix); LR(geom_solref);
    LR(geom_solimp); LR(geom_margin); LR(geom_gap); LR(geom_rbound); LR(site_pos); LR(site_quat);
    LI(actuator_trnid); LI(actuator_trntype); LI(actuator_biastype); LI(actuator_ctrllimited); LI(actuator_forcelimited);
    LR(actuator_gear); LR(actuator_gainprm); LR(actuator_biasprm); LR(actuator_ctrlrange); LR(actuator_forcerange);
    LI(pair_geom1); LI(pair_geom2); LI(ray_geoms);
    LR(key_qpos); LR(key_qvel); LR(key_ctrl); LR(key_mpos); LR(key_mquat);
    LI(tendon_adr); LI(tendon_num); LI(tendon_limited); LI(wrap_dof); LI(wrap_qposadr);
    LR(wrap_coef); LR(tendon_range); LR(tendon_margin); LR(tendon_solref); LR(tendon_solimp); LR(tendon_invweight0);
#undef LI
#undef LR
    num_term = b.i("task_num_term"); num_residual = b.i("task_num_residual"); num_trace = b.i("task_num_trace");
    residual_id = b.i("task_residual_id"); risk = (T)b.r("task_risk");
    dim_norm_residual = b.ints("task_dim_norm_residual"); norm = b.ints("task_norm");
    num_norm_parameter = b.ints("task_num_norm_parameter"); trace_objtype = b.ints("task_trace_objtype");
    trace_objid = b.ints("task_trace_objid"); task_ids = b.ints("task_ids");
    weight = cv(b.reals("task_weight")); norm_parameter = cv(b.reals("task_norm_parameter"));
    parameters = cv(b.reals("task_parameters")); task_state = cv(b.reals("task_state"));
  }
};

}  // namespace oracle
