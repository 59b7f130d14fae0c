// TEST INFRASTRUCTURE - CPU oracle (see oracle/model.h header).
// ilqg.h: iLQG pieces of the hot path.
//   iLQGPolicy::Action (time-interpolated feedback policy)  <- mjpc/planners/ilqg/policy.cc:82-161,
//       interpolation helpers mjpc/utilities.cc:303-422, FindInterval mjpc/utilities.h:125-144
//   discrete line-search policy (ActionRollouts)             <- mjpc/planners/ilqg/planner.cc:630-692
//   ModelDerivatives::Compute                                <- mjpc/planners/model_derivatives.cc:45-165
//       ([EXT] mjd_transitionFD restated: one-sided / clamped control differences, tangent-space state diff)
//   CostDerivatives::DerivativeStep/Compute                  <- mjpc/planners/cost_derivatives.cc:77-230
//   iLQGBackwardPass::RiccatiStep + driver                   <- mjpc/planners/ilqg/backward_pass.cc:65-250,
//       mjpc/planners/ilqg/planner.cc:429-520  ([EXT] mju_boxQP restated: projected Newton)
//   Pinned by tests/test_oracle_golden.py against mjpc/test/ilqg_planner/backward_pass_test.cc:101-108.
#pragma once
#include <algorithm>
#include <cmath>
#include <thread>
#include <vector>

#include "rollout.h"

namespace oracle {

// ------------------------------------------------------------------------------------------ policy
template <class T>
struct ILQGPolicyData {
  int H = 0;
  std::vector<T> u, x, t, K, du;  // [H][nu], [H][dim_state], [H], [H][nu][n], [H][nu]
};

template <class T>
void find_interval(int* bounds, const T* seq, T value, int length) {
  int upper = 0;
  while (upper < length && !(value < seq[upper])) upper++;
  int lower = upper - 1;
  if (lower < 0) { bounds[0] = bounds[1] = 0; }
  else if (lower > length - 1) { bounds[0] = bounds[1] = length - 1; }
  else { bounds[0] = mm::max(lower, 0); bounds[1] = mm::min(upper, length - 1); }
}
template <class T>
T fd_slope(T x, const T* xs, const T* ys, int dim, int length, int i) {
  int b[2];
  find_interval(b, xs, x, length);
  if (b[0] == 0 && b[1] == 0) {
    if (length > 2) return (ys[dim * (b[1] + 1) + i] - ys[dim * b[1] + i]) / (xs[b[1] + 1] - xs[b[1]]);
    return 0;
  } else if (b[0] == length - 1 && b[1] == length - 1) {
    if (length > 2) return (ys[dim * b[0] + i] - ys[dim * (b[0] - 1) + i]) / (xs[b[0]] - xs[b[0] - 1]);
    return 0;
  } else if (b[0] == 0) {
    return (ys[dim * b[1] + i] - ys[dim * b[0] + i]) / (xs[b[1]] - xs[b[0]]);
  }
  return (T)0.5 * (ys[dim * b[1] + i] - ys[dim * b[0] + i]) / (xs[b[1]] - xs[b[0]]) +
         (T)0.5 * (ys[dim * b[0] + i] - ys[dim * (b[0] - 1) + i]) / (xs[b[0]] - xs[b[0] - 1]);
}
// representation 0 zero-order, 1 linear, 2 cubic
template <class T>
void interpolate(T* out, T x, const T* xs, const T* ys, int dim, int length, int rep) {
  int b[2];
  find_interval(b, xs, x, length);
  if (rep == 0 || b[0] == b[1]) {
    for (int i = 0; i < dim; i++) out[i] = ys[dim * b[0] + i];
    return;
  }
  T t = (x - xs[b[0]]) / (xs[b[1]] - xs[b[0]]);
  if (rep == 1) {
    for (int i = 0; i < dim; i++) out[i] = ys[dim * b[0] + i] * (1 - t) + ys[dim * b[1] + i] * t;
    return;
  }
  T dt = xs[b[1]] - xs[b[0]];
  T c0 = 2 * t * t * t - 3 * t * t + 1, c1 = (t * t * t - 2 * t * t + t) * dt, c2 = -2 * t * t * t + 3 * t * t,
    c3 = (t * t * t - t * t) * dt;
  for (int i = 0; i < dim; i++) {
    T p0 = ys[b[0] * dim + i], p1 = ys[b[1] * dim + i];
    T m0 = fd_slope(xs[b[0]], xs, ys, dim, length, i), m1 = fd_slope(xs[b[1]], xs, ys, dim, length, i);
    out[i] = c0 * p0 + c1 * m0 + c2 * p1 + c3 * m1;
  }
}

template <class T>
void normalize_state_quats(const Model<T>& m, T* qpos) {
  for (int j = 0; j < m.njnt; j++) {
    if (m.jnt_type[j] == JNT_FREE) quat_normalize(qpos + m.jnt_qposadr[j] + 3);
    else if (m.jnt_type[j] == JNT_BALL) quat_normalize(qpos + m.jnt_qposadr[j]);
  }
}

// mode 0/1/2: time-indexed with that interpolation, feedback scaled by `step`;  mode 3: step-indexed,
// action = u[t] + step*du[t] + K[t] * (x (-) x_nom[t])
template <class T>
Policy<T> ilqg_policy(const Model<T>& m, const ILQGPolicyData<T>& pd, T step, int mode) {
  return [&m, &pd, step, mode](T* action, const T* state, T time, int index) {
    int ds = m.nq + m.nv + m.na, n = 2 * m.nv + m.na, nu = m.nu, H = pd.H;
    std::vector<T> xi(ds), K(nu * n), dx(n);
    T scale;
    if (mode == 3) {
      for (int i = 0; i < nu; i++) action[i] = pd.u[index * nu + i] + step * pd.du[index * nu + i];
      for (int i = 0; i < ds; i++) xi[i] = pd.x[index * ds + i];
      for (int i = 0; i < nu * n; i++) K[i] = pd.K[index * nu * n + i];
      scale = 1;
    } else {
      int b[2];
      find_interval(b, pd.t.data(), time, H);
      int rep = (b[0] == b[1]) ? 0 : mode;
      interpolate(action, time, pd.t.data(), pd.u.data(), nu, H - 1, rep);
      interpolate(xi.data(), time, pd.t.data(), pd.x.data(), ds, H, rep);
      if (rep != 0) normalize_state_quats(m, xi.data());
      interpolate(K.data(), time, pd.t.data(), pd.K.data(), nu * n, H - 1, rep);
      scale = step;
    }
    state_diff(m, dx.data(), xi.data(), state, (T)1);
    for (int i = 0; i < nu; i++) {
      T a = 0;
      for (int j = 0; j < n; j++) a += K[i * n + j] * dx[j];
      action[i] += scale * a;
    }
    clamp_ctrl(action, m.actuator_ctrlrange.data(), nu);
  };
}

// ------------------------------------------------------------------------------------------ model derivatives
// qpos <- qpos (+) eps * e_i   ([EXT] mj_integratePos with a unit tangent vector)
template <class T>
void perturb_pos(const Model<T>& m, T* qpos, int dof, T eps) {
  int j = m.dof_jntid[dof];
  int qa = m.jnt_qposadr[j], da = m.jnt_dofadr[j], k = dof - da;
  switch (m.jnt_type[j]) {
    case JNT_FREE:
      if (k < 3) { qpos[qa + k] += eps; break; }
      { T w[3] = {0, 0, 0}; w[k - 3] = 1; quat_integrate(qpos + qa + 3, w, eps); }
      break;
    case JNT_BALL: { T w[3] = {0, 0, 0}; w[k] = 1; quat_integrate(qpos + qa, w, eps); break; }
    default: qpos[qa] += eps;
  }
}

// Evaluate / interpolate time indices of ModelDerivatives::Compute (model_derivatives.cc:56-72): every (skip+1)-th step,
// plus T-2 and T-1; everything else is linearly interpolated between its evaluated neighbours (:109-164).
inline void derivative_indices(int T, int skip, std::vector<int>& evaluate, std::vector<int>& interpolate) {
  evaluate.clear(); interpolate.clear();
  const int s = skip + 1;
  evaluate.push_back(0);
  for (int t = s; t < T - s; t += s) evaluate.push_back(t);
  evaluate.push_back(T - 2);
  evaluate.push_back(T - 1);
  for (int t = 0, e = 0; t < T; t++) {
    if (e == (int)evaluate.size() || evaluate[e] > t) interpolate.push_back(t);
    else e++;
  }
}

// A[H][n][n], B[H][n][nu], C[H][nr][n], D[H][nr][nu]; rows of C/D are the task residual rows only
// (the reference differentiates all nsensordata rows but CostDerivatives reads only the first num_residual).
// mode 0: one-sided differences; mode 1: centred ([EXT] mjd_transitionFD flg_centered; a control whose forward or
// backward nudge would leave ctrlrange falls back to the one-sided difference that stays inside).
// skip: ModelDerivatives::Compute's derivative_skip (evaluate list + linear interpolation).  nthreads: the reference
// schedules one task per evaluated time step on its ThreadPool (model_derivatives.cc:76-104).
template <class T>
int model_derivatives(const Model<T>& m, const double* states, const double* actions, const double* times,
                      const double* mocap, int H, double tol, double* A, double* B, double* C, double* D, int skip = 0,
                      int mode = 0, int nthreads = 1) {
  int nq = m.nq, nv = m.nv, nu = m.nu, ds = nq + nv, n = 2 * nv, nr = m.num_residual;
  ResidualCallback<T> cb = residual_by_id<T>(m.residual_id);
  const T eps = (T)tol;
  std::vector<int> evaluate, interpolate_t;
  if (H >= 2) derivative_indices(H, skip, evaluate, interpolate_t);
  else evaluate.push_back(0);
  std::fill(A, A + (size_t)H * n * n, 0.0); std::fill(B, B + (size_t)H * n * nu, 0.0);
  std::fill(C, C + (size_t)H * nr * n, 0.0); std::fill(D, D + (size_t)H * nr * nu, 0.0);
  auto worker = [&](int w) {
    Data<T> d(m);
    for (int i = 0; i < m.nmocap; i++) {
      for (int c = 0; c < 3; c++) d.mocap_pos[3 * i + c] = (T)mocap[7 * i + c];
      for (int c = 0; c < 4; c++) d.mocap_quat[4 * i + c] = (T)mocap[7 * i + 3 + c];
    }
    std::vector<T> x0(ds), u0(nu), y0(ds), r0(nr), y(ds), ym(ds), dy(n), warm(nv), r(nr), rm(nr);
    auto run = [&](const std::vector<T>& x, const std::vector<T>& u, T time, std::vector<T>& ynext, std::vector<T>& rr,
                   bool do_step) {
      for (int i = 0; i < nq; i++) d.qpos[i] = x[i];
      for (int i = 0; i < nv; i++) d.qvel[i] = x[nq + i];
      for (int i = 0; i < nu; i++) d.ctrl[i] = u[i];
      d.qacc_warmstart = warm;
      d.time = time;
      d.warning = false;
      forward(m, d, cb);
      for (int i = 0; i < nr; i++) rr[i] = d.residual[i];
      if (do_step) {
        euler(m, d);
        for (int i = 0; i < nq; i++) ynext[i] = d.qpos[i];
        for (int i = 0; i < nv; i++) ynext[nq + i] = d.qvel[i];
      }
    };
    for (size_t ei = w; ei < evaluate.size(); ei += nthreads) {
      const int t = evaluate[ei];
      const bool last = t == H - 1;
      for (int i = 0; i < ds; i++) x0[i] = (T)states[(size_t)t * ds + i];
      for (int i = 0; i < nu; i++) u0[i] = (T)actions[(size_t)t * nu + i];
      const T time = (T)times[t];
      std::fill(warm.begin(), warm.end(), (T)0);
      run(x0, u0, time, y0, r0, !last);
      warm = d.qacc;  // every perturbed evaluation restarts the solver from the centre solution
      double* At = A + (size_t)t * n * n; double* Bt = B + (size_t)t * n * nu;
      double* Ct = C + (size_t)t * nr * n; double* Dt = D + (size_t)t * nr * nu;
      // one column: plus (and, centred, minus) evaluation -> state / residual differences
      auto column = [&](const std::vector<T>& xp, const std::vector<T>& up, const std::vector<T>* xm,
                        const std::vector<T>* um, T hp, double* Scol, int sw, double* Rcol, int rw, int cc, bool want_state) {
        run(xp, up, time, y, r, want_state);
        if (xm) {
          run(*xm, *um, time, ym, rm, want_state);
          if (want_state) {
            state_diff(m, dy.data(), ym.data(), y.data(), 2 * eps);
            for (int k = 0; k < n; k++) Scol[k * sw + cc] = dy[k];
          }
          for (int k = 0; k < nr; k++) Rcol[k * rw + cc] = (r[k] - rm[k]) / (2 * eps);
        } else {
          if (want_state) {
            state_diff(m, dy.data(), y0.data(), y.data(), hp);
            for (int k = 0; k < n; k++) Scol[k * sw + cc] = dy[k];
          }
          for (int k = 0; k < nr; k++) Rcol[k * rw + cc] = (r[k] - r0[k]) / hp;
        }
      };
      // controls (nudge only where the result stays inside ctrlrange)
      if (!last)
        for (int i = 0; i < nu; i++) {
          const bool limited = m.actuator_ctrllimited[i];
          const T lo = m.actuator_ctrlrange[2 * i], hi = m.actuator_ctrlrange[2 * i + 1];
          auto in_range = [&](T a, T b) { return a >= lo && a <= hi && b >= lo && b <= hi; };
          const bool fwd = !limited || in_range(u0[i], u0[i] + eps);
          const bool back = !limited || in_range(u0[i] - eps, u0[i]);
          if (!fwd && !back) continue;   // column stays zero
          std::vector<T> up = u0, um = u0;
          if (mode == 1 && fwd && back) {
            up[i] += eps; um[i] -= eps;
            column(x0, up, &x0, &um, eps, Bt, nu, Dt, nu, i, true);
          } else {
            const T h = fwd ? eps : -eps;
            up[i] += h;
            column(x0, up, nullptr, nullptr, h, Bt, nu, Dt, nu, i, true);
          }
        }
      // velocities, then positions (tangent space)
      for (int pass = 0; pass < 2; pass++)
        for (int i = 0; i < nv; i++) {
          std::vector<T> xp = x0, xm = x0;
          if (pass == 0) { xp[nq + i] += eps; xm[nq + i] -= eps; }
          else { perturb_pos(m, xp.data(), i, eps); perturb_pos(m, xm.data(), i, -eps); }
          const int cc = pass == 0 ? nv + i : i;
          if (mode == 1) column(xp, u0, &xm, &u0, eps, At, n, Ct, n, cc, !last);
          else column(xp, u0, nullptr, nullptr, eps, At, n, Ct, n, cc, !last);
        }
    }
  };
  if (nthreads <= 1) worker(0);
  else {
    std::vector<std::thread> th;
    for (int w = 0; w < nthreads; w++) th.emplace_back(worker, w);
    for (auto& x : th) x.join();
  }
  // linear interpolation of the skipped steps (model_derivatives.cc:109-164)
  for (int t : interpolate_t) {
    int b[2];
    { // FindInterval over the evaluate list (utilities.h:125-144)
      int upper = 0;
      const int len = (int)evaluate.size();
      while (upper < len && !(t < evaluate[upper])) upper++;
      const int lower = upper - 1;
      if (lower < 0) b[0] = b[1] = 0;
      else if (lower > len - 1) b[0] = b[1] = len - 1;
      else { b[0] = std::max(lower, 0); b[1] = std::min(upper, len - 1); }
    }
    const int e0 = evaluate[b[0]], e1 = evaluate[b[1]];
    const double tt = b[0] == b[1] ? 0.0 : double(t - e0) / double(e1 - e0);
    auto lerp = [&](double* X, size_t sz) {
      for (size_t k = 0; k < sz; k++) X[(size_t)t * sz + k] = (1.0 - tt) * X[(size_t)e0 * sz + k] + tt * X[(size_t)e1 * sz + k];
    };
    lerp(A, (size_t)n * n); lerp(B, (size_t)n * nu); lerp(C, (size_t)nr * n); lerp(D, (size_t)nr * nu);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------ cost derivatives
// residual[H][nr], C[H][nr][n], D[H][nr][m] -> cx[H][n], cu[H][m], cxx[H][n][n], cuu[H][m][m], cxu[H][n][m]
template <class T>
void cost_derivatives(const CostSpec<T>& cs, const T* residual, const T* C, const T* D, int H, int n, int m, T* cx,
                      T* cu, T* cxx, T* cuu, T* cxu) {
  int nr = cs.num_residual;
  for (int t = 0; t < H; t++) {
    T* Cx = cx + (size_t)t * n; T* Cu = cu + (size_t)t * m; T* Cxx = cxx + (size_t)t * n * n;
    T* Cuu = cuu + (size_t)t * m * m; T* Cxu = cxu + (size_t)t * n * m;
    std::fill(Cx, Cx + n, (T)0); std::fill(Cu, Cu + m, (T)0); std::fill(Cxx, Cxx + n * n, (T)0);
    std::fill(Cuu, Cuu + m * m, (T)0); std::fill(Cxu, Cxu + n * m, (T)0);
    int f = 0, p = 0;
    T c = 0;
    for (int i = 0; i < cs.num_term; i++) {
      int k = cs.dim_norm_residual[i];
      T w = cs.weight[i] / (T)H;
      std::vector<T> g(k), Hn(k * k), Sx(k * n), Su(k * m);
      const T* r = residual + (size_t)t * nr + f;
      const T* rx = C + (size_t)t * nr * n + (size_t)f * n;
      const T* ru = D + (size_t)t * nr * m + (size_t)f * m;
      T val = Norm<T>(g.data(), Hn.data(), r, cs.norm_parameter.data() + p, k, cs.norm[i]);
      c += w * val;
      for (int a = 0; a < n; a++) { T s = 0; for (int b = 0; b < k; b++) s += rx[b * n + a] * g[b]; Cx[a] += w * s; }
      for (int a = 0; a < m; a++) { T s = 0; for (int b = 0; b < k; b++) s += ru[b * m + a] * g[b]; Cu[a] += w * s; }
      for (int a = 0; a < k; a++)
        for (int b = 0; b < n; b++) { T s = 0; for (int q = 0; q < k; q++) s += Hn[a * k + q] * rx[q * n + b]; Sx[a * n + b] = s; }
      for (int a = 0; a < k; a++)
        for (int b = 0; b < m; b++) { T s = 0; for (int q = 0; q < k; q++) s += Hn[a * k + q] * ru[q * m + b]; Su[a * m + b] = s; }
      for (int a = 0; a < n; a++)
        for (int b = 0; b < n; b++) { T s = 0; for (int q = 0; q < k; q++) s += Sx[q * n + a] * rx[q * n + b]; Cxx[a * n + b] += w * s; }
      for (int a = 0; a < n; a++)
        for (int b = 0; b < m; b++) { T s = 0; for (int q = 0; q < k; q++) s += Sx[q * n + a] * ru[q * m + b]; Cxu[a * m + b] += w * s; }
      for (int a = 0; a < m; a++)
        for (int b = 0; b < m; b++) { T s = 0; for (int q = 0; q < k; q++) s += Su[q * m + a] * ru[q * m + b]; Cuu[a * m + b] += w * s; }
      f += k;
      p += cs.num_norm_parameter[i];
    }
    if (mm::fabs(cs.risk) < (T)kRiskNeutralTolerance) continue;
    T s = mm::exp(cs.risk * c);
    for (int a = 0; a < n; a++) Cx[a] *= s;
    for (int a = 0; a < m; a++) Cu[a] *= s;
    // note the order: the reference scales cx/cu first and then uses the *scaled* gradients in the outer products
    for (int a = 0; a < n; a++) for (int b = 0; b < n; b++) Cxx[a * n + b] = Cxx[a * n + b] * s + cs.risk * s * Cx[a] * Cx[b];
    for (int a = 0; a < n; a++) for (int b = 0; b < m; b++) Cxu[a * m + b] = Cxu[a * m + b] * s + cs.risk * s * Cx[a] * Cu[b];
    for (int a = 0; a < m; a++) for (int b = 0; b < m; b++) Cuu[a * m + b] = Cuu[a * m + b] * s + cs.risk * s * Cu[a] * Cu[b];
  }
}

// ------------------------------------------------------------------------------------------ box QP
// min 0.5 x'Hx + g'x, lower <= x <= upper. Returns number of free dims (R = Cholesky of the free block,
// row-major nfree x nfree; index = free dims) or -1 if the free Hessian is not positive definite.
template <class T>
int box_qp(T* res, T* R, int* index, const T* Hm, const T* g, int n, const T* lower, const T* upper) {
  const int maxiter = 100;
  const T mingrad = (T)1e-16, backtrack = (T)0.5, minstep = (T)1e-22, armijo = (T)0.01;
  std::vector<T> grad(n), search(n), cand(n), tmp(n), rhs(n), sol(n);
  std::vector<int> clamped(n, 0), oldclamped(n, 0);
  auto value_of = [&](const T* x) {
    T v = 0;
    for (int i = 0; i < n; i++) { T a = 0; for (int j = 0; j < n; j++) a += Hm[i * n + j] * x[j]; v += x[i] * ((T)0.5 * a + g[i]); }
    return v;
  };
  for (int i = 0; i < n; i++) res[i] = mm::max(lower[i], mm::min(upper[i], res[i]));
  T value = value_of(res);
  int nfree = 0;
  for (int iter = 0; iter < maxiter; iter++) {
    for (int i = 0; i < n; i++) { T a = g[i]; for (int j = 0; j < n; j++) a += Hm[i * n + j] * res[j]; grad[i] = a; }
    oldclamped = clamped;
    bool changed = iter == 0;
    nfree = 0;
    for (int i = 0; i < n; i++) {
      clamped[i] = (res[i] == lower[i] && grad[i] > 0) || (res[i] == upper[i] && grad[i] < 0);
      if (clamped[i] != oldclamped[i]) changed = true;
      if (!clamped[i]) index[nfree++] = i;
    }
    if (nfree == 0) break;
    if (changed) {
      for (int a = 0; a < nfree; a++)
        for (int b = 0; b < nfree; b++) R[a * nfree + b] = Hm[index[a] * n + index[b]];
      T minp = chol_factor(R, nfree);
      if (!(minp > kMinVal<T>())) return -1;
    }
    T norm2 = 0;
    for (int a = 0; a < nfree; a++) norm2 += grad[index[a]] * grad[index[a]];
    if (norm2 < mingrad * mingrad) break;
    for (int i = 0; i < n; i++) tmp[i] = clamped[i] ? res[i] : (T)0;
    for (int a = 0; a < nfree; a++) {
      int i = index[a];
      T s = g[i];
      for (int j = 0; j < n; j++) s += Hm[i * n + j] * tmp[j];
      rhs[a] = s;
    }
    chol_solve(sol.data(), R, rhs.data(), nfree);
    std::fill(search.begin(), search.end(), (T)0);
    for (int a = 0; a < nfree; a++) search[index[a]] = -sol[a] - res[index[a]];
    T sdotg = 0;
    for (int i = 0; i < n; i++) sdotg += search[i] * grad[i];
    if (sdotg >= 0) break;
    T step = 1, vc = value;
    bool accepted = false;
    while (step > minstep) {
      for (int i = 0; i < n; i++) cand[i] = mm::max(lower[i], mm::min(upper[i], res[i] + step * search[i]));
      vc = value_of(cand.data());
      if ((vc - value) / (step * sdotg) >= armijo) { accepted = true; break; }
      step *= backtrack;
    }
    if (!accepted) break;
    for (int i = 0; i < n; i++) res[i] = cand[i];
    value = vc;
  }
  return nfree;
}

// ------------------------------------------------------------------------------------------ Riccati backward pass
// One pass at fixed regularisation mu. Returns 1 on success, 0 on failure (caller scales mu and retries:
// mjpc/planners/ilqg/planner.cc:429-520). Outputs time-major row-major: Vx[H][n], Vxx[H][n][n], du[H][m],
// K[H][m][n], Qx[H][n], Qu[H][m], Qxx[H][n][n], Qxu[H][n][m], Quu[H][m][m], dV[2].
template <class T>
int backward_pass(const T* A, const T* B, const T* cx, const T* cu, const T* cxx, const T* cxu, const T* cuu,
                  const T* actions, const T* ctrlrange, int n, int m, int H, T mu, int reg_type, int limits, T* Vx,
                  T* Vxx, T* du, T* K, T* dV, T* Qx, T* Qu, T* Qxx, T* Qxu, T* Quu) {
  dV[0] = dV[1] = 0;
  for (int i = 0; i < n; i++) Vx[(size_t)(H - 1) * n + i] = cx[(size_t)(H - 1) * n + i];
  for (int i = 0; i < n * n; i++) Vxx[(size_t)(H - 1) * n * n + i] = cxx[(size_t)(H - 1) * n * n + i];
  std::vector<T> qp_res(m, 0), qp_R(m * m), qp_lower(m), qp_upper(m), Vreg(n * n), Qxu_reg(n * m), Quu_reg(m * m),
      tmp(n * n), tmp2(mm::max(n, m) * mm::max(n, m)), L(m * m), sol(m), rhs(m), Qd(m);
  std::vector<int> qp_index(m);
  auto AtW = [&](T* out, const T* At, const T* W) {  // out[n][n] = At' * W
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) { T s = 0; for (int k = 0; k < n; k++) s += At[k * n + i] * W[k * n + j]; out[i * n + j] = s; }
  };
  for (int t = H - 2; t >= 0; t--) {
    const T* At = A + (size_t)t * n * n; const T* Bt = B + (size_t)t * n * m;
    const T* Wx = Vx + (size_t)(t + 1) * n; const T* Wxx = Vxx + (size_t)(t + 1) * n * n;
    T* Qxt = Qx + (size_t)t * n; T* Qut = Qu + (size_t)t * m; T* Qxxt = Qxx + (size_t)t * n * n;
    T* Qxut = Qxu + (size_t)t * n * m; T* Quut = Quu + (size_t)t * m * m;
    T* Kt = K + (size_t)t * m * n; T* dut = du + (size_t)t * m;
    T* Vxt = Vx + (size_t)t * n; T* Vxxt = Vxx + (size_t)t * n * n;
    AtW(tmp.data(), At, Wxx);
    for (int i = 0; i < n; i++) { T s = cx[(size_t)t * n + i]; for (int k = 0; k < n; k++) s += At[k * n + i] * Wx[k]; Qxt[i] = s; }
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) { T s = cxx[(size_t)t * n * n + i * n + j]; for (int k = 0; k < n; k++) s += tmp[i * n + k] * At[k * n + j]; Qxxt[i * n + j] = s; }
    for (int i = 0; i < m; i++) { T s = cu[(size_t)t * m + i]; for (int k = 0; k < n; k++) s += Bt[k * m + i] * Wx[k]; Qut[i] = s; }
    for (int i = 0; i < n; i++)
      for (int j = 0; j < m; j++) { T s = cxu[(size_t)t * n * m + i * m + j]; for (int k = 0; k < n; k++) s += tmp[i * n + k] * Bt[k * m + j]; Qxut[i * m + j] = s; }
    auto BtWB = [&](T* out, const T* W, const T* base) {  // out[m][m] = base + Bt' W Bt
      for (int i = 0; i < m; i++)
        for (int j = 0; j < n; j++) { T s = 0; for (int k = 0; k < n; k++) s += Bt[k * m + i] * W[k * n + j]; tmp2[i * n + j] = s; }
      for (int i = 0; i < m; i++)
        for (int j = 0; j < m; j++) { T s = base[i * m + j]; for (int k = 0; k < n; k++) s += tmp2[i * n + k] * Bt[k * m + j]; out[i * m + j] = s; }
    };
    BtWB(Quut, Wxx, cuu + (size_t)t * m * m);
    // regularisation: 0 control, 1 state-control (feedback), 2 value, 3 none (ilqg/settings.h:30)
    if (reg_type == 2) {
      for (int i = 0; i < n * n; i++) Vreg[i] = Wxx[i];
      for (int i = 0; i < n; i++) Vreg[i * n + i] += mu;
      AtW(tmp.data(), At, Vreg.data());
      for (int i = 0; i < n; i++)
        for (int j = 0; j < m; j++) { T s = cxu[(size_t)t * n * m + i * m + j]; for (int k = 0; k < n; k++) s += tmp[i * n + k] * Bt[k * m + j]; Qxu_reg[i * m + j] = s; }
      BtWB(Quu_reg.data(), Vreg.data(), cuu + (size_t)t * m * m);
    } else {
      for (int i = 0; i < n * m; i++) Qxu_reg[i] = Qxut[i];
      for (int i = 0; i < m * m; i++) Quu_reg[i] = Quut[i];
    }
    if (mu != 0) {
      if (reg_type == 0) {
        for (int i = 0; i < m; i++) Quu_reg[i * m + i] += mu;
      } else if (reg_type == 1) {
        for (int i = 0; i < n; i++)
          for (int j = 0; j < m; j++) { T s = 0; for (int k = 0; k < n; k++) s += At[k * n + i] * Bt[k * m + j]; Qxu_reg[i * m + j] += mu * s; }
        for (int i = 0; i < m; i++)
          for (int j = 0; j < m; j++) { T s = 0; for (int k = 0; k < n; k++) s += Bt[k * m + i] * Bt[k * m + j]; Quu_reg[i * m + j] += mu * s; }
      }
    }
    for (int i = 0; i < m * n; i++) Kt[i] = 0;
    if (limits == 1) {
      for (int i = 0; i < m; i++) {
        qp_lower[i] = ctrlrange[2 * i] - actions[(size_t)t * m + i];
        qp_upper[i] = ctrlrange[2 * i + 1] - actions[(size_t)t * m + i];
      }
      int mf = box_qp(qp_res.data(), qp_R.data(), qp_index.data(), Quu_reg.data(), Qut, m, qp_lower.data(), qp_upper.data());
      if (mf < 0) return 0;
      // K on free dims: -H_free^-1 Qxu_free' (the reference uses the unregularised Qxu here; backward_pass.cc:176-192)
      for (int j = 0; j < n; j++) {
        for (int i = 0; i < mf; i++) rhs[i] = Qxut[j * m + qp_index[i]];
        chol_solve(sol.data(), qp_R.data(), rhs.data(), mf);
        for (int i = 0; i < mf; i++) Kt[qp_index[i] * n + j] = -sol[i];
      }
      for (int i = 0; i < m; i++) dut[i] = qp_res[i];
    } else {
      for (int i = 0; i < m * m; i++) L[i] = Quu_reg[i];
      T minp = chol_factor(L.data(), m);
      if (!(minp > kMinVal<T>())) return 0;
      for (int j = 0; j < n; j++) {
        for (int i = 0; i < m; i++) rhs[i] = Qxut[j * m + i];
        chol_solve(sol.data(), L.data(), rhs.data(), m);
        for (int i = 0; i < m; i++) Kt[i * n + j] = -sol[i];
      }
      chol_solve(sol.data(), L.data(), Qut, m);
      for (int i = 0; i < m; i++) dut[i] = -sol[i];
    }
    // cost-to-go update
    for (int i = 0; i < m; i++) { T s = 0; for (int j = 0; j < m; j++) s += Quut[i * m + j] * dut[j]; Qd[i] = s; }
    for (int i = 0; i < m; i++) { dV[0] += dut[i] * Qut[i]; dV[1] += (T)0.5 * dut[i] * Qd[i]; }
    for (int i = 0; i < n; i++) {
      T s = Qxt[i];
      for (int k = 0; k < m; k++) s += Kt[k * n + i] * (Qd[k] + Qut[k]) + Qxut[i * m + k] * dut[k];
      Vxt[i] = s;
    }
    // Vxx = Qxx + K' Quu K + Qxu K + K' Qxu', symmetrised
    for (int i = 0; i < m; i++)
      for (int j = 0; j < n; j++) { T s = 0; for (int k = 0; k < m; k++) s += Quut[i * m + k] * Kt[k * n + j]; tmp2[i * n + j] = s; }
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        T s = Qxxt[i * n + j];
        for (int k = 0; k < m; k++) s += Kt[k * n + i] * tmp2[k * n + j] + Qxut[i * m + k] * Kt[k * n + j] + Kt[k * n + i] * Qxut[j * m + k];
        tmp[i * n + j] = s;
      }
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) Vxxt[i * n + j] = (T)0.5 * (tmp[i * n + j] + tmp[j * n + i]);
  }
  if (H >= 2) {
    for (int i = 0; i < m * n; i++) K[(size_t)(H - 1) * m * n + i] = K[(size_t)(H - 2) * m * n + i];
    for (int i = 0; i < m; i++) du[(size_t)(H - 1) * m + i] = du[(size_t)(H - 2) * m + i];
  }
  return 1;
}

}  // namespace oracle
