// TEST INFRASTRUCTURE - CPU oracle (see oracle/model.h header).
// task.h: norms, cost terms, risk transform and the time-indexed spline policy.
//   Norm            <- mjpc/norm.cc:50-210   (pinned by tests/test_oracle_golden.py vs norm_test.cc:42-109)
//   CostTerms/Value <- mjpc/task.cc:71-110   (pinned vs task_test.cc:77-95)
//   spline sample   <- mjpc/spline/spline.cc:103-156, 250-287 (pinned vs spline_test.cc:115-158)
//   policy clamp    <- mjpc/planners/sampling/policy.cc:52-59, mjpc/utilities.cc:112-116
#pragma once
#include <cmath>
#include <vector>

#include "counted.h"
#include "model.h"

namespace oracle {

enum NormType { kNull = -1, kQuadratic = 0, kL22 = 1, kL2 = 2, kCosh = 3, kPowerLoss = 5, kSmoothAbsLoss = 6,
                kSmoothAbs2Loss = 7, kRectifyLoss = 8 };
constexpr double kRiskNeutralTolerance = 1e-6;  // mjpc/task.h

// value; optional gradient g[n] and Hessian H[n*n]
template <class T>
T Norm(T* g, T* H, const T* x, const T* params, int n, int type) {
  T y = 0;
  T p = params ? params[0] : 0, q = params ? params[1] : 0;
  if (H) for (int i = 0; i < n * n; i++) H[i] = 0;
  switch (type) {
    case kNull:
      y = x[0];
      if (g) g[0] = 1;
      if (H) H[0] = 0;
      break;
    case kQuadratic:
      for (int i = 0; i < n; i++) y += x[i] * x[i];
      y *= (T)0.5;
      if (g) for (int i = 0; i < n; i++) g[i] = x[i];
      if (H) for (int i = 0; i < n; i++) H[i * n + i] = 1;
      break;
    case kL22: {
      T c = 0;
      for (int i = 0; i < n; i++) c += x[i] * x[i];
      T a = mm::pow(c, q / 2) + mm::pow(p, q);
      T s = mm::pow(a, 1 / q);
      y = s - p;
      T dd = mm::pow(c, q / 2 - 1);
      T b = s / a * dd;
      if (g) for (int i = 0; i < n; i++) g[i] = b * x[i];
      if (H) {
        c = (1 - q) * dd / a + (q - 2) / mm::max(c, (T)1e-15);
        for (int i = 0; i < n; i++)
          for (int j = 0; j < n; j++) H[i + j * n] = b * ((i == j ? (T)1 : (T)0) + x[i] * x[j] * c);
      }
      break;
    }
    case kL2: {
      T dsum = 0;
      for (int i = 0; i < n; i++) dsum += x[i] * x[i];
      T s = mm::sqrt(dsum + p * p);
      y = s - p;
      if (g) for (int i = 0; i < n; i++) g[i] = s ? x[i] / s : (T)0;
      if (H && s)
        for (int i = 0; i < n; i++)
          for (int j = 0; j < n; j++) H[i + j * n] = ((i == j ? 1 : 0) - g[i] * g[j]) / s;
      break;
    }
    case kCosh:
      for (int i = 0; i < n; i++) {
        y += p * p * (mm::cosh(x[i] / p) - 1);
        if (g) g[i] = p * mm::sinh(x[i] / p);
        if (H) H[i * n + i] = mm::cosh(x[i] / p);
      }
      break;
    case kPowerLoss:
      for (int i = 0; i < n; i++) {
        T s = mm::fabs(x[i]);
        y += mm::pow(s, p);
        if (g) g[i] = (x[i] > 0 ? 1 : (x[i] < 0 ? -1 : 0)) * p * mm::pow(s, p - 1);
        if (H) H[i * n + i] = (p - 1) * p * mm::pow(s, p - 2);
      }
      break;
    case kSmoothAbsLoss:
      for (int i = 0; i < n; i++) {
        T s = mm::sqrt(x[i] * x[i] + p * p);
        y += s - p;
        if (g) g[i] = s ? x[i] / s : (T)0;
        if (H) H[n * i + i] = s ? (1 - g[i] * g[i]) / s : (T)0;
      }
      break;
    case kSmoothAbs2Loss:
      for (int i = 0; i < n; i++) {
        T a = mm::fabs(x[i]);
        T dd = mm::pow(a, q);
        T e = dd + mm::pow(p, q);
        T s = mm::pow(e, 1 / q);
        y += s - p;
        T c = s * mm::pow(a, q - 2) / e;
        if (g) g[i] = c * x[i];
        if (H) H[i * n + i] = c * (q - 1) * (1 - dd / e);
      }
      break;
    case kRectifyLoss:
      for (int i = 0; i < n; i++) {
        if (p > 0) {
          T s = mm::exp(x[i] / p);
          y += p * mm::log(1 + s);
          if (g) g[i] = s / (1 + s);
          if (H) H[i * n + i] = s / (p * (1 + s) * (1 + s));
        } else {
          y += x[i] > 0 ? x[i] : 0;
          if (g) g[i] = x[i] > 0 ? 1 : 0;
        }
      }
      break;
  }
  return y;
}

// live cost specification (what mjpc::BaseResidualFn::Update() snapshots: task.cc:112-123)
template <class T>
struct CostSpec {
  int num_term = 0, num_residual = 0;
  std::vector<int> dim_norm_residual, norm, num_norm_parameter;
  std::vector<T> weight, norm_parameter;
  T risk = 0;
  CostSpec() = default;
  explicit CostSpec(const Model<T>& m)
      : num_term(m.num_term), num_residual(m.num_residual), dim_norm_residual(m.dim_norm_residual), norm(m.norm),
        num_norm_parameter(m.num_norm_parameter), weight(m.weight), norm_parameter(m.norm_parameter), risk(m.risk) {}
};

template <class T>
void CostTerms(const CostSpec<T>& c, T* terms, const T* residual, bool weighted) {
  int f = 0, p = 0;
  for (int k = 0; k < c.num_term; k++) {
    terms[k] = (weighted ? c.weight[k] : (T)1) *
               Norm<T>(nullptr, nullptr, residual + f, c.norm_parameter.data() + p, c.dim_norm_residual[k], c.norm[k]);
    f += c.dim_norm_residual[k];
    p += c.num_norm_parameter[k];
  }
}

template <class T>
T CostValue(const CostSpec<T>& c, const T* residual) {
  std::vector<T> terms(mm::max(c.num_term, 1));
  CostTerms(c, terms.data(), residual, true);
  T cost = 0;
  for (int i = 0; i < c.num_term; i++) cost += terms[i];
  if (mm::fabs(c.risk) < (T)kRiskNeutralTolerance) return cost;
  return (mm::exp(c.risk * cost) - 1) / c.risk;
}

// ---- time spline (zero / linear / cubic-Hermite with finite-difference slopes)
enum { kZeroSpline = 0, kLinearSpline = 1, kCubicSpline = 2 };

template <class T>
T spline_slope(const T* times, const T* values, int P, int dim, int node, int k) {
  auto v = [&](int n) { return values[n * dim + k]; };
  if (node == 0) return (v(1) - v(0)) / (times[1] - times[0]);
  if (node == P - 1) return (v(node) - v(node - 1)) / (times[node] - times[node - 1]);
  return (T)0.5 * (v(node + 1) - v(node)) / (times[node + 1] - times[node]) +
         (T)0.5 * (v(node) - v(node - 1)) / (times[node] - times[node - 1]);
}

template <class T>
void spline_sample(T* out, const T* times, const T* values, int P, int dim, int interp, T time) {
  if (P == 0) { for (int i = 0; i < dim; i++) out[i] = 0; return; }
  int upper = 0;  // std::upper_bound: first node with time_node > time
  while (upper < P && !(time < times[upper])) upper++;
  if (upper == P) { for (int i = 0; i < dim; i++) out[i] = values[(P - 1) * dim + i]; return; }
  if (upper == 0) { for (int i = 0; i < dim; i++) out[i] = values[i]; return; }
  int lower = upper - 1;
  T t = (time - times[lower]) / (times[upper] - times[lower]);
  if (interp == kZeroSpline) {
    for (int i = 0; i < dim; i++) out[i] = values[lower * dim + i];
  } else if (interp == kLinearSpline) {
    for (int i = 0; i < dim; i++) out[i] = values[lower * dim + i] * (1 - t) + values[upper * dim + i] * t;
  } else {
    T dt = times[upper] - times[lower];
    T c0 = 2 * t * t * t - 3 * t * t + 1, c1 = (t * t * t - 2 * t * t + t) * dt, c2 = -2 * t * t * t + 3 * t * t,
      c3 = (t * t * t - t * t) * dt;
    for (int i = 0; i < dim; i++) {
      T p0 = values[lower * dim + i], p1 = values[upper * dim + i];
      T m0 = spline_slope(times, values, P, dim, lower, i), m1 = spline_slope(times, values, P, dim, upper, i);
      out[i] = c0 * p0 + c1 * m0 + c2 * p1 + c3 * m1;
    }
  }
}

template <class T>
void clamp_ctrl(T* x, const T* bounds, int n) {
  for (int i = 0; i < n; i++) x[i] = mm::max(bounds[2 * i], mm::min(bounds[2 * i + 1], x[i]));
}

}  // namespace oracle
