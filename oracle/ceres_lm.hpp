// ORACLE — TEST INFRASTRUCTURE ONLY (see types.hpp header).  PARITY UNPINNED.
//
// ceres_lm.hpp: restatement of what `ceres::Solve` does for the reference's
// three call sites (laser_odometry.cpp:217-221,257-258,458-463;
// laser_mapping.cpp:461-467,610-617; visual_odometry.cpp:67-68,258-259,423).
// Third party: Ceres Solver 2.0.0 (README.md:24; not vendored, not installed
// here) — the published algorithm of its TrustRegionMinimizer +
// LevenbergMarquardtStrategy + DenseQRSolver with default options is restated:
//
//   * residual blocks evaluated with their loss; HuberLoss(a): rho(s)=s for
//     s<=a^2 else 2a*sqrt(s)-a^2; cost = 1/2 sum rho(s); because rho''<=0 the
//     Corrector only scales residual and Jacobian block by sqrt(rho').
//   * EigenQuaternionParameterization on (x,y,z,w): Plus(q,d) =
//     [sin|d|/|d| d, cos|d|] (x) q; ComputeJacobian rows
//     [w z -y; -z w x; y -x w; -x -y -z]; local Jacobian = global * that.
//   * Jacobi scaling 1/(1+||col||) fixed at iteration 0.
//   * LM: D = sqrt(clamp(diag(J'J),1e-6,1e32)/radius); min ||J y - r||^2+||D y||^2
//     by dense QR; step=-y; model_cost_change = -(J step).(r + J step/2);
//     rho = (cost - cost_new)/model_cost_change; accept iff rho > 1e-3;
//     accepted: radius /= max(1/3, 1-(2rho-1)^3) (capped 1e16), decrease_factor=2;
//     rejected: radius /= decrease_factor, decrease_factor *= 2, diagonal reused.
//   * loop order per iteration: step -> (invalid step handling) -> candidate
//     cost -> parameter tolerance -> function tolerance -> accept/reject;
//     failed steps count toward max_num_iterations.
//   * tolerances: function 1e-6, gradient 1e-10, parameter 1e-8; radius
//     initial 1e4, min 1e-32.
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>
#include <vector>

namespace oracle {

// One residual block: up to 3 residuals, ambient parameter size 7 (q xyzw + t)
// or 6 (angle-axis + t).  J is nres x ambient, row-major.
struct CostBlock {
  virtual ~CostBlock() {}
  virtual int num_residuals() const = 0;
  virtual void evaluate(const double* x, double* r, double* J) const = 0;
};

struct LMOptions {
  int max_num_iterations = 4;
  bool quaternion_manifold = true;  // true: x = [q(4), t(3)]; false: x = [aa(3), t(3)]
  bool use_huber = true;
  double huber_a = 0.1;
  double initial_trust_region_radius = 1e4;
  double max_trust_region_radius = 1e16;
  double min_trust_region_radius = 1e-32;
  double min_relative_decrease = 1e-3;
  double min_lm_diagonal = 1e-6;
  double max_lm_diagonal = 1e32;
  double function_tolerance = 1e-6;
  double gradient_tolerance = 1e-10;
  double parameter_tolerance = 1e-8;
  int max_num_consecutive_invalid_steps = 5;
};

struct LMIteration {
  int iteration = 0;
  double cost = 0;           // cost at the accepted point after this iteration
  double candidate_cost = 0;
  double model_cost_change = 0;
  double relative_decrease = 0;
  double radius = 0;
  int step_is_valid = 0;
  int step_is_successful = 0;
};

enum LMTermination { LM_NO_CONVERGENCE = 0, LM_CONVERGENCE_GRADIENT = 1, LM_CONVERGENCE_PARAMETER = 2,
                     LM_CONVERGENCE_FUNCTION = 3, LM_FAILURE = 4 };

struct LMSummary {
  std::vector<LMIteration> iterations;
  int termination = LM_NO_CONVERGENCE;
  double initial_cost = 0, final_cost = 0;
  int num_successful_steps = 0;
};

namespace detail {

inline void quat_plus(const double* x, const double* d, double* out) {
  // EigenQuaternionParameterization::Plus
  const double nd = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  if (nd > 0.0) {
    const double s = std::sin(nd) / nd;
    const double qx = s * d[0], qy = s * d[1], qz = s * d[2], qw = std::cos(nd);
    // delta_q * x  (Eigen product, coefficients x,y,z,w)
    const double ax = x[0], ay = x[1], az = x[2], aw = x[3];
    out[0] = qw * ax + qx * aw + qy * az - qz * ay;
    out[1] = qw * ay + qy * aw + qz * ax - qx * az;
    out[2] = qw * az + qz * aw + qx * ay - qy * ax;
    out[3] = qw * aw - qx * ax - qy * ay - qz * az;
  } else {
    out[0] = x[0]; out[1] = x[1]; out[2] = x[2]; out[3] = x[3];
  }
}

inline void plus(const LMOptions& o, const double* x, const double* d, double* out) {
  if (o.quaternion_manifold) {
    quat_plus(x, d, out);
    out[4] = x[4] + d[3]; out[5] = x[5] + d[4]; out[6] = x[6] + d[5];
  } else {
    for (int i = 0; i < 6; ++i) out[i] = x[i] + d[i];
  }
}

// Dense evaluation.  Returns cost; fills residuals (R) and Jacobian (R x 6,
// row-major, already in the local/tangent parameterisation and loss-corrected)
// when `jac` is non-null.
inline double evaluate(const std::vector<const CostBlock*>& blocks, const LMOptions& o, const double* x,
                       std::vector<double>* res, std::vector<double>* jac) {
  const int amb = o.quaternion_manifold ? 7 : 6;
  double cost = 0.0;
  size_t row = 0;
  if (res) res->clear();
  if (jac) jac->clear();
  for (const CostBlock* b : blocks) {
    const int nr = b->num_residuals();
    double r[3];
    double Jg[3 * 7];
    b->evaluate(x, r, jac ? Jg : nullptr);
    double Jl[3 * 6];
    if (jac) {
      if (o.quaternion_manifold) {
        // global (nr x 4) * plus-Jacobian (4 x 3)
        const double qx = x[0], qy = x[1], qz = x[2], qw = x[3];
        const double P[12] = {qw, qz, -qy, -qz, qw, qx, qy, -qx, qw, -qx, -qy, -qz};
        for (int i = 0; i < nr; ++i) {
          for (int c = 0; c < 3; ++c) {
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s += Jg[i * 7 + k] * P[k * 3 + c];
            Jl[i * 6 + c] = s;
          }
          for (int c = 0; c < 3; ++c) Jl[i * 6 + 3 + c] = Jg[i * 7 + 4 + c];
        }
      } else {
        for (int i = 0; i < nr * 6; ++i) Jl[i] = Jg[i];
      }
    }
    double sq = 0.0;
    for (int i = 0; i < nr; ++i) sq += r[i] * r[i];
    double rho0 = sq, rho1 = 1.0;
    if (o.use_huber) {
      const double a = o.huber_a, bb = a * a;
      if (sq > bb) {
        const double rr = std::sqrt(sq);
        rho0 = 2.0 * a * rr - bb;
        rho1 = std::max(std::numeric_limits<double>::min(), a / rr);
      }
    }
    cost += 0.5 * rho0;
    const double sc = std::sqrt(rho1);
    for (int i = 0; i < nr; ++i) {
      if (res) res->push_back(r[i] * sc);
      if (jac) for (int c = 0; c < 6; ++c) jac->push_back(Jl[i * 6 + c] * sc);
    }
    row += nr;
  }
  return cost;
}

// Least squares min ||A y - b|| for A (m x 6, row-major, overwritten) by Householder QR.
inline bool householder_qr_solve(std::vector<double>& A, std::vector<double>& b, int m, double* y) {
  const int n = 6;
  for (int k = 0; k < n; ++k) {
    double nrm = 0.0;
    for (int i = k; i < m; ++i) nrm += A[i * n + k] * A[i * n + k];
    nrm = std::sqrt(nrm);
    if (nrm == 0.0) return false;
    const double alpha = A[k * n + k] > 0 ? -nrm : nrm;
    // v = x - alpha e1
    std::vector<double> v(m - k);
    for (int i = k; i < m; ++i) v[i - k] = A[i * n + k];
    v[0] -= alpha;
    double vn = 0.0;
    for (double t : v) vn += t * t;
    if (vn == 0.0) continue;
    for (int c = k; c < n; ++c) {
      double s = 0.0;
      for (int i = k; i < m; ++i) s += v[i - k] * A[i * n + c];
      s = 2.0 * s / vn;
      for (int i = k; i < m; ++i) A[i * n + c] -= s * v[i - k];
    }
    double s = 0.0;
    for (int i = k; i < m; ++i) s += v[i - k] * b[i];
    s = 2.0 * s / vn;
    for (int i = k; i < m; ++i) b[i] -= s * v[i - k];
  }
  for (int k = n - 1; k >= 0; --k) {
    double s = b[k];
    for (int c = k + 1; c < n; ++c) s -= A[k * n + c] * y[c];
    if (A[k * n + k] == 0.0) return false;
    y[k] = s / A[k * n + k];
  }
  for (int k = 0; k < n; ++k) if (!std::isfinite(y[k])) return false;
  return true;
}

}  // namespace detail

inline void lm_solve(const std::vector<const CostBlock*>& blocks, const LMOptions& o, double* x, LMSummary* sum) {
  using namespace detail;
  const int amb = o.quaternion_manifold ? 7 : 6;
  LMSummary local;
  LMSummary& S = sum ? *sum : local;
  S = LMSummary();

  std::vector<double> r, J;
  double x_cost = evaluate(blocks, o, x, &r, &J);
  const int R = (int)r.size();
  S.initial_cost = x_cost;
  S.final_cost = x_cost;
  {
    LMIteration it0; it0.iteration = 0; it0.cost = x_cost; it0.radius = o.initial_trust_region_radius;
    S.iterations.push_back(it0);
  }
  if (R == 0) { S.termination = LM_CONVERGENCE_GRADIENT; return; }

  // gradient (unscaled) and Jacobi scaling at iteration 0
  double scale[6];
  auto gradient_max_norm = [&]() {
    double g = 0.0;
    for (int c = 0; c < 6; ++c) {
      double s = 0.0;
      for (int i = 0; i < R; ++i) s += J[i * 6 + c] * r[i];
      g = std::max(g, std::fabs(s));
    }
    return g;
  };
  double gmax = gradient_max_norm();
  for (int c = 0; c < 6; ++c) {
    double s = 0.0;
    for (int i = 0; i < R; ++i) s += J[i * 6 + c] * J[i * 6 + c];
    scale[c] = 1.0 / (1.0 + std::sqrt(s));
  }
  auto scale_columns = [&]() {
    for (int i = 0; i < R; ++i) for (int c = 0; c < 6; ++c) J[i * 6 + c] *= scale[c];
  };
  scale_columns();
  if (gmax <= o.gradient_tolerance) { S.termination = LM_CONVERGENCE_GRADIENT; return; }

  double radius = o.initial_trust_region_radius;
  double decrease_factor = 2.0;
  bool reuse_diagonal = false;
  double diagonal[6];
  int num_consecutive_invalid = 0;
  double x_norm = 0.0;
  for (int i = 0; i < amb; ++i) x_norm += x[i] * x[i];
  x_norm = std::sqrt(x_norm);

  int iteration = 0;
  while (true) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue (checks on the summary of the previous iteration)
    if (iteration >= o.max_num_iterations) { S.termination = LM_NO_CONVERGENCE; break; }
    if (gmax <= o.gradient_tolerance) { S.termination = LM_CONVERGENCE_GRADIENT; break; }
    if (radius <= o.min_trust_region_radius) { S.termination = LM_CONVERGENCE_PARAMETER; break; }
    ++iteration;
    LMIteration it; it.iteration = iteration;

    // --- LevenbergMarquardtStrategy::ComputeStep
    if (!reuse_diagonal) {
      for (int c = 0; c < 6; ++c) {
        double s = 0.0;
        for (int i = 0; i < R; ++i) s += J[i * 6 + c] * J[i * 6 + c];
        diagonal[c] = std::min(std::max(s, o.min_lm_diagonal), o.max_lm_diagonal);
      }
    }
    double D[6];
    for (int c = 0; c < 6; ++c) D[c] = std::sqrt(diagonal[c] / radius);
    std::vector<double> A((size_t)(R + 6) * 6, 0.0), b((size_t)R + 6, 0.0);
    for (int i = 0; i < R * 6; ++i) A[i] = J[i];
    for (int c = 0; c < 6; ++c) A[(size_t)(R + c) * 6 + c] = D[c];
    for (int i = 0; i < R; ++i) b[i] = r[i];
    double step[6];
    bool ok = householder_qr_solve(A, b, R + 6, step);
    reuse_diagonal = true;
    double model_cost_change = 0.0;
    if (ok) {
      for (int c = 0; c < 6; ++c) step[c] = -step[c];
      // model_cost_change = -(J step).(r + J step / 2)
      for (int i = 0; i < R; ++i) {
        double m = 0.0;
        for (int c = 0; c < 6; ++c) m += J[i * 6 + c] * step[c];
        model_cost_change -= m * (r[i] + m / 2.0);
      }
    }
    it.model_cost_change = model_cost_change;
    it.step_is_valid = ok && model_cost_change > 0.0;
    if (!it.step_is_valid) {
      // HandleInvalidStep
      ++num_consecutive_invalid;
      it.cost = x_cost;
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
      it.radius = radius;
      S.iterations.push_back(it);
      if (num_consecutive_invalid >= o.max_num_consecutive_invalid_steps) { S.termination = LM_FAILURE; break; }
      continue;
    }
    num_consecutive_invalid = 0;
    double delta[6];
    for (int c = 0; c < 6; ++c) delta[c] = step[c] * scale[c];
    double cand[7];
    plus(o, x, delta, cand);
    const double cand_cost = evaluate(blocks, o, cand, nullptr, nullptr);
    it.candidate_cost = cand_cost;

    // ParameterToleranceReached
    double step_norm = 0.0;
    for (int i = 0; i < amb; ++i) step_norm += (x[i] - cand[i]) * (x[i] - cand[i]);
    step_norm = std::sqrt(step_norm);
    if (step_norm <= o.parameter_tolerance * (x_norm + o.parameter_tolerance)) {
      it.cost = x_cost; it.radius = radius; S.iterations.push_back(it);
      S.termination = LM_CONVERGENCE_PARAMETER; break;
    }
    // FunctionToleranceReached
    const double cost_change = x_cost - cand_cost;
    if (std::fabs(cost_change) <= o.function_tolerance * x_cost) {
      it.cost = x_cost; it.radius = radius; S.iterations.push_back(it);
      S.termination = LM_CONVERGENCE_FUNCTION; break;
    }
    const double rel = cost_change / model_cost_change;
    it.relative_decrease = rel;
    if (rel > o.min_relative_decrease) {
      // HandleSuccessfulStep
      for (int i = 0; i < amb; ++i) x[i] = cand[i];
      x_norm = 0.0;
      for (int i = 0; i < amb; ++i) x_norm += x[i] * x[i];
      x_norm = std::sqrt(x_norm);
      x_cost = evaluate(blocks, o, x, &r, &J);
      gmax = gradient_max_norm();
      scale_columns();
      it.step_is_successful = 1;
      ++S.num_successful_steps;
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rel - 1.0, 3));
      radius = std::min(o.max_trust_region_radius, radius);
      decrease_factor = 2.0;
      reuse_diagonal = false;
    } else {
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
      reuse_diagonal = true;
    }
    it.cost = x_cost; it.radius = radius;
    S.iterations.push_back(it);
  }
  S.final_cost = x_cost;
}

}  // namespace oracle
