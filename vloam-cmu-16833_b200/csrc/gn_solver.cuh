// vloam_b200 — Gauss-Newton / Levenberg-Marquardt machinery shared by the laser odometry, laser mapping and visual
// odometry solve kernels: per-residual analytic Jacobians, the 28-double normal-equation accumulator with its
// warp-shuffle block reduction, and the Ceres-2.0 trust-region bookkeeping (restated in oracle/ceres_lm.hpp).
#pragma once
#include "common.cuh"

namespace vb {

// ---------------------------------------------------------------------------------------------
// Normal-equation accumulation helpers.  acc[0..20] = upper triangle of J'J (row-major), acc[21..26] = J'r,
// acc[27] = cost (1/2 rho).
__device__ __forceinline__ void accum_row(double acc[28], const double J[6], double r, double w) {
  // J and r already loss-corrected when w == 1; otherwise scale here: contributes w * J'J and w * J'r
  int k = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const double wi = w * J[i];
#pragma unroll
    for (int j = i; j < 6; ++j) acc[k++] += wi * J[j];
    acc[21 + i] += wi * r;
  }
}

// Huber(0.1) as ceres::HuberLoss + Corrector (rho'' <= 0 branch): returns rho'(s) and adds 1/2 rho(s) to cost.
__device__ __forceinline__ double huber_weight(double s, double* cost) {
  const double a = 0.1, b = a * a;  // ceres::HuberLoss(a): b_ = a * a
  if (s > b) {
    const double r = sqrt(s);
    *cost += 0.5 * (2.0 * a * r - b);
    return fmax(2.2250738585072014e-308, a / r);
  }
  *cost += 0.5 * s;
  return 1.0;
}

// One edge residual block (lidarFactor.hpp:14-56, s == 1).  lp = q*p + t.
__device__ __forceinline__ void edge_block(const double q[4], const double t[3], const float4 p, const double a[3],
                                           const double bb[3], double acc[28]) {
  double Rp[3];
  quat_rotate(q, (double)p.x, (double)p.y, (double)p.z, Rp);
  const double lp[3] = {Rp[0] + t[0], Rp[1] + t[1], Rp[2] + t[2]};
  const double ux = lp[0] - a[0], uy = lp[1] - a[1], uz = lp[2] - a[2];
  const double vx = lp[0] - bb[0], vy = lp[1] - bb[1], vz = lp[2] - bb[2];
  const double dx = a[0] - bb[0], dy = a[1] - bb[1], dz = a[2] - bb[2];
  const double den = sqrt(dx * dx + dy * dy + dz * dz);
  const double r[3] = {(uy * vz - uz * vy) / den, (uz * vx - ux * vz) / den, (ux * vy - uy * vx) / den};
  const double s = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
  const double w = huber_weight(s, &acc[27]);
  // d r / d lp = -[d]x / den ; d lp / d delta = -2 [Rp]x  (EigenQuaternionParameterization, see oracle/ceres_lm.hpp)
  const double ex = dx / den, ey = dy / den, ez = dz / den;
  // A = -[e]x
  const double A[3][3] = {{0.0, ez, -ey}, {-ez, 0.0, ex}, {ey, -ex, 0.0}};
  // G = -2 [Rp]x
  const double G[3][3] = {{0.0, 2.0 * Rp[2], -2.0 * Rp[1]}, {-2.0 * Rp[2], 0.0, 2.0 * Rp[0]}, {2.0 * Rp[1], -2.0 * Rp[0], 0.0}};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double J[6];
#pragma unroll
    for (int c = 0; c < 3; ++c) J[c] = A[i][0] * G[0][c] + A[i][1] * G[1][c] + A[i][2] * G[2][c];
    J[3] = A[i][0]; J[4] = A[i][1]; J[5] = A[i][2];
    accum_row(acc, J, r[i], w);
  }
}

// One plane residual block: r = n . lp + d0  (LidarPlaneFactor: d0 = -n . j; LidarPlaneNormFactor: d0 given).
__device__ __forceinline__ void plane_block(const double q[4], const double t[3], const float4 p, const double n[3],
                                            double d0, double acc[28]) {
  double Rp[3];
  quat_rotate(q, (double)p.x, (double)p.y, (double)p.z, Rp);
  const double lp[3] = {Rp[0] + t[0], Rp[1] + t[1], Rp[2] + t[2]};
  const double r = n[0] * lp[0] + n[1] * lp[1] + n[2] * lp[2] + d0;
  const double w = huber_weight(r * r, &acc[27]);
  // n' * (-2 [Rp]x) = -2 (n x Rp)'
  const double J[6] = {-2.0 * (n[1] * Rp[2] - n[2] * Rp[1]), -2.0 * (n[2] * Rp[0] - n[0] * Rp[2]),
                       -2.0 * (n[0] * Rp[1] - n[1] * Rp[0]), n[0], n[1], n[2]};
  accum_row(acc, J, r, w);
}

// Block-wide sum of 28 doubles -> red[0..27] (valid for all threads after the call).  blockDim.x <= 1024.
static __device__ void block_reduce28(double acc[28], double* red /*[28]*/, double* scratch /*[32][28]*/) {
  const int w = threadIdx.x >> 5, l = lane_id(), nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < 28; ++k) acc[k] = warp_sum(acc[k]);
  if (l == 0) {
#pragma unroll
    for (int k = 0; k < 28; ++k) scratch[w * 28 + k] = acc[k];
  }
  __syncthreads();
  if (threadIdx.x < 28) {
    double s = 0.0;
    for (int i = 0; i < nw; ++i) s += scratch[i * 28 + threadIdx.x];
    red[threadIdx.x] = s;
  }
  __syncthreads();
}


// ---------------------------------------------------------------------------------------------
// shard_allreduce: sum vec[0..n) (n <= 32, shared memory) over the ranks of a point-sharded group from inside a running
// kernel.  Called by every thread of the CTA that owns stream b.  Each rank writes its vector into its OWN slot (local
// stores), publishes a sequence number with release semantics, polls the other ranks' sequence numbers through peer
// memory and then reads their vectors; the sum is formed in rank order, so every rank ends up with identical bits and
// the redundant LM bookkeeping that follows stays in lock-step without any further communication.  A rank can be at
// most one exchange ahead of its peers (it needs their next publication to advance), hence two generations per slot.
constexpr long long kShardSpinLimit = 1ll << 22;   // ~1 s of polling before giving up (reported, never hangs)
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
static __device__ void shard_allreduce(const ShardView& sv, int b, double* vec, int n) {
  __shared__ unsigned long long s_seq;
  if (threadIdx.x == 0) s_seq = sv.seq[b] + 1;
  __syncthreads();
  const unsigned long long seq = s_seq;
  ShardSlot* mine = sv.peer[sv.rank] + b;
  if ((int)threadIdx.x < n) mine->v[seq & 1][threadIdx.x] = vec[threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    st_release_sys_u64(&mine->flag, seq);
    sv.seq[b] = seq;
  }
  if (threadIdx.x < 32) {   // lane r polls rank r
    bool ok = true;
    const int r = threadIdx.x;
    if (r < sv.world && r != sv.rank && *reinterpret_cast<volatile int*>(sv.error) == 0) {
      const unsigned long long* f = &(sv.peer[r] + b)->flag;
      long long spins = 0;
      while (ld_acquire_sys_u64(f) < seq) {
        __nanosleep(100);
        if (++spins > kShardSpinLimit) { ok = false; break; }
      }
    }
    if (!__all_sync(0xffffffffu, ok) && threadIdx.x == 0) atomicOr(sv.error, 1);
  }
  __syncthreads();
  if ((int)threadIdx.x < n) {
    double s = 0.0;
    for (int r = 0; r < sv.world; ++r) s += ld_relaxed_sys_f64(&(sv.peer[r] + b)->v[seq & 1][threadIdx.x]);
    vec[threadIdx.x] = s;
  }
  __syncthreads();
}

// EigenQuaternionParameterization::Plus / Euclidean plus.  x = [q(4), t(3)], delta[6].
__device__ __forceinline__ void manifold_plus(const double x[7], const double d[6], double out[7]) {
  const double nd = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  if (nd > 0.0) {
    const double s = sin(nd) / nd;
    const double dq[4] = {s * d[0], s * d[1], s * d[2], cos(nd)};
    quat_mul(dq, x, out);
  } else {
    out[0] = x[0]; out[1] = x[1]; out[2] = x[2]; out[3] = x[3];
  }
  out[4] = x[4] + d[3]; out[5] = x[5] + d[4]; out[6] = x[6] + d[5];
}

// Solve (H + diag(D2)) y = g for symmetric positive definite 6x6 by Cholesky.  H: upper triangle (21).
// Fully unrolled with compile-time indices so the factor lives in registers (this runs on one thread per problem and
// sits on the critical path of every LM iteration).
static __device__ __forceinline__ bool chol_solve6(const double Hu[21], const double D2[6], const double g[6], double y[6]) {
  double L[6][6];
  bool ok = true;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      if (j > i) continue;
      // A[i][j] with j <= i is stored at upper-triangle position (j, i): index = j*6 - j*(j-1)/2 + (i - j)
      double s = Hu[j * 6 - (j * (j - 1)) / 2 + (i - j)];
      if (i == j) s += D2[i];
#pragma unroll
      for (int q = 0; q < 6; ++q) if (q < j) s -= L[i][q] * L[j][q];
      if (i == j) {
        if (!(s > 0.0)) ok = false;
        L[i][i] = sqrt(s);
      } else {
        L[i][j] = s / L[j][j];
      }
    }
  }
  double z[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double s = g[i];
#pragma unroll
    for (int q = 0; q < 6; ++q) if (q < i) s -= L[i][q] * z[q];
    z[i] = s / L[i][i];
  }
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    double s = z[i];
#pragma unroll
    for (int q = 0; q < 6; ++q) if (q > i) s -= L[q][i] * y[q];
    y[i] = s / L[i][i];
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) if (!isfinite(y[i])) ok = false;
  return ok;
}

// Trust-region LM state kept in shared memory by the solve kernels (Ceres 2.0 defaults, oracle/ceres_lm.hpp).
struct LMCore {
  double x[7], cand[7];
  double H[21], g[6], cost;     // at x, unscaled
  double scale[6], diagonal[6];
  double radius, decrease_factor, model_cost_change, x_norm, gmax;
  int reuse_diagonal, iteration, done, termination, invalid_run;
  int euclid;  // 1: x = [angle-axis(3), t(3)] with plain addition (visual odometry); 0: x = [q(4), t(3)] on the quaternion manifold
  double red[28];               // (J'J, J'r, cost) of the latest evaluation
};
// ... plus the block-reduction scratch of the one-CTA-per-problem kernels
struct LMShared : LMCore {
  double scratch[32 * 28];
};

enum { TERM_NO_CONVERGENCE = 0, TERM_GRADIENT = 1, TERM_PARAMETER = 2, TERM_FUNCTION = 3, TERM_FAILURE = 4 };

__device__ __forceinline__ void lm_record(SolveTrace* tr, double cost, double cand, double mcc, double rel, double radius,
                                          int valid, int succ) {
  const int n = tr->n_records;
  if (n < kMaxLMRecords) {
    LMRecord& R = tr->rec[n];
    R.cost = cost; R.candidate_cost = cand; R.model_cost_change = mcc; R.relative_decrease = rel; R.radius = radius;
    R.step_is_valid = valid; R.step_is_successful = succ;
  }
  tr->n_records = n + 1;
}

// Thread 0: given (H, g, cost) at x, compute the next LM step and candidate; handles invalid steps by shrinking the
// radius (each invalid step is one iteration).  Returns with S.done set, or with S.cand ready for evaluation.
static __device__ void lm_prepare_step(LMCore& S, SolveTrace* tr, int max_iterations) {
  while (true) {
    if (S.iteration >= max_iterations) { S.done = 1; S.termination = TERM_NO_CONVERGENCE; return; }
    if (S.gmax <= 1e-10) { S.done = 1; S.termination = TERM_GRADIENT; return; }
    if (S.radius <= 1e-32) { S.done = 1; S.termination = TERM_PARAMETER; return; }
    S.iteration++;
    double Hs[21], gs[6];
    {
      int k = 0;
      for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) { Hs[k] = S.H[k] * S.scale[i] * S.scale[j]; ++k; }
      for (int i = 0; i < 6; ++i) gs[i] = S.g[i] * S.scale[i];
    }
    if (!S.reuse_diagonal) {
      int k = 0;
      for (int i = 0; i < 6; ++i) { S.diagonal[i] = fmin(fmax(Hs[k], 1e-6), 1e32); k += 6 - i; }
    }
    double D2[6];
    for (int i = 0; i < 6; ++i) D2[i] = S.diagonal[i] / S.radius;
    double y[6];
    const bool ok = chol_solve6(Hs, D2, gs, y);
    S.reuse_diagonal = 1;
    double mcc = 0.0;
    if (ok) {
      // step = -y ; model_cost_change = -(J step).(r + J step / 2) = y'g - y'Hy/2
      double yHy = 0.0, yg = 0.0;
      int k = 0;
      for (int i = 0; i < 6; ++i) {
        yg += y[i] * gs[i];
        for (int j = i; j < 6; ++j) { yHy += (i == j ? 1.0 : 2.0) * y[i] * Hs[k] * y[j]; ++k; }
      }
      mcc = yg - 0.5 * yHy;
    }
    if (!(ok && mcc > 0.0)) {
      S.invalid_run++;
      S.radius = S.radius / S.decrease_factor; S.decrease_factor *= 2.0; S.reuse_diagonal = 1;
      lm_record(tr, S.cost, 0.0, mcc, 0.0, S.radius, 0, 0);
      if (S.invalid_run >= 5) { S.done = 1; S.termination = TERM_FAILURE; return; }
      continue;
    }
    S.invalid_run = 0;
    S.model_cost_change = mcc;
    double delta[6];
    for (int i = 0; i < 6; ++i) delta[i] = -y[i] * S.scale[i];
    if (S.euclid) { for (int i = 0; i < 6; ++i) S.cand[i] = S.x[i] + delta[i]; S.cand[6] = 0.0; }
    else manifold_plus(S.x, delta, S.cand);
    return;
  }
}

// Thread 0: red[] holds (H, g, cost) evaluated at S.cand.  Accept / reject, update the trust region.
static __device__ void lm_finish_step(LMCore& S, SolveTrace* tr) {
  const double cand_cost = S.red[27];
  double step_norm = 0.0;
  for (int i = 0; i < 7; ++i) step_norm += (S.x[i] - S.cand[i]) * (S.x[i] - S.cand[i]);
  step_norm = sqrt(step_norm);
  if (step_norm <= 1e-8 * (S.x_norm + 1e-8)) {
    lm_record(tr, S.cost, cand_cost, S.model_cost_change, 0.0, S.radius, 1, 0);
    S.done = 1; S.termination = TERM_PARAMETER; return;
  }
  const double cost_change = S.cost - cand_cost;
  if (fabs(cost_change) <= 1e-6 * S.cost) {
    lm_record(tr, S.cost, cand_cost, S.model_cost_change, 0.0, S.radius, 1, 0);
    S.done = 1; S.termination = TERM_FUNCTION; return;
  }
  const double rel = cost_change / S.model_cost_change;
  if (rel > 1e-3) {
    for (int i = 0; i < 7; ++i) S.x[i] = S.cand[i];
    double xn = 0.0;
    for (int i = 0; i < 7; ++i) xn += S.x[i] * S.x[i];
    S.x_norm = sqrt(xn);
    for (int i = 0; i < 21; ++i) S.H[i] = S.red[i];
    double gm = 0.0;
    for (int i = 0; i < 6; ++i) { S.g[i] = S.red[21 + i]; gm = fmax(gm, fabs(S.g[i])); }
    S.gmax = gm;
    S.cost = cand_cost;
    const double t = 2.0 * rel - 1.0;
    S.radius = S.radius / fmax(1.0 / 3.0, 1.0 - t * t * t);
    S.radius = fmin(1e16, S.radius);
    S.decrease_factor = 2.0;
    S.reuse_diagonal = 0;
    lm_record(tr, S.cost, cand_cost, S.model_cost_change, rel, S.radius, 1, 1);
  } else {
    S.radius = S.radius / S.decrease_factor;
    S.decrease_factor *= 2.0;
    S.reuse_diagonal = 1;
    lm_record(tr, S.cost, cand_cost, S.model_cost_change, rel, S.radius, 1, 0);
  }
}

// Thread 0: red[] holds (H, g, cost) at the initial x (iteration 0).
static __device__ void lm_begin(LMCore& S, SolveTrace* tr, const double x0[7]) {
  for (int i = 0; i < 7; ++i) S.x[i] = x0[i];
  double xn = 0.0;
  for (int i = 0; i < 7; ++i) xn += x0[i] * x0[i];
  S.x_norm = sqrt(xn);
  for (int i = 0; i < 21; ++i) S.H[i] = S.red[i];
  double gm = 0.0;
  for (int i = 0; i < 6; ++i) { S.g[i] = S.red[21 + i]; gm = fmax(gm, fabs(S.g[i])); }
  S.gmax = gm;
  S.cost = S.red[27];
  int k = 0;
  for (int i = 0; i < 6; ++i) { S.scale[i] = 1.0 / (1.0 + sqrt(S.H[k])); k += 6 - i; }
  S.radius = 1e4; S.decrease_factor = 2.0; S.reuse_diagonal = 0; S.iteration = 0; S.done = 0;
  S.termination = TERM_NO_CONVERGENCE; S.invalid_run = 0;
  tr->n_records = 0;
  lm_record(tr, S.cost, 0.0, 0.0, 0.0, S.radius, 0, 0);
}


// Runs the whole trust-region loop for one problem inside the calling CTA.  `evaluate(x)` must be a block-wide
// callable that leaves (J'J, J'r, cost) at x in S.red[0..27] (e.g. via block_reduce28) and ends with a barrier.
// On return S.x holds the solution (valid for all threads after the final barrier).
template <typename Eval>
__device__ void lm_solve_block(LMShared& S, SolveTrace* tr, int max_iterations, bool empty, Eval evaluate, int euclid = 0) {
  if (threadIdx.x == 0) S.euclid = euclid;
  double x0[7];
  for (int i = 0; i < 7; ++i) x0[i] = S.x[i];
  evaluate(x0);
  if (threadIdx.x == 0) {
    lm_begin(S, tr, x0);
    if (empty) { S.done = 1; S.termination = TERM_GRADIENT; }
    else lm_prepare_step(S, tr, max_iterations);
  }
  __syncthreads();
  while (!S.done) {
    double xc[7];
    for (int i = 0; i < 7; ++i) xc[i] = S.cand[i];
    evaluate(xc);
    if (threadIdx.x == 0) {
      lm_finish_step(S, tr);
      if (!S.done) lm_prepare_step(S, tr, max_iterations);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 7; ++i) tr->para[i] = S.x[i];
    tr->termination = S.termination;
  }
}

}  // namespace vb
