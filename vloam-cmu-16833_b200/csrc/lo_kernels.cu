// vloam_b200 — laserOdometry on sm_100a.
//
// Replaces vloam::LaserOdometry::solveLO
// (reference src/lidar_odometry_mapping/src/laser_odometry.cpp:187-536) and the Ceres
// problem it builds from lidarFactor.hpp:14-106:
//
//   lo_associate  :266-444  TransformToStart (:149-167), exact 1-NN in the last scan's
//                           less-sharp / less-flat cloud, ring-window 2nd / 3rd neighbour
//   lo_solve      :457-463 + :477-478  the whole ceres::Solve (Huber 0.1, quaternion manifold,
//                           Levenberg-Marquardt, <= 4 iterations) for one outer pass in ONE
//                           launch: per-residual analytic Jacobians, warp-shuffle reduction of
//                           the 6x6 J'J / 6x1 J'r / cost (28 doubles), 6x6 Cholesky and the
//                           trust-region bookkeeping on-chip — no host round trip per iteration.
//
// The Ceres semantics restated here are documented in oracle/ceres_lm.hpp.
#include "common.cuh"
#include "internal.h"

namespace vb {

// ---------------------------------------------------------------------------------------------
// lo_associate: one warp per query.  grid (ceil((kMaxSharp + kMaxFlat) / 8), B), block 256.
//   query slots [0, kMaxSharp)            : corner features vs cornerLast
//   query slots [kMaxSharp, +kMaxFlat)    : plane  features vs surfLast
// corr[b][slot] = (closest, ind2, ind3, valid)
__global__ void __launch_bounds__(256) lo_associate(const SRHeader* __restrict__ hdrCur, const SRHeader* __restrict__ hdrLast,
                                                     const LOState* __restrict__ lo,
                                                     const float4* __restrict__ sharp, const float4* __restrict__ flat,
                                                     const float4* __restrict__ cornerLast, const float4* __restrict__ surfLast,
                                                     int cap, int4* __restrict__ corr) {
  const int b = blockIdx.y;
  const int slot = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int l = lane_id();
  const SRHeader& hc = hdrCur[b];
  const SRHeader& hl = hdrLast[b];
  const bool isCorner = slot < kMaxSharp;
  const int qi = isCorner ? slot : slot - kMaxSharp;
  const int nq = isCorner ? hc.nSharp : hc.nFlat;
  if (slot >= kMaxSharp + kMaxFlat) return;
  int4 out = make_int4(-1, -1, -1, 0);
  if (qi < nq) {
    const float4 p = isCorner ? sharp[(size_t)b * kMaxSharp + qi] : flat[(size_t)b * kMaxFlat + qi];
    // TransformToStart, DISTORTION == false: un = q_last_curr * p + t_last_curr, rounded to float (:158-165)
    double un[3];
    quat_rotate(lo[b].para_q, (double)p.x, (double)p.y, (double)p.z, un);
    const float sx = (float)(un[0] + lo[b].para_t[0]);
    const float sy = (float)(un[1] + lo[b].para_t[1]);
    const float sz = (float)(un[2] + lo[b].para_t[2]);
    const float4* T = isCorner ? cornerLast + (size_t)b * kMaxLessSharp : surfLast + (size_t)b * cap;
    const int nT = isCorner ? hl.nLessSharp : hl.nLessFlat;
    // exact 1-NN (kd-tree replacement): ties -> lower index
    unsigned long long best = 0xffffffffffffffffull;
    for (int j = l; j < nT; j += 32) {
      const float4 t = T[j];
      const float d = sqdist_f(sx, sy, sz, t.x, t.y, t.z);
      const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)j;
      best = key < best ? key : best;
    }
    best = warp_min_u64(best);
    if (nT > 0 && (double)__uint_as_float((unsigned)(best >> 32)) < 25.0) {  // DISTANCE_SQ_THRESHOLD (:272)
      const int closest = (int)(unsigned)best;
      const int id = (int)T[closest].w;  // closestPointScanID (:275)
      // The reference walks the ring-major target cloud away from `closest` in both directions, classifying
      // every point by int(intensity) and stopping at the first point more than NEARBY_SCAN = 2.5 rings away
      // (:279-324 / :368-417).  Restated literally, 32 points per step: a later candidate replaces the
      // incumbent only if strictly closer, so the winner is the minimum over (distance, visiting order) with
      // the forward walk (ascending j) visited before the backward walk (descending j).
      unsigned long long k2 = 0xffffffffffffffffull, k3 = 0xffffffffffffffffull;
      for (int base = closest + 1; base < nT; base += 32) {  // forward
        const int j = base + l;
        bool brk = false, live = j < nT;
        int rid = 0;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live) { t = T[j]; rid = (int)t.w; brk = (double)rid > (double)id + 2.5; }
        const unsigned bm = __ballot_sync(0xffffffffu, brk);
        if (bm) live = live && l < __ffs(bm) - 1;
        if (live) {
          const float d = sqdist_f(t.x, t.y, t.z, sx, sy, sz);
          if ((double)d < 25.0) {
            const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)j;
            if (isCorner) { if (rid > id) k2 = key < k2 ? key : k2; }
            else if (rid <= id) k2 = key < k2 ? key : k2;
            else k3 = key < k3 ? key : k3;
          }
        }
        if (bm) break;
      }
      for (int base = closest - 1; base >= 0; base -= 32) {  // backward
        const int j = base - l;
        bool brk = false, live = j >= 0;
        int rid = 0;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live) { t = T[j]; rid = (int)t.w; brk = (double)rid < (double)id - 2.5; }
        const unsigned bm = __ballot_sync(0xffffffffu, brk);
        if (bm) live = live && l < __ffs(bm) - 1;
        if (live) {
          const float d = sqdist_f(t.x, t.y, t.z, sx, sy, sz);
          if ((double)d < 25.0) {
            const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (0x80000000u + (unsigned)(nT - j));
            if (isCorner) { if (rid < id) k2 = key < k2 ? key : k2; }
            else if (rid >= id) k2 = key < k2 ? key : k2;
            else k3 = key < k3 ? key : k3;
          }
        }
        if (bm) break;
      }
      k2 = warp_min_u64(k2);
      k3 = warp_min_u64(k3);
      auto decode = [&](unsigned long long k) {
        const unsigned o = (unsigned)k;
        return (o & 0x80000000u) ? nT - (int)(o & 0x7fffffffu) : (int)o;
      };
      if (isCorner) {
        if (k2 != 0xffffffffffffffffull) out = make_int4(closest, decode(k2), -1, 1);
      } else if (k2 != 0xffffffffffffffffull && k3 != 0xffffffffffffffffull) {
        out = make_int4(closest, decode(k2), decode(k3), 1);
      }
    }
  }
  if (l == 0) corr[(size_t)b * (kMaxSharp + kMaxFlat) + slot] = out;
}

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
__device__ void block_reduce28(double acc[28], double* red /*[28]*/, double* scratch /*[32][28]*/) {
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
__device__ bool chol_solve6(const double Hu[21], const double D2[6], const double g[6], double y[6]) {
  double A[6][6];
  int k = 0;
  for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) { A[i][j] = Hu[k]; A[j][i] = Hu[k]; ++k; }
  for (int i = 0; i < 6; ++i) A[i][i] += D2[i];
  double L[6][6];
  for (int i = 0; i < 6; ++i) {
    for (int j = 0; j <= i; ++j) {
      double s = A[i][j];
      for (int q = 0; q < j; ++q) s -= L[i][q] * L[j][q];
      if (i == j) {
        if (!(s > 0.0)) return false;
        L[i][i] = sqrt(s);
      } else {
        L[i][j] = s / L[j][j];
      }
    }
  }
  double z[6];
  for (int i = 0; i < 6; ++i) { double s = g[i]; for (int q = 0; q < i; ++q) s -= L[i][q] * z[q]; z[i] = s / L[i][i]; }
  for (int i = 5; i >= 0; --i) { double s = z[i]; for (int q = i + 1; q < 6; ++q) s -= L[q][i] * y[q]; y[i] = s / L[i][i]; }
  for (int i = 0; i < 6; ++i) if (!isfinite(y[i])) return false;
  return true;
}

// Trust-region LM state kept in shared memory by the solve kernels (Ceres 2.0 defaults, oracle/ceres_lm.hpp).
struct LMShared {
  double x[7], cand[7];
  double H[21], g[6], cost;     // at x, unscaled
  double scale[6], diagonal[6];
  double radius, decrease_factor, model_cost_change, x_norm, gmax;
  int reuse_diagonal, iteration, done, termination, invalid_run, eval_target;  // eval_target: 0 = x, 1 = cand
  double red[28];
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
__device__ void lm_prepare_step(LMShared& S, SolveTrace* tr, int max_iterations) {
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
    manifold_plus(S.x, delta, S.cand);
    return;
  }
}

// Thread 0: red[] holds (H, g, cost) evaluated at S.cand.  Accept / reject, update the trust region.
__device__ void lm_finish_step(LMShared& S, SolveTrace* tr) {
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
__device__ void lm_begin(LMShared& S, SolveTrace* tr, const double x0[7]) {
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

// ---------------------------------------------------------------------------------------------
// lo_solve: grid (B), block 256.  One CTA solves one stream's pass.
__global__ void __launch_bounds__(256) lo_solve(const SRHeader* __restrict__ hdrCur, LOState* __restrict__ lo,
                                                 const float4* __restrict__ sharp, const float4* __restrict__ flat,
                                                 const float4* __restrict__ cornerLast, const float4* __restrict__ surfLast,
                                                 int cap, const int4* __restrict__ corr, int pass, int max_iterations,
                                                 int integrate) {
  __shared__ LMShared S;
  const int b = blockIdx.x;
  LOState& st = lo[b];
  SolveTrace* tr = &st.trace[pass];
  const int4* cr = corr + (size_t)b * (kMaxSharp + kMaxFlat);
  const float4* sh = sharp + (size_t)b * kMaxSharp;
  const float4* fl = flat + (size_t)b * kMaxFlat;
  const float4* CL = cornerLast + (size_t)b * kMaxLessSharp;
  const float4* SL = surfLast + (size_t)b * cap;
  const int nSharp = hdrCur[b].nSharp, nFlat = hdrCur[b].nFlat;

  auto evaluate = [&](const double* x) {
    double acc[28];
#pragma unroll
    for (int k = 0; k < 28; ++k) acc[k] = 0.0;
    const double q[4] = {x[0], x[1], x[2], x[3]};
    const double t[3] = {x[4], x[5], x[6]};
    for (int s = threadIdx.x; s < kMaxSharp + kMaxFlat; s += blockDim.x) {
      const bool isCorner = s < kMaxSharp;
      const int qi = isCorner ? s : s - kMaxSharp;
      if (qi >= (isCorner ? nSharp : nFlat)) continue;
      const int4 c = cr[s];
      if (!c.w) continue;
      if (isCorner) {
        const float4 p = sh[qi], A = CL[c.x], Bp = CL[c.y];
        const double a[3] = {(double)A.x, (double)A.y, (double)A.z};
        const double bb[3] = {(double)Bp.x, (double)Bp.y, (double)Bp.z};
        edge_block(q, t, p, a, bb, acc);
      } else {
        const float4 p = fl[qi], Jp = SL[c.x], Lp = SL[c.y], Mp = SL[c.z];
        // ljm_norm = normalize((j - l) x (j - m))  (lidarFactor.hpp:68-69)
        const double ax = (double)Jp.x - (double)Lp.x, ay = (double)Jp.y - (double)Lp.y, az = (double)Jp.z - (double)Lp.z;
        const double bx = (double)Jp.x - (double)Mp.x, by = (double)Jp.y - (double)Mp.y, bz = (double)Jp.z - (double)Mp.z;
        double n[3] = {ay * bz - az * by, az * bx - ax * bz, ax * by - ay * bx};
        const double z2 = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
        if (z2 > 0.0) { const double nn = sqrt(z2); n[0] /= nn; n[1] /= nn; n[2] /= nn; }
        const double d0 = -(n[0] * (double)Jp.x + n[1] * (double)Jp.y + n[2] * (double)Jp.z);
        plane_block(q, t, p, n, d0, acc);
      }
    }
    block_reduce28(acc, S.red, S.scratch);
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) S.x[i] = st.para_q[i];
    for (int i = 0; i < 3; ++i) S.x[4 + i] = st.para_t[i];
  }
  __syncthreads();
  // correspondence counts (:348, :441)
  {
    int nc = 0, np = 0;
    for (int s = threadIdx.x; s < kMaxSharp + kMaxFlat; s += blockDim.x) {
      const bool isCorner = s < kMaxSharp;
      const int qi = isCorner ? s : s - kMaxSharp;
      if (qi < (isCorner ? nSharp : nFlat) && cr[s].w) { if (isCorner) nc++; else np++; }
    }
    nc = __reduce_add_sync(0xffffffffu, nc);
    np = __reduce_add_sync(0xffffffffu, np);
    __shared__ int s_nc, s_np;
    if (threadIdx.x == 0) { s_nc = 0; s_np = 0; }
    __syncthreads();
    if (lane_id() == 0) { atomicAdd(&s_nc, nc); atomicAdd(&s_np, np); }
    __syncthreads();
    if (threadIdx.x == 0) { st.corner_correspondence = s_nc; st.plane_correspondence = s_np; tr->n_corner = s_nc; tr->n_plane = s_np; }
  }
  double x0[7];
  for (int i = 0; i < 7; ++i) x0[i] = S.x[i];
  evaluate(x0);
  if (threadIdx.x == 0) {
    lm_begin(S, tr, x0);
    if (tr->n_corner + tr->n_plane == 0) { S.done = 1; S.termination = TERM_GRADIENT; }
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
    for (int i = 0; i < 4; ++i) st.para_q[i] = S.x[i];
    for (int i = 0; i < 3; ++i) st.para_t[i] = S.x[4 + i];
    for (int i = 0; i < 7; ++i) tr->para[i] = S.x[i];
    tr->termination = S.termination;
    if (integrate) {  // :477-478
      double rt[3];
      quat_rotate(st.q_w, st.para_t[0], st.para_t[1], st.para_t[2], rt);
      st.t_w[0] += rt[0]; st.t_w[1] += rt[1]; st.t_w[2] += rt[2];
      double qn[4];
      quat_mul(st.q_w, st.para_q, qn);
      for (int i = 0; i < 4; ++i) st.q_w[i] = qn[i];
    }
  }
}

__global__ void lo_init_state(LOState* lo, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  LOState& s = lo[b];
  s.para_q[0] = s.para_q[1] = s.para_q[2] = 0.0; s.para_q[3] = 1.0;
  s.para_t[0] = s.para_t[1] = s.para_t[2] = 0.0;
  s.q_w[0] = s.q_w[1] = s.q_w[2] = 0.0; s.q_w[3] = 1.0;
  s.t_w[0] = s.t_w[1] = s.t_w[2] = 0.0;
  s.corner_correspondence = s.plane_correspondence = 0;
  s.trace[0].n_records = s.trace[1].n_records = 0;
}

// !detach_VO_LO: para_q / para_t are overwritten with the VO prior at the top of BOTH outer passes, before the
// association (laser_odometry.cpp:223-236).  prior: [B][7] = q(xyzw), t.
__global__ void lo_set_motion(LOState* lo, const double* __restrict__ prior, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  for (int i = 0; i < 4; ++i) lo[b].para_q[i] = prior[b * 7 + i];
  for (int i = 0; i < 3; ++i) lo[b].para_t[i] = prior[b * 7 + 4 + i];
}

// pose[b][16] = q_last_curr(4) t_last_curr(3) q_w_curr(4) t_w_curr(3) corner_correspondence plane_correspondence
__global__ void lo_export_pose(const LOState* __restrict__ lo, double* __restrict__ pose, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const LOState& s = lo[b];
  double* o = pose + (size_t)b * 16;
  for (int i = 0; i < 4; ++i) { o[i] = s.para_q[i]; o[7 + i] = s.q_w[i]; }
  for (int i = 0; i < 3; ++i) { o[4 + i] = s.para_t[i]; o[11 + i] = s.t_w[i]; }
  o[14] = s.corner_correspondence; o[15] = s.plane_correspondence;
}
void launch_lo_export(Profiler* prof, cudaStream_t st, const LOState* lo, double* pose, int B) {
  VB_LAUNCH(prof, K_LO_EXPORT, st, lo_export_pose<<<(B + 127) / 128, 128, 0, st>>>(lo, pose, B));
}
void launch_lo_set_motion(Profiler* prof, cudaStream_t st, LOState* lo, const double* motion, int B) {
  VB_LAUNCH(prof, K_LO_SET_MOTION, st, lo_set_motion<<<(B + 127) / 128, 128, 0, st>>>(lo, motion, B));
}

void launch_lo_init(Profiler* prof, cudaStream_t st, LOState* lo, int B) {
  VB_LAUNCH(prof, K_LO_INIT, st, lo_init_state<<<(B + 127) / 128, 128, 0, st>>>(lo, B));
}

void launch_lo_pass(Profiler* prof, cudaStream_t st, int B, int cap, const SRHeader* hdrCur, const SRHeader* hdrLast, LOState* lo,
                    const float4* sharp, const float4* flat, const float4* cornerLast, const float4* surfLast,
                    int4* corr, int pass, int max_iterations, int integrate, const double* prior) {
  if (prior) VB_LAUNCH(prof, K_LO_SET_MOTION, st, lo_set_motion<<<(B + 127) / 128, 128, 0, st>>>(lo, prior, B));
  VB_LAUNCH(prof, K_LO_ASSOCIATE, st, lo_associate<<<dim3((kMaxSharp + kMaxFlat + 7) / 8, B), 256, 0, st>>>(
                                          hdrCur, hdrLast, lo, sharp, flat, cornerLast, surfLast, cap, corr));
  VB_LAUNCH(prof, K_LO_SOLVE, st, lo_solve<<<B, 256, 0, st>>>(hdrCur, lo, sharp, flat, cornerLast, surfLast, cap, corr, pass,
                                                               max_iterations, integrate));
}

}  // namespace vb
