// vloam_b200 — residual blocks with motion-distortion interpolation (laser_odometry.h:90 DISTORTION == true): the pose is
// applied as Identity.slerp(s, q_last_curr), s * t_last_curr with a per-point ratio s (lidarFactor.hpp:28-35, 78-83).
// The shipped configuration has s == 1 and uses the closed-form Jacobians of gn_solver.cuh; for s != 1 the Jacobian goes
// through Eigen's slerp (acos / sin of the quaternion's w), so the block is evaluated with forward-mode dual numbers over
// the six local increments — the same chain rule ceres::AutoDiffCostFunction + EigenQuaternionParameterization apply
// (7 ambient partials times the 7 x 6 plus-Jacobian), seeded directly in the local frame.
#pragma once
#include "gn_solver.cuh"

namespace vb {

struct DJet {
  double a;
  double v[6];
};
__device__ __forceinline__ DJet jconst(double s) { DJet h; h.a = s; for (int i = 0; i < 6; ++i) h.v[i] = 0.0; return h; }
__device__ __forceinline__ DJet operator+(const DJet& f, const DJet& g) { DJet h; h.a = f.a + g.a; for (int i = 0; i < 6; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
__device__ __forceinline__ DJet operator-(const DJet& f, const DJet& g) { DJet h; h.a = f.a - g.a; for (int i = 0; i < 6; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
__device__ __forceinline__ DJet operator-(const DJet& f) { DJet h; h.a = -f.a; for (int i = 0; i < 6; ++i) h.v[i] = -f.v[i]; return h; }
__device__ __forceinline__ DJet operator*(const DJet& f, const DJet& g) { DJet h; h.a = f.a * g.a; for (int i = 0; i < 6; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
__device__ __forceinline__ DJet operator/(const DJet& f, const DJet& g) {      // ceres::Jet: (f / g)' = (f' - (f / g) g') / g
  DJet h; const double gi = 1.0 / g.a; const double fg = f.a * gi; h.a = fg;
  for (int i = 0; i < 6; ++i) h.v[i] = (f.v[i] - fg * g.v[i]) * gi;
  return h;
}
__device__ __forceinline__ DJet jsin(const DJet& f) { DJet h; h.a = sin(f.a); const double c = cos(f.a); for (int i = 0; i < 6; ++i) h.v[i] = c * f.v[i]; return h; }
__device__ __forceinline__ DJet jacos(const DJet& f) { DJet h; h.a = acos(f.a); const double d = -1.0 / sqrt(1.0 - f.a * f.a); for (int i = 0; i < 6; ++i) h.v[i] = d * f.v[i]; return h; }

// lp = Identity.slerp(s, q (+) delta) * p + s * (t + dt) as a function of the six local increments at zero
// (Eigen::QuaternionBase::slerp restated as in oracle/lidar_factors.hpp; q (+) delta = [sin|d| / |d| d, cos|d|] * q).
__device__ __forceinline__ void transform_point_slerp(const double q[4], const double t[3], double s, const float4 p, DJet lp[3]) {
  // seeds: d(q (+) delta) / d delta_i at 0 = (e_i, 0) * q
  DJet qx = jconst(q[0]), qy = jconst(q[1]), qz = jconst(q[2]), qw = jconst(q[3]);
  qx.v[0] = q[3];  qx.v[1] = q[2];  qx.v[2] = -q[1];
  qy.v[0] = -q[2]; qy.v[1] = q[3];  qy.v[2] = q[0];
  qz.v[0] = q[1];  qz.v[1] = -q[0]; qz.v[2] = q[3];
  qw.v[0] = -q[0]; qw.v[1] = -q[1]; qw.v[2] = -q[2];
  const double one = 1.0 - 2.220446049250313e-16;
  const DJet d = qw;                                  // dot(identity, q)
  const DJet absD = d.a < 0.0 ? -d : d;
  DJet scale0, scale1;
  if (absD.a >= one) {
    scale0 = jconst(1.0 - s); scale1 = jconst(s);
  } else {
    const DJet theta = jacos(absD);
    const DJet sinTheta = jsin(theta);
    scale0 = jsin(jconst(1.0 - s) * theta) / sinTheta;
    scale1 = jsin(jconst(s) * theta) / sinTheta;
  }
  if (d.a < 0.0) scale1 = -scale1;
  const DJet sw = scale0 + scale1 * qw, sx = scale1 * qx, sy = scale1 * qy, sz = scale1 * qz;
  // Eigen quaternion * vector: v + w (2 u x v) + u x (2 u x v)
  const DJet vx = jconst((double)p.x), vy = jconst((double)p.y), vz = jconst((double)p.z);
  DJet ux = sy * vz - sz * vy, uy = sz * vx - sx * vz, uz = sx * vy - sy * vx;
  ux = ux + ux; uy = uy + uy; uz = uz + uz;
  lp[0] = vx + sw * ux + (sy * uz - sz * uy);
  lp[1] = vy + sw * uy + (sz * ux - sx * uz);
  lp[2] = vz + sw * uz + (sx * uy - sy * ux);
  for (int k = 0; k < 3; ++k) { DJet tk = jconst(s * t[k]); tk.v[3 + k] = s; lp[k] = lp[k] + tk; }
}

// LidarEdgeFactor with ratio s (lidarFactor.hpp:14-56)
__device__ __forceinline__ void edge_block_slerp(const double q[4], const double t[3], const float4 p, const double a[3], const double bb[3],
                                                 double s, double acc[28]) {
  DJet lp[3];
  transform_point_slerp(q, t, s, p, lp);
  const DJet ux = lp[0] - jconst(a[0]), uy = lp[1] - jconst(a[1]), uz = lp[2] - jconst(a[2]);
  const DJet vx = lp[0] - jconst(bb[0]), vy = lp[1] - jconst(bb[1]), vz = lp[2] - jconst(bb[2]);
  const double dx = a[0] - bb[0], dy = a[1] - bb[1], dz = a[2] - bb[2];
  const DJet den = jconst(sqrt(dx * dx + dy * dy + dz * dz));
  const DJet r[3] = {(uy * vz - uz * vy) / den, (uz * vx - ux * vz) / den, (ux * vy - uy * vx) / den};
  const double w = huber_weight(r[0].a * r[0].a + r[1].a * r[1].a + r[2].a * r[2].a, &acc[27]);
#pragma unroll
  for (int i = 0; i < 3; ++i) accum_row(acc, r[i].v, r[i].a, w);
}
// LidarPlaneFactor with ratio s (lidarFactor.hpp:58-106): r = (lp - j) . n = lp . n + d0
__device__ __forceinline__ void plane_block_slerp(const double q[4], const double t[3], const float4 p, const double n[3], double d0, double s,
                                                  double acc[28]) {
  DJet lp[3];
  transform_point_slerp(q, t, s, p, lp);
  const DJet r = jconst(n[0]) * lp[0] + jconst(n[1]) * lp[1] + jconst(n[2]) * lp[2] + jconst(d0);
  const double w = huber_weight(r.a * r.a, &acc[27]);
  accum_row(acc, r.v, r.a, w);
}

}  // namespace vb
