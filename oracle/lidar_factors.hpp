// ORACLE — TEST INFRASTRUCTURE ONLY (see types.hpp header).  PARITY UNPINNED.
//
// lidar_factors.hpp: literal restatement of the three live Ceres functors of
// reference include/lidar_odometry_mapping/lidarFactor.hpp — LidarEdgeFactor
// (:14-56), LidarPlaneFactor (:58-106), LidarPlaneNormFactor (:108-139) —
// evaluated with dual numbers exactly as ceres::AutoDiffCostFunction<.,N,4,3>
// would (7 partials).  Eigen pieces restated: Quaternion::slerp (Eigen 3.3
// Geometry/Quaternion.h), quaternion * vector (_transformVector), cross, norm.
#pragma once
#include <limits>

#include "ceres_lm.hpp"
#include "jet.hpp"
#include "types.hpp"

namespace oracle {

template <typename T> struct V3 { T x, y, z; };
template <typename T> inline V3<T> vadd(const V3<T>& a, const V3<T>& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <typename T> inline V3<T> vsub(const V3<T>& a, const V3<T>& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <typename T> inline V3<T> vcross(const V3<T>& a, const V3<T>& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <typename T> inline T vdot(const V3<T>& a, const V3<T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename T> struct Q4 { T w, x, y, z; };

// Eigen::QuaternionBase::slerp(t, other) called on the identity quaternion.
template <typename T>
inline Q4<T> slerp_from_identity(const T& t, const Q4<T>& other) {
  const T one = T(1.0) - T(std::numeric_limits<double>::epsilon());
  const Q4<T> self{T(1.0), T(0.0), T(0.0), T(0.0)};
  T d = self.w * other.w + self.x * other.x + self.y * other.y + self.z * other.z;
  T absD = jabs(d);
  T scale0, scale1;
  if (absD >= one) {
    scale0 = T(1.0) - t;
    scale1 = t;
  } else {
    T theta = jacos(absD);
    T sinTheta = jsin(theta);
    scale0 = jsin((T(1.0) - t) * theta) / sinTheta;
    scale1 = jsin((t * theta)) / sinTheta;
  }
  if (d < T(0.0)) scale1 = -scale1;
  return {scale0 * self.w + scale1 * other.w, scale0 * self.x + scale1 * other.x,
          scale0 * self.y + scale1 * other.y, scale0 * self.z + scale1 * other.z};
}

// Eigen quaternion * vector: v + w*(2 u x v) + u x (2 u x v)
template <typename T>
inline V3<T> qrot(const Q4<T>& q, const V3<T>& v) {
  V3<T> u{q.x, q.y, q.z};
  V3<T> uv = vcross(u, v);
  uv = vadd(uv, uv);
  V3<T> wuv{q.w * uv.x, q.w * uv.y, q.w * uv.z};
  return vadd(vadd(v, wuv), vcross(u, uv));
}

struct LidarEdgeFunctor {  // lidarFactor.hpp:14-56
  Vec3 curr_point, last_point_a, last_point_b;
  double s;
  template <typename T>
  bool operator()(const T* q, const T* t, T* residual) const {
    V3<T> cp{T(curr_point.x), T(curr_point.y), T(curr_point.z)};
    V3<T> lpa{T(last_point_a.x), T(last_point_a.y), T(last_point_a.z)};
    V3<T> lpb{T(last_point_b.x), T(last_point_b.y), T(last_point_b.z)};
    Q4<T> q_last_curr{q[3], q[0], q[1], q[2]};
    q_last_curr = slerp_from_identity(T(s), q_last_curr);
    V3<T> t_last_curr{T(s) * t[0], T(s) * t[1], T(s) * t[2]};
    V3<T> lp = vadd(qrot(q_last_curr, cp), t_last_curr);
    V3<T> nu = vcross(vsub(lp, lpa), vsub(lp, lpb));
    V3<T> de = vsub(lpa, lpb);
    T den = jsqrt(vdot(de, de));
    residual[0] = nu.x / den;
    residual[1] = nu.y / den;
    residual[2] = nu.z / den;
    return true;
  }
};

struct LidarPlaneFunctor {  // lidarFactor.hpp:58-106
  Vec3 curr_point, last_point_j, ljm_norm;
  double s;
  LidarPlaneFunctor(const Vec3& c, const Vec3& j, const Vec3& l, const Vec3& m, double s_) : curr_point(c), last_point_j(j), s(s_) {
    ljm_norm = cross(j - l, j - m);                // :68
    const double z = dot(ljm_norm, ljm_norm);       // Eigen normalize(): divide by sqrt(squaredNorm) if > 0
    if (z > 0) { const double n = std::sqrt(z); ljm_norm = {ljm_norm.x / n, ljm_norm.y / n, ljm_norm.z / n}; }
  }
  template <typename T>
  bool operator()(const T* q, const T* t, T* residual) const {
    V3<T> cp{T(curr_point.x), T(curr_point.y), T(curr_point.z)};
    V3<T> lpj{T(last_point_j.x), T(last_point_j.y), T(last_point_j.z)};
    V3<T> ljm{T(ljm_norm.x), T(ljm_norm.y), T(ljm_norm.z)};
    Q4<T> q_last_curr{q[3], q[0], q[1], q[2]};
    q_last_curr = slerp_from_identity(T(s), q_last_curr);
    V3<T> t_last_curr{T(s) * t[0], T(s) * t[1], T(s) * t[2]};
    V3<T> lp = vadd(qrot(q_last_curr, cp), t_last_curr);
    residual[0] = vdot(vsub(lp, lpj), ljm);
    return true;
  }
};

struct LidarPlaneNormFunctor {  // lidarFactor.hpp:108-139
  Vec3 curr_point, plane_unit_norm;
  double negative_OA_dot_norm;
  template <typename T>
  bool operator()(const T* q, const T* t, T* residual) const {
    Q4<T> q_w_curr{q[3], q[0], q[1], q[2]};
    V3<T> t_w_curr{t[0], t[1], t[2]};
    V3<T> cp{T(curr_point.x), T(curr_point.y), T(curr_point.z)};
    V3<T> point_w = vadd(qrot(q_w_curr, cp), t_w_curr);
    V3<T> norm{T(plane_unit_norm.x), T(plane_unit_norm.y), T(plane_unit_norm.z)};
    residual[0] = vdot(norm, point_w) + T(negative_OA_dot_norm);
    return true;
  }
};

// ceres::AutoDiffCostFunction<Functor, NR, 4, 3>
template <typename Functor, int NR>
struct AutoDiffBlock43 : CostBlock {
  Functor f;
  explicit AutoDiffBlock43(const Functor& f_) : f(f_) {}
  int num_residuals() const override { return NR; }
  void evaluate(const double* x, double* r, double* J) const override {
    if (!J) {
      f(x, x + 4, r);
      return;
    }
    typedef Jet<7> JT;
    JT q[4], t[3], res[NR];
    for (int i = 0; i < 4; ++i) q[i] = JT(x[i], i);
    for (int i = 0; i < 3; ++i) t[i] = JT(x[4 + i], 4 + i);
    f(q, t, res);
    for (int i = 0; i < NR; ++i) {
      r[i] = res[i].a;
      for (int c = 0; c < 7; ++c) J[i * 7 + c] = res[i].v[c];
    }
  }
};

}  // namespace oracle
