// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
//
// CPU restatement of the VLOAM per-scan hot path (reference:
// YukunXia/VLOAM-CMU-16833 @ /root/reference).  Only tests/, bench.py's
// cpu_baseline / --impl reference leg and __graft_entry__.smoke() may use it,
// and only as the checker / reported baseline.
//
// PARITY UNPINNED: the reference ships no tests or golden vectors and its
// third-party arithmetic (PCL VoxelGrid / KdTreeFLANN, Ceres 2.0 Solve, Eigen
// 3.3) is not vendored and not installed here, so this restatement could not be
// run against the real libraries.  Its fidelity rests on the cross-checks in
// tests/ (dual-number Jacobians vs analytic, kNN vs brute force and scipy,
// converged pose vs scipy.optimize and ground truth, voxel grid vs numpy).
//
// types.hpp: plain point / pose types shared by every oracle stage.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace oracle {

// pcl::PointXYZI as used through `typedef pcl::PointXYZI PointType`
// (reference include/lidar_odometry_mapping/common.h:43): 16-byte record.
struct PointXYZI {
  float x, y, z, intensity;
};
using Cloud = std::vector<PointXYZI>;

struct Vec3 {
  double x = 0, y = 0, z = 0;
};
inline Vec3 operator+(const Vec3& a, const Vec3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(const Vec3& a, const Vec3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator*(double s, const Vec3& a) { return {s * a.x, s * a.y, s * a.z}; }
inline double dot(const Vec3& a, const Vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross(const Vec3& a, const Vec3& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline double norm(const Vec3& a) { return std::sqrt(dot(a, a)); }

// Eigen::Quaterniond semantics, coefficient storage order (x, y, z, w).
struct Quat {
  double x = 0, y = 0, z = 0, w = 1;
};
// Eigen quaternion product (Hamilton).
inline Quat operator*(const Quat& a, const Quat& b) {
  return {a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
          a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x,
          a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
// Eigen QuaternionBase::_transformVector: v + 2w(u x v) + 2 u x (u x v).
inline Vec3 rotate(const Quat& q, const Vec3& v) {
  Vec3 u{q.x, q.y, q.z};
  Vec3 uv = cross(u, v);
  uv = uv + uv;
  return v + q.w * uv + cross(u, uv);
}
// Eigen Quaternion::inverse(): conjugate / squaredNorm.
inline Quat inverse(const Quat& q) {
  double n2 = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
  return {-q.x / n2, -q.y / n2, -q.z / n2, q.w / n2};
}

}  // namespace oracle
