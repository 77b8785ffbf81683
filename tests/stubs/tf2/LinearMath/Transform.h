// stub of the tf2 linear algebra the adapters touch: quaternion (x, y, z, w), vector, rigid transform with *, inverse()
#pragma once
#include <cmath>
namespace tf2 {
class Vector3 {
 public:
  Vector3(double x = 0, double y = 0, double z = 0) : v_{x, y, z} {}
  double x() const { return v_[0]; } double y() const { return v_[1]; } double z() const { return v_[2]; }
  double getX() const { return v_[0]; } double getY() const { return v_[1]; } double getZ() const { return v_[2]; }
  Vector3 operator+(const Vector3& o) const { return {v_[0] + o.v_[0], v_[1] + o.v_[1], v_[2] + o.v_[2]}; }
  Vector3 operator-() const { return {-v_[0], -v_[1], -v_[2]}; }
 private:
  double v_[3];
};
class Quaternion {
 public:
  Quaternion(double x = 0, double y = 0, double z = 0, double w = 1) : q_{x, y, z, w} {}
  double x() const { return q_[0]; } double y() const { return q_[1]; } double z() const { return q_[2]; } double w() const { return q_[3]; }
  Quaternion inverse() const { return {-q_[0], -q_[1], -q_[2], q_[3]}; }
  void setRotation(const Vector3& axis, double angle) {        // tf2::Quaternion::setRotation: the axis is normalised by its length
    const double d = std::sqrt(axis.x() * axis.x() + axis.y() * axis.y() + axis.z() * axis.z()), s = std::sin(angle * 0.5) / d;
    q_[0] = axis.x() * s; q_[1] = axis.y() * s; q_[2] = axis.z() * s; q_[3] = std::cos(angle * 0.5);
  }
  double getAngle() const { return 2.0 * std::acos(q_[3]); }
  Vector3 getAxis() const {
    const double s2 = 1.0 - q_[3] * q_[3];
    if (s2 < 1e-14) return Vector3(1.0, 0.0, 0.0);
    const double s = 1.0 / std::sqrt(s2);
    return Vector3(q_[0] * s, q_[1] * s, q_[2] * s);
  }
  Quaternion operator*(const Quaternion& b) const {
    return {q_[3] * b.q_[0] + q_[0] * b.q_[3] + q_[1] * b.q_[2] - q_[2] * b.q_[1], q_[3] * b.q_[1] + q_[1] * b.q_[3] + q_[2] * b.q_[0] - q_[0] * b.q_[2],
            q_[3] * b.q_[2] + q_[2] * b.q_[3] + q_[0] * b.q_[1] - q_[1] * b.q_[0], q_[3] * b.q_[3] - q_[0] * b.q_[0] - q_[1] * b.q_[1] - q_[2] * b.q_[2]};
  }
  Vector3 rotate(const Vector3& v) const {
    const double ux = q_[0], uy = q_[1], uz = q_[2], w = q_[3];
    double cx = uy * v.z() - uz * v.y(), cy = uz * v.x() - ux * v.z(), cz = ux * v.y() - uy * v.x();
    cx += cx; cy += cy; cz += cz;
    return {v.x() + w * cx + (uy * cz - uz * cy), v.y() + w * cy + (uz * cx - ux * cz), v.z() + w * cz + (ux * cy - uy * cx)};
  }
 private:
  double q_[4];
};
class Transform {
 public:
  Transform() = default;
  Transform(const Quaternion& q, const Vector3& t) : q_(q), t_(t) {}
  void setIdentity() { q_ = Quaternion(); t_ = Vector3(); }
  void setOrigin(const Vector3& t) { t_ = t; }
  void setRotation(const Quaternion& q) { q_ = q; }
  const Vector3& getOrigin() const { return t_; }
  Quaternion getRotation() const { return q_; }
  Transform inverse() const { const Quaternion qi = q_.inverse(); return Transform(qi, -qi.rotate(t_)); }
  Transform operator*(const Transform& o) const { return Transform(q_ * o.q_, q_.rotate(o.t_) + t_); }
 private:
  Quaternion q_;
  Vector3 t_;
};
}  // namespace tf2
