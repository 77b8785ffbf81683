#pragma once
#include <memory>
#include <vector>
namespace pcl {
template <typename T>
struct PointCloud {
  using Ptr = std::shared_ptr<PointCloud<T>>;
  std::vector<T> points;
  size_t size() const { return points.size(); }
  void clear() { points.clear(); }
  void push_back(const T& p) { points.push_back(p); }
  T& operator[](size_t i) { return points[i]; }
  const T& operator[](size_t i) const { return points[i]; }
};
}  // namespace pcl
