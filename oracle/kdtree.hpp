// ORACLE — TEST INFRASTRUCTURE ONLY (see types.hpp header).  PARITY UNPINNED.
//
// kdtree.hpp: exact k-nearest-neighbour search with the result contract of
// pcl::KdTreeFLANN<PointXYZI>::nearestKSearch (third party PCL + FLANN, not
// vendored; call sites laser_odometry.cpp:269,356,525-526 and
// laser_mapping.cpp:452-453,477,543): exact search (eps = 0) over x,y,z only,
// squared distances accumulated in float exactly like flann::L2_Simple
// (result = 0; for each dim: diff = a-b; result += diff*diff), results sorted
// ascending.  Ties (measure-zero for the noisy synthetic data) are broken by
// the lower point index — FLANN's own tie order depends on its tree layout.
//
// Also the structure timed as the CPU baseline: a kd-tree is (re)built for
// every target cloud, as the reference does every scan.
#pragma once
#include <algorithm>
#include <cstdint>
#include <limits>
#include <numeric>
#include <vector>

#include "types.hpp"

namespace oracle {

inline float l2_simple(const float* a, const PointXYZI& b) {
  float result = 0.f;
  float diff = a[0] - b.x; result += diff * diff;
  diff = a[1] - b.y; result += diff * diff;
  diff = a[2] - b.z; result += diff * diff;
  return result;
}

class KdTree {
 public:
  void set_input(const Cloud* cloud) {
    cloud_ = cloud;
    const int n = static_cast<int>(cloud->size());
    idx_.resize(n);
    std::iota(idx_.begin(), idx_.end(), 0);
    nodes_.clear();
    nodes_.reserve(n / 4 + 8);
    if (n > 0) build(0, n);
  }
  // k nearest, sorted by (distance, index).  Returns the number found (min(k, size)).
  int nearest_k(const PointXYZI& q, int k, int* out_idx, float* out_d) const {
    const float qq[3] = {q.x, q.y, q.z};
    Heap h{k, 0, out_idx, out_d};
    if (!nodes_.empty()) search(0, qq, h);
    return h.n;
  }

 private:
  struct Node { int lo, hi, left, right, dim; float split; };
  struct Heap {  // small sorted array, ascending by (d, idx)
    int k, n; int* idx; float* d;
    float worst() const { return n < k ? std::numeric_limits<float>::infinity() : d[n - 1]; }
    void add(float dist, int i) {
      if (n == k && !(dist < d[n - 1] || (dist == d[n - 1] && i < idx[n - 1]))) return;
      int pos = n < k ? n++ : k - 1;
      while (pos > 0 && (d[pos - 1] > dist || (d[pos - 1] == dist && idx[pos - 1] > i))) {
        d[pos] = d[pos - 1]; idx[pos] = idx[pos - 1]; --pos;
      }
      d[pos] = dist; idx[pos] = i;
    }
  };
  static constexpr int kLeaf = 12;

  int build(int lo, int hi) {
    const int id = static_cast<int>(nodes_.size());
    nodes_.push_back({lo, hi, -1, -1, -1, 0.f});
    if (hi - lo <= kLeaf) return id;
    float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
    for (int i = lo; i < hi; ++i) {
      const PointXYZI& p = (*cloud_)[idx_[i]];
      mn[0] = std::min(mn[0], p.x); mx[0] = std::max(mx[0], p.x);
      mn[1] = std::min(mn[1], p.y); mx[1] = std::max(mx[1], p.y);
      mn[2] = std::min(mn[2], p.z); mx[2] = std::max(mx[2], p.z);
    }
    int dim = 0;
    if (mx[1] - mn[1] > mx[dim] - mn[dim]) dim = 1;
    if (mx[2] - mn[2] > mx[dim] - mn[dim]) dim = 2;
    if (!(mx[dim] - mn[dim] > 0.f)) return id;  // all coincident: keep as leaf
    const int mid = (lo + hi) / 2;
    auto coord = [&](int i) { const PointXYZI& p = (*cloud_)[i]; return dim == 0 ? p.x : dim == 1 ? p.y : p.z; };
    std::nth_element(idx_.begin() + lo, idx_.begin() + mid, idx_.begin() + hi,
                     [&](int a, int b) { return coord(a) < coord(b); });
    const float split = coord(idx_[mid]);
    const int l = build(lo, mid);
    const int r = build(mid, hi);
    nodes_[id].left = l; nodes_[id].right = r; nodes_[id].dim = dim; nodes_[id].split = split;
    return id;
  }
  void search(int id, const float* q, Heap& h) const {
    const Node& nd = nodes_[id];
    if (nd.left < 0) {
      for (int i = nd.lo; i < nd.hi; ++i) h.add(l2_simple(q, (*cloud_)[idx_[i]]), idx_[i]);
      return;
    }
    const double diff = static_cast<double>(q[nd.dim]) - static_cast<double>(nd.split);
    const int first = diff < 0 ? nd.left : nd.right;
    const int second = diff < 0 ? nd.right : nd.left;
    search(first, q, h);
    // lower bound on the exact distance to anything on the far side; small slack so
    // float rounding of the accumulated distance can never make the prune unsafe.
    if (diff * diff * (1.0 - 1e-6) <= static_cast<double>(h.worst())) search(second, q, h);
  }

  const Cloud* cloud_ = nullptr;
  std::vector<int> idx_;
  std::vector<Node> nodes_;
};

// Brute-force reference used by the tests to validate KdTree.
inline int brute_nearest_k(const Cloud& cloud, const PointXYZI& q, int k, int* out_idx, float* out_d) {
  const float qq[3] = {q.x, q.y, q.z};
  int n = 0;
  for (int i = 0; i < static_cast<int>(cloud.size()); ++i) {
    const float dist = l2_simple(qq, cloud[i]);
    if (n == k && !(dist < out_d[n - 1])) continue;  // ascending index => ties keep the earlier
    int pos = n < k ? n++ : k - 1;
    while (pos > 0 && out_d[pos - 1] > dist) { out_d[pos] = out_d[pos - 1]; out_idx[pos] = out_idx[pos - 1]; --pos; }
    out_d[pos] = dist; out_idx[pos] = i;
  }
  return n;
}

}  // namespace oracle
