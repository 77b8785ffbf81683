// ORACLE — TEST INFRASTRUCTURE ONLY (see types.hpp header).  PARITY UNPINNED.
//
// voxel_grid.hpp: restatement of pcl::VoxelGrid<pcl::PointXYZI>::applyFilter
// with the defaults the reference uses (downsample_all_data=true,
// min_points_per_voxel=0, no filter field) — third party PCL 1.8-1.10 (README.md:25,
// not vendored).  Call sites: scan_registration.cpp:433-437 (leaf 0.2),
// laser_mapping.cpp:100-101,433-439,694-700 (leaf = line/plane resolution).
//
//   bbox (float min/max) -> overflow guard on dx*dy*dz (returns the input
//   unfiltered) -> min_b = floor(min*inv_leaf) -> per point
//   ijk = int(floor(p*inv_leaf) - float(min_b)); key = i + j*dx + k*dx*dy ->
//   sort by key -> per key the float sum of x,y,z,intensity divided by float(n),
//   emitted in ascending key order.
//
// Deviation (documented, DESIGN.md "Q-VG"): PCL sorts with std::sort on the key
// only, so the float summation order inside a voxel is implementation-defined.
// Here (and in the CUDA path) points of one voxel are summed in ascending input
// index.  `literal_unstable=true` reproduces the literal std::sort order of
// this toolchain instead, for the test that bounds the difference (<= a few ulp).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

#include "types.hpp"

namespace oracle {

inline void voxel_grid_filter(const Cloud& in, float leaf, Cloud* out, bool literal_unstable = false) {
  out->clear();
  if (in.empty()) return;
  const float inv = 1.0f / leaf;  // Eigen::Array4f::Ones() / leaf_size_.array()
  float mn[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
  float mx[3] = {-std::numeric_limits<float>::max(), -std::numeric_limits<float>::max(), -std::numeric_limits<float>::max()};
  for (const PointXYZI& p : in) {
    mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
    mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
  }
  const int64_t dx = static_cast<int64_t>((mx[0] - mn[0]) * inv) + 1;
  const int64_t dy = static_cast<int64_t>((mx[1] - mn[1]) * inv) + 1;
  const int64_t dz = static_cast<int64_t>((mx[2] - mn[2]) * inv) + 1;
  if ((dx * dy * dz) > static_cast<int64_t>(std::numeric_limits<int32_t>::max())) {
    *out = in;  // "Leaf size is too small for the input dataset": PCL returns the input
    return;
  }
  int min_b[3], max_b[3], div_b[3];
  for (int a = 0; a < 3; ++a) {
    min_b[a] = static_cast<int>(std::floor(mn[a] * inv));
    max_b[a] = static_cast<int>(std::floor(mx[a] * inv));
    div_b[a] = max_b[a] - min_b[a] + 1;
  }
  const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
  struct Item { unsigned int idx; unsigned int pt; };
  std::vector<Item> items;
  items.reserve(in.size());
  for (size_t i = 0; i < in.size(); ++i) {
    const PointXYZI& p = in[i];
    const int ijk0 = static_cast<int>(std::floor(p.x * inv) - static_cast<float>(min_b[0]));
    const int ijk1 = static_cast<int>(std::floor(p.y * inv) - static_cast<float>(min_b[1]));
    const int ijk2 = static_cast<int>(std::floor(p.z * inv) - static_cast<float>(min_b[2]));
    const int idx = ijk0 * mul[0] + ijk1 * mul[1] + ijk2 * mul[2];
    items.push_back({static_cast<unsigned int>(idx), static_cast<unsigned int>(i)});
  }
  if (literal_unstable) {
    std::sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.idx < b.idx; });
  } else {
    std::sort(items.begin(), items.end(),
              [](const Item& a, const Item& b) { return a.idx != b.idx ? a.idx < b.idx : a.pt < b.pt; });
  }
  size_t index = 0;
  while (index < items.size()) {
    size_t i = index + 1;
    while (i < items.size() && items[i].idx == items[index].idx) ++i;
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    for (size_t li = index; li < i; ++li) {
      const PointXYZI& p = in[items[li].pt];
      sx += p.x; sy += p.y; sz += p.z; si += p.intensity;
    }
    const float n = static_cast<float>(i - index);
    out->push_back({sx / n, sy / n, sz / n, si / n});
    index = i;
  }
}

}  // namespace oracle
