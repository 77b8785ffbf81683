#pragma once
#include <pcl/point_cloud.h>
#include <sensor_msgs/PointCloud2.h>
#include <cstring>
namespace pcl {
template <typename T>
void toROSMsg(const PointCloud<T>& c, sensor_msgs::PointCloud2& m) {
  m.width = (uint32_t)c.size(); m.point_step = sizeof(T); m.row_step = m.width * m.point_step;
  m.data.resize((size_t)m.row_step);
  if (!c.points.empty()) std::memcpy(m.data.data(), c.points.data(), m.data.size());
}
}  // namespace pcl
