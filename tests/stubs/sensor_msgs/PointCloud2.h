#pragma once
#include <std_msgs/Header.h>
#include <cstdint>
#include <vector>
#include <memory>
namespace sensor_msgs {
struct PointCloud2;
typedef std::shared_ptr<const PointCloud2> PointCloud2ConstPtr;
struct PointCloud2 { std_msgs::Header header; uint32_t height = 1, width = 0, point_step = 0, row_step = 0; std::vector<uint8_t> data; };
}
