#pragma once
#include <std_msgs/Header.h>
#include <memory>
namespace sensor_msgs {
struct CameraInfo { std_msgs::Header header; double K[9] = {0}, R[9] = {0}, P[12] = {0}; };
typedef std::shared_ptr<const CameraInfo> CameraInfoConstPtr;
}
