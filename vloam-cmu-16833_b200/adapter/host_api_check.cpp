// Compile/link check of the ROS-free host mirror against libvloam_b200.so (built by __graft_entry__.build()).
#include <cstdio>

#include "host_api.hpp"
#include "lidar_odometry_mapping_b200.h"

int main() {
  try {
    vloam_b200::LidarOdometryMapping lom(0);
    lom.params().max_points = 4096;
    lom.init();
    std::printf("context created\n");
  } catch (const std::exception& e) {
    std::printf("no device: %s\n", e.what());  // expected on a CPU-only box: the library refuses to run, no fallback
  }
  return 0;
}
