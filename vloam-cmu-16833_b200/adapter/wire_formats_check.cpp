// Exercises adapter/wire_formats.hpp without ROS or CUDA (built by __graft_entry__.build(), run by tests/test_wire_formats.py).
//   wire_formats_check bin   FILE             -> "<n> <sum x> <sum y> <sum z> <sum i>"
//   wire_formats_check pc2   FILE STEP OX OY OZ N  -> "<zero_copy> <stride> <sum x> <sum y> <sum z>"
//   wire_formats_check poses FILE             -> reads lines "qx qy qz qw tx ty tz" (world_T_base), the first line is base_T_cam0;
//                                               prints the KITTI-format dump
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>

#include "wire_formats.hpp"

using namespace vloam_b200;

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  const std::string mode = argv[1];
  if (mode == "bin") {
    std::vector<float> p;
    const int n = load_kitti_bin(argv[2], p);
    double s[4] = {0, 0, 0, 0};
    for (int i = 0; i < n; ++i) for (int k = 0; k < 4; ++k) s[k] += p[4 * i + k];
    std::printf("%d %.6f %.6f %.6f %.6f\n", n, s[0], s[1], s[2], s[3]);
    return 0;
  }
  if (mode == "pc2" && argc >= 8) {
    std::ifstream f(argv[2], std::ios::binary);
    std::vector<uint8_t> data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    const uint32_t step = (uint32_t)std::atoi(argv[3]);
    std::vector<PointFieldDesc> fields = {{"x", (uint32_t)std::atoi(argv[4]), 7}, {"y", (uint32_t)std::atoi(argv[5]), 7},
                                          {"z", (uint32_t)std::atoi(argv[6]), 7}, {"intensity", 12, 7}};
    PointCloud2View v;
    if (!v.bind(data.data(), (size_t)std::atoi(argv[7]), step, fields)) return 3;
    double s[3] = {0, 0, 0};
    for (int i = 0; i < v.n; ++i) for (int k = 0; k < 3; ++k) s[k] += v.points[(size_t)i * v.stride_floats + k];
    std::printf("%d %d %.6f %.6f %.6f\n", v.zero_copy ? 1 : 0, v.stride_floats, s[0], s[1], s[2]);
    return 0;
  }
  if (mode == "poses") {
    std::ifstream f(argv[2]);
    std::string line;
    bool first = true;
    int count = 0;
    Cam0StartFrameWriter* w = nullptr;
    while (std::getline(f, line)) {
      std::istringstream is(line);
      double q[4], t[3];
      is >> q[0] >> q[1] >> q[2] >> q[3] >> t[0] >> t[1] >> t[2];
      if (first) { w = new Cam0StartFrameWriter(Mat4::from_qt(q, t)); first = false; continue; }
      std::fputs(w->write(nullptr, count++, Mat4::from_qt(q, t)).c_str(), stdout);
    }
    delete w;
    return 0;
  }
  return 2;
}
