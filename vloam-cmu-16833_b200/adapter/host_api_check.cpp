// The ROS-free C++ host mirror (host_api.hpp, wire_formats.hpp) against libvloam_b200.so.
//   host_api_check                         compile / link / load check (built and run by __graft_entry__.build(); on a CPU-only
//                                          box the library refuses to create a context: there is no fallback)
//   host_api_check run A.bin B.bin ...     the frame loop of vloam_main_node.cpp:125-180 over KITTI .bin scans, in C++:
//                                          LidarOdometryMapping::{reset, scanRegistrationIO, laserOdometryIO, laserMappingIO} and
//                                          VisualOdometry::{reset, processPointCloud, queryDepth}; prints one line per frame
//                                          (tests/test_gpu_host_api.py compares them with the Python mirror's results)
#include <cstdio>
#include <cstring>

#include "host_api.hpp"
#include "lidar_odometry_mapping_b200.h"
#include "visual_odometry_b200.h"
#include "wire_formats.hpp"

int main(int argc, char** argv) {
  try {
    vloam_b200::LidarOdometryMapping lom(0);
    lom.params().max_points = 1 << 17;
    lom.params().map_capacity_points = 1 << 17;
    lom.init();
    vloam_b200::VisualOdometry vo(1 << 17, 1024, 0);
    std::printf("context created\n");
    if (argc < 3 || std::strcmp(argv[1], "run") != 0) return 0;
    // KITTI-like calibration (vloam_b200/synth.py: kitti_like_calibration)
    const float cam_T_velo[16] = {0, -1, 0, 0, 0, 0, -1, -0.08f, 1, 0, 0, -0.27f, 0, 0, 0, 1};
    const float rect0_T_cam[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    const float P_rect0[12] = {718.856f, 0, 607.1928f, 0, 0, 718.856f, 185.2157f, 0, 0, 0, 1, 0};
    vo.setUpPointCloud(cam_T_velo, rect0_T_cam, P_rect0);
    vloam_b200::Cam0StartFrameWriter dump(vloam_b200::Mat4::identity());
    for (int k = 2; k < argc; ++k) {
      std::vector<float> xyzi;
      const int n = vloam_b200::load_kitti_bin(argv[k], xyzi);
      if (n <= 0) { std::printf("cannot read %s\n", argv[k]); return 3; }
      vo.reset();
      lom.reset();
      vo.processPointCloud(xyzi.data(), n, 4);
      const float z = vo.queryDepth(620.f, 250.f);
      lom.scanRegistrationIO(xyzi.data(), n, 4);
      lom.laserOdometryIO();
      lom.laserMappingIO();
      std::printf("frame %d n %d depth %.9g lo_t %.17g %.17g %.17g lo_q %.17g %.17g %.17g %.17g mo_t %.17g %.17g %.17g corr %d %d less_flat %zu\n", k - 2,
                  n, (double)z, lom.odom.t[0], lom.odom.t[1], lom.odom.t[2], lom.odom.q[0], lom.odom.q[1], lom.odom.q[2], lom.odom.q[3],
                  lom.mapped.t[0], lom.mapped.t[1], lom.mapped.t[2], lom.corner_correspondence, lom.plane_correspondence,
                  lom.cloud(VLOAM_CLOUD_LESS_FLAT).size() / 4);
      std::fputs(("pose " + dump.write(nullptr, k - 2, vloam_b200::Mat4::from_qt(lom.mapped.q.data(), lom.mapped.t.data()))).c_str(), stdout);
    }
    {  // ImageUtil::detKeypoints / matchDescriptors through the mirror, on a deterministic checkerboard-with-ramp image
      const int H = 120, W = 200;
      std::vector<uint8_t> img((size_t)H * W);
      for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) img[(size_t)y * W + x] = (uint8_t)((x * 3 + y * 5 + (((x / 9) + (y / 7)) % 2) * 110) % 256);
      const std::vector<float> xy = vo.detKeypoints(img.data(), H, W);
      std::printf("corners %zu", xy.size() / 2);
      for (size_t i = 0; i < xy.size() && i < 12; ++i) std::printf(" %g", (double)xy[i]);
      std::printf("\n");
      std::vector<uint8_t> d0(40 * 32), d1(50 * 32);
      for (size_t i = 0; i < d0.size(); ++i) d0[i] = (uint8_t)((i * 37 + (i / 32) * 11) % 251);
      for (size_t i = 0; i < d1.size(); ++i) d1[i] = (uint8_t)(((i % (40 * 32)) * 37 + ((i % (40 * 32)) / 32) * 11) % 251 ^ ((i / 32) % 3 == 0 ? 1 : 0));
      const std::vector<int> m = vo.matchDescriptors(d0.data(), 40, d1.data(), 50);
      std::printf("matches %zu", m.size() / 3);
      for (size_t i = 0; i < m.size() && i < 9; ++i) std::printf(" %d", m[i]);
      std::printf("\n");
      // ImageUtil::descKeypoints (ORB) on those corners, then VisualOdometry::processImage over two frames (the second one moved by 3 px)
      const vloam_b200::VisualOdometry::Features f = vo.descKeypoints(img.data(), H, W, xy.data(), (int)(xy.size() / 2));
      unsigned sum = 0;
      for (size_t i = 0; i < f.descriptors.size(); ++i) sum = sum * 31u + f.descriptors[i];
      std::printf("described %d checksum %u\n", f.size(), sum);
      std::vector<uint8_t> img2((size_t)H * W);
      for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) img2[(size_t)y * W + x] = img[(size_t)y * W + (x + 3) % W];
      vo.reset();
      const vloam_b200::VisualOdometry::Frame f0 = vo.processImage(img.data(), H, W);
      vo.reset();
      const vloam_b200::VisualOdometry::Frame f1 = vo.processImage(img2.data(), H, W);
      unsigned msum = 0;
      for (size_t i = 0; i < f1.matches.size(); ++i) msum = msum * 31u + (unsigned)f1.matches[i];
      std::printf("chain %d %zu %d %zu checksum %u\n", f0.features.size(), f0.matches.size() / 3, f1.features.size(), f1.matches.size() / 3, msum);
    }
  } catch (const std::exception& e) {
    std::printf("no device: %s\n", e.what());  // expected on a CPU-only box: the library refuses to run, no fallback
    return argc >= 3 ? 4 : 0;
  }
  return 0;
}
