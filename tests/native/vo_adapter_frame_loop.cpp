// CI build of adapter/visual_odometry_b200.h (the drop-in vloam::VisualOdometry) against tests/stubs — no ROS / PCL / OpenCV in this
// image — driven through the call sequence of the reference's caller, vloam_main_node.cpp:134-166:
// reset, processImage, setUpPointCloud, processPointCloud, solveNlsAll, publish.  The stub ImageUtil aborts when called, so a
// run that completes never left the device path of processImage.
//   vo_adapter_frame_loop compile-only                               -> prints the topic the constructor advertises (no GPU needed)
//   vo_adapter_frame_loop run H W calib.txt img0 scan0 img1 scan1 .. -> runs the frames on the GPU, prints features / matches / motion
// img*: H * W bytes; scan*: packed float32 x y z; calib.txt: cam_T_velo (16), R (9), P (12) as text.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <vector>

#include "../../vloam-cmu-16833_b200/adapter/visual_odometry_b200.h"

static std::vector<char> slurp(const char* path) {
  std::ifstream f(path, std::ios::binary);
  return std::vector<char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

int main(int argc, char** argv) {
  auto& p = ros::stub::params();
  p["loam_verbose_level"] = 0; p["reset_VO_to_identity"] = 1; p["remove_VO_outlier"] = 100; p["keypoint_NMS"] = 0; p["CLAHE"] = 0;
  p["visualize_optical_flow"] = 0; p["optical_flow_match"] = 0;
  auto tf = std::make_shared<vloam::VloamTF>();
  vloam::VisualOdometry vo;
  if (argc < 6 || std::strcmp(argv[1], "compile-only") == 0) {
    for (const auto& t : ros::stub::topics()) std::printf("topic %s\n", t.first.c_str());
    return 0;
  }
  const int H = std::atoi(argv[2]), W = std::atoi(argv[3]);
  auto info = std::make_shared<sensor_msgs::CameraInfo>();
  {
    std::ifstream c(argv[4]);
    for (int r = 0; r < 4; ++r) for (int k = 0; k < 4; ++k) c >> tf->imu_eigen_T_velo.matrix()(r, k);     // imu_T_cam0 = I: cam_T_velo = imu_T_velo
    for (int k = 0; k < 9; ++k) c >> info->R[k];
    for (int k = 0; k < 12; ++k) c >> info->P[k];
  }
  vo.init(tf);
  for (int a = 5, k = 0; a + 1 < argc; a += 2, ++k) {
    std::vector<char> img = slurp(argv[a]), raw = slurp(argv[a + 1]);
    pcl::PointCloud<pcl::PointXYZ> cloud;
    cloud.points.resize(raw.size() / 12);
    for (size_t i = 0; i < cloud.points.size(); ++i) { float v[3]; std::memcpy(v, raw.data() + 12 * i, 12); cloud.points[i].x = v[0]; cloud.points[i].y = v[1]; cloud.points[i].z = v[2]; }
    auto msg = std::make_shared<sensor_msgs::PointCloud2>();
    vo.reset();                                                                                  // vloam_main_node.cpp:134
    vo.processImage(cv::Mat(H, W, CV_8UC1, img.data()));                                         // :139
    vo.setUpPointCloud(info);                                                                    // :140
    vo.processPointCloud(msg, cloud, false, true);                                               // :149
    if (vo.count > 0) vo.solveNlsAll();                                                          // :158
    vo.publish();
    unsigned sum = 0;
    const cv::Mat& d = vo.descriptors[vo.i];
    for (int i = 0; i < d.rows * d.cols; ++i) sum = sum * 31u + d.data[i];
    unsigned msum = 0;
    for (const cv::DMatch& m : vo.matches) msum = (msum * 31u + (unsigned)m.queryIdx) * 31u + (unsigned)m.trainIdx;
    std::printf("vo %d keypoints %zu rows %d desc %u matches %zu msum %u motion %.17g %.17g %.17g %.17g %.17g %.17g\n", k, vo.keypoints[vo.i].size(), d.rows, sum,
                vo.matches.size(), msum, vo.angles_0to1[0], vo.angles_0to1[1], vo.angles_0to1[2], vo.t_0to1[0], vo.t_0to1[1], vo.t_0to1[2]);
  }
  for (const auto& t : ros::stub::topics()) std::printf("published %s %zu\n", t.first.c_str(), t.second.count);
  return 0;
}
