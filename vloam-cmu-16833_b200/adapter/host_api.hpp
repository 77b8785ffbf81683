// vloam_b200 — ROS-free C++ host mirror of the reference's LiDAR façade over the C ABI.
//
// Same method names, argument meaning and call order as vloam::LidarOdometryMapping
// (reference include/lidar_odometry_mapping/lidar_odometry_mapping.h:45-86), with pcl::PointCloud replaced by plain
// float arrays (x, y, z[, pad] per point) so it compiles without ROS / PCL.  The ROS adapter (lidar_odometry_mapping_b200.h)
// is a thin wrapper around this class.
#pragma once
#include <array>
#include <algorithm>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/vloam_b200.h"

namespace vloam_b200 {

struct Pose {
  std::array<double, 4> q{{0, 0, 0, 1}};  // x, y, z, w (Eigen::Quaterniond::coeffs order)
  std::array<double, 3> t{{0, 0, 0}};
};

class LidarOdometryMapping {
 public:
  explicit LidarOdometryMapping(int device = 0) {
    if (vloam_ctx_create(device, &ctx_) != VLOAM_OK) throw std::runtime_error("vloam_ctx_create failed (no CUDA device?)");
    vloam_lidar_params_default(&params_);
  }
  ~LidarOdometryMapping() {
    if (h_) vloam_lidar_destroy(h_);
    if (ctx_) vloam_ctx_destroy(ctx_);
  }
  LidarOdometryMapping(const LidarOdometryMapping&) = delete;
  LidarOdometryMapping& operator=(const LidarOdometryMapping&) = delete;

  vloam_lidar_params& params() { return params_; }  // fill from the ROS parameter server before init()

  // lidar_odometry_mapping.cpp:40-63
  void init() {
    if (h_) { vloam_lidar_destroy(h_); h_ = nullptr; }
    check(vloam_lidar_create(ctx_, &params_, &h_));
  }
  // lidar_odometry_mapping.cpp:65-71
  void reset() { check(vloam_lidar_reset(h_)); }
  // lidar_odometry_mapping.cpp:73-94: `points` = n records of `stride_floats` floats (4 for pcl::PointXYZ)
  void scanRegistrationIO(const float* points, int n, int stride_floats = 4) {
    check(vloam_scan_registration(h_, points, &n, stride_floats, (size_t)n));
  }
  // lidar_odometry_mapping.cpp:96-123: vo_prior = velo_last_VOT_velo_curr (q xyzw, t) or nullptr
  void laserOdometryIO(const double* vo_prior = nullptr) {
    double p[14];
    int c[2];
    check(vloam_laser_odometry(h_, vo_prior, p, c));
    for (int i = 0; i < 4; ++i) { last_curr.q[i] = p[i]; odom.q[i] = p[7 + i]; }
    for (int i = 0; i < 3; ++i) { last_curr.t[i] = p[4 + i]; odom.t[i] = p[11 + i]; }
    corner_correspondence = c[0]; plane_correspondence = c[1];
  }
  // lidar_odometry_mapping.cpp:125-154
  void laserMappingIO() {
    double p[14];
    check(vloam_laser_mapping(h_, p));
    for (int i = 0; i < 4; ++i) { mapped.q[i] = p[i]; wmap_wodom.q[i] = p[7 + i]; }
    for (int i = 0; i < 3; ++i) { mapped.t[i] = p[4 + i]; wmap_wodom.t[i] = p[11 + i]; }
  }
  // per-stream status of the last laserMappingIO (VLOAM_LM_* bits; 0 = the scan was mapped and inserted)
  int mappingStatus() {
    int st = 0;
    check(vloam_get_lm_status(h_, &st));
    return st;
  }
  // ScanRegistration::output / LaserOdometry::output / LaserMapping::publish clouds as XYZI records
  std::vector<float> cloud(int which) {
    int n = 0;
    check(vloam_get_cloud(h_, 0, which, nullptr, 0, &n));
    std::vector<float> out((size_t)n * 4);
    if (n) check(vloam_get_cloud(h_, 0, which, out.data(), n, &n));
    return out;
  }

  Pose last_curr;   // q_last_curr / t_last_curr  -> vloam_tf->base_prev_LOT_base_curr   (laser_odometry.cpp:563-567)
  Pose odom;        // q_w_curr / t_w_curr        -> vloam_tf->world_LOT_base_last        (laser_odometry.cpp:570-571)
  Pose mapped;      // mapping q_w_curr / t_w_curr -> vloam_tf->world_MOT_base_last       (laser_mapping.cpp:728-729)
  Pose wmap_wodom;
  int corner_correspondence = 0, plane_correspondence = 0;

 private:
  void check(int rc) {
    if (rc != VLOAM_OK) throw std::runtime_error(std::string("vloam_b200: ") + vloam_last_error(ctx_));
  }
  vloam_ctx* ctx_ = nullptr;
  vloam_lidar* h_ = nullptr;
  vloam_lidar_params params_{};
};

// ROS-free mirror of vloam::VisualOdometry's depth-association / solve half
// (reference include/visual_odometry/visual_odometry.h:40-121) and of processImage in the reference's configuration
// (Shi-Tomasi + ORB + BF / kNN: detKeypoints, descKeypoints, matchDescriptors, or the whole chain in one call).  One instance per sensor; shares the context of a LidarOdometryMapping when
// given one, so that the scan uploaded by scanRegistrationIO is also VO's cloud (vloam_main_node.cpp:148-166).
class VisualOdometry {
 public:
  explicit VisualOdometry(int max_points = 1 << 18, int max_matches = 2048, int device = 0) : max_matches_(max_matches) {
    if (vloam_ctx_create(device, &ctx_) != VLOAM_OK) throw std::runtime_error("vloam_ctx_create failed (no CUDA device?)");
    check(vloam_vo_create(ctx_, 1, max_points, max_matches, &h_));
  }
  ~VisualOdometry() {
    if (h_) vloam_vo_destroy(h_);
    if (ctx_) vloam_ctx_destroy(ctx_);
  }
  VisualOdometry(const VisualOdometry&) = delete;
  VisualOdometry& operator=(const VisualOdometry&) = delete;

  // visual_odometry.cpp:86-90
  void reset() { check(vloam_vo_reset(h_)); ++count; }
  // visual_odometry.cpp:132-155: row-major cam_T_velo (4x4), rect0_T_cam (4x4; the ROS path leaves (3,3) = 0), P_rect0 (3x4)
  void setUpPointCloud(const float cam_T_velo[16], const float rect0_T_cam[16], const float P_rect0[12]) {
    check(vloam_vo_set_calibration(h_, cam_T_velo, rect0_T_cam, P_rect0));
  }
  // visual_odometry.cpp:157-186: `points` = n records of `stride_floats` floats starting with x, y, z
  void processPointCloud(const float* points, int n, int stride_floats = 4) {
    check(vloam_vo_process_cloud(h_, points, &n, stride_floats, (size_t)n));
  }
  // visual_odometry.cpp:254-450.  prev_uv / curr_uv: m matched keypoint pixels (cv::KeyPoint::pt of the previous / current
  // image), init = (angle-axis, t) of cam0_curr_LOT_cam0_prev or nullptr (reset_VO_to_identity).  Fills angles_0to1, t_0to1.
  void solveNlsAll(const float* prev_uv, const float* curr_uv, int m, const double* init = nullptr, int remove_VO_outlier = 100,
                   int max_iterations = 100) {
    if (m > max_matches_) m = max_matches_;
    double out[8];
    check(vloam_vo_solve(h_, prev_uv, curr_uv, &m, init, remove_VO_outlier, max_iterations, out));
    for (int i = 0; i < 3; ++i) { angles_0to1[i] = out[i]; t_0to1[i] = out[3 + i]; }
    counter32 = (int)out[6]; counter22 = (int)out[7];
  }
  // ImageUtil::detKeypoints, DetectorType::ShiTomasi (image_util.cpp:11-37): 8-bit grey image, rows packed -> (x, y) pairs in
  // cv::goodFeaturesToTrack's order (the adapter wraps them into cv::KeyPoint with size = 5 like :29-35)
  std::vector<float> detKeypoints(const uint8_t* image, int height, int width, int max_corners = 1024, double quality_level = 0.03,
                                  double min_distance = 7.5) {
    std::vector<float> xy((size_t)max_corners * 2);
    int n = 0;
    check(vloam_vo_detect_corners(h_, image, height, width, max_corners, quality_level, min_distance, xy.data(), &n));
    xy.resize((size_t)n * 2);
    return xy;
  }
  // ImageUtil::matchDescriptors (image_util.cpp:214-296, BF + NORM_HAMMING + kNN + 0.8 ratio test): rows of the ORB descriptor
  // matrices (32 bytes each) -> (queryIdx, trainIdx, distance) triples in query order
  std::vector<int> matchDescriptors(const uint8_t* desc_query, int n_query, const uint8_t* desc_train, int n_train, double ratio = 0.8) {
    std::vector<uint8_t> q((size_t)max_matches_ * 32, 0), t((size_t)max_matches_ * 32, 0);
    n_query = n_query < max_matches_ ? n_query : max_matches_; n_train = n_train < max_matches_ ? n_train : max_matches_;
    std::copy(desc_query, desc_query + (size_t)n_query * 32, q.begin());
    std::copy(desc_train, desc_train + (size_t)n_train * 32, t.begin());
    std::vector<int> m((size_t)max_matches_ * 3);
    int nm = 0;
    check(vloam_vo_match_descriptors(h_, q.data(), &n_query, t.data(), &n_train, nullptr, nullptr, ratio, m.data(), &nm));
    m.resize((size_t)nm * 3);
    return m;
  }
  // ImageUtil::descKeypoints, DescriptorType::ORB (image_util.cpp:162-212): n key points (x, y) on an 8-bit grey image -> the key
  // points cv::ORB keeps (the reference's vector after the call) and their 32-byte descriptor rows
  struct Features {
    std::vector<float> keypoints_xy;      // 2 floats per key point
    std::vector<uint8_t> descriptors;     // 32 bytes per key point (the rows of the cv::Mat)
    int size() const { return (int)(keypoints_xy.size() / 2); }
  };
  Features descKeypoints(const uint8_t* image, int height, int width, const float* keypoints_xy, int n) {
    n = n < max_matches_ ? n : max_matches_;
    std::vector<float> in((size_t)max_matches_ * 2, 0.f);
    std::copy(keypoints_xy, keypoints_xy + (size_t)n * 2, in.begin());
    Features f;
    f.keypoints_xy.resize((size_t)max_matches_ * 2); f.descriptors.resize((size_t)max_matches_ * 32);
    int kept = 0;
    check(vloam_vo_describe_orb(h_, image, height, width, in.data(), &n, f.keypoints_xy.data(), nullptr, f.descriptors.data(), &kept));
    f.keypoints_xy.resize((size_t)kept * 2); f.descriptors.resize((size_t)kept * 32);
    return f;
  }
  // VisualOdometry::processImage (visual_odometry.cpp:92-130) with ShiTomasi + ORB + BF / kNN: one image upload, detection,
  // description and (after the first frame) matching on the device.  Call reset() first.  Returns keypoints[i], descriptors[i]
  // and `matches` = (queryIdx into the previous frame's key points, trainIdx into this frame's, distance) triples.
  struct Frame {
    Features features;
    std::vector<int> matches;
  };
  Frame processImage(const uint8_t* image, int height, int width) {
    Frame fr;
    int nk = 0, nm = 0;
    check(vloam_vo_process_image(h_, image, height, width, &nk, &nm));
    fr.features.keypoints_xy.resize((size_t)max_matches_ * 2); fr.features.descriptors.resize((size_t)max_matches_ * 32);
    check(vloam_vo_get_frame_features(h_, 0, fr.features.keypoints_xy.data(), fr.features.descriptors.data(), &nk));
    fr.features.keypoints_xy.resize((size_t)nk * 2); fr.features.descriptors.resize((size_t)nk * 32);
    fr.matches.resize((size_t)max_matches_ * 3);
    check(vloam_vo_get_matches(h_, fr.matches.data(), &nm));
    fr.matches.resize((size_t)nm * 3);
    return fr;
  }
  // PointCloudUtil::queryDepth (point_cloud_util.cpp:302-407); slot 0 = current frame, 1 = previous
  float queryDepth(float x, float y, int slot = 0) {
    const float xy[2] = {x, y};
    float z = -1.f;
    check(vloam_vo_query_depth(h_, 0, slot, xy, 1, &z));
    return z;
  }

  int count = -1;
  double angles_0to1[3] = {0, 0, 0}, t_0to1[3] = {0, 0, 0};   // cam0_curr_T_cam0_last as angle-axis + translation
  int counter32 = 0, counter22 = 0;

 private:
  void check(int rc) {
    if (rc != VLOAM_OK) throw std::runtime_error(std::string("vloam_b200: ") + vloam_last_error(ctx_));
  }
  vloam_ctx* ctx_ = nullptr;
  vloam_vo* h_ = nullptr;
  int max_matches_;
};

}  // namespace vloam_b200
