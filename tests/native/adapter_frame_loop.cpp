// CI build of the ROS-facing adapters against tests/stubs (no ROS / PCL / Eigen in this image): both the façade
// (adapter/lidar_odometry_mapping_b200.h) and the three stage classes (adapter/lidar_stages_b200.h) are driven through the
// call sequence of the reference's caller — vloam_main_node.cpp:134,148,165-167 for the façade and
// lidar_odometry_mapping.cpp:65-154 for the stage classes — on scans read from stdin-named files.
//   adapter_frame_loop compile-only            -> prints the topics the classes advertise (no GPU needed)
//   adapter_frame_loop run scan0.bin scan1.bin -> runs the frames on the GPU, prints the published poses as text
#include <cstdio>
#include <cstring>
#include <fstream>
#include <vector>

#include "../../vloam-cmu-16833_b200/adapter/lidar_odometry_mapping_b200.h"
#include "../../vloam-cmu-16833_b200/adapter/lidar_stages_b200.h"

static pcl::PointCloud<pcl::PointXYZ> load(const char* path) {   // packed float32 x y z records
  std::ifstream f(path, std::ios::binary);
  std::vector<char> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  pcl::PointCloud<pcl::PointXYZ> c;
  const size_t n = raw.size() / 12;
  c.points.resize(n);
  for (size_t i = 0; i < n; ++i) { float v[3]; std::memcpy(v, raw.data() + 12 * i, 12); c.points[i].x = v[0]; c.points[i].y = v[1]; c.points[i].z = v[2]; }
  return c;
}

static void set_params() {
  auto& p = ros::stub::params();
  p["loam_verbose_level"] = 0; p["scan_line"] = 64; p["minimum_range"] = 5.0; p["mapping_skip_frame"] = 1; p["detach_VO_LO"] = 1;
  p["mapping_line_resolution"] = 0.4; p["mapping_plane_resolution"] = 0.8; p["map_pub_number"] = 2;
}

int main(int argc, char** argv) {
  set_params();
  if (argc < 2 || std::strcmp(argv[1], "compile-only") == 0) {
    for (const char* t : {"/velodyne_cloud_2", "/laser_cloud_sharp", "/laser_cloud_less_sharp", "/laser_cloud_flat", "/laser_cloud_less_flat",
                          "/laser_cloud_corner_last", "/laser_cloud_surf_last", "/velodyne_cloud_3", "/laser_odom_to_init", "/laser_odom_path",
                          "/laser_cloud_surround", "/laser_cloud_map", "/velodyne_cloud_registered", "/aft_mapped_to_init", "/aft_mapped_path"})
      std::printf("topic %s\n", t);
    return 0;
  }
  auto tf = std::make_shared<vloam::VloamTF>();
  // ---- the façade, as vloam_main_node.cpp drives it
  {
    vloam::LidarOdometryMapping lom;
    lom.init(tf);
    for (int k = 2; k < argc; ++k) {
      const pcl::PointCloud<pcl::PointXYZ> cloud = load(argv[k]);
      lom.reset();
      lom.scanRegistrationIO(cloud);
      lom.laserOdometryIO();
      lom.laserMappingIO();
      const tf2::Vector3 t = tf->world_MOT_base_last.getOrigin();
      const tf2::Quaternion q = tf->world_MOT_base_last.getRotation();
      std::printf("facade %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", k - 2, q.x(), q.y(), q.z(), q.w(), t.x(), t.y(), t.z());
    }
  }
  // ---- the stage classes, as lidar_odometry_mapping.cpp drives them
  {
    vloam::ScanRegistration scan_registration;
    vloam::LaserOdometry laser_odometry;
    vloam::LaserMapping laser_mapping;
    scan_registration.init();
    laser_odometry.init(tf);
    laser_mapping.init(tf);
    pcl::PointCloud<vloam::PointType>::Ptr laserCloud, sharp, lessSharp, flat, lessFlat, cornerLast, surfLast, fullRes;
    Eigen::Quaterniond q_wodom_curr;
    Eigen::Vector3d t_wodom_curr;
    bool skip_frame = false;
    for (int k = 2; k < argc; ++k) {
      const pcl::PointCloud<pcl::PointXYZ> cloud = load(argv[k]);
      scan_registration.reset(); laser_mapping.reset();                                   // :65-71
      scan_registration.input(cloud); scan_registration.publish();                        // :73-94
      scan_registration.output(laserCloud, sharp, lessSharp, flat, lessFlat);
      laser_odometry.input(laserCloud, sharp, lessSharp, flat, lessFlat);                 // :96-123
      laser_odometry.solveLO(); laser_odometry.publish();
      laser_odometry.output(q_wodom_curr, t_wodom_curr, cornerLast, surfLast, fullRes, skip_frame);
      laser_mapping.input(cornerLast, surfLast, fullRes, q_wodom_curr, t_wodom_curr, skip_frame);   // :125-154
      if (!skip_frame) laser_mapping.solveMapping();
      laser_mapping.publish();
      const tf2::Vector3 t = tf->world_MOT_base_last.getOrigin();
      const tf2::Quaternion q = tf->world_MOT_base_last.getRotation();
      std::printf("stages %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g  sharp %zu lessFlat %zu\n", k - 2, q.x(), q.y(), q.z(), q.w(), t.x(), t.y(), t.z(),
                  sharp->size(), lessFlat->size());
    }
    for (const auto& kv : ros::stub::topics()) std::printf("published %s %zu\n", kv.first.c_str(), kv.second.count);
  }
  return 0;
}
