// Drop-in replacement for include/lidar_odometry_mapping/lidar_odometry_mapping.h of YukunXia/VLOAM-CMU-16833:
// the same class name, namespace and public methods, so src/vloam_main/src/vloam_main_node.cpp compiles unchanged.
// Built exactly like the reference's façade (lidar_odometry_mapping.cpp:40-154) from the three stage classes — here the
// B200-backed ones of lidar_stages_b200.h, which share one device-resident pipeline over the C ABI.
// Needs ROS + PCL headers; the development image has neither, so CI builds it against tests/stubs.
#pragma once
#if __has_include(<ros/ros.h>) && __has_include(<pcl/point_cloud.h>)
#include "lidar_stages_b200.h"

namespace vloam {

class LidarOdometryMapping {
 public:
  LidarOdometryMapping() : nh("lidar_odometry_mapping_node") {}

  void init(std::shared_ptr<VloamTF>& vloam_tf_) {   // lidar_odometry_mapping.cpp:40-63
    vloam_tf = vloam_tf_;
    if (!ros::param::get("loam_verbose_level", verbose_level)) ROS_BREAK();
    scan_registration.init();
    laser_odometry.init(vloam_tf);
    laser_mapping.init(vloam_tf);
    skip_frame = false;
    // The reference publishes every feature cloud of every frame (scan_registration.cpp:451-499).  That costs a device ->
    // host copy per cloud here; set the (new, optional) parameter publish_feature_clouds to false to keep them on the GPU.
    ros::param::get("publish_feature_clouds", publish_feature_clouds);
  }

  void reset() {   // :65-71
    scan_registration.reset();
    laser_mapping.reset();
  }

  void scanRegistrationIO(const pcl::PointCloud<pcl::PointXYZ>& laserCloudIn) {   // :73-94
    scan_registration.input(laserCloudIn);
    if (publish_feature_clouds) {
      scan_registration.publish();
      scan_registration.output(laserCloud, cornerPointsSharp, cornerPointsLessSharp, surfPointsFlat, surfPointsLessFlat);
    }
  }

  void laserOdometryIO() {   // :96-123
    laser_odometry.input(laserCloud, cornerPointsSharp, cornerPointsLessSharp, surfPointsFlat, surfPointsLessFlat);
    laser_odometry.solveLO();
    laser_odometry.publish();
    if (publish_feature_clouds) {
      laser_odometry.output(q_wodom_curr, t_wodom_curr, laserCloudCornerLast, laserCloudSurfLast, laserCloudFullRes, skip_frame);
    } else {
      pcl::PointCloud<PointType>::Ptr none;                      // output() fills clouds only for frames that are mapped
      bool skip = true;
      laser_odometry.output(q_wodom_curr, t_wodom_curr, none, none, none, skip);
      skip_frame = skip;
    }
  }

  void laserMappingIO() {   // :125-154
    laser_mapping.input(laserCloudCornerLast, laserCloudSurfLast, laserCloudFullRes, q_wodom_curr, t_wodom_curr, skip_frame);
    if (!skip_frame) laser_mapping.solveMapping();
    laser_mapping.publish();
  }

 private:
  std::shared_ptr<VloamTF> vloam_tf;
  ros::NodeHandle nh;
  int verbose_level = 0;
  bool publish_feature_clouds = true;

  ScanRegistration scan_registration;
  pcl::PointCloud<PointType>::Ptr laserCloud, cornerPointsSharp, cornerPointsLessSharp, surfPointsFlat, surfPointsLessFlat;
  LaserOdometry laser_odometry;
  Eigen::Quaterniond q_wodom_curr;
  Eigen::Vector3d t_wodom_curr;
  pcl::PointCloud<PointType>::Ptr laserCloudCornerLast, laserCloudSurfLast, laserCloudFullRes;
  bool skip_frame = false;
  LaserMapping laser_mapping;
};

}  // namespace vloam
#endif
