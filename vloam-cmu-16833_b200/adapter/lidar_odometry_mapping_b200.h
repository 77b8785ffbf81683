// Drop-in replacement for include/lidar_odometry_mapping/lidar_odometry_mapping.h of YukunXia/VLOAM-CMU-16833:
// the same class name, namespace and public methods, so src/vloam_main/src/vloam_main_node.cpp compiles unchanged.
// Only built where ROS + PCL exist (they do not in the development image).
#pragma once
#if __has_include(<ros/ros.h>) && __has_include(<pcl/point_cloud.h>)
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <ros/ros.h>
#include <vloam_tf/vloam_tf.h>

#include <memory>

#include "host_api.hpp"

namespace vloam {

class LidarOdometryMapping {
 public:
  LidarOdometryMapping() : nh("lidar_odometry_mapping_node") {}

  void init(std::shared_ptr<VloamTF>& vloam_tf_) {
    vloam_tf = vloam_tf_;
    vloam_lidar_params& p = impl.params();
    // the parameters the three reference classes read, with the same fatal-on-missing behaviour
    if (!ros::param::get("loam_verbose_level", verbose_level)) ROS_BREAK();
    if (!ros::param::get("scan_line", p.scan_line)) ROS_BREAK();                               // scan_registration.cpp:48
    if (!ros::param::get("minimum_range", p.minimum_range)) ROS_BREAK();                       // :51
    if (!ros::param::get("mapping_skip_frame", p.mapping_skip_frame)) ROS_BREAK();             // laser_odometry.cpp:53
    bool detach = true;
    if (!ros::param::get("detach_VO_LO", detach)) ROS_BREAK();                                 // laser_odometry.cpp:47
    p.detach_VO_LO = detach ? 1 : 0;
    if (!ros::param::get("mapping_line_resolution", p.mapping_line_resolution)) ROS_BREAK();   // laser_mapping.cpp:95
    if (!ros::param::get("mapping_plane_resolution", p.mapping_plane_resolution)) ROS_BREAK(); // :97
    p.batch = 1;
    p.max_points = 1 << 18;
    try { impl.init(); } catch (const std::exception& e) { ROS_ERROR("%s", e.what()); ROS_BREAK(); }
  }

  void reset() { impl.reset(); }

  void scanRegistrationIO(const pcl::PointCloud<pcl::PointXYZ>& laserCloudIn) {
    // pcl::PointXYZ is 4 floats (x, y, z, pad): handed over without a copy
    impl.scanRegistrationIO(reinterpret_cast<const float*>(laserCloudIn.points.data()), (int)laserCloudIn.points.size(), 4);
  }

  void laserOdometryIO() {
    double prior[7];
    const tf2::Quaternion q = vloam_tf->velo_last_VOT_velo_curr.getRotation();                 // laser_odometry.cpp:225-232
    const tf2::Vector3 t = vloam_tf->velo_last_VOT_velo_curr.getOrigin();
    prior[0] = q.x(); prior[1] = q.y(); prior[2] = q.z(); prior[3] = q.w();
    prior[4] = t.x(); prior[5] = t.y(); prior[6] = t.z();
    impl.laserOdometryIO(prior);
    // laser_odometry.cpp:563-571
    const auto& f = impl.last_curr; const auto& w = impl.odom;
    vloam_tf->base_prev_LOT_base_curr.setOrigin(tf2::Vector3(f.t[0], f.t[1], f.t[2]));
    vloam_tf->base_prev_LOT_base_curr.setRotation(tf2::Quaternion(f.q[0], f.q[1], f.q[2], f.q[3]));
    vloam_tf->cam0_curr_LOT_cam0_prev = vloam_tf->base_T_cam0.inverse() * vloam_tf->base_prev_LOT_base_curr.inverse() * vloam_tf->base_T_cam0;
    vloam_tf->world_LOT_base_last.setOrigin(tf2::Vector3(w.t[0], w.t[1], w.t[2]));
    vloam_tf->world_LOT_base_last.setRotation(tf2::Quaternion(w.q[0], w.q[1], w.q[2], w.q[3]));
  }

  void laserMappingIO() {
    impl.laserMappingIO();
    const auto& m = impl.mapped;                                                                // laser_mapping.cpp:728-729
    vloam_tf->world_MOT_base_last.setOrigin(tf2::Vector3(m.t[0], m.t[1], m.t[2]));
    vloam_tf->world_MOT_base_last.setRotation(tf2::Quaternion(m.q[0], m.q[1], m.q[2], m.q[3]));
  }

 private:
  std::shared_ptr<VloamTF> vloam_tf;
  ros::NodeHandle nh;
  int verbose_level = 0;
  vloam_b200::LidarOdometryMapping impl;
};

}  // namespace vloam
#endif
