// stub of the reference's <vloam_tf/vloam_tf.h>: the transform members LidarOdometryMapping reads and writes
// (laser_odometry.cpp:225-232, 563-571; laser_mapping.cpp:728-729).  On a ROS machine the reference's own header is used.
#pragma once
#include <Eigen/Dense>
#include <tf2/LinearMath/Transform.h>
namespace vloam {
class VloamTF {
 public:
  Eigen::Isometry3f imu_eigen_T_velo, imu_eigen_T_cam0;       // vloam_tf.h:35, read by VisualOdometry::setUpPointCloud
  tf2::Transform world_VOT_base_last;                          // vloam_tf.h:39, read by VisualOdometry::publish
  tf2::Transform velo_last_VOT_velo_curr, base_prev_LOT_base_curr, cam0_curr_LOT_cam0_prev, base_T_cam0, world_LOT_base_last, world_MOT_base_last;
  VloamTF() { velo_last_VOT_velo_curr.setIdentity(); base_prev_LOT_base_curr.setIdentity(); cam0_curr_LOT_cam0_prev.setIdentity(); base_T_cam0.setIdentity();
              world_LOT_base_last.setIdentity(); world_MOT_base_last.setIdentity(); world_VOT_base_last.setIdentity(); }
};
}  // namespace vloam
