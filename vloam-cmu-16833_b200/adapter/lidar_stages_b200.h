// Drop-in replacements for the three stage classes behind vloam::LidarOdometryMapping —
//   include/lidar_odometry_mapping/scan_registration.h:64-121   vloam::ScanRegistration
//   include/lidar_odometry_mapping/laser_odometry.h:68-148      vloam::LaserOdometry
//   include/lidar_odometry_mapping/laser_mapping.h:66-190       vloam::LaserMapping
// — with the reference's public methods (init / reset / input / solve* / publish / output) and the topics they advertise
// (scan_registration.cpp:64-68, laser_odometry.cpp:108-112, laser_mapping.cpp:103-108), so that code written against the
// stage classes (the reference's own lidar_odometry_mapping.cpp:40-154) keeps compiling.  The three objects of one
// sensor share one device-resident pipeline (vloam_b200::LidarOdometryMapping over the C ABI): the clouds the reference
// copies from stage to stage (pcl::PointCloud::Ptr arguments of input()) stay on the GPU and are only materialised by
// output() / publish().  Compiled in CI against tests/stubs (no ROS in the development image), see INTEGRATION.md.
#pragma once
#if __has_include(<ros/ros.h>) && __has_include(<pcl/point_cloud.h>)
#include <nav_msgs/Odometry.h>
#include <nav_msgs/Path.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl_conversions/pcl_conversions.h>
#include <ros/ros.h>
#include <sensor_msgs/PointCloud2.h>
#if __has_include(<tf/transform_broadcaster.h>)
#include <tf/transform_broadcaster.h>
#endif
#include <vloam_tf/vloam_tf.h>

#if __has_include(<eigen3/Eigen/Dense>)
#include <eigen3/Eigen/Dense>
#else
#include <Eigen/Dense>
#endif

#include <memory>

#include "host_api.hpp"

namespace vloam {

using PointType = pcl::PointXYZI;

namespace b200_detail {
// The pipeline the three stage objects of the calling thread share (the reference keeps one ScanRegistration,
// LaserOdometry and LaserMapping per LidarOdometryMapping, lidar_odometry_mapping.h:66-85).
inline std::shared_ptr<vloam_b200::LidarOdometryMapping>& shared_core() {
  static thread_local std::shared_ptr<vloam_b200::LidarOdometryMapping> core;
  if (!core) core = std::make_shared<vloam_b200::LidarOdometryMapping>();
  return core;
}
inline void fill_cloud(vloam_b200::LidarOdometryMapping& core, int which, pcl::PointCloud<PointType>::Ptr& out) {
  if (!out) out = std::make_shared<pcl::PointCloud<PointType>>();
  const std::vector<float> v = core.cloud(which);
  out->points.resize(v.size() / 4);
  for (size_t i = 0; i < out->points.size(); ++i) {
    out->points[i].x = v[4 * i]; out->points[i].y = v[4 * i + 1]; out->points[i].z = v[4 * i + 2]; out->points[i].intensity = v[4 * i + 3];
  }
}
template <typename Pub>
inline void publish_cloud(const Pub& pub, vloam_b200::LidarOdometryMapping& core, int which, const char* frame) {
  pcl::PointCloud<PointType>::Ptr c;
  fill_cloud(core, which, c);
  sensor_msgs::PointCloud2 msg;
  pcl::toROSMsg(*c, msg);
  msg.header.stamp = ros::Time::now();
  msg.header.frame_id = frame;
  pub.publish(msg);
}
inline void fill_odometry(nav_msgs::Odometry& o, const vloam_b200::Pose& p, const char* child) {
  o.header.frame_id = "map";
  o.child_frame_id = child;
  o.header.stamp = ros::Time::now();
  o.pose.pose.orientation.x = p.q[0]; o.pose.pose.orientation.y = p.q[1]; o.pose.pose.orientation.z = p.q[2]; o.pose.pose.orientation.w = p.q[3];
  o.pose.pose.position.x = p.t[0]; o.pose.pose.position.y = p.t[1]; o.pose.pose.position.z = p.t[2];
}
}  // namespace b200_detail

class ScanRegistration {
 public:
  ScanRegistration() : nh("scan_registration_node") {}

  void init() {   // scan_registration.cpp:40-87
    core = b200_detail::shared_core();
    vloam_lidar_params& p = core->params();
    if (!ros::param::get("loam_verbose_level", verbose_level)) ROS_BREAK();
    if (!ros::param::get("scan_line", p.scan_line)) ROS_BREAK();
    if (!ros::param::get("minimum_range", p.minimum_range)) ROS_BREAK();
    if (p.scan_line != 16 && p.scan_line != 32 && p.scan_line != 64) { ROS_ERROR("only support velodyne with 16, 32 or 64 scan line!"); ROS_BREAK(); }
    pubLaserCloud = nh.advertise<sensor_msgs::PointCloud2>("/velodyne_cloud_2", 100);
    pubCornerPointsSharp = nh.advertise<sensor_msgs::PointCloud2>("/laser_cloud_sharp", 100);
    pubCornerPointsLessSharp = nh.advertise<sensor_msgs::PointCloud2>("/laser_cloud_less_sharp", 100);
    pubSurfPointsFlat = nh.advertise<sensor_msgs::PointCloud2>("/laser_cloud_flat", 100);
    pubSurfPointsLessFlat = nh.advertise<sensor_msgs::PointCloud2>("/laser_cloud_less_flat", 100);
  }
  void reset() {}   // scan_registration.cpp:89-98: the per-scan clouds are device buffers overwritten by the next input()
  void input(const pcl::PointCloud<pcl::PointXYZ>& laserCloudIn_) {   // :131-449
    core->scanRegistrationIO(reinterpret_cast<const float*>(laserCloudIn_.points.data()), (int)laserCloudIn_.points.size(),
                             (int)(sizeof(pcl::PointXYZ) / sizeof(float)));
  }
  void publish() {   // :451-499
    b200_detail::publish_cloud(pubLaserCloud, *core, VLOAM_CLOUD_FULL, "velo");
    b200_detail::publish_cloud(pubCornerPointsSharp, *core, VLOAM_CLOUD_SHARP, "velo");
    b200_detail::publish_cloud(pubCornerPointsLessSharp, *core, VLOAM_CLOUD_LESS_SHARP, "velo");
    b200_detail::publish_cloud(pubSurfPointsFlat, *core, VLOAM_CLOUD_FLAT, "velo");
    b200_detail::publish_cloud(pubSurfPointsLessFlat, *core, VLOAM_CLOUD_LESS_FLAT, "velo");
  }
  void output(pcl::PointCloud<PointType>::Ptr& laserCloud_, pcl::PointCloud<PointType>::Ptr& cornerPointsSharp_,
              pcl::PointCloud<PointType>::Ptr& cornerPointsLessSharp_, pcl::PointCloud<PointType>::Ptr& surfPointsFlat_,
              pcl::PointCloud<PointType>::Ptr& surfPointsLessFlat_) {   // :501-512
    b200_detail::fill_cloud(*core, VLOAM_CLOUD_FULL, laserCloud_);
    b200_detail::fill_cloud(*core, VLOAM_CLOUD_SHARP, cornerPointsSharp_);
    b200_detail::fill_cloud(*core, VLOAM_CLOUD_LESS_SHARP, cornerPointsLessSharp_);
    b200_detail::fill_cloud(*core, VLOAM_CLOUD_FLAT, surfPointsFlat_);
    b200_detail::fill_cloud(*core, VLOAM_CLOUD_LESS_FLAT, surfPointsLessFlat_);
  }

 private:
  std::shared_ptr<vloam_b200::LidarOdometryMapping> core;
  ros::NodeHandle nh;
  int verbose_level = 0;
  ros::Publisher pubLaserCloud, pubCornerPointsSharp, pubCornerPointsLessSharp, pubSurfPointsFlat, pubSurfPointsLessFlat;
};

class LaserOdometry {
 public:
  LaserOdometry() : nh("laser_odometry_node") {}

  void init(std::shared_ptr<VloamTF>& vloam_tf_) {   // laser_odometry.cpp:41-117
    vloam_tf = vloam_tf_;
    core = b200_detail::shared_core();
    vloam_lidar_params& p = core->params();
    if (!ros::param::get("loam_verbose_level", verbose_level)) ROS_BREAK();
    bool detach = true;
    if (!ros::param::get("detach_VO_LO", detach)) ROS_BREAK();
    p.detach_VO_LO = detach ? 1 : 0;
    if (!ros::param::get("mapping_skip_frame", p.mapping_skip_frame)) ROS_BREAK();
    pubLaserCloudCornerLast = nh.advertise<sensor_msgs::PointCloud2>("/laser_cloud_corner_last", 100);
    pubLaserCloudSurfLast = nh.advertise<sensor_msgs::PointCloud2>("/laser_cloud_surf_last", 100);
    pubLaserCloudFullRes = nh.advertise<sensor_msgs::PointCloud2>("/velodyne_cloud_3", 100);
    pubLaserOdometry = nh.advertise<nav_msgs::Odometry>("/laser_odom_to_init", 100);
    pubLaserPath = nh.advertise<nav_msgs::Path>("/laser_odom_path", 100);
  }
  // :135-146.  The five clouds are the ones ScanRegistration::output handed out; their device-resident originals are used.
  void input(const pcl::PointCloud<PointType>::Ptr&, const pcl::PointCloud<PointType>::Ptr&, const pcl::PointCloud<PointType>::Ptr&,
             const pcl::PointCloud<PointType>::Ptr&, const pcl::PointCloud<PointType>::Ptr&) {}
  void solveLO() {   // :187-536
    double prior[7];
    const tf2::Quaternion q = vloam_tf->velo_last_VOT_velo_curr.getRotation();   // :225-232
    const tf2::Vector3 t = vloam_tf->velo_last_VOT_velo_curr.getOrigin();
    prior[0] = q.x(); prior[1] = q.y(); prior[2] = q.z(); prior[3] = q.w();
    prior[4] = t.x(); prior[5] = t.y(); prior[6] = t.z();
    core->laserOdometryIO(prior);
    ++frameCount;
  }
  void publish() {   // :538-608
    nav_msgs::Odometry laserOdometry;
    b200_detail::fill_odometry(laserOdometry, core->odom, "laser_odom");
    pubLaserOdometry.publish(laserOdometry);
    const auto& f = core->last_curr; const auto& w = core->odom;
    vloam_tf->base_prev_LOT_base_curr.setOrigin(tf2::Vector3(f.t[0], f.t[1], f.t[2]));
    vloam_tf->base_prev_LOT_base_curr.setRotation(tf2::Quaternion(f.q[0], f.q[1], f.q[2], f.q[3]));
    vloam_tf->cam0_curr_LOT_cam0_prev = vloam_tf->base_T_cam0.inverse() * vloam_tf->base_prev_LOT_base_curr.inverse() * vloam_tf->base_T_cam0;
    vloam_tf->world_LOT_base_last.setOrigin(tf2::Vector3(w.t[0], w.t[1], w.t[2]));
    vloam_tf->world_LOT_base_last.setRotation(tf2::Quaternion(w.q[0], w.q[1], w.q[2], w.q[3]));
    geometry_msgs::PoseStamped laserPose;
    laserPose.header = laserOdometry.header;
    laserPose.pose = laserOdometry.pose.pose;
    laserPath.header.stamp = laserOdometry.header.stamp;
    laserPath.poses.push_back(laserPose);
    laserPath.header.frame_id = "map";
    pubLaserPath.publish(laserPath);
  }
  void output(Eigen::Quaterniond& q_w_curr_, Eigen::Vector3d& t_w_curr_, pcl::PointCloud<PointType>::Ptr& laserCloudCornerLast_,
              pcl::PointCloud<PointType>::Ptr& laserCloudSurfLast_, pcl::PointCloud<PointType>::Ptr& laserCloudFullRes_, bool& skip_frame) {   // :610-629
    const auto& w = core->odom;
    q_w_curr_ = Eigen::Quaterniond(w.q[3], w.q[0], w.q[1], w.q[2]);
    t_w_curr_ = Eigen::Vector3d(w.t[0], w.t[1], w.t[2]);
    const bool want_clouds = !skip_frame;      // (in: a caller that keeps the clouds on the GPU passes true; out: the skip flag)
    skip_frame = frameCount % core->params().mapping_skip_frame != 0;
    if (!skip_frame && want_clouds) {
      b200_detail::fill_cloud(*core, VLOAM_CLOUD_CORNER_LAST, laserCloudCornerLast_);
      b200_detail::fill_cloud(*core, VLOAM_CLOUD_SURF_LAST, laserCloudSurfLast_);
      b200_detail::fill_cloud(*core, VLOAM_CLOUD_FULL, laserCloudFullRes_);
    }
  }

 private:
  std::shared_ptr<VloamTF> vloam_tf;
  std::shared_ptr<vloam_b200::LidarOdometryMapping> core;
  ros::NodeHandle nh;
  int verbose_level = 0, frameCount = 0;
  nav_msgs::Path laserPath;
  ros::Publisher pubLaserCloudCornerLast, pubLaserCloudSurfLast, pubLaserCloudFullRes, pubLaserOdometry, pubLaserPath;
};

class LaserMapping {
 public:
  LaserMapping() : nh("laser_mapping_node") {}

  void init(std::shared_ptr<VloamTF>& vloam_tf_) {   // laser_mapping.cpp:40-125
    vloam_tf = vloam_tf_;
    core = b200_detail::shared_core();
    vloam_lidar_params& p = core->params();
    if (!ros::param::get("loam_verbose_level", verbose_level)) ROS_BREAK();
    if (!ros::param::get("mapping_line_resolution", p.mapping_line_resolution)) ROS_BREAK();
    if (!ros::param::get("mapping_plane_resolution", p.mapping_plane_resolution)) ROS_BREAK();
    if (!ros::param::get("map_pub_number", map_pub_number)) ROS_BREAK();   // :123-124
    if (!ros::param::get("mapping_skip_frame", mapping_skip_frame)) mapping_skip_frame = 1;
    pubLaserCloudSurround = nh.advertise<sensor_msgs::PointCloud2>("/laser_cloud_surround", 100);
    pubLaserCloudMap = nh.advertise<sensor_msgs::PointCloud2>("/laser_cloud_map", 100);
    pubLaserCloudFullRes = nh.advertise<sensor_msgs::PointCloud2>("/velodyne_cloud_registered", 100);
    pubOdomAftMapped = nh.advertise<nav_msgs::Odometry>("/aft_mapped_to_init", 100);
    pubLaserAfterMappedPath = nh.advertise<nav_msgs::Path>("/aft_mapped_path", 100);
    // the last of the three init() calls (lidar_odometry_mapping.cpp:52-62) creates the device pipeline with the parameters read so far
    p.batch = 1;
    p.max_points = 1 << 18;
    try { core->init(); } catch (const std::exception& e) { ROS_ERROR("%s", e.what()); ROS_BREAK(); }
    b200_detail::shared_core().reset();   // the next trio of stage objects (another sensor) gets a pipeline of its own
  }
  void reset() { core->reset(); }   // :127-131
  // :167-196.  Clouds and odometry pose are taken from the device-resident laser-odometry state; skip_frame is recomputed
  // there from frameCount % mapping_skip_frame exactly like LaserOdometry::output did.
  void input(const pcl::PointCloud<PointType>::Ptr&, const pcl::PointCloud<PointType>::Ptr&, const pcl::PointCloud<PointType>::Ptr&,
             const Eigen::Quaterniond&, const Eigen::Vector3d&, const bool& skip_frame_) { skip_frame = skip_frame_; }
  void solveMapping() {   // :198-708
    core->laserMappingIO();
    solved = true;
    ++frameCount;           // :707
    // the reference's map grows without bound; this one lives in a pool of vloam_lidar_params::map_capacity_points
    const int st = core->mappingStatus();
    if (st & (VLOAM_LM_CORNER_MAP_FULL | VLOAM_LM_SURF_MAP_FULL))
      ROS_WARN("laser mapping: the map pool is full (status 0x%x) - this scan was not inserted; raise map_capacity_points", st);
  }
  void publish() {   // :710-814
    if (!solved) core->laserMappingIO();      // a skipped frame: only the high-frequency pose is refreshed (:186-190, 742-756)
    solved = false;
    nav_msgs::Odometry odomAftMapped;
    b200_detail::fill_odometry(odomAftMapped, core->mapped, "aft_mapped");
    const auto& m = core->mapped;
    vloam_tf->world_MOT_base_last.setOrigin(tf2::Vector3(m.t[0], m.t[1], m.t[2]));
    vloam_tf->world_MOT_base_last.setRotation(tf2::Quaternion(m.q[0], m.q[1], m.q[2], m.q[3]));
    pubOdomAftMapped.publish(odomAftMapped);
    geometry_msgs::PoseStamped laserAfterMappedPose;
    laserAfterMappedPose.header = odomAftMapped.header;
    laserAfterMappedPose.pose = odomAftMapped.pose.pose;
    laserAfterMappedPath.header.stamp = odomAftMapped.header.stamp;
    laserAfterMappedPath.header.frame_id = "map";
    laserAfterMappedPath.poses.push_back(laserAfterMappedPose);
    pubLaserAfterMappedPath.publish(laserAfterMappedPath);
#if __has_include(<tf/transform_broadcaster.h>)
    {   // :767-776 map -> aft_mapped
      static tf::TransformBroadcaster br;
      tf::Transform transform;
      transform.setOrigin(tf::Vector3(m.t[0], m.t[1], m.t[2]));
      transform.setRotation(tf::Quaternion(m.q[0], m.q[1], m.q[2], m.q[3]));
      br.sendTransform(tf::StampedTransform(transform, odomAftMapped.header.stamp, "map", "aft_mapped"));
    }
#endif
    if ((frameCount * mapping_skip_frame) % map_pub_number == 0)   // :778-790: the whole map, every cube's corner then surf points
      b200_detail::publish_cloud(pubLaserCloudMap, *core, VLOAM_CLOUD_MAP, "velo_origin");
    // :792-805 the full-resolution scan in the map frame.  (On a skipped frame the reference transforms its already
    // transformed copy a second time, :175-180 + :797-801; here the scan is always taken in the sensor frame.)
    b200_detail::publish_cloud(pubLaserCloudFullRes, *core, VLOAM_CLOUD_FULL_REGISTERED, "map");
  }
  void output() {}

 private:
  std::shared_ptr<VloamTF> vloam_tf;
  std::shared_ptr<vloam_b200::LidarOdometryMapping> core;
  ros::NodeHandle nh;
  int verbose_level = 0;
  bool skip_frame = false, solved = false;
  int frameCount = 0, map_pub_number = 20, mapping_skip_frame = 1;
  nav_msgs::Path laserAfterMappedPath;
  ros::Publisher pubLaserCloudSurround, pubLaserCloudMap, pubLaserCloudFullRes, pubOdomAftMapped, pubLaserAfterMappedPath;
};

}  // namespace vloam
#endif
