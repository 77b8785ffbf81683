// Drop-in replacement for include/visual_odometry/visual_odometry.h of YukunXia/VLOAM-CMU-16833: the same class name,
// namespace, public methods and the public members the caller reads (cam0_curr_T_cam0_last, vloam_main_node.cpp:160), so
// src/vloam_main/src/vloam_main_node.cpp compiles unchanged.  The image front end (processImage: Shi-Tomasi detection, ORB
// description, descriptor matching), the LiDAR depth association, residual construction and solve run through libvloam_b200.so.  Only built where ROS + PCL + OpenCV exist.
#pragma once
#if __has_include(<ros/ros.h>) && __has_include(<pcl/point_cloud.h>) && __has_include(<opencv2/opencv.hpp>)
#include <geometry_msgs/PoseStamped.h>
#include <nav_msgs/Odometry.h>
#include <nav_msgs/Path.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <ros/ros.h>
#include <sensor_msgs/CameraInfo.h>
#include <sensor_msgs/PointCloud2.h>
#include <tf2/LinearMath/Transform.h>
#include <visual_odometry/image_util.h>
#include <vloam_tf/vloam_tf.h>

#include <cmath>
#include <memory>
#include <opencv2/opencv.hpp>
#include <vector>

#include "host_api.hpp"

namespace vloam {

class VisualOdometry {
 public:
  VisualOdometry() { pub_point_cloud = nh.advertise<sensor_msgs::PointCloud2>("/point_cloud_follow_VO", 5); }   // visual_odometry.cpp:5-8

  void init(std::shared_ptr<VloamTF>& vloam_tf_) {                                                             // :10-84
    vloam_tf = vloam_tf_;
    if (!ros::param::get("loam_verbose_level", verbose_level)) ROS_BREAK();
    if (!ros::param::get("reset_VO_to_identity", reset_VO_to_identity)) ROS_BREAK();
    if (!ros::param::get("remove_VO_outlier", remove_VO_outlier)) ROS_BREAK();
    if (!ros::param::get("keypoint_NMS", keypoint_NMS)) ROS_BREAK();
    if (!ros::param::get("CLAHE", CLAHE)) ROS_BREAK();
    if (!ros::param::get("visualize_optical_flow", visualize_optical_flow)) ROS_BREAK();
    if (!ros::param::get("optical_flow_match", optical_flow_match)) ROS_BREAK();
    ros::param::get("vo_frontend_on_device", frontend_on_device);
    count = -1;
    images.resize(2); keypoints.resize(2); descriptors.resize(2); keypoints_2f.resize(2);
    if (CLAHE) clahe = cv::createCLAHE(2.0, cv::Size(8, 8));
    try { impl.reset(new vloam_b200::VisualOdometry()); } catch (const std::exception& e) { ROS_ERROR("%s", e.what()); ROS_BREAK(); }
    cam0_curr_T_cam0_last.setIdentity();
    pubvisualOdometry = nh.advertise<nav_msgs::Odometry>("/visual_odom_to_init", 100);
    pubvisualPath = nh.advertise<nav_msgs::Path>("/visual_odom_path", 100);
  }

  void reset() { ++count; i = count % 2; impl->reset(); }                                                      // :86-90

  // :92-130.  With `vo_frontend_on_device` (new, optional parameter, default true), an 8-bit grey image and the descriptor
  // branch (optical_flow_match false) the whole chain — detection (image_util.cpp:11-37), ORB description (:162-212), matching
  // (:214-296) — is one device call whose results are mirrored into keypoints[i] / descriptors[i] / matches (after a frame that
  // OpenCV processed the device holds no previous descriptors: that one frame is matched from the host copies).
  void processImage(const cv::Mat& img00) {
    if (CLAHE) clahe->apply(img00, images[i]); else images[i] = img00;
    const bool on_device = frontend_on_device && images[i].type() == CV_8UC1 && images[i].isContinuous();
    if (on_device && !optical_flow_match && images[i].rows >= 3 && images[i].cols >= 3) {
      const vloam_b200::VisualOdometry::Frame fr = impl->processImage(images[i].data, images[i].rows, images[i].cols);
      const int n = fr.features.size();
      keypoints[i].clear();
      for (int k = 0; k < n; ++k) {                               // image_util.cpp:29-35, filtered by cv::ORB (:204)
        cv::KeyPoint kp;
        kp.pt = cv::Point2f(fr.features.keypoints_xy[2 * k], fr.features.keypoints_xy[2 * k + 1]);
        kp.size = 5;
        keypoints[i].push_back(kp);
      }
      descriptors[i] = n ? cv::Mat(n, 32, CV_8UC1, const_cast<uint8_t*>(fr.features.descriptors.data())).clone() : cv::Mat();
      std::vector<int> m = fr.matches;
      const cv::Mat& dq = descriptors[1 - i];
      if (count > 0 && !device_prev)
        m = (n && dq.type() == CV_8UC1 && dq.cols == 32 && dq.isContinuous()) ? impl->matchDescriptors(dq.data, dq.rows, descriptors[i].data, n)
                                                                              : std::vector<int>();
      matches.clear();
      for (size_t k = 0; k + 2 < m.size(); k += 3) matches.emplace_back(m[k], m[k + 1], (float)m[k + 2]);       // cv::DMatch(queryIdx, trainIdx, distance)
      device_prev = true;
      return;
    }
    device_prev = false;
    if (on_device) {
      const std::vector<float> xy = impl->detKeypoints(images[i].data, images[i].rows, images[i].cols);
      keypoints[i].clear();
      for (size_t k = 0; k + 1 < xy.size(); k += 2) {            // image_util.cpp:29-35
        cv::KeyPoint kp;
        kp.pt = cv::Point2f(xy[k], xy[k + 1]);
        kp.size = 5;
        keypoints[i].push_back(kp);
      }
    } else {
      keypoints[i] = image_util.detKeypoints(images[i]);
    }
    if (!optical_flow_match) descriptors[i] = image_util.descKeypoints(keypoints[i], images[i]);
    if (count > 0) {
      const cv::Mat &dq = descriptors[1 - i], &dt = descriptors[i];
      if (!optical_flow_match && frontend_on_device && dq.type() == CV_8UC1 && dt.type() == CV_8UC1 && dq.cols == 32 && dt.cols == 32 &&
          dq.isContinuous() && dt.isContinuous()) {
        const std::vector<int> m = impl->matchDescriptors(dq.data, dq.rows, dt.data, dt.rows);
        matches.clear();
        for (size_t k = 0; k + 2 < m.size(); k += 3) matches.emplace_back(m[k], m[k + 1], (float)m[k + 2]);   // cv::DMatch(queryIdx, trainIdx, distance)
      } else if (!optical_flow_match) matches = image_util.matchDescriptors(dq, dt);
      else std::tie(keypoints_2f[1 - i], keypoints_2f[i], optical_flow_status) = image_util.calculateOpticalFlow(images[1 - i], images[i], keypoints[i]);
    }
  }

  void setUpPointCloud(const sensor_msgs::CameraInfoConstPtr& camera_info_msg) {                              // :132-155
    const Eigen::Matrix4f T = (vloam_tf->imu_eigen_T_cam0.matrix().inverse() * vloam_tf->imu_eigen_T_velo.matrix()).cast<float>();
    float cam_T_velo[16], rect0_T_cam[16] = {0}, P_rect0[12];
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) cam_T_velo[4 * r + c] = T(r, c);
    for (int k = 0; k < 9; ++k) rect0_T_cam[4 * (k / 3) + k % 3] = (float)camera_info_msg->R[k];   // (3,3) stays 0 like the reference
    for (int k = 0; k < 12; ++k) P_rect0[k] = (float)camera_info_msg->P[k];
    impl->setUpPointCloud(cam_T_velo, rect0_T_cam, P_rect0);
  }

  void processPointCloud(const sensor_msgs::PointCloud2ConstPtr& point_cloud_msg, const pcl::PointCloud<pcl::PointXYZ>& point_cloud_pcl,
                         const bool& /*visualize_depth*/, const bool& publish_point_cloud) {                    // :157-186
    impl->processPointCloud(reinterpret_cast<const float*>(point_cloud_pcl.points.data()), (int)point_cloud_pcl.points.size(), 4);
    if (publish_point_cloud) {
      sensor_msgs::PointCloud2 m = *point_cloud_msg;
      m.header.frame_id = "velo"; m.header.stamp = ros::Time::now();
      pub_point_cloud.publish(m);
    }
  }

  void solveNlsAll() {                                                                                          // :254-450
    double init[6];
    const double* init_p = nullptr;
    if (!reset_VO_to_identity) {                                                                                // :269-281, init from LO
      const tf2::Transform& T = vloam_tf->cam0_curr_LOT_cam0_prev;
      const tf2::Vector3 ax = T.getRotation().getAxis();
      const double ang = T.getRotation().getAngle();
      init[0] = ax.getX() * ang; init[1] = ax.getY() * ang; init[2] = ax.getZ() * ang;
      init[3] = T.getOrigin().getX(); init[4] = T.getOrigin().getY(); init[5] = T.getOrigin().getZ();
      init_p = init;
    }
    std::vector<float> prev_uv, curr_uv;                                                                        // :286-307 (pixel truncation happens on the device)
    if (!optical_flow_match) {
      for (const cv::DMatch& m : matches) {
        const cv::Point2f a = keypoints[1 - i][m.queryIdx].pt, b = keypoints[i][m.trainIdx].pt;
        prev_uv.push_back(a.x); prev_uv.push_back(a.y); curr_uv.push_back(b.x); curr_uv.push_back(b.y);
      }
    } else {
      for (size_t k = 0; k < keypoints_2f[i].size(); ++k) {
        if (optical_flow_status[k] != 1) continue;
        prev_uv.push_back(keypoints_2f[1 - i][k].x); prev_uv.push_back(keypoints_2f[1 - i][k].y);
        curr_uv.push_back(keypoints_2f[i][k].x); curr_uv.push_back(keypoints_2f[i][k].y);
      }
    }
    impl->solveNlsAll(prev_uv.data(), curr_uv.data(), (int)(prev_uv.size() / 2), init_p, remove_VO_outlier, 100);
    for (int k = 0; k < 3; ++k) { angles_0to1[k] = impl->angles_0to1[k]; t_0to1[k] = impl->t_0to1[k]; }
    cam0_curr_T_cam0_last.setOrigin(tf2::Vector3(t_0to1[0], t_0to1[1], t_0to1[2]));                             // :426-430 (NaN when angle == 0: the
    angle = std::sqrt(angles_0to1[0] * angles_0to1[0] + angles_0to1[1] * angles_0to1[1] + angles_0to1[2] * angles_0to1[2]);   // caller guards, Q17)
    cam0_curr_q_cam0_last.setRotation(tf2::Vector3(angles_0to1[0] / angle, angles_0to1[1] / angle, angles_0to1[2] / angle), angle);
    cam0_curr_T_cam0_last.setRotation(cam0_curr_q_cam0_last);
  }

  void publish() {                                                                                              // :452-487
    visualOdometry.header.frame_id = "map"; visualOdometry.child_frame_id = "visual_odom"; visualOdometry.header.stamp = ros::Time::now();
    const tf2::Quaternion q = vloam_tf->world_VOT_base_last.getRotation();
    const tf2::Vector3 t = vloam_tf->world_VOT_base_last.getOrigin();
    visualOdometry.pose.pose.orientation.x = q.x(); visualOdometry.pose.pose.orientation.y = q.y();
    visualOdometry.pose.pose.orientation.z = q.z(); visualOdometry.pose.pose.orientation.w = q.w();
    visualOdometry.pose.pose.position.x = t.x(); visualOdometry.pose.pose.position.y = t.y(); visualOdometry.pose.pose.position.z = t.z();
    pubvisualOdometry.publish(visualOdometry);
    geometry_msgs::PoseStamped ps; ps.header = visualOdometry.header; ps.pose = visualOdometry.pose.pose;
    visualPath.header = visualOdometry.header; visualPath.poses.push_back(ps);
    pubvisualPath.publish(visualPath);
  }

  // public members of the reference class that the caller reads
  std::shared_ptr<VloamTF> vloam_tf;
  int i = 0, j = 0, count = -1;
  vloam::ImageUtil image_util;
  std::vector<cv::Mat> images, descriptors;
  std::vector<std::vector<cv::KeyPoint>> keypoints;
  std::vector<cv::DMatch> matches;
  std::vector<std::vector<cv::Point2f>> keypoints_2f;
  std::vector<uchar> optical_flow_status;
  double angles_0to1[3] = {0, 0, 0}, t_0to1[3] = {0, 0, 0};
  float angle = 0.f;
  tf2::Transform cam0_curr_T_cam0_last;
  tf2::Quaternion cam0_curr_q_cam0_last;
  ros::Publisher pub_point_cloud;

 private:
  ros::NodeHandle nh;
  int verbose_level = 0, remove_VO_outlier = 100;
  bool reset_VO_to_identity = false, keypoint_NMS = false, CLAHE = false, visualize_optical_flow = false, optical_flow_match = false;
  bool frontend_on_device = true, device_prev = false;
  cv::Ptr<cv::CLAHE> clahe;
  nav_msgs::Odometry visualOdometry;
  nav_msgs::Path visualPath;
  ros::Publisher pubvisualOdometry, pubvisualPath;
  std::unique_ptr<vloam_b200::VisualOdometry> impl;
};

}  // namespace vloam
#endif
