#pragma once
#include <ros/ros.h>
namespace std_msgs { struct Header { std::string frame_id; ros::Time stamp; }; }
