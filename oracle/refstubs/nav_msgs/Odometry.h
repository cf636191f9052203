#pragma once
#include <memory>
#include <std_msgs/Header.h>
#include <geometry_msgs/PoseStamped.h>
namespace nav_msgs {
struct Odometry { typedef std::shared_ptr<Odometry const> ConstPtr; typedef std::shared_ptr<Odometry> Ptr;
  std_msgs::Header header; std::string child_frame_id; struct { geometry_msgs::Pose pose; } pose; struct { struct { geometry_msgs::Point linear, angular; } twist; } twist; };
typedef std::shared_ptr<Odometry const> OdometryConstPtr;
}
