#pragma once
#include <memory>
#include <std_msgs/Header.h>
#include <geometry_msgs/PoseStamped.h>
namespace nav_msgs {
struct Odometry { std_msgs::Header header; std::string child_frame_id; struct { geometry_msgs::Pose pose; } pose; };
typedef std::shared_ptr<Odometry const> OdometryConstPtr;
}
