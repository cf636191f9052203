#pragma once
#include <string>
#include <ros/time.h>
namespace std_msgs { struct Header { unsigned seq = 0; ros::Time stamp; std::string frame_id; }; }
