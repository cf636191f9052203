#pragma once
#include <std_msgs/Header.h>
namespace sensor_msgs { struct Imu { std_msgs::Header header; }; }
