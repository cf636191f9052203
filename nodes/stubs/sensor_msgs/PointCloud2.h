#pragma once
#include <std_msgs/Header.h>
namespace sensor_msgs { struct PointCloud2 { std_msgs::Header header; unsigned height = 0, width = 0, point_step = 0, row_step = 0; bool is_bigendian = false, is_dense = true; std::vector<unsigned char> data; };
typedef boost::shared_ptr<PointCloud2 const> PointCloud2ConstPtr; }
