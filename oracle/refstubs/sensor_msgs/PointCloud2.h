#pragma once
// stand-in: a point record is point_step bytes, x y z at 0 4 8, intensity (when the step allows) at 16 -- PCL's layouts
#include <memory>
#include <vector>
#include <std_msgs/Header.h>
namespace boost { template <class T> using shared_ptr = std::shared_ptr<T>; }
namespace sensor_msgs {
struct PointCloud2 { std_msgs::Header header; unsigned height = 1, width = 0, point_step = 0, row_step = 0; bool is_bigendian = false, is_dense = true;
  std::vector<unsigned char> data; };
typedef boost::shared_ptr<PointCloud2 const> PointCloud2ConstPtr;
typedef boost::shared_ptr<PointCloud2> PointCloud2Ptr;
}
