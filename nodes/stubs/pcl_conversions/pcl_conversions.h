#pragma once
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <sensor_msgs/PointCloud2.h>
namespace pcl { template <class P> void fromROSMsg(const sensor_msgs::PointCloud2&, pcl::PointCloud<P>&) {} template <class P> void toROSMsg(const pcl::PointCloud<P>&, sensor_msgs::PointCloud2&) {} }
