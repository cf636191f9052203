#pragma once
#include <cstring>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <sensor_msgs/PointCloud2.h>
namespace pcl {
template <class P> void toROSMsg(const PointCloud<P>& c, sensor_msgs::PointCloud2& m) {
  m.height = 1; m.width = (unsigned)c.points.size(); m.point_step = sizeof(P); m.row_step = m.width * m.point_step; m.is_dense = c.is_dense;
  m.data.resize((std::size_t)m.row_step); if (m.row_step) std::memcpy(m.data.data(), c.points.data(), m.row_step); }
template <class P> void fromROSMsg(const sensor_msgs::PointCloud2& m, PointCloud<P>& c) {     // fields matched by name in PCL: x y z (and intensity) here
  const std::size_t n = (std::size_t)m.width * m.height; c.points.assign(n, P()); c.width = m.width; c.height = m.height; c.is_dense = m.is_dense;
  const std::size_t take = sizeof(P) < m.point_step ? sizeof(P) : m.point_step;
  for (std::size_t i = 0; i < n; ++i) { P p; std::memcpy((void*)&p, m.data.data() + i * m.point_step, take >= 20 ? take : 12); c.points[i] = p; } }
}
