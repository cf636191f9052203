#pragma once
#include <cmath>
#include <vector>
#include <pcl/point_cloud.h>
namespace pcl {
// filter.hpp removeNaNFromPointCloud: order-preserving compaction of the points whose x, y, z are all finite
template <class P> void removeNaNFromPointCloud(const PointCloud<P>& in, PointCloud<P>& out, std::vector<int>& index) {
  if (&in != &out) { out.header = in.header; out.points.resize(in.points.size()); }
  index.resize(in.points.size());
  std::size_t j = 0;
  for (std::size_t i = 0; i < in.points.size(); ++i) {
    if (!std::isfinite(in.points[i].x) || !std::isfinite(in.points[i].y) || !std::isfinite(in.points[i].z)) continue;
    out.points[j] = in.points[i]; index[j] = (int)i; ++j; }
  if (j != in.points.size()) { out.points.resize(j); index.resize(j); }
  out.height = 1; out.width = (std::uint32_t)j; out.is_dense = true;
}
}
