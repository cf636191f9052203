// Glue shared by the patched catkin nodes: views over PCL clouds for the lmono C ABI.
// pcl::PointXYZI is a 32-byte record with x@0, y@4, z@8, intensity@16
// (Aloam/include/aloam_velodyne/common.h:43), which lmono_cloud_view describes directly, so
// no repacking happens on the host.
#pragma once
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <stdexcept>
#include <string>
#include "lmono.h"

namespace lmono_glue {

template <typename PointT> struct Layout;
template <> struct Layout<pcl::PointXYZI> { static constexpr int stride = sizeof(pcl::PointXYZI), ioff = 16; };
template <> struct Layout<pcl::PointXYZ>  { static constexpr int stride = sizeof(pcl::PointXYZ),  ioff = -1; };

template <typename PointT>
inline lmono_cloud_view view(const pcl::PointCloud<PointT>& c) {
  lmono_cloud_view v;
  v.base = c.points.empty() ? nullptr : c.points.data();
  v.n = static_cast<int32_t>(c.points.size());
  v.stride_bytes = Layout<PointT>::stride;
  v.intensity_offset = Layout<PointT>::ioff;
  return v;
}

// Output cloud backed by a PCL cloud that is resized to `capacity` first and trimmed after the call.
template <typename PointT>
inline lmono_cloud_out out(pcl::PointCloud<PointT>& c, size_t capacity) {
  c.points.resize(capacity);
  lmono_cloud_out o;
  o.base = c.points.data();
  o.capacity = static_cast<int32_t>(capacity);
  o.stride_bytes = Layout<PointT>::stride;
  o.intensity_offset = Layout<PointT>::ioff;
  o.n_out = 0;
  return o;
}
template <typename PointT>
inline void trim(pcl::PointCloud<PointT>& c, const lmono_cloud_out& o) {
  c.points.resize(static_cast<size_t>(o.n_out));
  c.width = static_cast<uint32_t>(o.n_out); c.height = 1; c.is_dense = true;
}

inline void check(int rc, const char* what) {
  if (rc != LMONO_OK) throw std::runtime_error(std::string(what) + ": " + lmono_strerror(rc));
}

}  // namespace lmono_glue
