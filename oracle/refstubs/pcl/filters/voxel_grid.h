#pragma once
// stand-in for pcl::VoxelGrid<PointT>::applyFilter (PCL 1.8 voxel_grid.hpp, downsample_all_data = true, no minimum
// points per voxel): bounding box over the finite points, voxel index from floor(p * inverse_leaf) - min_b, the
// (index, point) pairs sorted with std::sort on the index alone, one centroid of ALL fields (x y z intensity) per run
// of equal indices, float sums divided by the float count.  This is a restatement of the library, not reference source.
#include <algorithm>
#include <cmath>
#include <limits>
#include <vector>
#include <pcl/filters/filter.h>
#include <pcl/point_cloud.h>
namespace pcl {
namespace refstub_detail { struct IdxPt { unsigned idx; unsigned pt; bool operator<(const IdxPt& o) const { return idx < o.idx; } };
template <class P> inline float field4(const P&) { return 0.0f; }
template <class P> inline void set_field4(P&, float) {}
}
template <class P> class VoxelGrid {
 public:
  void setInputCloud(const typename PointCloud<P>::ConstPtr& c) { in_ = c; }
  void setLeafSize(float lx, float ly, float lz) { leaf_[0] = lx; leaf_[1] = ly; leaf_[2] = lz; for (int a = 0; a < 3; ++a) inv_[a] = 1.0f / leaf_[a]; }
  void filter(PointCloud<P>& out) {
    const std::vector<P>& pts = in_->points;
    out.points.clear(); out.height = 1; out.is_dense = true; out.header = in_->header;
    if (pts.empty()) { out.width = 0; return; }
    float mn[3] = { std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max() };
    float mx[3] = { -mn[0], -mn[1], -mn[2] };
    for (const P& p : pts) { if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
      const float v[3] = { p.x, p.y, p.z }; for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], v[a]); mx[a] = std::max(mx[a], v[a]); } }
    int minb[3], maxb[3], div[3];
    for (int a = 0; a < 3; ++a) { minb[a] = (int)std::floor(mn[a] * inv_[a]); maxb[a] = (int)std::floor(mx[a] * inv_[a]); div[a] = maxb[a] - minb[a] + 1; }
    const int mul[3] = { 1, div[0], div[0] * div[1] };
    std::vector<refstub_detail::IdxPt> v; v.reserve(pts.size());
    for (std::size_t i = 0; i < pts.size(); ++i) { const P& p = pts[i];
      if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
      const int i0 = (int)(std::floor(p.x * inv_[0]) - (float)minb[0]);
      const int i1 = (int)(std::floor(p.y * inv_[1]) - (float)minb[1]);
      const int i2 = (int)(std::floor(p.z * inv_[2]) - (float)minb[2]);
      v.push_back({ (unsigned)(i0 * mul[0] + i1 * mul[1] + i2 * mul[2]), (unsigned)i }); }
    std::sort(v.begin(), v.end(), std::less<refstub_detail::IdxPt>());
    for (std::size_t f = 0; f < v.size();) {
      std::size_t l = f + 1; while (l < v.size() && v[l].idx == v[f].idx) ++l;
      float s[4] = { 0, 0, 0, 0 };
      for (std::size_t k = f; k < l; ++k) { const P& p = pts[v[k].pt]; s[0] += p.x; s[1] += p.y; s[2] += p.z; s[3] += refstub_detail::field4(p); }
      const float cnt = (float)(l - f);
      P o; o.x = s[0] / cnt; o.y = s[1] / cnt; o.z = s[2] / cnt; refstub_detail::set_field4(o, s[3] / cnt);
      out.points.push_back(o); f = l; }
    out.width = (std::uint32_t)out.points.size();
  }
 private:
  typename PointCloud<P>::ConstPtr in_; float leaf_[3] = { 1, 1, 1 }, inv_[3] = { 1, 1, 1 };
};
}
#include <pcl/point_types.h>
namespace pcl { namespace refstub_detail {
template <> inline float field4<PointXYZI>(const PointXYZI& p) { return p.intensity; }
template <> inline void set_field4<PointXYZI>(PointXYZI& p, float v) { p.intensity = v; }
} }
