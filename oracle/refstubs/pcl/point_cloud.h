#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>
namespace boost { template <class T> using shared_ptr = std::shared_ptr<T>; }
namespace pcl {
struct PCLHeader { std::uint32_t seq = 0; std::uint64_t stamp = 0; std::string frame_id; };
template <class P> struct PointCloud {
  typedef boost::shared_ptr<PointCloud<P>> Ptr; typedef boost::shared_ptr<const PointCloud<P>> ConstPtr;
  typedef typename std::vector<P>::iterator iterator; typedef typename std::vector<P>::const_iterator const_iterator;
  PCLHeader header; std::vector<P> points; std::uint32_t width = 0, height = 0; bool is_dense = true;
  PointCloud& operator+=(const PointCloud& o) {          // point_cloud.h: append, width = size, height = 1
    points.insert(points.end(), o.points.begin(), o.points.end()); width = (std::uint32_t)points.size(); height = 1;
    if (!o.is_dense) is_dense = false; return *this; }
  void clear() { points.clear(); width = height = 0; }
  void push_back(const P& p) { points.push_back(p); width = (std::uint32_t)points.size(); height = 1; }
  std::size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  P& operator[](std::size_t i) { return points[i]; } const P& operator[](std::size_t i) const { return points[i]; }
};
}
