#pragma once
#include <memory>
#include <vector>
#include <cstdint>
namespace pcl { template <class P> struct PointCloud { std::vector<P> points; std::uint32_t width = 0, height = 0; bool is_dense = true;
  typedef std::shared_ptr<PointCloud<P>> Ptr; typedef P* iterator;
  PointCloud& operator+=(const PointCloud& o) { points.insert(points.end(), o.points.begin(), o.points.end()); return *this; }
  void clear() { points.clear(); } void push_back(const P& p) { points.push_back(p); } std::size_t size() const { return points.size(); } }; }
