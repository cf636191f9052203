#pragma once
// stand-in for pcl::transformPointCloud (PCL 1.8 transforms.hpp, dense cloud, Matrix<Scalar, 4, 4>): the three rows of
// the transform applied to (x, y, z) in Scalar, each coordinate stored as float; the other fields are copied
#include <eigen3/Eigen/Dense>
#include <pcl/point_cloud.h>
namespace pcl {
template <class P, class S> void transformPointCloud(const PointCloud<P>& in, PointCloud<P>& out, const Eigen::Matrix<S, 4, 4>& T) {
  if (&in != &out) { out.header = in.header; out.is_dense = in.is_dense; out.width = in.width; out.height = in.height; out.points = in.points; }
  for (std::size_t i = 0; i < out.points.size(); ++i) {
    const S x = in.points[i].x, y = in.points[i].y, z = in.points[i].z;
    out.points[i].x = static_cast<float>(T(0, 0) * x + T(0, 1) * y + T(0, 2) * z + T(0, 3));
    out.points[i].y = static_cast<float>(T(1, 0) * x + T(1, 1) * y + T(1, 2) * z + T(1, 3));
    out.points[i].z = static_cast<float>(T(2, 0) * x + T(2, 1) * y + T(2, 2) * z + T(2, 3));
  }
}
}
