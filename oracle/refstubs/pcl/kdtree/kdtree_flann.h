#pragma once
// stand-in for pcl::KdTreeFLANN<PointT>: exact k nearest neighbours through the oracle's restatement of FLANN's
// KDTreeSingleIndex (oracle/kdtree.c).  Library stand-in, not reference source.
#include <memory>
#include <vector>
#include <pcl/point_cloud.h>
#include "lmono_oracle.h"
namespace pcl {
template <class P> class KdTreeFLANN {
 public:
  typedef boost::shared_ptr<KdTreeFLANN<P>> Ptr;
  KdTreeFLANN() {}
  ~KdTreeFLANN() { if (t_) lmono_cpu_kdtree_free(t_); }
  KdTreeFLANN(const KdTreeFLANN&) = delete; KdTreeFLANN& operator=(const KdTreeFLANN&) = delete;
  void setInputCloud(const typename PointCloud<P>::ConstPtr& c) {
    if (t_) { lmono_cpu_kdtree_free(t_); t_ = nullptr; }
    std::vector<o_pt> pts(c->points.size());
    for (std::size_t i = 0; i < pts.size(); ++i) { pts[i].x = c->points[i].x; pts[i].y = c->points[i].y; pts[i].z = c->points[i].z; pts[i].i = 0; }
    n_ = (int)pts.size();
    if (n_ > 0) t_ = lmono_cpu_kdtree_build(pts.data(), n_);
  }
  int nearestKSearch(const P& p, int k, std::vector<int>& idx, std::vector<float>& d2) const {
    idx.assign((std::size_t)k, 0); d2.assign((std::size_t)k, 0.0f);
    if (!t_ || n_ < k) { idx.clear(); d2.clear(); return 0; }
    o_pt q; q.x = p.x; q.y = p.y; q.z = p.z; q.i = 0;
    std::vector<int32_t> ii((std::size_t)k);
    lmono_cpu_kdtree_knn(t_, &q, 1, k, ii.data(), d2.data());
    for (int j = 0; j < k; ++j) idx[(std::size_t)j] = ii[(std::size_t)j];
    return k;
  }
 private:
  o_kdtree* t_ = nullptr; int n_ = 0;
};
}
