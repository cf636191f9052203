// TEST INFRASTRUCTURE.  Compiles the reference's laserMapping node AS IT LIES under /root/reference (the translation
// unit, with its lidarFactor.hpp, is #included below; nothing is copied) against the functional stand-ins of
// oracle/refstubs/.  The node's process() (Aloam/src/laserMapping.cpp:232-903) runs unchanged on the caller's thread:
// main() is called for its set-up (parameters, publishers, the 4851 cube clouds), its std::thread is a stand-in that only
// remembers the function, and the 2 ms sleep at the bottom of process()'s outer loop hands control back once the message
// buffers are drained.  Window shifts, cube bookkeeping, association, line / plane gates, factor construction, insertion
// and the per-cube refilter are reference code; VoxelGrid, kd-tree, Eigen arithmetic / decompositions and the Ceres
// minimiser are the stand-ins (the minimiser being oracle/lm.c driven by the reference's cost functors).
// Built by `make -C oracle ref` into oracle/_ref/libref_mapping.so only where /root/reference exists.
#include <cmath>
#include <math.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <ctime>
#include <iostream>
#include <mutex>
#include <queue>
#include <string>
#include <thread>
#include <vector>
#include <ros/ros.h>
namespace refstub { struct Idle {}; typedef void (*ProcessFn)(); inline ProcessFn& process_fn() { static ProcessFn f = nullptr; return f; } }
namespace std {                                /* stand-ins reached through the two macros below */
struct refstub_thread { template <class F> refstub_thread(F f) { refstub::process_fn() = f; } };
namespace this_thread { template <class D> void refstub_sleep_for(const D&) { throw refstub::Idle(); } }
}
#define printf(...) ((void)0)                 /* the node reports timings on stdout */
#define thread refstub_thread
#define sleep_for refstub_sleep_for
#define main ref_mapping_main
#include "laserMapping.cpp"                   /* -I/root/reference/Aloam/src */
#undef main
#undef sleep_for
#undef thread
#undef printf

static sensor_msgs::PointCloud2ConstPtr make_msg(const float* xyzi, int n, double stamp) {
  pcl::PointCloud<PointType> c;
  c.points.resize((size_t)n);
  for (int i = 0; i < n; ++i) { c.points[i].x = xyzi[4 * i]; c.points[i].y = xyzi[4 * i + 1]; c.points[i].z = xyzi[4 * i + 2]; c.points[i].intensity = xyzi[4 * i + 3]; }
  sensor_msgs::PointCloud2Ptr m(new sensor_msgs::PointCloud2());
  pcl::toROSMsg(c, *m);
  m->header.stamp = ros::Time().fromSec(stamp);
  return m;
}

// a fresh node: globals back to their initial values (laserMapping.cpp:64-121), then main() for the set-up
extern "C" int ref_mapping_reset(double line_res, double plane_res) {
  std::cout.setstate(std::ios_base::failbit);
  refstub::State& S = refstub::state();
  S.params["mapping_line_resolution"] = line_res;
  S.params["mapping_plane_resolution"] = plane_res;
  frameCount = 0;
  laserCloudCenWidth = 10; laserCloudCenHeight = 10; laserCloudCenDepth = 5;
  parameters[0] = parameters[1] = parameters[2] = 0; parameters[3] = 1; parameters[4] = parameters[5] = parameters[6] = 0;
  q_wmap_wodom = Eigen::Quaterniond(1, 0, 0, 0); t_wmap_wodom = Eigen::Vector3d(0, 0, 0);
  q_wodom_curr = Eigen::Quaterniond(1, 0, 0, 0); t_wodom_curr = Eigen::Vector3d(0, 0, 0);
  while (!cornerLastBuf.empty()) cornerLastBuf.pop();
  while (!surfLastBuf.empty()) surfLastBuf.pop();
  while (!fullResBuf.empty()) fullResBuf.pop();
  while (!odometryBuf.empty()) odometryBuf.pop();
  laserAfterMappedPath = nav_msgs::Path();
  int argc = 1; char arg0[] = "alaserMapping"; char* argv[] = { arg0, nullptr };
  ref_mapping_main(argc, argv);
  return refstub::process_fn() ? 0 : -1;
}

// one sweep through the node's four callbacks and process().  pose_out: q_w_curr (x y z w), t_w_curr as published on
// /aft_mapped_to_init, then q_wmap_wodom, t_wmap_wodom; info: laserCloudCenWidth/Height/Depth, frameCount.
// full_inout (may be NULL): the full-resolution sweep, overwritten with /velodyne_cloud_registered.
extern "C" int ref_mapping_step(const float* corner_last, int nc, const float* surf_last, int ns, float* full_inout, int nfull,
                                const double q_odom[4], const double t_odom[3], double stamp, double pose_out[14], int32_t info[4]) {
  refstub::State& S = refstub::state();
  S.clouds.clear(); S.odoms.clear();
  nav_msgs::Odometry::Ptr od(new nav_msgs::Odometry());
  od->header.stamp = ros::Time().fromSec(stamp);
  od->pose.pose.orientation.x = q_odom[0]; od->pose.pose.orientation.y = q_odom[1]; od->pose.pose.orientation.z = q_odom[2]; od->pose.pose.orientation.w = q_odom[3];
  od->pose.pose.position.x = t_odom[0]; od->pose.pose.position.y = t_odom[1]; od->pose.pose.position.z = t_odom[2];
  refstub::Delivery d;
  d.clouds.push_back({ "/laser_cloud_corner_last", make_msg(corner_last, nc, stamp) });
  d.clouds.push_back({ "/laser_cloud_surf_last", make_msg(surf_last, ns, stamp) });
  d.clouds.push_back({ "/velodyne_cloud_3", make_msg(full_inout, full_inout ? nfull : 0, stamp) });
  d.odoms.push_back({ "/laser_odom_to_init", od });
  refstub::deliver(d);
  try { refstub::process_fn()(); } catch (const refstub::Idle&) {}
  auto it = S.odoms.find("/aft_mapped_to_init");
  if (it == S.odoms.end() || it->second.size() != 1) return -1;
  const nav_msgs::Odometry& o = it->second[0];
  pose_out[0] = o.pose.pose.orientation.x; pose_out[1] = o.pose.pose.orientation.y; pose_out[2] = o.pose.pose.orientation.z; pose_out[3] = o.pose.pose.orientation.w;
  pose_out[4] = o.pose.pose.position.x; pose_out[5] = o.pose.pose.position.y; pose_out[6] = o.pose.pose.position.z;
  pose_out[7] = q_wmap_wodom.x(); pose_out[8] = q_wmap_wodom.y(); pose_out[9] = q_wmap_wodom.z(); pose_out[10] = q_wmap_wodom.w();
  pose_out[11] = t_wmap_wodom.x(); pose_out[12] = t_wmap_wodom.y(); pose_out[13] = t_wmap_wodom.z();
  info[0] = laserCloudCenWidth; info[1] = laserCloudCenHeight; info[2] = laserCloudCenDepth; info[3] = frameCount;
  if (full_inout) {
    auto fr = S.clouds.find("/velodyne_cloud_registered");
    if (fr == S.clouds.end() || fr->second.size() != 1 || (int)fr->second[0].width != nfull) return -2;
    const sensor_msgs::PointCloud2& m = fr->second[0];
    for (int i = 0; i < nfull; ++i) { const unsigned char* p = m.data.data() + (size_t)i * m.point_step; std::memcpy(full_inout + 4 * i, p, 12); std::memcpy(full_inout + 4 * i + 3, p + 16, 4); }
  }
  return 0;
}

// the cube clouds in cube-index order (which 0 corner, 1 surf); returns the point count (out may be NULL to size it)
extern "C" int ref_mapping_export(int which, float* out, int cap) {
  int n = 0;
  for (int i = 0; i < laserCloudNum; ++i) {
    const pcl::PointCloud<PointType>& c = which == 0 ? *laserCloudCornerArray[i] : *laserCloudSurfArray[i];
    for (const PointType& p : c.points) { if (out && n < cap) { out[4 * n] = p.x; out[4 * n + 1] = p.y; out[4 * n + 2] = p.z; out[4 * n + 3] = p.intensity; } ++n; }
  }
  return n;
}

// harness conveniences for the timed arm (bench.py --impl reference), not node code: a prior map is dealt into the node's
// cube clouds by the node's cube rule (:741-757) and passed once through the node's own per-cube filters (:788-801);
// q/t_wmap_wodom can be put back to a given value between registrations.
extern "C" int ref_mapping_import(int which, const float* xyzi, int n) {
  pcl::PointCloud<PointType>::Ptr* arr = which == 0 ? laserCloudCornerArray : laserCloudSurfArray;
  for (int i = 0; i < n; ++i) {
    PointType p; p.x = xyzi[4 * i]; p.y = xyzi[4 * i + 1]; p.z = xyzi[4 * i + 2]; p.intensity = xyzi[4 * i + 3];
    int c[3]; const float v[3] = { p.x, p.y, p.z }; const int cen[3] = { laserCloudCenWidth, laserCloudCenHeight, laserCloudCenDepth };
    for (int a = 0; a < 3; ++a) { c[a] = int((v[a] + 25.0) / 50.0) + cen[a]; if (v[a] + 25.0 < 0) c[a]--; }
    if (c[0] < 0 || c[0] >= laserCloudWidth || c[1] < 0 || c[1] >= laserCloudHeight || c[2] < 0 || c[2] >= laserCloudDepth) continue;
    arr[c[0] + laserCloudWidth * c[1] + laserCloudWidth * laserCloudHeight * c[2]]->push_back(p);
  }
  for (int i = 0; i < laserCloudNum; ++i) {
    if (arr[i]->points.empty()) continue;
    pcl::PointCloud<PointType>::Ptr tmp(new pcl::PointCloud<PointType>());
    pcl::VoxelGrid<PointType>& f = which == 0 ? downSizeFilterCorner : downSizeFilterSurf;
    f.setInputCloud(arr[i]); f.filter(*tmp); arr[i] = tmp;
  }
  return 0;
}
extern "C" void ref_mapping_set_state(const double q[4], const double t[3]) {
  q_wmap_wodom = Eigen::Quaterniond(q[3], q[0], q[1], q[2]); t_wmap_wodom = Eigen::Vector3d(t[0], t[1], t[2]);
}
