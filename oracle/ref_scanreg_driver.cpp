// TEST INFRASTRUCTURE.  Compiles the reference's scanRegistration node AS IT LIES under /root/reference (the
// translation unit is #included below, nothing is copied) against the functional stand-ins of oracle/refstubs/, and
// exposes one C entry point that pushes a cloud through the node's own callback (laserCloudHandler,
// Aloam/src/scanRegistration.cpp:113-459) and hands back what the node published plus its label / curvature arrays.
// Used by tests/test_oracle_vs_ref.py to pin oracle/scan_registration.c; built by `make -C oracle ref` into
// oracle/_ref/libref_scanreg.so only where /root/reference exists.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <ctime>
#include <string>
#include <vector>
#include <ros/ros.h>
#define printf(...) ((void)0)                 /* the node reports timings on stdout */
#define main ref_scanreg_main
#include "scanRegistration.cpp"               /* -I/root/reference/Aloam/src */
#undef main
#undef printf

static int copy_out(const char* topic, float* dst, int cap) {
  auto it = refstub::state().clouds.find(topic);
  if (it == refstub::state().clouds.end() || it->second.empty()) return -1;
  const sensor_msgs::PointCloud2& m = it->second.back();
  const int n = (int)m.width;
  if (n > cap) return -2;
  for (int i = 0; i < n; ++i) {
    const unsigned char* p = m.data.data() + (size_t)i * m.point_step;
    std::memcpy(dst + 4 * i, p, 12);
    std::memcpy(dst + 4 * i + 3, p + 16, 4);
  }
  return n;
}

extern "C" int ref_scan_register(const float* xyz, int n, int stride_floats, int n_scans, double minimum_range,
                                 float* full, float* sharp, float* less_sharp, float* flat, float* less_flat, int cap,
                                 int32_t* counts /*5*/, int32_t* labels, float* curvature) {
  refstub::state().params["scan_line"] = n_scans;
  refstub::state().params["minimum_range"] = minimum_range;
  refstub::state().clouds.clear();
  int argc = 1; char arg0[] = "ascanRegistration"; char* argv[] = { arg0, nullptr };
  ref_scanreg_main(argc, argv);                                   // parameters, advertise, subscribe; the stand-in spin() returns
  auto sub = refstub::state().cloud_subs.find("/velodyne_points");
  if (sub == refstub::state().cloud_subs.end()) return -1;
  std::memset(cloudLabel, 0, sizeof(cloudLabel));                 // a fresh process starts from zeroed globals
  std::memset(cloudCurvature, 0, sizeof(cloudCurvature));
  std::memset(cloudNeighborPicked, 0, sizeof(cloudNeighborPicked));
  std::memset(cloudSortInd, 0, sizeof(cloudSortInd));
  sensor_msgs::PointCloud2Ptr msg(new sensor_msgs::PointCloud2());
  msg->width = (unsigned)n; msg->height = 1; msg->point_step = 16; msg->row_step = 16u * (unsigned)n;
  msg->data.assign((size_t)n * 16, 0);
  for (int i = 0; i < n; ++i) std::memcpy(msg->data.data() + (size_t)i * 16, xyz + (size_t)i * stride_floats, 12);
  sub->second(msg);
  counts[0] = copy_out("/velodyne_cloud_2", full, cap);
  counts[1] = copy_out("/laser_cloud_sharp", sharp, cap);
  counts[2] = copy_out("/laser_cloud_less_sharp", less_sharp, cap);
  counts[3] = copy_out("/laser_cloud_flat", flat, cap);
  counts[4] = copy_out("/laser_cloud_less_flat", less_flat, cap);
  for (int k = 0; k < 5; ++k) if (counts[k] < 0) return -2;
  for (int i = 0; i < counts[0]; ++i) { labels[i] = cloudLabel[i]; curvature[i] = cloudCurvature[i]; }
  return 0;
}
