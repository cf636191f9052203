// TEST INFRASTRUCTURE.  Compiles the reference's colour mapper AS IT LIES under /root/reference
// (mono_lidar_mapping/src/map_builder/Map_Builder.cc is #included below; nothing is copied) against the functional
// stand-ins of oracle/refstubs/.  MapBuilder::associateToMap (:213-334: projection raster with its implicit float ->
// int and double -> uchar conversions, depthFill's sequence of morphology / blur calls with the configured kernels,
// the per-pixel lift with its depth window and the |x| > 20 && y > 1.8 rule, the transform to the world frame) is
// reference code; OpenCV's image operators, camodocal's pinhole model and pcl::transformPointCloud underneath are
// stand-ins (the image operators being the oracle's restatements of OpenCV).
// Built by `make -C oracle ref` into oracle/_ref/libref_color.so only where /root/reference exists.
#include <cmath>
#include <math.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <iostream>
#include <mutex>
#include <queue>
#include <string>
#include <thread>
#include <vector>
#include <ros/ros.h>
#define printf(...) ((void)0)
#include "map_builder/Map_Builder.cc"         /* -I/root/reference/mono_lidar_mapping/src */
#undef printf

// the globals of mapping_parameter.h (defined in the node's parameter reader, which is not on the path)
Eigen::Vector3d tlc; Eigen::Matrix3d rlc; std::string CAM0; camodocal::CameraPtr m_camera; std::string IMAGE_TOPIC_0;
ros::Publisher pub_depth_map_, pub_rgb_points_, pub_pro_img_, pub_rgb_map_;
int SAVE_MAP = 0; double DELAY_TIME = 0; int KERNEL_SIZE = 5; double SKIP_DIS = 0; int FILTER_SIZE = 0;
std::string KERNEL_TYPE = "FULL", BLUR_TYPE = "bilateral";

// one frame through MapBuilder::associateToMap.  pts_cam: camera-frame XYZ; bgr: h x w x 3; QT: world pose of the camera.
// Out: the raster before depthFill, the depth image after it, the lifted cloud in the camera frame (as published on
// pub_rgb_points_) and in the world frame (as queued in rgb_points_buf), its colours (r g b).
extern "C" int ref_color_frame(const float* pts_cam, int n, int stride_floats, const uint8_t* bgr, const o_camera* cam, const o_pose* QT,
                               uint8_t* depth_raw, uint8_t* depth_filled, float* cloud_cam_xyz, float* cloud_world_xyz, uint8_t* cloud_rgb,
                               int cap, int* n_out) {
  m_camera.reset(new camodocal::PinholeCamera(cam->fx, cam->fy, cam->cx, cam->cy, cam->k1, cam->k2, cam->p1, cam->p2));
  KERNEL_SIZE = cam->kernel_size;
  KERNEL_TYPE = cam->kernel_type == 0 ? "FULL" : cam->kernel_type == 1 ? "CROSS" : "ELLIPSE";
  BLUR_TYPE = cam->blur_type == 0 ? "bilateral" : "gaussian";
  pub_rgb_points_.topic = "rgb_points";
  refstub::state().clouds.clear();
  cv::refstub_cv::log().dilate_inputs.clear();
  pcl::PointCloud<pcl::PointXYZ>::Ptr cloud(new pcl::PointCloud<pcl::PointXYZ>());
  cloud->points.resize((size_t)n);
  for (int i = 0; i < n; ++i) { cloud->points[i].x = pts_cam[(size_t)i * stride_floats]; cloud->points[i].y = pts_cam[(size_t)i * stride_floats + 1]; cloud->points[i].z = pts_cam[(size_t)i * stride_floats + 2]; }
  cv::Mat frame(cam->height, cam->width, CV_8UC3);
  std::memcpy(frame.data(), bgr, (size_t)cam->height * cam->width * 3);
  const Eigen::Quaterniond Q(QT->q[3], QT->q[0], QT->q[1], QT->q[2]);
  const Eigen::Vector3d T(QT->t[0], QT->t[1], QT->t[2]);
  MapBuilder mb;
  mb.associateToMap(Q, T, cloud, frame, 0.0);
  const size_t npix = (size_t)cam->height * cam->width;
  if (cv::refstub_cv::log().dilate_inputs.empty()) return -1;
  std::memcpy(depth_raw, cv::refstub_cv::log().dilate_inputs[0].data(), npix);
  std::memcpy(depth_filled, cv::refstub_cv::log().colormap_input.data(), npix);
  auto it = refstub::state().clouds.find("rgb_points");
  if (it == refstub::state().clouds.end() || it->second.size() != 1 || mb.rgb_points_buf.size() != 1) return -2;
  const sensor_msgs::PointCloud2& m = it->second[0];
  const pcl::PointCloud<pcl::PointXYZRGB>& w = mb.rgb_points_buf.back().second;
  const int np = (int)m.width;
  if ((int)w.points.size() != np) return -3;
  *n_out = np;
  for (int i = 0; i < np && i < cap; ++i) {
    const unsigned char* p = m.data.data() + (size_t)i * m.point_step;
    std::memcpy(cloud_cam_xyz + 3 * (size_t)i, p, 12);
    cloud_rgb[3 * (size_t)i] = p[18]; cloud_rgb[3 * (size_t)i + 1] = p[17]; cloud_rgb[3 * (size_t)i + 2] = p[16];
    cloud_world_xyz[3 * (size_t)i] = w.points[i].x; cloud_world_xyz[3 * (size_t)i + 1] = w.points[i].y; cloud_world_xyz[3 * (size_t)i + 2] = w.points[i].z;
  }
  return 0;
}
