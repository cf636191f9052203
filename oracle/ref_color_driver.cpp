// TEST INFRASTRUCTURE.  Compiles the reference's colour-map node AS IT LIES under /root/reference
// (mono_lidar_mapping/src/map_build_node.cc and src/map_builder/Map_Builder.cc are #included below; nothing is copied)
// against the functional stand-ins of oracle/refstubs/.  Reference code on the path: the node's message handlers and
// process() (map_build_node.cc:73-238: extrinsic handling, buffer synchronisation, T = [rlc^T | -rlc^T tlc], the cloud
// transform call) and MapBuilder::associateToMap (Map_Builder.cc:213-334: projection raster with its implicit float ->
// int and double -> uchar conversions, depthFill's sequence of morphology / blur calls with the configured kernels, the
// per-pixel lift with its depth window and the |x| > 20 && y > 1.8 rule, the transform to the world frame) and camodocal's
// pinhole model, which lives in the reference tree too (camera_models/src/camera_models/Camera.cc + PinholeCamera.cc:
// constructor, spaceToPlane, liftProjective with its 8-step undistortion, distortion).  OpenCV's image operators, cv_bridge
// and pcl::transformPointCloud underneath are stand-ins (the image operators being the oracle's restatements of OpenCV).
// main() and the calibration half of the camera model are compiled but not run (they go through cv::FileStorage /
// calib3d, declaration-level stand-ins); the driver sets the node's globals itself.
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
namespace refstub { struct Idle {}; }
namespace std {                                /* stand-ins reached through the macros below */
struct refstub_thread { template <class F, class... A> refstub_thread(F, A...) {} };
namespace this_thread { template <class D> void refstub_sleep_for(const D&) { throw refstub::Idle(); } }
}
#define printf(...) ((void)0)
#define fprintf(...) ((void)0)                /* the node logs timings into a file under the author's home directory */
#define fflush(...) ((void)0)
#define thread refstub_thread
#define sleep_for refstub_sleep_for
#define main ref_mapnode_main
#include "map_build_node.cc"                  /* -I/root/reference/mono_lidar_mapping/src */
#undef main
#include "map_builder/Map_Builder.cc"
#include "camera_models/Camera.cc"            /* -I/root/reference/camera_models/src */
#include "camera_models/PinholeCamera.cc"
#undef sleep_for
#undef thread
#undef fflush
#undef fprintf
#undef printf

// CameraFactory.cc drags in every camera model of camodocal; only main() names the factory, and main() is not run
namespace camodocal {
boost::shared_ptr<CameraFactory> CameraFactory::m_instance;
CameraFactory::CameraFactory() {}
boost::shared_ptr<CameraFactory> CameraFactory::instance(void) { if (!m_instance) m_instance.reset(new CameraFactory()); return m_instance; }
CameraPtr CameraFactory::generateCameraFromYamlFile(const std::string&) { return CameraPtr(); }
}

static void configure(const o_camera* cam) {
  m_camera.reset(new camodocal::PinholeCamera("cam0", cam->width, cam->height, cam->k1, cam->k2, cam->p1, cam->p2, cam->fx, cam->fy, cam->cx, cam->cy));
  KERNEL_SIZE = cam->kernel_size;
  KERNEL_TYPE = cam->kernel_type == 0 ? "FULL" : cam->kernel_type == 1 ? "CROSS" : "ELLIPSE";
  BLUR_TYPE = cam->blur_type == 0 ? "bilateral" : "gaussian";
  DELAY_TIME = 0.05; SKIP_DIS = 0.0; SAVE_MAP = 0; FILTER_SIZE = 0;
  pub_rgb_points_.topic = "rgb_points";
  refstub::state().clouds.clear();
  cv::refstub_cv::log().dilate_inputs.clear();
}

static int collect(MapBuilder& mb, const o_camera* cam, uint8_t* depth_raw, uint8_t* depth_filled, float* cloud_cam_xyz, float* cloud_world_xyz,
                   uint8_t* cloud_rgb, int cap, int* n_out) {
  const size_t npix = (size_t)cam->height * cam->width;
  if (cv::refstub_cv::log().dilate_inputs.empty()) return -1;
  std::memcpy(depth_raw, cv::refstub_cv::log().dilate_inputs[0].data(), npix);
  std::memcpy(depth_filled, cv::refstub_cv::log().colormap_input.data(), npix);
  auto it = refstub::state().clouds.find("rgb_points");
  if (it == refstub::state().clouds.end() || it->second.size() != 1 || mb.rgb_points_buf.empty()) return -2;
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

// one frame straight through MapBuilder::associateToMap.  pts_cam: camera-frame XYZ; bgr: h x w x 3; QT: world pose of the
// camera.  Out: the raster before depthFill, the depth image after it, the lifted cloud in the camera frame (as published
// on pub_rgb_points_) and in the world frame (as queued in rgb_points_buf), its colours (r g b).
extern "C" int ref_color_frame(const float* pts_cam, int n, int stride_floats, const uint8_t* bgr, const o_camera* cam, const o_pose* QT,
                               uint8_t* depth_raw, uint8_t* depth_filled, float* cloud_cam_xyz, float* cloud_world_xyz, uint8_t* cloud_rgb,
                               int cap, int* n_out) {
  configure(cam);
  pcl::PointCloud<pcl::PointXYZ>::Ptr cloud(new pcl::PointCloud<pcl::PointXYZ>());
  cloud->points.resize((size_t)n);
  for (int i = 0; i < n; ++i) { cloud->points[i].x = pts_cam[(size_t)i * stride_floats]; cloud->points[i].y = pts_cam[(size_t)i * stride_floats + 1]; cloud->points[i].z = pts_cam[(size_t)i * stride_floats + 2]; }
  cv::Mat frame(cam->height, cam->width, CV_8UC3);
  std::memcpy(frame.data(), bgr, (size_t)cam->height * cam->width * 3);
  const Eigen::Quaterniond Q(QT->q[3], QT->q[0], QT->q[1], QT->q[2]);
  const Eigen::Vector3d T(QT->t[0], QT->t[1], QT->t[2]);
  MapBuilder mb;
  mb.associateToMap(Q, T, cloud, frame, 0.0);
  return collect(mb, cam, depth_raw, depth_filled, cloud_cam_xyz, cloud_world_xyz, cloud_rgb, cap, n_out);
}

// one frame through the NODE: /fused/extrinsic (lidar-to-camera extrinsic q_lc, t_lc), /compact_data (lidar-frame cloud),
// the image and /fused/new_camera_odometry go through the node's own handlers, then process() runs until it idles.
extern "C" int ref_mapnode_frame(const float* pts_lidar, int n, int stride_floats, const uint8_t* bgr, const o_camera* cam,
                                 const o_pose* extrinsic_lc, const o_pose* QT, double stamp,
                                 uint8_t* depth_raw, uint8_t* depth_filled, float* cloud_cam_xyz, float* cloud_world_xyz, uint8_t* cloud_rgb,
                                 int cap, int* n_out) {
  configure(cam);
  while (!map_builder.rgb_points_buf.empty()) map_builder.rgb_points_buf.pop();
  auto odo = [stamp](const o_pose* p) { nav_msgs::Odometry::Ptr m(new nav_msgs::Odometry()); m->header.stamp = ros::Time().fromSec(stamp);
    m->pose.pose.orientation.x = p->q[0]; m->pose.pose.orientation.y = p->q[1]; m->pose.pose.orientation.z = p->q[2]; m->pose.pose.orientation.w = p->q[3];
    m->pose.pose.position.x = p->t[0]; m->pose.pose.position.y = p->t[1]; m->pose.pose.position.z = p->t[2]; return m; };
  extrinsicHandler(odo(extrinsic_lc));
  sensor_msgs::PointCloud2Ptr pm(new sensor_msgs::PointCloud2());
  pm->header.stamp = ros::Time().fromSec(stamp); pm->width = (unsigned)n; pm->height = 1; pm->point_step = 16; pm->row_step = 16u * (unsigned)n;
  pm->data.assign((size_t)n * 16, 0);
  for (int i = 0; i < n; ++i) std::memcpy(pm->data.data() + (size_t)i * 16, pts_lidar + (size_t)i * stride_floats, 12);
  pointsHandler(pm);
  sensor_msgs::ImagePtr im(new sensor_msgs::Image());
  im->header.stamp = ros::Time().fromSec(stamp); im->height = (unsigned)cam->height; im->width = (unsigned)cam->width; im->encoding = "bgr8"; im->step = 3u * (unsigned)cam->width;
  im->data.assign(bgr, bgr + (size_t)cam->height * cam->width * 3);
  imageHandler(im);
  odomHandler(odo(QT));
  try { process(); } catch (const refstub::Idle&) {}
  return collect(map_builder, cam, depth_raw, depth_filled, cloud_cam_xyz, cloud_world_xyz, cloud_rgb, cap, n_out);
}
