// monolio_map_build_node with MapBuilder::associateToMap (mono_lidar_mapping/src/map_builder/Map_Builder.cc:213-332) and
// the extrinsic transform in front of it (src/map_build_node.cc:216-225) replaced by ONE lmono_project_color() call:
// raster, depthFill and the per-pixel lift run on the GPU; the node keeps the message pairing rules (:129-170), the
// skip distance (:181), every topic, frame id and YAML key (:276-297), the rgb_map accumulation thread
// (Map_Builder.cc:8-106) and the two debug images.  ~pro_map (r = 3 HSV discs, :240-265) is drawn here with the
// reference's own cv::circle call from the projection the library returns (lmono_color_projection).
//
// The reference publishes nothing on /compact_data: launch/map_build_b200.launch remaps it to
// /velodyne_cloud_3, the full-resolution sweep of the odometry node.
#include <ros/ros.h>
#include <cv_bridge/cv_bridge.h>
#include <nav_msgs/Odometry.h>
#include <sensor_msgs/Image.h>
#include <sensor_msgs/PointCloud2.h>
#include <sensor_msgs/image_encodings.h>
#include <pcl/io/ply_io.h>
#include <pcl_conversions/pcl_conversions.h>
#include <opencv2/opencv.hpp>
#include <Eigen/Geometry>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <mutex>
#include <queue>
#include <thread>
#include <vector>
#include "camodocal/camera_models/CameraFactory.h"
#include "camodocal/camera_models/PinholeCamera.h"
#include "lmono_ros_glue.hpp"

namespace {
lmono_ctx* g_ctx = nullptr;
lmono_pinhole g_cam;
Eigen::Vector3d tlc(0, 0, 0);
Eigen::Matrix3d rlc = Eigen::Matrix3d::Identity();
std::string IMAGE_TOPIC_0;
double SKIP_DIS = 0, DELAY_TIME = 0;
int SAVE_MAP = 0;
std::queue<sensor_msgs::ImageConstPtr> image_buf;
std::queue<sensor_msgs::PointCloud2ConstPtr> point_buf;
std::queue<nav_msgs::Odometry::ConstPtr> pose_buf;
std::queue<std::pair<double, pcl::PointCloud<pcl::PointXYZRGB>>> rgb_points_buf;
std::mutex buf_mutex, process_mutex, map_mutex;
Eigen::Vector3d last_T(0, 0, 0);
ros::Publisher pub_depth_map, pub_rgb_points, pub_pro_img, pub_rgb_map;

double clip(double n, double lower, double upper) { return std::max(lower, std::min(n, upper)); }

// Map_Builder.cc:8-106
void process_mapping() {
  pcl::PointCloud<pcl::PointXYZRGB>::Ptr rgb_map;
  int map_index = 0;
  while (ros::ok()) {
    std::pair<double, pcl::PointCloud<pcl::PointXYZRGB>> item;
    bool have = false;
    { std::lock_guard<std::mutex> l(map_mutex); if (!rgb_points_buf.empty()) { item = rgb_points_buf.front(); rgb_points_buf.pop(); have = true; } }
    if (have) {
      if (rgb_map) *rgb_map += item.second; else { rgb_map.reset(new pcl::PointCloud<pcl::PointXYZRGB>); *rgb_map = item.second; }
      printf("mapping at %d\n", map_index);
      map_index++;
      sensor_msgs::PointCloud2 msg; pcl::toROSMsg(*rgb_map, msg);
      msg.header.frame_id = "camera_init"; msg.header.stamp = ros::Time(item.first);
      pub_rgb_map.publish(msg);
      if (map_index > 0 && map_index % 10 == 0) {
        if (SAVE_MAP) pcl::io::savePLYFileBinary("/home/bo/raw_data/map/rgb_map" + std::to_string(map_index) + ".ply", *rgb_map);
        rgb_map->clear();
      }
    }
    std::this_thread::sleep_for(std::chrono::milliseconds(10));
  }
}

// replaces MapBuilder::associateToMap
void associate_to_map(const Eigen::Quaterniond& Q, const Eigen::Vector3d& T, const pcl::PointCloud<pcl::PointXYZ>& cloud, const cv::Mat& frame, double t) {
  const int W = frame.cols, H = frame.rows, npix = W * H;
  // map_build_node.cc:216-222: the 3x4 [rlc^T | -rlc^T tlc] the reference hands to pcl::transformPointCloud
  const Eigen::Matrix3d Rt = rlc.transpose();
  const Eigen::Vector3d tt = -1.0 * (Rt * tlc);
  double T34[12];
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) T34[4 * r + c] = Rt(r, c); T34[4 * r + 3] = tt(r); }
  lmono_pose QT = {{Q.x(), Q.y(), Q.z(), Q.w()}, {T.x(), T.y(), T.z()}};
  static std::vector<uint8_t> depth_raw, depth_filled, rgb;
  static std::vector<float> cam_xyz, world_xyz, uvz;
  depth_raw.resize(npix); depth_filled.resize(npix); rgb.resize((size_t)npix * 3); cam_xyz.resize((size_t)npix * 3); world_xyz.resize((size_t)npix * 3);
  int32_t n_out = 0;
  g_cam.width = W; g_cam.height = H;
  lmono_glue::check(lmono_project_color(g_ctx, lmono_glue::view(cloud), T34, frame.data, (int32_t)frame.step, &g_cam, &QT,
                                        depth_raw.data(), depth_filled.data(), cam_xyz.data(), world_xyz.data(), rgb.data(), npix, &n_out),
                    "lmono_project_color");
  // ~pro_map (Map_Builder.cc:219-221, 240-248): discs in cloud order on the HSV copy of the frame
  cv::Mat HSV_MAT, SHOW_MAT;
  cv::cvtColor(frame, HSV_MAT, cv::COLOR_BGR2HSV);
  const int n = (int)cloud.points.size();
  uvz.resize((size_t)n * 3);
  lmono_glue::check(lmono_color_projection(g_ctx, uvz.data(), n), "lmono_color_projection");
  for (int i = 0; i < n; ++i) {
    if (std::isnan(uvz[3 * i])) continue;
    const double new_depth = clip((double)uvz[3 * i + 2], 0, 100);
    cv::circle(HSV_MAT, cv::Point2f(uvz[3 * i], uvz[3 * i + 1]), 3, cv::Scalar(int(new_depth * 6), 255, 255), -1);
  }
  cv::cvtColor(HSV_MAT, SHOW_MAT, cv::COLOR_HSV2BGR);
  cv::Mat depth_map(H, W, CV_8UC1, depth_filled.data()), heat_map;
  cv::applyColorMap(depth_map, heat_map, cv::COLORMAP_JET);                                // :250-251
  std_msgs::Header header; header.frame_id = "camera"; header.stamp = ros::Time(t);
  cv_bridge::CvImage pro; pro.header = header; pro.encoding = sensor_msgs::image_encodings::BGR8; pro.image = SHOW_MAT;
  pub_pro_img.publish(pro.toImageMsg());
  cv_bridge::CvImage dep; dep.header = header; dep.encoding = sensor_msgs::image_encodings::BGR8; dep.image = heat_map;
  pub_depth_map.publish(dep.toImageMsg());
  // :275-332 rgb_cloud (camera frame) and w_cloud (world frame), row-major pixel order
  pcl::PointCloud<pcl::PointXYZRGB> rgb_cloud, w_cloud;
  rgb_cloud.points.resize(n_out); w_cloud.points.resize(n_out);
  for (int i = 0; i < n_out; ++i) {
    pcl::PointXYZRGB a, b;
    a.x = cam_xyz[3 * i]; a.y = cam_xyz[3 * i + 1]; a.z = cam_xyz[3 * i + 2];
    b.x = world_xyz[3 * i]; b.y = world_xyz[3 * i + 1]; b.z = world_xyz[3 * i + 2];
    a.r = b.r = rgb[3 * i]; a.g = b.g = rgb[3 * i + 1]; a.b = b.b = rgb[3 * i + 2];
    rgb_cloud.points[i] = a; w_cloud.points[i] = b;
  }
  rgb_cloud.width = w_cloud.width = n_out; rgb_cloud.height = w_cloud.height = 1;
  { std::lock_guard<std::mutex> l(map_mutex); rgb_points_buf.push(std::make_pair(t, w_cloud)); }
  sensor_msgs::PointCloud2 pm; pcl::toROSMsg(rgb_cloud, pm);
  pm.header.frame_id = "camera"; pm.header.stamp = ros::Time(t);
  pub_rgb_points.publish(pm);
}

// map_build_node.cc:75-118
void on_extrinsic(const nav_msgs::Odometry::ConstPtr& m) {
  std::lock_guard<std::mutex> l(process_mutex);
  tlc = Eigen::Vector3d(m->pose.pose.position.x, m->pose.pose.position.y, m->pose.pose.position.z);
  rlc = Eigen::Quaterniond(m->pose.pose.orientation.w, m->pose.pose.orientation.x, m->pose.pose.orientation.y, m->pose.pose.orientation.z).toRotationMatrix();
}
void on_points(const sensor_msgs::PointCloud2ConstPtr& m) { std::lock_guard<std::mutex> l(buf_mutex); point_buf.push(m); }
void on_image(const sensor_msgs::ImageConstPtr& m) { std::lock_guard<std::mutex> l(buf_mutex); image_buf.push(m); }
void on_pose(const nav_msgs::Odometry::ConstPtr& m) { std::lock_guard<std::mutex> l(buf_mutex); pose_buf.push(m); }

// map_build_node.cc:120-232
void process() {
  while (ros::ok()) {
    sensor_msgs::ImageConstPtr image_msg; sensor_msgs::PointCloud2ConstPtr point_msg; nav_msgs::Odometry::ConstPtr pose_msg;
    {
      std::lock_guard<std::mutex> l(buf_mutex);
      if (!image_buf.empty() && !point_buf.empty() && !pose_buf.empty()) {
        const double ti = image_buf.front()->header.stamp.toSec();
        if (ti > pose_buf.front()->header.stamp.toSec()) pose_buf.pop();
        else if (ti > point_buf.front()->header.stamp.toSec() + DELAY_TIME) point_buf.pop();
        else if (ti < point_buf.front()->header.stamp.toSec() - DELAY_TIME) image_buf.pop();
        else if (image_buf.back()->header.stamp.toSec() >= pose_buf.front()->header.stamp.toSec()) {
          pose_msg = pose_buf.front(); pose_buf.pop();
          while (!pose_buf.empty()) pose_buf.pop();
          while (image_buf.front()->header.stamp.toSec() < pose_msg->header.stamp.toSec()) { image_buf.pop(); point_buf.pop(); }
          image_msg = image_buf.front(); image_buf.pop();
          point_msg = point_buf.front(); point_buf.pop();
        }
      }
    }
    if (point_msg) {
      const Eigen::Vector3d T(pose_msg->pose.pose.position.x, pose_msg->pose.pose.position.y, pose_msg->pose.pose.position.z);
      const Eigen::Quaterniond Q(pose_msg->pose.pose.orientation.w, pose_msg->pose.pose.orientation.x, pose_msg->pose.pose.orientation.y, pose_msg->pose.pose.orientation.z);
      if ((T - last_T).norm() < SKIP_DIS) continue;
      cv_bridge::CvImageConstPtr ptr;
      if (image_msg->encoding == "8UC1") {
        sensor_msgs::Image img = *image_msg; img.encoding = "mono8";
        ptr = cv_bridge::toCvCopy(img, sensor_msgs::image_encodings::MONO8);
      } else ptr = cv_bridge::toCvCopy(image_msg, sensor_msgs::image_encodings::BGR8);
      cv::Mat image = ptr->image.clone();
      if (image.channels() == 1) cv::cvtColor(image, image, cv::COLOR_GRAY2BGR);      // the colour fetch of :303-305 reads three channels
      pcl::PointCloud<pcl::PointXYZ> cloud;
      pcl::fromROSMsg(*point_msg, cloud);
      std::lock_guard<std::mutex> l(process_mutex);
      associate_to_map(Q, T, cloud, image, image_msg->header.stamp.toSec());
    }
    std::this_thread::sleep_for(std::chrono::milliseconds(5));
  }
}
}  // namespace

int main(int argc, char** argv) {
  ros::init(argc, argv, "map_builder");
  ros::NodeHandle nh("~");
  std::string config_file, CAM0, KERNEL_TYPE, BLUR_TYPE;
  int KERNEL_SIZE = 5, FILTER_SIZE = 5;
  nh.getParam("map_config_file", config_file);
  cv::FileStorage fs(config_file, cv::FileStorage::READ);
  if (!fs.isOpened()) { ROS_WARN("config_file dosen't exist; wrong config_file path"); ROS_BREAK(); return 0; }
  fs["cam0_calib"] >> CAM0; fs["delay_time"] >> DELAY_TIME; fs["kernel_type"] >> KERNEL_TYPE; fs["blur_type"] >> BLUR_TYPE;
  fs["kernel_size"] >> KERNEL_SIZE; fs["skip_dis"] >> SKIP_DIS; fs["filter_size"] >> FILTER_SIZE; fs["image0_topic"] >> IMAGE_TOPIC_0; fs["save_map"] >> SAVE_MAP;
  // the pinhole parameters camodocal reads from cam0_calib (PinholeCamera::Parameters): only the model the hot path implements
  camodocal::CameraPtr cam = camodocal::CameraFactory::instance()->generateCameraFromYamlFile(CAM0);
  boost::shared_ptr<camodocal::PinholeCamera> pin = boost::dynamic_pointer_cast<camodocal::PinholeCamera>(cam);
  if (!pin) { ROS_ERROR("lmono_b200 implements the PINHOLE model only (config/kitti00_cam.yaml)"); return 1; }
  const camodocal::PinholeCamera::Parameters& P = pin->getParameters();
  g_cam.fx = P.fx(); g_cam.fy = P.fy(); g_cam.cx = P.cx(); g_cam.cy = P.cy(); g_cam.k1 = P.k1(); g_cam.k2 = P.k2(); g_cam.p1 = P.p1(); g_cam.p2 = P.p2();
  g_cam.width = P.imageWidth(); g_cam.height = P.imageHeight();
  g_cam.kernel_type = KERNEL_TYPE == "CROSS" ? 1 : (KERNEL_TYPE == "ELLIPSE" ? 2 : 0);      // Map_Builder.cc:344-356
  g_cam.kernel_size = KERNEL_SIZE;
  g_cam.blur_type = BLUR_TYPE == "gaussian" ? 1 : 0;                                          // :393-400
  lmono_params prm; lmono_default_params(&prm);
  prm.image_width = g_cam.width; prm.image_height = g_cam.height;
  prm.max_cubes_corner = prm.max_cubes_surf = 1;           // this node never touches the cube map
  lmono_glue::check(lmono_create(0, &prm, nullptr, &g_ctx), "lmono_create");
  ros::Subscriber s0 = nh.subscribe<nav_msgs::Odometry>("/fused/extrinsic", 2000, on_extrinsic);
  ros::Subscriber s1 = nh.subscribe<sensor_msgs::PointCloud2>("/compact_data", 2000, on_points);
  ros::Subscriber s2 = nh.subscribe<sensor_msgs::Image>(IMAGE_TOPIC_0, 2000, on_image);
  ros::Subscriber s3 = nh.subscribe<nav_msgs::Odometry>("/fused/new_camera_odometry", 2000, on_pose);
  pub_depth_map = nh.advertise<sensor_msgs::Image>("depth_map", 1000);
  pub_rgb_points = nh.advertise<sensor_msgs::PointCloud2>("rgb_points", 1000);
  pub_pro_img = nh.advertise<sensor_msgs::Image>("pro_map", 1000);
  pub_rgb_map = nh.advertise<sensor_msgs::PointCloud2>("rgb_map", 5);
  std::thread measurement_process{process};
  std::thread map_manager{process_mapping};
  ros::Rate r(5);
  while (ros::ok()) { ros::spinOnce(); r.sleep(); }
  measurement_process.join(); map_manager.join();
  lmono_destroy(g_ctx);
  return 0;
}
