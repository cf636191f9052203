// alaserMapping with the body of process() (Aloam/src/laserMapping.cpp:307-801, 838-842) replaced by
// lmono_map_step(); the rolling cube map stays resident on the GPU inside the lmono_ctx and is read
// back only for the /laser_cloud_surround (every 5th sweep) and /laser_cloud_map (every 20th)
// publishers.  Queue alignment, frame dropping and all topics follow laserMapping.cpp:235-305, 806-886,
// 908-926; the high-frequency odometry callback (:197-229) uses the wmap_wodom pose returned by the
// last step.
#include <ros/ros.h>
#include <nav_msgs/Odometry.h>
#include <nav_msgs/Path.h>
#include <geometry_msgs/PoseStamped.h>
#include <sensor_msgs/PointCloud2.h>
#include <tf/transform_broadcaster.h>
#include <pcl_conversions/pcl_conversions.h>
#include <Eigen/Geometry>
#include <chrono>
#include <mutex>
#include <queue>
#include <thread>
#include "lmono_ros_glue.hpp"

namespace {
lmono_ctx* g_ctx = nullptr;
std::mutex m_buf, m_pose;
std::queue<sensor_msgs::PointCloud2ConstPtr> q_corner, q_surf, q_full;
std::queue<nav_msgs::Odometry::ConstPtr> q_odom;
lmono_pose g_wmap_wodom = {{0, 0, 0, 1}, {0, 0, 0}};
ros::Publisher pub_surround, pub_map, pub_registered, pub_aft, pub_aft_hf, pub_path;
nav_msgs::Path g_path;

nav_msgs::Odometry to_odom(const lmono_pose& p, const ros::Time& stamp) {
  nav_msgs::Odometry o;
  o.header.frame_id = "/camera_init"; o.child_frame_id = "/aft_mapped"; o.header.stamp = stamp;
  o.pose.pose.orientation.x = p.q[0]; o.pose.pose.orientation.y = p.q[1]; o.pose.pose.orientation.z = p.q[2]; o.pose.pose.orientation.w = p.q[3];
  o.pose.pose.position.x = p.t[0]; o.pose.pose.position.y = p.t[1]; o.pose.pose.position.z = p.t[2];
  return o;
}

// named handlers: a lambda is ambiguous between roscpp's function-pointer and boost::function overloads of subscribe()
void on_corner(const sensor_msgs::PointCloud2ConstPtr& m) { std::lock_guard<std::mutex> l(m_buf); q_corner.push(m); }
void on_surf(const sensor_msgs::PointCloud2ConstPtr& m) { std::lock_guard<std::mutex> l(m_buf); q_surf.push(m); }
void on_full(const sensor_msgs::PointCloud2ConstPtr& m) { std::lock_guard<std::mutex> l(m_buf); q_full.push(m); }

void on_odom(const nav_msgs::Odometry::ConstPtr& m) {
  { std::lock_guard<std::mutex> l(m_buf); q_odom.push(m); }
  lmono_pose wm; { std::lock_guard<std::mutex> l(m_pose); wm = g_wmap_wodom; }
  const Eigen::Quaterniond q_wm(wm.q[3], wm.q[0], wm.q[1], wm.q[2]);
  const Eigen::Quaterniond q_o(m->pose.pose.orientation.w, m->pose.pose.orientation.x, m->pose.pose.orientation.y, m->pose.pose.orientation.z);
  const Eigen::Vector3d t_o(m->pose.pose.position.x, m->pose.pose.position.y, m->pose.pose.position.z);
  const Eigen::Quaterniond q = q_wm * q_o;
  const Eigen::Vector3d t = q_wm * t_o + Eigen::Vector3d(wm.t[0], wm.t[1], wm.t[2]);
  lmono_pose w = {{q.x(), q.y(), q.z(), q.w()}, {t.x(), t.y(), t.z()}};
  pub_aft_hf.publish(to_odom(w, m->header.stamp));
}

// /laser_cloud_surround and /laser_cloud_map: corner and surf clouds interleaved cube by cube, the byte order of the
// reference's messages (laserMapping.cpp:808-816, 826-830) -- lmono_map_export(which = 2)
void publish_map(ros::Publisher& pub, int scope, const ros::Time& stamp) {
  pcl::PointCloud<pcl::PointXYZI> all;
  size_t cap = 1 << 20;
  for (;;) {
    lmono_cloud_out o = lmono_glue::out(all, cap);
    const int rc = lmono_map_export(g_ctx, 2, scope, &o);
    if (rc == LMONO_E_CAPACITY) { cap = static_cast<size_t>(o.n_out); continue; }
    if (rc != LMONO_OK) { ROS_WARN("lmono_map_export: %s", lmono_strerror(rc)); return; }
    lmono_glue::trim(all, o);
    break;
  }
  sensor_msgs::PointCloud2 msg; pcl::toROSMsg(all, msg);
  msg.header.stamp = stamp; msg.header.frame_id = "/camera_init";
  pub.publish(msg);
}

void process() {
  int frame_count = 0;
  while (ros::ok()) {
    sensor_msgs::PointCloud2ConstPtr mc, ms, mf; nav_msgs::Odometry::ConstPtr mo;
    {
      std::lock_guard<std::mutex> l(m_buf);
      if (!q_corner.empty() && !q_surf.empty() && !q_full.empty() && !q_odom.empty()) {
        const double tc = q_corner.front()->header.stamp.toSec();
        while (!q_odom.empty() && q_odom.front()->header.stamp.toSec() < tc) q_odom.pop();
        while (!q_surf.empty() && q_surf.front()->header.stamp.toSec() < tc) q_surf.pop();
        while (!q_full.empty() && q_full.front()->header.stamp.toSec() < tc) q_full.pop();
        if (!q_odom.empty() && !q_surf.empty() && !q_full.empty()) {
          if (q_surf.front()->header.stamp.toSec() != q_odom.front()->header.stamp.toSec() ||
              q_full.front()->header.stamp.toSec() != q_odom.front()->header.stamp.toSec() || tc != q_odom.front()->header.stamp.toSec()) {
            printf("unsync messeage!");
          } else {
            mc = q_corner.front(); ms = q_surf.front(); mf = q_full.front(); mo = q_odom.front();
            q_corner.pop(); q_surf.pop(); q_full.pop(); q_odom.pop();
            while (!q_corner.empty()) { q_corner.pop(); printf("drop lidar frame in mapping for real time performance \n"); }
          }
        }
      }
    }
    if (mc) {
      pcl::PointCloud<pcl::PointXYZI> corner, surf, full, registered;
      pcl::fromROSMsg(*mc, corner); pcl::fromROSMsg(*ms, surf); pcl::fromROSMsg(*mf, full);
      lmono_pose odom = {{mo->pose.pose.orientation.x, mo->pose.pose.orientation.y, mo->pose.pose.orientation.z, mo->pose.pose.orientation.w},
                         {mo->pose.pose.position.x, mo->pose.pose.position.y, mo->pose.pose.position.z}};
      lmono_pose w_curr, wm; lmono_map_report rep;
      lmono_cloud_out o_reg = lmono_glue::out(registered, full.points.size());
      const int rc_step = lmono_map_step(g_ctx, lmono_glue::view(corner), lmono_glue::view(surf), &odom, &w_curr, &wm, &rep,
                                         lmono_glue::view(full), &o_reg);
      if (rc_step == LMONO_E_DEVICE) {      // a map capacity limit was hit (slab pool / cube slab): the pose is valid, some new points were dropped
        uint32_t bits = 0; lmono_last_fault(g_ctx, &bits);
        int32_t freed = 0;
        if (bits & LMONO_FAULT_POOL_EXHAUSTED) lmono_map_evict(g_ctx, 4, &freed);      // slab pool exhausted: drop the cubes more than 4 cubes (200 m) from the window centre
        ROS_WARN("lmono_map_step: map capacity reached (fault bits 0x%x), %d far cubes evicted: raise max_cubes_* / cube_capacity_* (INTEGRATION.md)", bits, freed);
      } else if (rc_step != LMONO_OK) {
        ROS_ERROR("lmono_map_step: %s -- sweep skipped", lmono_strerror(rc_step));
        continue;
      }
      lmono_glue::trim(registered, o_reg);
      { std::lock_guard<std::mutex> l(m_pose); g_wmap_wodom = wm; }
      printf("map corner num %d  surf num %d \n", rep.corner_from_map, rep.surf_from_map);
      if (!rep.optimized) ROS_WARN("time Map corner and surf num are not enough");
      printf("whole mapping time %f ms +++++\n", rep.ms_gpu);
      const ros::Time stamp = ros::Time().fromSec(mo->header.stamp.toSec());
      if (frame_count % 5 == 0) publish_map(pub_surround, 0, stamp);
      if (frame_count % 20 == 0) publish_map(pub_map, 1, stamp);
      sensor_msgs::PointCloud2 reg_msg; pcl::toROSMsg(registered, reg_msg);
      reg_msg.header.stamp = stamp; reg_msg.header.frame_id = "/camera_init";
      pub_registered.publish(reg_msg);
      nav_msgs::Odometry aft = to_odom(w_curr, stamp);
      pub_aft.publish(aft);
      geometry_msgs::PoseStamped ps; ps.header = aft.header; ps.pose = aft.pose.pose;
      g_path.header.stamp = aft.header.stamp; g_path.header.frame_id = "/camera_init"; g_path.poses.push_back(ps);
      pub_path.publish(g_path);
      static tf::TransformBroadcaster br;
      tf::Transform tr; tr.setOrigin(tf::Vector3(w_curr.t[0], w_curr.t[1], w_curr.t[2]));
      tf::Quaternion q; q.setW(w_curr.q[3]); q.setX(w_curr.q[0]); q.setY(w_curr.q[1]); q.setZ(w_curr.q[2]);
      tr.setRotation(q);
      br.sendTransform(tf::StampedTransform(tr, aft.header.stamp, "/camera_init", "/aft_mapped"));
      frame_count++;
    }
    std::this_thread::sleep_for(std::chrono::milliseconds(2));
  }
}
}  // namespace

int main(int argc, char** argv) {
  ros::init(argc, argv, "laserMapping");
  ros::NodeHandle nh;
  float line_res = 0, plane_res = 0;
  nh.param<float>("mapping_line_resolution", line_res, 0.4);
  nh.param<float>("mapping_plane_resolution", plane_res, 0.8);
  printf("line resolution %f plane resolution %f \n", line_res, plane_res);
  lmono_params prm; lmono_default_params(&prm);
  prm.mapping_line_resolution = line_res; prm.mapping_plane_resolution = plane_res;
  lmono_glue::check(lmono_create(0, &prm, nullptr, &g_ctx), "lmono_create");
  ros::Subscriber s0 = nh.subscribe<sensor_msgs::PointCloud2>("/laser_cloud_corner_last", 100, on_corner);
  ros::Subscriber s1 = nh.subscribe<sensor_msgs::PointCloud2>("/laser_cloud_surf_last", 100, on_surf);
  ros::Subscriber s2 = nh.subscribe<nav_msgs::Odometry>("/laser_odom_to_init", 100, on_odom);
  ros::Subscriber s3 = nh.subscribe<sensor_msgs::PointCloud2>("/velodyne_cloud_3", 100, on_full);
  pub_surround = nh.advertise<sensor_msgs::PointCloud2>("/laser_cloud_surround", 100);
  pub_map = nh.advertise<sensor_msgs::PointCloud2>("/laser_cloud_map", 100);
  pub_registered = nh.advertise<sensor_msgs::PointCloud2>("/velodyne_cloud_registered", 100);
  pub_aft = nh.advertise<nav_msgs::Odometry>("/aft_mapped_to_init", 100);
  pub_aft_hf = nh.advertise<nav_msgs::Odometry>("/aft_mapped_to_init_high_frec", 100);
  pub_path = nh.advertise<nav_msgs::Path>("/aft_mapped_path", 100);
  std::thread worker{process};
  ros::spin();
  worker.join();
  lmono_destroy(g_ctx);
  return 0;
}
