// alaserOdometry with the association + ceres::Solve loop of Aloam/src/laserOdometry.cpp:265-568
// replaced by lmono_odom_step().  The previous sweep's clouds and the warm-start pose
// (para_q / para_t, :97-98) live in the lmono_ctx.  Topics, stamps, the unsync ROS_BREAK and the
// mapping_skip_frame cadence follow laserOdometry.cpp:195-213, 224-241, 511-530, 570-591.
#include <ros/ros.h>
#include <nav_msgs/Odometry.h>
#include <nav_msgs/Path.h>
#include <geometry_msgs/PoseStamped.h>
#include <sensor_msgs/PointCloud2.h>
#include <pcl_conversions/pcl_conversions.h>
#include <mutex>
#include <queue>
#include "lmono_ros_glue.hpp"

namespace {
std::mutex m_buf;
std::queue<sensor_msgs::PointCloud2ConstPtr> q_sharp, q_less_sharp, q_flat, q_less_flat, q_full;
template <std::queue<sensor_msgs::PointCloud2ConstPtr>* Q>
void push(const sensor_msgs::PointCloud2ConstPtr& m) { std::lock_guard<std::mutex> l(m_buf); Q->push(m); }
}  // namespace

int main(int argc, char** argv) {
  ros::init(argc, argv, "laserOdometry");
  ros::NodeHandle nh;
  int skip_frame_num = 2;
  nh.param<int>("mapping_skip_frame", skip_frame_num, 2);
  printf("Mapping %d Hz \n", 10 / skip_frame_num);
  lmono_params prm; lmono_default_params(&prm); prm.mapping_skip_frame = skip_frame_num; prm.stages = LMONO_STAGE_ODOMETRY;
  lmono_ctx* ctx = nullptr;
  prm.max_cubes_corner = prm.max_cubes_surf = 1;      // this node never touches the cube map: no slab pool
  lmono_glue::check(lmono_create(0, &prm, nullptr, &ctx), "lmono_create");

  ros::Subscriber s0 = nh.subscribe<sensor_msgs::PointCloud2>("/laser_cloud_sharp", 100, push<&q_sharp>);
  ros::Subscriber s1 = nh.subscribe<sensor_msgs::PointCloud2>("/laser_cloud_less_sharp", 100, push<&q_less_sharp>);
  ros::Subscriber s2 = nh.subscribe<sensor_msgs::PointCloud2>("/laser_cloud_flat", 100, push<&q_flat>);
  ros::Subscriber s3 = nh.subscribe<sensor_msgs::PointCloud2>("/laser_cloud_less_flat", 100, push<&q_less_flat>);
  ros::Subscriber s4 = nh.subscribe<sensor_msgs::PointCloud2>("/velodyne_cloud_2", 100, push<&q_full>);
  ros::Publisher pub_corner_last = nh.advertise<sensor_msgs::PointCloud2>("/laser_cloud_corner_last", 100);
  ros::Publisher pub_surf_last = nh.advertise<sensor_msgs::PointCloud2>("/laser_cloud_surf_last", 100);
  ros::Publisher pub_full = nh.advertise<sensor_msgs::PointCloud2>("/velodyne_cloud_3", 100);
  ros::Publisher pub_odom = nh.advertise<nav_msgs::Odometry>("/laser_odom_to_init", 100);
  ros::Publisher pub_path = nh.advertise<nav_msgs::Path>("/laser_odom_path", 100);
  nav_msgs::Path path;
  int frame_count = 0;
  ros::Rate rate(100);
  while (ros::ok()) {
    ros::spinOnce();
    sensor_msgs::PointCloud2ConstPtr m_sharp, m_ls, m_flat, m_lf, m_full;
    {
      std::lock_guard<std::mutex> l(m_buf);
      if (!q_sharp.empty() && !q_less_sharp.empty() && !q_flat.empty() && !q_less_flat.empty() && !q_full.empty()) {
        m_sharp = q_sharp.front(); m_ls = q_less_sharp.front(); m_flat = q_flat.front(); m_lf = q_less_flat.front(); m_full = q_full.front();
        const double t = m_full->header.stamp.toSec();
        if (m_sharp->header.stamp.toSec() != t || m_ls->header.stamp.toSec() != t || m_flat->header.stamp.toSec() != t ||
            m_lf->header.stamp.toSec() != t) { printf("unsync messeage!"); ROS_BREAK(); }
        q_sharp.pop(); q_less_sharp.pop(); q_flat.pop(); q_less_flat.pop(); q_full.pop();
      }
    }
    if (m_full) {
      pcl::PointCloud<pcl::PointXYZI> sharp, less_sharp, flat, less_flat;
      pcl::fromROSMsg(*m_sharp, sharp); pcl::fromROSMsg(*m_ls, less_sharp);
      pcl::fromROSMsg(*m_flat, flat); pcl::fromROSMsg(*m_lf, less_flat);
      lmono_pose last_curr, w_curr; lmono_odom_report rep;
      lmono_glue::check(lmono_odom_step(ctx, lmono_glue::view(sharp), lmono_glue::view(less_sharp), lmono_glue::view(flat),
                                        lmono_glue::view(less_flat), &last_curr, &w_curr, &rep), "lmono_odom_step");
      if (!rep.inited) std::cout << "Initialization finished \n";
      else if (rep.corner_corr[1] + rep.plane_corr[1] < 10) printf("less correspondence! *************************************************\n");
      const ros::Time stamp = ros::Time().fromSec(m_lf->header.stamp.toSec());
      nav_msgs::Odometry odom;
      odom.header.frame_id = "/camera_init"; odom.child_frame_id = "/laser_odom"; odom.header.stamp = stamp;
      odom.pose.pose.orientation.x = w_curr.q[0]; odom.pose.pose.orientation.y = w_curr.q[1];
      odom.pose.pose.orientation.z = w_curr.q[2]; odom.pose.pose.orientation.w = w_curr.q[3];
      odom.pose.pose.position.x = w_curr.t[0]; odom.pose.pose.position.y = w_curr.t[1]; odom.pose.pose.position.z = w_curr.t[2];
      pub_odom.publish(odom);
      geometry_msgs::PoseStamped ps; ps.header = odom.header; ps.pose = odom.pose.pose;
      path.header.stamp = odom.header.stamp; path.header.frame_id = "/camera_init"; path.poses.push_back(ps);
      pub_path.publish(path);
      if (frame_count % skip_frame_num == 0) {
        frame_count = 0;
        sensor_msgs::PointCloud2 c = *m_ls, s = *m_lf, f = *m_full;   // this sweep's less-sharp / less-flat become "last"
        c.header.stamp = s.header.stamp = f.header.stamp = stamp;
        c.header.frame_id = s.header.frame_id = f.header.frame_id = "/camera";
        pub_corner_last.publish(c); pub_surf_last.publish(s); pub_full.publish(f);
      }
      printf("whole laserOdometry time %f ms \n \n", rep.ms_gpu);
      if (rep.ms_gpu > 100) ROS_WARN("odometry process over 100ms");
      frame_count++;
    }
    rate.sleep();
  }
  lmono_destroy(ctx);
  return 0;
}
