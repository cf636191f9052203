// ascanRegistration with the per-sweep loops of Aloam/src/scanRegistration.cpp:132-408 replaced by
// one lmono_scan_register() call.  Node name, topics, message types, frame ids, queue sizes and
// parameters are those of the reference (scanRegistration.cpp:461-500, 413-441); the build needs
// ROS + PCL exactly like the original node and links liblmono_b200.so instead of doing the work on
// the CPU.  Not compilable in the graft container (no ROS); see INTEGRATION.md.
#include <ros/ros.h>
#include <sensor_msgs/PointCloud2.h>
#include <pcl_conversions/pcl_conversions.h>
#include "lmono_ros_glue.hpp"

namespace {
lmono_ctx* g_ctx = nullptr;
ros::Publisher pub_full, pub_sharp, pub_less_sharp, pub_flat, pub_less_flat, pub_removed;

void publish(ros::Publisher& pub, const pcl::PointCloud<pcl::PointXYZI>& cloud, const std_msgs::Header& in) {
  sensor_msgs::PointCloud2 msg;
  pcl::toROSMsg(cloud, msg);
  msg.header.stamp = in.stamp;
  msg.header.frame_id = "/camera_init";
  pub.publish(msg);
}

void on_sweep(const sensor_msgs::PointCloud2ConstPtr& in) {
  pcl::PointCloud<pcl::PointXYZ> raw;                       // intensity of the sensor is ignored (:132-133)
  pcl::fromROSMsg(*in, raw);
  pcl::PointCloud<pcl::PointXYZI> full, sharp, less_sharp, flat, less_flat;
  const size_t n = raw.points.size();
  lmono_cloud_out o_full = lmono_glue::out(full, n), o_sharp = lmono_glue::out(sharp, 64 * 12),
                  o_ls = lmono_glue::out(less_sharp, 64 * 120), o_flat = lmono_glue::out(flat, 64 * 24),
                  o_lf = lmono_glue::out(less_flat, n);
  lmono_scan_report rep;
  lmono_glue::check(lmono_scan_register(g_ctx, lmono_glue::view(raw), &o_full, &o_sharp, &o_ls, &o_flat, &o_lf, nullptr, &rep),
                    "lmono_scan_register");
  lmono_glue::trim(full, o_full); lmono_glue::trim(sharp, o_sharp); lmono_glue::trim(less_sharp, o_ls);
  lmono_glue::trim(flat, o_flat); lmono_glue::trim(less_flat, o_lf);
  printf("points size %d \n", rep.n_kept);
  printf("scan registration time %f ms *************\n", rep.ms_gpu);
  if (rep.ms_gpu > 100) ROS_WARN("scan registration process over 100ms");
  publish(pub_full, full, in->header);
  publish(pub_sharp, sharp, in->header);
  publish(pub_less_sharp, less_sharp, in->header);
  publish(pub_flat, flat, in->header);
  publish(pub_less_flat, less_flat, in->header);
}
}  // namespace

int main(int argc, char** argv) {
  ros::init(argc, argv, "scanRegistration");
  ros::NodeHandle nh;
  lmono_params prm;
  lmono_default_params(&prm);
  prm.stages = LMONO_STAGE_SCAN;       // this node never touches the cube map: no slab pools
  int scan_line = 16; double minimum_range = 0.1;
  nh.param<int>("scan_line", scan_line, 16);
  nh.param<double>("minimum_range", minimum_range, 0.1);
  printf("scan line number %d \n", scan_line);
  if (scan_line != 16 && scan_line != 32 && scan_line != 64) { printf("only support velodyne with 16, 32 or 64 scan line!"); return 0; }
  prm.scan_line = scan_line; prm.minimum_range = static_cast<float>(minimum_range);
  prm.max_cubes_corner = prm.max_cubes_surf = 1;      // this node never touches the cube map: no slab pool
  lmono_glue::check(lmono_create(0, &prm, nullptr, &g_ctx), "lmono_create");
  ros::Subscriber sub = nh.subscribe<sensor_msgs::PointCloud2>("/velodyne_points", 100, on_sweep);
  pub_full = nh.advertise<sensor_msgs::PointCloud2>("/velodyne_cloud_2", 100);
  pub_sharp = nh.advertise<sensor_msgs::PointCloud2>("/laser_cloud_sharp", 100);
  pub_less_sharp = nh.advertise<sensor_msgs::PointCloud2>("/laser_cloud_less_sharp", 100);
  pub_flat = nh.advertise<sensor_msgs::PointCloud2>("/laser_cloud_flat", 100);
  pub_less_flat = nh.advertise<sensor_msgs::PointCloud2>("/laser_cloud_less_flat", 100);
  pub_removed = nh.advertise<sensor_msgs::PointCloud2>("/laser_remove_points", 100);
  ros::spin();
  lmono_destroy(g_ctx);
  return 0;
}
