#pragma once
// stand-in ROS: parameters come from a table the driver fills, publishers record the last message per topic,
// subscriptions are recorded so that the driver can deliver messages, spin() returns at once
#include <cstdio>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include <ros/time.h>
#include <sensor_msgs/PointCloud2.h>
namespace refstub {
inline std::map<std::string, double>& params() { static std::map<std::string, double> p; return p; }
inline std::map<std::string, sensor_msgs::PointCloud2>& published() { static std::map<std::string, sensor_msgs::PointCloud2> p; return p; }
inline std::map<std::string, std::function<void(const sensor_msgs::PointCloud2ConstPtr&)>>& cloud_subs() {
  static std::map<std::string, std::function<void(const sensor_msgs::PointCloud2ConstPtr&)>> s; return s; }
}
namespace ros {
struct Publisher { std::string topic;
  void publish(const sensor_msgs::PointCloud2& m) const { refstub::published()[topic] = m; }
  template <class M> void publish(const M&) const {} };
struct Subscriber {};
struct NodeHandle {
  NodeHandle() {} explicit NodeHandle(const std::string&) {}
  template <class T> bool param(const std::string& k, T& v, const T& dflt) const {
    auto it = refstub::params().find(k); v = it == refstub::params().end() ? dflt : (T)it->second; return it != refstub::params().end(); }
  template <class M> Publisher advertise(const std::string& topic, unsigned) { Publisher p; p.topic = topic; return p; }
  template <class M> Subscriber subscribe(const std::string& topic, unsigned, void (*cb)(const boost::shared_ptr<M const>&)) { reg(topic, cb); return Subscriber(); }
 private:
  static void reg(const std::string& topic, void (*cb)(const sensor_msgs::PointCloud2ConstPtr&)) { refstub::cloud_subs()[topic] = cb; }
  template <class F> static void reg(const std::string&, F) {}
};
inline void init(int&, char**, const std::string&) {}
inline bool ok() { return false; }
inline void spin() {}
inline void spinOnce() {}
}
#define ROS_INFO(...) ((void)0)
#define ROS_WARN(...) ((void)0)
#define ROS_ERROR(...) ((void)0)
#define ROS_BREAK() std::abort()
