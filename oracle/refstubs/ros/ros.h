#pragma once
// stand-in ROS: parameters come from a table the driver fills, publishers record every message per topic, subscriptions
// are recorded so that the driver (or spinOnce, from the delivery list) can hand messages to the node's own callbacks,
// spin() returns at once, ok() is true while deliveries are pending
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include <ros/time.h>
#include <sensor_msgs/PointCloud2.h>
#include <nav_msgs/Odometry.h>
#include <sensor_msgs/Image.h>
#define ROSCONSOLE_DEFAULT_NAME "ros"
namespace refstub {
typedef std::function<void(const sensor_msgs::PointCloud2ConstPtr&)> CloudCb;
typedef std::function<void(const nav_msgs::Odometry::ConstPtr&)> OdomCb;
struct Delivery { std::vector<std::pair<std::string, sensor_msgs::PointCloud2ConstPtr>> clouds; std::vector<std::pair<std::string, nav_msgs::Odometry::ConstPtr>> odoms; };
struct State {
  std::map<std::string, double> params;
  std::map<std::string, std::vector<sensor_msgs::PointCloud2>> clouds;     // published, per topic, in order
  std::map<std::string, std::vector<nav_msgs::Odometry>> odoms;
  std::map<std::string, CloudCb> cloud_subs; std::map<std::string, OdomCb> odom_subs;
  std::vector<Delivery> deliveries; std::size_t next = 0;
};
inline State& state() { static State s; return s; }
inline void deliver(const Delivery& d) {
  for (auto& c : d.clouds) { auto it = state().cloud_subs.find(c.first); if (it == state().cloud_subs.end()) std::abort(); it->second(c.second); }
  for (auto& o : d.odoms) { auto it = state().odom_subs.find(o.first); if (it == state().odom_subs.end()) std::abort(); it->second(o.second); }
}
}
namespace ros {
struct Publisher { std::string topic;
  void publish(const sensor_msgs::PointCloud2& m) const { refstub::state().clouds[topic].push_back(m); }
  void publish(const nav_msgs::Odometry& m) const { refstub::state().odoms[topic].push_back(m); }
  template <class M> void publish(const M&) const {} };
struct Subscriber {};
struct NodeHandle {
  NodeHandle() {} explicit NodeHandle(const std::string&) {}
  bool getParam(const std::string&, std::string& v) const { v.clear(); return false; }
  template <class T> bool param(const std::string& k, T& v, const T& dflt) const {
    auto& p = refstub::state().params; auto it = p.find(k); v = it == p.end() ? dflt : (T)it->second; return it != p.end(); }
  template <class M> Publisher advertise(const std::string& topic, unsigned) { Publisher p; p.topic = topic; return p; }
  template <class M> Subscriber subscribe(const std::string& topic, unsigned, void (*cb)(const boost::shared_ptr<M const>&)) { reg(topic, cb); return Subscriber(); }
 private:
  static void reg(const std::string& topic, void (*cb)(const sensor_msgs::PointCloud2ConstPtr&)) { refstub::state().cloud_subs[topic] = cb; }
  static void reg(const std::string& topic, void (*cb)(const nav_msgs::Odometry::ConstPtr&)) { refstub::state().odom_subs[topic] = cb; }
  static void reg(const std::string&, void (*)(const sensor_msgs::ImageConstPtr&)) {}
};
inline void init(int&, char**, const std::string&) {}
namespace console { namespace levels { enum Level { Debug, Info, Warn, Error }; } inline bool set_logger_level(const char*, levels::Level) { return true; } }
inline bool ok() { return refstub::state().next < refstub::state().deliveries.size(); }
inline void spinOnce() { auto& s = refstub::state(); if (s.next < s.deliveries.size()) refstub::deliver(s.deliveries[s.next++]); }
inline void spin() {}
}
#define ROS_INFO(...) ((void)0)
#define ROS_WARN(...) ((void)0)
#define ROS_ERROR(...) ((void)0)
#define ROS_INFO_STREAM(x) ((void)0)
#define ROS_BREAK() std::abort()
