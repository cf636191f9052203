#pragma once
#include <cstdio>
#include <functional>
#include <iostream>
#include <memory>
#include <string>
#include <vector>
namespace boost { template <class T> using shared_ptr = std::shared_ptr<T>; template <class T, class U> shared_ptr<T> dynamic_pointer_cast(const shared_ptr<U>& p) { return std::dynamic_pointer_cast<T>(p); } }
namespace ros {
struct Time { Time() {} explicit Time(double) {} double toSec() const { return 0; } unsigned long long toNSec() const { return 0; } Time& fromSec(double) { return *this; } static Time now() { return Time(); } };
struct Rate { explicit Rate(double) {} bool sleep() { return true; } };
struct Publisher { template <class M> void publish(const M&) const {} };
struct Subscriber {};
struct NodeHandle {
  NodeHandle() {} explicit NodeHandle(const std::string&) {}
  template <class T> bool param(const std::string&, T&, const T&) const { return true; }
  template <class T> bool getParam(const std::string&, T&) const { return true; }
  template <class M> Publisher advertise(const std::string&, unsigned) { return Publisher(); }
  template <class M> Subscriber subscribe(const std::string&, unsigned, std::function<void(const boost::shared_ptr<M const>&)>) { return Subscriber(); }
  template <class M> Subscriber subscribe(const std::string&, unsigned, void (*)(const boost::shared_ptr<M const>&)) { return Subscriber(); }
};
inline void init(int&, char**, const std::string&) {}
inline bool ok() { return true; }
inline void spin() {}
inline void spinOnce() {}
}  // namespace ros
#define ROS_INFO(...) std::printf(__VA_ARGS__)
#define ROS_WARN(...) std::printf(__VA_ARGS__)
#define ROS_ERROR(...) std::printf(__VA_ARGS__)
#define ROS_INFO_STREAM(x) (std::cout << x)
#define ROS_BREAK() std::abort()
