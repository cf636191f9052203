#pragma once
// declaration-level stand-ins: the debug viewer of the colour mapper is never started by the tests
#include <memory>
#include <mutex>
#include <string>
#include <pcl/point_cloud.h>
namespace boost {
struct mutex { std::mutex m; struct scoped_lock { explicit scoped_lock(mutex& mm) : l(mm.m) {} std::lock_guard<std::mutex> l; }; };
template <class T, class... A> shared_ptr<T> make_shared(A&&... a) { return std::make_shared<T>(std::forward<A>(a)...); }
namespace posix_time { struct microseconds { explicit microseconds(long) {} }; }
namespace this_thread { template <class D> void sleep(const D&) {} }
}
namespace pcl { namespace visualization {
struct PCLVisualizer {
  explicit PCLVisualizer(const std::string&) {}
  template <class C> bool addPointCloud(const C&, const std::string&) { return true; }
  template <class C> bool updatePointCloud(const C&, const std::string&) { return true; }
  void createViewPort(double, double, double, double, int&) {}
  void setBackgroundColor(double, double, double, int = 0) {}
  void addCoordinateSystem(double, int = 0) {}
  void addText(const std::string&, int, int, const std::string&, int = 0) {}
  void setCameraPosition(double, double, double, double, double, double, int = 0) {}
  bool wasStopped() const { return true; }
  void spinOnce(int = 1) {}
}; } }
