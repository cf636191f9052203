#pragma once
#include <string>
#include <pcl/point_cloud.h>
namespace pcl { namespace io { template <class P> int savePLYFileBinary(const std::string&, const PointCloud<P>&) { return 0; } } }
