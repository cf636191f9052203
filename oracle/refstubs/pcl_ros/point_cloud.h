#pragma once
#include <pcl/point_cloud.h>
#include <pcl_conversions/pcl_conversions.h>
