#pragma once
// PCL's layouts (point_types.hpp): 16-byte xyz block with data[3] = 1, then the 16-byte block holding intensity
#include <cstdint>
namespace pcl {
struct alignas(16) PointXYZ { float x = 0, y = 0, z = 0, data3 = 1.0f; };
struct alignas(16) PointXYZI { float x = 0, y = 0, z = 0, data3 = 1.0f; float intensity = 0, data_c1 = 0, data_c2 = 0, data_c3 = 0; };
struct alignas(16) PointXYZRGB { float x = 0, y = 0, z = 0, data3 = 1.0f; std::uint8_t b = 0, g = 0, r = 0, a = 255; float pad_[3] = { 0, 0, 0 }; };
static_assert(sizeof(PointXYZ) == 16 && sizeof(PointXYZI) == 32 && sizeof(PointXYZRGB) == 32, "PCL layouts");
}
