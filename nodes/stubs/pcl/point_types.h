#pragma once
#include <cstdint>
namespace pcl {
struct alignas(16) PointXYZ { float x, y, z, _pad; };                                 // 16 bytes
struct alignas(16) PointXYZI { float x, y, z, _pad0; float intensity; float _pad1[3]; };   // 32 bytes, intensity at 16 (common.h:43)
struct alignas(16) PointXYZRGB { float x, y, z, _pad0; union { struct { std::uint8_t b, g, r, a; }; float rgb; }; float _pad1[3]; };
static_assert(sizeof(PointXYZ) == 16 && sizeof(PointXYZI) == 32 && sizeof(PointXYZRGB) == 32, "PCL layouts");
}
