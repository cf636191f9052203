#pragma once
#include <cstddef>
#include <string>
#define CV_8UC1 0
#define CV_8UC3 16
namespace cv {
struct Size { int width = 0, height = 0; };
struct Point2f { float x = 0, y = 0; Point2f() {} Point2f(float a, float b) : x(a), y(b) {} };
struct Scalar { Scalar(double = 0, double = 0, double = 0, double = 0) {} };
struct Mat { int rows = 0, cols = 0; unsigned char* data = nullptr; std::size_t step = 0; Mat() {} Mat(int, int, int, void*) {} Mat clone() const { return *this; } int channels() const { return 3; } Size size() const { return Size(); } };
enum { COLOR_BGR2HSV = 40, COLOR_HSV2BGR = 54, COLOR_GRAY2BGR = 8, COLORMAP_JET = 2 };
inline void cvtColor(const Mat&, Mat&, int) {}
inline void circle(Mat&, Point2f, int, const Scalar&, int) {}
inline void applyColorMap(const Mat&, Mat&, int) {}
struct FileNode { template <class T> void operator>>(T&) const {} };
struct FileStorage { enum { READ = 0 }; FileStorage(const std::string&, int) {} bool isOpened() const { return true; } FileNode operator[](const char*) const { return FileNode(); } };
}
