#pragma once
// stand-in for the OpenCV API the reference's colour mapper uses: an 8-bit Mat with shared storage, and the imgproc
// calls of MapBuilder::depthFill mapped one to one onto the oracle's restatements of OpenCV 3.2 (oracle/color.c:
// getStructuringElement, dilate / erode with the default border, medianBlur 5, bilateralFilter 5, GaussianBlur 5x5).
// The colour-space conversions, cv::circle and applyColorMap only feed the two debug images: they keep sizes and types
// and record their input (the first dilate of a frame sees the raw depth raster, applyColorMap the filled one).
// Library stand-in, not reference source.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include "lmono_oracle.h"
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_32F 5
#define CV_64F 6
#define CV_32FC1 5
#define CV_32FC2 13
typedef unsigned char uchar;                     /* OpenCV's cvdef.h declares it globally */
namespace cv {
using ::uchar;
struct Size { int width, height; Size(int w = 0, int h = 0) : width(w), height(h) {} bool operator==(const Size& o) const { return width == o.width && height == o.height; } };
template <class T> struct Point_ { T x, y; Point_(T x_ = 0, T y_ = 0) : x(x_), y(y_) {} template <class U> Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
  Point_ operator-(const Point_& o) const { return Point_(x - o.x, y - o.y); } Point_ operator+(const Point_& o) const { return Point_(x + o.x, y + o.y); } };
typedef Point_<float> Point2f; typedef Point_<double> Point2d; typedef Point_<int> Point2i;
template <class T> struct Point3_ { T x, y, z; Point3_(T x_ = 0, T y_ = 0, T z_ = 0) : x(x_), y(y_), z(z_) {} };
typedef Point3_<float> Point3f; typedef Point3_<double> Point3d;
template <class T> inline double norm(const Point_<T>& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y); }
struct Scalar { double v[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { v[0] = a; v[1] = b; v[2] = c; v[3] = d; } };
struct Vec3b { uchar v[3]; uchar& operator[](int i) { return v[i]; } const uchar& operator[](int i) const { return v[i]; } };
class Mat {
 public:
  int rows = 0, cols = 0, type_ = CV_8UC1;
  std::shared_ptr<std::vector<uchar>> buf;                           // header copies share the pixels, clone() does not
  Mat() {}
  Mat(int r, int c, int type) : rows(r), cols(c), type_(type), buf(std::make_shared<std::vector<uchar>>((std::size_t)r * c * (type == CV_8UC3 ? 3 : 1), 0)) {}
  Mat(Size s, int type) : Mat(s.height, s.width, type) {}
  static Mat eye(int r, int c, int type) { return Mat(r, c, type); }          // declaration-level (calibration code only)
  template <class T> T& at(int i) { return *reinterpret_cast<T*>(buf->data() + (std::size_t)i * sizeof(T)); }
  template <class T> const T& at(int i) const { return *reinterpret_cast<const T*>(buf->data() + (std::size_t)i * sizeof(T)); }
  static Mat zeros(Size s, int type) { return Mat(s.height, s.width, type); }
  static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
  int channels() const { return type_ == CV_8UC3 ? 3 : 1; }
  Size size() const { return Size(cols, rows); }
  bool empty() const { return !buf || buf->empty(); }
  Mat clone() const { Mat m; m.rows = rows; m.cols = cols; m.type_ = type_; if (buf) m.buf = std::make_shared<std::vector<uchar>>(*buf); return m; }
  uchar* data() { return buf->data(); } const uchar* data() const { return buf->data(); }
  template <class T> T& at(int r, int c) { return *reinterpret_cast<T*>(buf->data() + ((std::size_t)r * cols + c) * sizeof(T)); }
  template <class T> const T& at(int r, int c) const { return *reinterpret_cast<const T*>(buf->data() + ((std::size_t)r * cols + c) * sizeof(T)); }
};
enum { COLOR_BGR2HSV = 40, COLOR_HSV2BGR = 54, COLORMAP_JET = 2 };
// declaration-level: yaml configuration and the calibration tool chain (camodocal's estimateIntrinsics / undistortion maps,
// the nodes' main()), none of which the tests run
struct FileNode { template <class T> void operator>>(T&) const {} FileNode operator[](const char*) const { return FileNode(); } bool isNone() const { return true; }
  operator int() const { return 0; } operator double() const { return 0.0; } operator std::string() const { return std::string(); } };
struct FileStorage { enum { READ = 0, WRITE = 1 }; FileStorage(const std::string&, int) {} bool isOpened() const { return false; } void release() {}
  FileNode operator[](const char*) const { return FileNode(); } FileNode operator[](const std::string&) const { return FileNode(); } };
template <class T> inline FileStorage& operator<<(FileStorage& f, const T&) { return f; }
struct _OutputArray { bool needed() const { return false; } void create(int, int, int) const {} Mat getMat() const { return Mat(); } };
typedef const _OutputArray& OutputArray; typedef const _OutputArray& InputArray;
inline _OutputArray noArray() { return _OutputArray(); }
enum { DECOMP_LU = 0, DECOMP_NORMAL = 16 };
template <class A, class B, class C, class D> inline bool solvePnP(const A&, const B&, const C&, const D&, Mat&, Mat&) { return false; }
inline bool solve(const Mat&, const Mat&, Mat&, int) { return false; }
template <class A, class B> inline Mat findHomography(const A&, const B&) { return Mat(3, 3, CV_64F); }
inline void Rodrigues(const Mat&, Mat&) {}
inline void convertMaps(const Mat&, const Mat&, Mat&, Mat&, int, bool) {}
template <class E> inline void cv2eigen(const Mat&, E&) {}
template <class E> inline void eigen2cv(const E&, Mat&) {}
enum { MORPH_RECT = 0, MORPH_CROSS = 1, MORPH_ELLIPSE = 2 };
enum { MORPH_ERODE = 0, MORPH_DILATE = 1, MORPH_OPEN = 2, MORPH_CLOSE = 3 };
namespace refstub_cv { struct Log { std::vector<Mat> dilate_inputs; Mat colormap_input; }; inline Log& log() { static Log l; return l; } }
inline void cvtColor(const Mat& src, Mat& dst, int) { dst = src.clone(); }
inline void circle(Mat&, Point2f, int, const Scalar&, int) {}
inline void applyColorMap(const Mat& src, Mat& dst, int) { refstub_cv::log().colormap_input = src.clone(); dst = Mat(src.rows, src.cols, CV_8UC3); }
inline Mat getStructuringElement(int shape, Size s) {
  if (s.width != s.height) std::abort();
  Mat k(s.height, s.width, CV_8UC1); lmono_cpu_cv_kernel(shape, s.width, k.data()); return k; }
inline void dilate(const Mat& src, Mat& dst, const Mat& kernel) {
  refstub_cv::log().dilate_inputs.push_back(src.clone());
  Mat o(src.rows, src.cols, CV_8UC1); lmono_cpu_cv_morph(src.data(), o.data(), src.cols, src.rows, kernel.data(), kernel.cols, 0); dst = o; }
inline void erode(const Mat& src, Mat& dst, const Mat& kernel) {
  Mat o(src.rows, src.cols, CV_8UC1); lmono_cpu_cv_morph(src.data(), o.data(), src.cols, src.rows, kernel.data(), kernel.cols, 1); dst = o; }
inline void morphologyEx(const Mat& src, Mat& dst, int op, const Mat& kernel) {
  if (op != MORPH_CLOSE) std::abort();
  Mat d(src.rows, src.cols, CV_8UC1), e(src.rows, src.cols, CV_8UC1);
  lmono_cpu_cv_morph(src.data(), d.data(), src.cols, src.rows, kernel.data(), kernel.cols, 0);
  lmono_cpu_cv_morph(d.data(), e.data(), src.cols, src.rows, kernel.data(), kernel.cols, 1); dst = e; }
inline void medianBlur(const Mat& src, Mat& dst, int ksize) {
  if (ksize != 5) std::abort();
  Mat o(src.rows, src.cols, CV_8UC1); lmono_cpu_cv_median5(src.data(), o.data(), src.cols, src.rows); dst = o; }
inline void bilateralFilter(const Mat& src, Mat& dst, int d, double sigma_color, double sigma_space) {
  if (d != 5) std::abort();
  Mat o(src.rows, src.cols, CV_8UC1); lmono_cpu_cv_bilateral5(src.data(), o.data(), src.cols, src.rows, sigma_color, sigma_space); dst = o; }
inline void GaussianBlur(const Mat& src, Mat& dst, Size k, double sigma) {
  if (k.width != 5 || k.height != 5 || sigma != 0) std::abort();
  Mat o(src.rows, src.cols, CV_8UC1); lmono_cpu_cv_gaussian5(src.data(), o.data(), src.cols, src.rows); dst = o; }
}
