#pragma once
// stand-in for the OpenCV API the reference's colour mapper uses: an 8-bit Mat with shared storage, and the imgproc
// calls of MapBuilder::depthFill mapped one to one onto the oracle's restatements of OpenCV 3.2 (oracle/color.c:
// getStructuringElement, dilate / erode with the default border, medianBlur 5, bilateralFilter 5, GaussianBlur 5x5).
// The colour-space conversions, cv::circle and applyColorMap only feed the two debug images: they keep sizes and types
// and record their input (the first dilate of a frame sees the raw depth raster, applyColorMap the filled one).
// Library stand-in, not reference source.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include "lmono_oracle.h"
#define CV_8UC1 0
#define CV_8UC3 16
typedef unsigned char uchar;                     /* OpenCV's cvdef.h declares it globally */
namespace cv {
using ::uchar;
struct Size { int width, height; Size(int w = 0, int h = 0) : width(w), height(h) {} };
struct Point2f { float x, y; Point2f(float x_ = 0, float y_ = 0) : x(x_), y(y_) {} };
struct Scalar { double v[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { v[0] = a; v[1] = b; v[2] = c; v[3] = d; } };
struct Vec3b { uchar v[3]; uchar& operator[](int i) { return v[i]; } const uchar& operator[](int i) const { return v[i]; } };
class Mat {
 public:
  int rows = 0, cols = 0, type_ = CV_8UC1;
  std::shared_ptr<std::vector<uchar>> buf;                           // header copies share the pixels, clone() does not
  Mat() {}
  Mat(int r, int c, int type) : rows(r), cols(c), type_(type), buf(std::make_shared<std::vector<uchar>>((std::size_t)r * c * (type == CV_8UC3 ? 3 : 1), 0)) {}
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
// declaration-level: the node reads its yaml configuration through these in main(), which the tests do not run
struct FileNode { template <class T> void operator>>(T&) const {} };
struct FileStorage { enum { READ = 0 }; FileStorage(const std::string&, int) {} bool isOpened() const { return false; } FileNode operator[](const char*) const { return FileNode(); } };
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
