#pragma once
// stand-in for camodocal's camera interface and its pinhole model (camera_models/src/camera_models/PinholeCamera.cc:
// constructor :292-295, liftProjective :450-510, spaceToPlane :520-542, distortion :646-662); the reference's own
// PinholeCamera.cc pulls in the whole calibration tool chain (OpenCV calib3d, FileStorage) and was not compiled.
// Library stand-in, not reference source.
#include <memory>
#include <string>
#include <eigen3/Eigen/Dense>
#include <pcl/point_cloud.h>     // boost::shared_ptr alias
namespace camodocal {
class Camera { public: virtual ~Camera() {}
  virtual void spaceToPlane(const Eigen::Vector3d& P, Eigen::Vector2d& p) const = 0;
  virtual void liftProjective(const Eigen::Vector2d& p, Eigen::Vector3d& P) const = 0; };
typedef boost::shared_ptr<Camera> CameraPtr;
class CameraFactory { public: static CameraFactory* instance() { static CameraFactory f; return &f; }   // declaration-level (main() only)
  CameraPtr generateCameraFromYamlFile(const std::string&) { return CameraPtr(); } };
class PinholeCamera : public Camera {
 public:
  PinholeCamera(double fx, double fy, double cx, double cy, double k1, double k2, double p1, double p2)
      : fx_(fx), fy_(fy), cx_(cx), cy_(cy), k1_(k1), k2_(k2), p1_(p1), p2_(p2) {
    nod_ = k1 == 0.0 && k2 == 0.0 && p1 == 0.0 && p2 == 0.0;
    ik11_ = 1.0 / fx; ik13_ = -cx / fx; ik22_ = 1.0 / fy; ik23_ = -cy / fy; }
  void distortion(double x, double y, double& dx, double& dy) const {
    const double mx2 = x * x, my2 = y * y, mxy = x * y, rho2 = mx2 + my2, rad = k1_ * rho2 + k2_ * rho2 * rho2;
    dx = x * rad + 2.0 * p1_ * mxy + p2_ * (rho2 + 2.0 * mx2);
    dy = y * rad + 2.0 * p2_ * mxy + p1_ * (rho2 + 2.0 * my2); }
  void spaceToPlane(const Eigen::Vector3d& P, Eigen::Vector2d& p) const override {
    double x = P(0) / P(2), y = P(1) / P(2);
    if (!nod_) { double dx, dy; distortion(x, y, dx, dy); x = x + dx; y = y + dy; }
    p = Eigen::Vector2d(fx_ * x + cx_, fy_ * y + cy_); }
  void liftProjective(const Eigen::Vector2d& p, Eigen::Vector3d& P) const override {
    const double mx_d = ik11_ * p(0) + ik13_, my_d = ik22_ * p(1) + ik23_;
    double mx_u = mx_d, my_u = my_d;
    if (!nod_) { double dx, dy; distortion(mx_d, my_d, dx, dy); mx_u = mx_d - dx; my_u = my_d - dy;
      for (int i = 1; i < 8; ++i) { distortion(mx_u, my_u, dx, dy); mx_u = mx_d - dx; my_u = my_d - dy; } }
    P = Eigen::Vector3d(mx_u, my_u, 1.0); }
 private:
  double fx_, fy_, cx_, cy_, k1_, k2_, p1_, p2_, ik11_, ik13_, ik22_, ik23_; bool nod_;
};
}
