#pragma once
#include "camodocal/camera_models/CameraFactory.h"
namespace camodocal { struct PinholeCamera : Camera { struct Parameters { double fx() const { return 0; } double fy() const { return 0; } double cx() const { return 0; } double cy() const { return 0; }
  double k1() const { return 0; } double k2() const { return 0; } double p1() const { return 0; } double p2() const { return 0; } int imageWidth() const { return 0; } int imageHeight() const { return 0; } };
  const Parameters& getParameters() const { return mParameters; } Parameters mParameters; }; }
