#pragma once
#include <ros/ros.h>
namespace camodocal { struct Camera { virtual ~Camera() {} }; typedef boost::shared_ptr<Camera> CameraPtr;
struct CameraFactory { static boost::shared_ptr<CameraFactory> instance() { return boost::shared_ptr<CameraFactory>(new CameraFactory); } CameraPtr generateCameraFromYamlFile(const std::string&) { return CameraPtr(); } }; }
