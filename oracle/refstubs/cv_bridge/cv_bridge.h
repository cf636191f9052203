#pragma once
// stand-in for cv_bridge: an image message is height x width x 3 (bgr8) or x 1 (mono8) bytes, toCvCopy copies them into a Mat
#include <cstring>
#include <memory>
#include <string>
#include <opencv2/opencv.hpp>
#include <sensor_msgs/Image.h>
#include <std_msgs/Header.h>
namespace cv_bridge {
struct CvImage { std_msgs::Header header; std::string encoding; cv::Mat image; };
typedef std::shared_ptr<CvImage const> CvImageConstPtr; typedef std::shared_ptr<CvImage> CvImagePtr;
inline CvImagePtr toCvCopy(const sensor_msgs::Image& m, const std::string& enc) {
  CvImagePtr p(new CvImage()); p->header = m.header; p->encoding = enc;
  p->image = cv::Mat((int)m.height, (int)m.width, enc == "mono8" ? CV_8UC1 : CV_8UC3);
  if (!m.data.empty()) std::memcpy(p->image.data(), m.data.data(), m.data.size());
  return p; }
inline CvImagePtr toCvCopy(const sensor_msgs::ImageConstPtr& m, const std::string& enc) { return toCvCopy(*m, enc); }
}
