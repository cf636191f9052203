#pragma once
#include <opencv2/opencv.hpp>
#include <sensor_msgs/Image.h>
namespace cv_bridge { struct CvImage { std_msgs::Header header; std::string encoding; cv::Mat image; sensor_msgs::ImagePtr toImageMsg() const { return sensor_msgs::ImagePtr(); } };
typedef boost::shared_ptr<CvImage const> CvImageConstPtr;
inline CvImageConstPtr toCvCopy(const sensor_msgs::ImageConstPtr&, const std::string&) { return CvImageConstPtr(); }
inline CvImageConstPtr toCvCopy(const sensor_msgs::Image&, const std::string&) { return CvImageConstPtr(); } }
