#pragma once
#include <string>
#include <opencv2/opencv.hpp>
#include <std_msgs/Header.h>
namespace cv_bridge { struct CvImage { std_msgs::Header header; std::string encoding; cv::Mat image; }; }
