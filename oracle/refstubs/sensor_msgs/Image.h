#pragma once
#include <memory>
#include <string>
#include <vector>
#include <std_msgs/Header.h>
namespace sensor_msgs {
struct Image { std_msgs::Header header; unsigned height = 0, width = 0; std::string encoding; unsigned char is_bigendian = 0; unsigned step = 0; std::vector<unsigned char> data; };
typedef std::shared_ptr<Image const> ImageConstPtr; typedef std::shared_ptr<Image> ImagePtr;
}
