#pragma once
#include <string>
namespace sensor_msgs { namespace image_encodings { const std::string BGR8 = "bgr8"; const std::string MONO8 = "mono8"; } }
