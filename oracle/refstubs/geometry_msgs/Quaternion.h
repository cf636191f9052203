#pragma once
#include <geometry_msgs/PoseStamped.h>
