#pragma once
#include "camodocal/camera_models/CameraFactory.h"
