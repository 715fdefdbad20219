#pragma once
#include "deplex/utils/depth_image.h"
#include "deplex/utils/eigen_io.h"
