// deplex/utils/eigen_io.h -- text I/O helpers (cpp/deplex/include/deplex/utils/eigen_io.h:30-46), Eigen-free.
#pragma once

#include <string>
#include <vector>

#include "deplex/utils/depth_image.h"

namespace deplex {
namespace utils {
/** Points [N x 3], row-major.  Throws "Error reading file: Invalid points shape" (eigen_io.cpp:35-37). */
std::vector<float> readPointCloudCSV(std::string const& path, char delimiter = ',');
/** 3x3 matrix without delimiters.  Throws "Error: Couldn't open intrinsics file <path>" (eigen_io.cpp:43-45). */
Intrinsics readIntrinsics(std::string const& intrinsics_path);
/** Row-major [N x 3] points, "x, y, z" per line (eigen_io.h:22 CSVFormat). */
void savePointCloudCSV(std::vector<float> const& pcd_points, std::string const& path);
}  // namespace utils
}  // namespace deplex
