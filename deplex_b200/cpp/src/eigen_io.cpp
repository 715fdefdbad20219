// eigen_io.cpp -- text I/O helpers (reference: cpp/deplex/src/deplex/utils/eigen_io.cpp:22-60).
#include "deplex/utils/eigen_io.h"

#include <fstream>
#include <iomanip>
#include <limits>
#include <sstream>
#include <stdexcept>

namespace deplex {
namespace utils {

std::vector<float> readPointCloudCSV(std::string const& path, char delimiter) {
  std::vector<float> points;
  std::ifstream file(path);
  std::string row, entry;
  while (std::getline(file, row)) {
    std::stringstream ss(row);
    while (std::getline(ss, entry, delimiter)) points.push_back(std::stof(entry));
  }
  if (points.size() % 3 != 0) throw std::runtime_error("Error reading file: Invalid points shape");
  return points;
}

Intrinsics readIntrinsics(std::string const& intrinsics_path) {
  std::ifstream in(intrinsics_path);
  if (!in.is_open()) throw std::runtime_error("Error: Couldn't open intrinsics file " + intrinsics_path);
  Intrinsics k{};
  for (float& v : k) in >> v;
  return k;
}

void savePointCloudCSV(std::vector<float> const& pcd_points, std::string const& path) {
  std::ofstream file(path);
  file << std::setprecision(std::numeric_limits<float>::max_digits10);
  const size_t n = pcd_points.size() / 3;
  for (size_t i = 0; i < n; ++i) {
    file << pcd_points[3 * i] << ", " << pcd_points[3 * i + 1] << ", " << pcd_points[3 * i + 2];
    if (i + 1 < n) file << "\n";
  }
}

}  // namespace utils
}  // namespace deplex
