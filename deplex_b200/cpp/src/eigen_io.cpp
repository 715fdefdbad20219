// eigen_io.cpp -- text I/O helpers (reference: cpp/deplex/src/deplex/utils/eigen_io.cpp:22-60).
#include "deplex/utils/eigen_io.h"

#include <cstdlib>
#include <fstream>
#include <iterator>
#include <iomanip>
#include <limits>
#include <sstream>
#include <stdexcept>

namespace deplex {
namespace utils {

// One pass over the whole file with strtof: fields are separated by `delimiter`, records by newlines.  Same contract as
// the reference reader (eigen_io.cpp:24-40): every field is a float, the total must be a multiple of three.
std::vector<float> readPointCloudCSV(std::string const& path, char delimiter) {
  std::ifstream file(path, std::ios::binary);
  std::string text((std::istreambuf_iterator<char>(file)), std::istreambuf_iterator<char>());
  std::vector<float> points;
  points.reserve(text.size() / 6);
  const char* cur = text.c_str();
  const char* const end = cur + text.size();
  while (cur < end) {
    if (*cur == delimiter || *cur == '\n' || *cur == '\r' || *cur == ' ' || *cur == '\t') {
      ++cur;
      continue;
    }
    char* after = nullptr;
    const float value = std::strtof(cur, &after);
    if (after == cur) throw std::invalid_argument("stof");  // what std::stof throws on a non-numeric field
    points.push_back(value);
    cur = after;
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
