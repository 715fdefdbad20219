// test_api.cpp -- the reference's gtest cases for the extractor (cpp/tests/test_plane_extractor.cpp:27-88,
// cpp/tests/test_config.cpp:24-29) restated as a plain executable over the C++ drop-in class.
//   test_api <depth.png> <intrinsics.K> <missing-parameters.ini> [expected_max_label]
// Needs a CUDA device.  Exit code 0 = all cases passed.
#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <string>

#include <deplex/deplex.h>

namespace {
int g_failed = 0;
#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) {                                                         \
      std::cerr << "FAILED " << __LINE__ << ": " #cond << "\n";            \
      ++g_failed;                                                          \
    }                                                                      \
  } while (0)

template <class F>
bool throws_runtime_error(F&& f, std::string* what = nullptr) {
  try {
    f();
  } catch (const std::runtime_error& e) {
    if (what) *what = e.what();
    return true;
  }
  return false;
}

int32_t max_of(const std::vector<int32_t>& v) {
  int32_t m = 0;
  for (int32_t x : v) m = x > m ? x : m;
  return m;
}
}  // namespace

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  using deplex::PlaneExtractor;
  using deplex::PointLayout;
  using deplex::config::Config;
  const int expected = argc > 4 ? std::atoi(argv[4]) : 34;

  deplex::utils::DepthImage image(argv[1]);
  const auto k = deplex::utils::readIntrinsics(argv[2]);
  const std::vector<float> points = image.toPointCloudRowMajor(k);
  const int h = image.getHeight(), w = image.getWidth();
  const int64_t n = static_cast<int64_t>(h) * w;
  CHECK(h == 480 && w == 640);

  {  // TUMPlaneExtraction.DefaultConfigExtraction
    PlaneExtractor algorithm(h, w);
    const auto labels = algorithm.process(points.data(), n, PointLayout::RowMajor);
    CHECK(static_cast<int64_t>(labels.size()) == n);
    CHECK(max_of(labels) == expected);
    // column-major input (Eigen::MatrixX3f order) gives the same labels
    std::vector<float> cm(points.size());
    for (int64_t i = 0; i < n; ++i)
      for (int a = 0; a < 3; ++a) cm[a * n + i] = points[3 * i + a];
    CHECK(algorithm.process(cm.data(), n, PointLayout::ColMajor) == labels);
    // raw depth in (toPointCloud evaluated on the device) gives the same labels
    std::vector<int32_t> from_depth(static_cast<size_t>(n));
    algorithm.processDepthBatch(image.data(), 1, k[0], k[4], k[2], k[5], from_depth.data());
    CHECK(from_depth == labels);
    // move operations keep the extractor usable
    PlaneExtractor moved(std::move(algorithm));
    CHECK(moved.process(points.data(), n, PointLayout::RowMajor) == labels);
    CHECK(!moved.planes().empty());
  }
  {  // ZeroLeadingConfigExtraction
    Config config;
    config.min_region_planarity_score = 5000;
    PlaneExtractor algorithm(h, w, config);
    const auto labels = algorithm.process(points.data(), n, PointLayout::RowMajor);
    CHECK(static_cast<int64_t>(labels.size()) == n && max_of(labels) == 0);
  }
  {  // ZeroPatchSize
    Config config;
    config.patch_size = 0;
    std::string what;
    CHECK(throws_runtime_error([&] { PlaneExtractor algorithm(h, w, config); }, &what));
    CHECK(what == "Error! Invalid config parameter: patchSize(0). patchSize has to be positive.");
  }
  {  // EnormousPatchSize
    Config config;
    config.patch_size = 1000000;
    PlaneExtractor algorithm(h, w, config);
    const auto labels = algorithm.process(points.data(), n, PointLayout::RowMajor);
    CHECK(static_cast<int64_t>(labels.size()) == n && max_of(labels) == 0);
  }
  {  // InvalidInput.ZeroValuePoints / EmptyPoints / WrongShape
    PlaneExtractor algorithm(h, w);
    std::vector<float> zeros(points.size(), 0.f);
    CHECK(max_of(algorithm.process(zeros.data(), n, PointLayout::RowMajor)) == 0);
    std::string what;
    CHECK(throws_runtime_error([&] { algorithm.process(nullptr, 0, PointLayout::RowMajor); }, &what));
    CHECK(what == "Error! Number of points doesn't match image shape: 0 != 480 x 640");
    CHECK(throws_runtime_error([&] { algorithm.process(points.data(), n / 2, PointLayout::RowMajor); }));
  }
  {  // ConfigInit: bad path throws, unknown / commented keys keep the defaults
    std::string what;
    CHECK(throws_runtime_error([&] { Config c(std::string("/nonexistent/dir/config.ini")); }, &what));
    CHECK(what == "Couldn't open ini file: /nonexistent/dir/config.ini");
    Config c{std::string(argv[3])};
    Config d;
    CHECK(c.patch_size == 12 && c.histogram_bins_per_coord == d.histogram_bins_per_coord && c.max_merge_dist == d.max_merge_dist);
  }
  {  // ReadImage: invalid files throw
    CHECK(throws_runtime_error([&] { deplex::utils::DepthImage bad{std::string(argv[3])}; }));
  }
  if (g_failed == 0) std::cout << "test_api: all cases passed\n";
  return g_failed == 0 ? 0 : 1;
}
