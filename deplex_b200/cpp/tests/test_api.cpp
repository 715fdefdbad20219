// test_api.cpp -- the reference's gtest cases for the extractor (cpp/tests/test_plane_extractor.cpp:27-88,
// cpp/tests/test_config.cpp:24-29) restated as a plain executable over the C++ drop-in class.
//   test_api <depth.png> <intrinsics.K> <missing-parameters.ini> [expected_max_label]
// Needs a CUDA device.  Exit code 0 = all cases passed.
#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <string>

#include <algorithm>

#include <deplex/deplex.h>
#include <deplex_b200.h>

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
  {  // BatchPipeline / SequenceExtractor (additions): the same labels as process(), frame by frame
    PlaneExtractor single(h, w);
    const auto want = single.process(points.data(), n, PointLayout::RowMajor);
    const int frames = 5;
    std::vector<float> clouds(points.size() * frames);
    for (int f = 0; f < frames; ++f) std::copy(points.begin(), points.end(), clouds.begin() + f * points.size());
    std::vector<int32_t> got(static_cast<size_t>(n) * frames, -1);
    deplex::SequenceExtractor seq(h, w, Config(), {}, /*max_batch=*/2);
    CHECK(seq.deviceCount() >= 1);
    seq.process(clouds.data(), frames, PointLayout::RowMajor, got.data());
    for (int f = 0; f < frames; ++f) CHECK(std::equal(want.begin(), want.end(), got.begin() + static_cast<size_t>(f) * n));
    int64_t b = -1, e = -1;
    seq.frameRange(frames, 0, &b, &e);
    CHECK(b == 0 && e >= 1 && e <= frames);
    // device-resident batches through a two-lane pipeline (device memory through the C-ABI helpers, no CUDA headers here)
    void *d_in = nullptr, *d_out = nullptr;
    CHECK(dpx_device_alloc(0, &d_in, clouds.size() * sizeof(float)) == DPX_OK);
    CHECK(dpx_device_alloc(0, &d_out, got.size() * sizeof(int32_t)) == DPX_OK);
    CHECK(dpx_memcpy_to_device(0, d_in, clouds.data(), clouds.size() * sizeof(float)) == DPX_OK);
    {
      deplex::BatchPipeline pipe(h, w, Config(), /*max_batch=*/2, /*lanes=*/2, /*device=*/0);
      CHECK(pipe.lanes() == 2);
      for (int f = 0; f < frames; f += 2) {
        const int nf = std::min(2, frames - f);
        pipe.submit(static_cast<const float*>(d_in) + static_cast<size_t>(f) * n * 3, nf, PointLayout::RowMajor,
                    static_cast<int32_t*>(d_out) + static_cast<size_t>(f) * n);
      }
      pipe.synchronize();
      CHECK(throws_runtime_error([&] { pipe.submit(static_cast<const float*>(d_in), 3, PointLayout::RowMajor, static_cast<int32_t*>(d_out)); }));
    }
    std::fill(got.begin(), got.end(), -1);
    CHECK(dpx_memcpy_to_host(0, got.data(), d_out, got.size() * sizeof(int32_t)) == DPX_OK);
    for (int f = 0; f < frames; ++f) CHECK(std::equal(want.begin(), want.end(), got.begin() + static_cast<size_t>(f) * n));
    dpx_device_free(0, d_in);
    dpx_device_free(0, d_out);
  }
  {  // ReadImage: invalid files throw
    CHECK(throws_runtime_error([&] { deplex::utils::DepthImage bad{std::string(argv[3])}; }));
  }
  if (g_failed == 0) std::cout << "test_api: all cases passed\n";
  return g_failed == 0 ? 0 : 1;
}
