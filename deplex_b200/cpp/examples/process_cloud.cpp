// process_cloud.cpp -- the reference's examples/process_cloud.cpp workflow on the B200 library:
// read a depth PNG + intrinsics + ini, run process() NUMBER_OF_RUNS times, print plane count and timing.
//   process_cloud <depth.png> <intrinsics.K> [config.ini] [runs]
#include <chrono>
#include <cstdlib>
#include <iostream>

#include <deplex/deplex.h>

int main(int argc, char** argv) {
  if (argc < 3) {
    std::cerr << "usage: " << argv[0] << " <depth.png> <intrinsics.K> [config.ini] [runs]\n";
    return 2;
  }
  try {
    deplex::utils::DepthImage image(argv[1]);
    const deplex::utils::Intrinsics k = deplex::utils::readIntrinsics(argv[2]);
    const deplex::config::Config config = argc > 3 ? deplex::config::Config(std::string(argv[3])) : deplex::config::Config();
    const int runs = argc > 4 ? std::atoi(argv[4]) : 10;

    const std::vector<float> points = image.toPointCloudRowMajor(k);
    deplex::PlaneExtractor algorithm(image.getHeight(), image.getWidth(), config);
    const int64_t n = static_cast<int64_t>(image.getHeight()) * image.getWidth();

    std::vector<int32_t> labels = algorithm.process(points.data(), n, deplex::PointLayout::RowMajor);  // warm-up
    const auto t0 = std::chrono::high_resolution_clock::now();
    for (int i = 0; i < runs; ++i) labels = algorithm.process(points.data(), n, deplex::PointLayout::RowMajor);
    const auto t1 = std::chrono::high_resolution_clock::now();
    int32_t max_label = 0;
    for (int32_t l : labels) max_label = l > max_label ? l : max_label;
    const double us = std::chrono::duration<double, std::micro>(t1 - t0).count() / (runs > 0 ? runs : 1);
    std::cout << "Number of found planes: " << max_label << "\n";
    std::cout << "Elapsed time (mean over " << runs << " runs): " << us << " us, FPS: " << 1e6 / us << "\n";
    for (const deplex::PlaneParams& p : algorithm.planes())
      if (p.n_points > 0)
        std::cout << "  plane n=(" << p.normal[0] << ", " << p.normal[1] << ", " << p.normal[2] << ") d=" << p.d
                  << " points=" << p.n_points << " merged_into=" << p.merge_label << "\n";
  } catch (const std::exception& e) {
    std::cerr << e.what() << "\n";
    return 1;
  }
  return 0;
}
