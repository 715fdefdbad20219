// process_sequence.cpp -- the reference's examples/process_sequence.cpp workflow (a directory of depth PNGs processed
// in order, min / max / mean time per frame and FPS) on the B200 library, in two modes:
//   latency   one frame per call, like the reference loop (decode + back-projection + process per frame)
//   batch     frames decoded up front, then processed B at a time from raw depth (toPointCloud runs on the device)
//   process_sequence <dir-with-png> <intrinsics.K> [config.ini] [batch]
#include <dirent.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include <deplex/deplex.h>

namespace {
std::vector<std::string> sorted_pngs(const std::string& dir) {
  std::vector<std::string> out;
  if (DIR* d = opendir(dir.c_str())) {
    while (dirent* e = readdir(d)) {
      const std::string name = e->d_name;
      if (name.size() > 4 && name.substr(name.size() - 4) == ".png") out.push_back(dir + "/" + name);
    }
    closedir(d);
  }
  std::sort(out.begin(), out.end());
  return out;
}
double us_since(std::chrono::high_resolution_clock::time_point t0) {
  return std::chrono::duration<double, std::micro>(std::chrono::high_resolution_clock::now() - t0).count();
}
}  // namespace

int main(int argc, char** argv) {
  if (argc < 3) {
    std::cerr << "usage: " << argv[0] << " <dir-with-png> <intrinsics.K> [config.ini] [batch]\n";
    return 2;
  }
  try {
    const std::vector<std::string> files = sorted_pngs(argv[1]);
    if (files.empty()) throw std::runtime_error(std::string("no .png files in ") + argv[1]);
    const deplex::utils::Intrinsics k = deplex::utils::readIntrinsics(argv[2]);
    const deplex::config::Config config = argc > 3 ? deplex::config::Config(std::string(argv[3])) : deplex::config::Config();
    const int batch = argc > 4 ? std::max(1, std::atoi(argv[4])) : 32;

    // ---- latency mode: the reference's loop, one extractor reused (the reference re-creates it per frame) ----
    deplex::utils::DepthImage image(files[0]);
    const int h = image.getHeight(), w = image.getWidth();
    const int64_t n = static_cast<int64_t>(h) * w;
    deplex::PlaneExtractor algorithm(h, w, config, batch);
    std::vector<double> us;
    std::vector<int32_t> labels;
    for (const std::string& f : files) {
      const auto t0 = std::chrono::high_resolution_clock::now();
      image.reset(f);
      const std::vector<float> points = image.toPointCloudRowMajor(k);
      labels = algorithm.process(points.data(), n, deplex::PointLayout::RowMajor);
      us.push_back(us_since(t0));
    }
    const double mean = [&] { double s = 0; for (double v : us) s += v; return s / us.size(); }();
    std::cout << "latency mode (decode + back-projection + process): " << files.size() << " frames, min "
              << *std::min_element(us.begin(), us.end()) << " us, max " << *std::max_element(us.begin(), us.end())
              << " us, mean " << mean << " us, FPS " << 1e6 / mean << "\n";

    // ---- batch mode: raw depth in, points generated on the device ----
    std::vector<uint16_t> depth(static_cast<size_t>(n) * files.size());
    for (size_t i = 0; i < files.size(); ++i) {
      image.reset(files[i]);
      std::copy(image.data(), image.data() + n, depth.begin() + i * n);
    }
    std::vector<int32_t> all(static_cast<size_t>(n) * files.size());
    const auto t0 = std::chrono::high_resolution_clock::now();
    for (size_t f0 = 0; f0 < files.size(); f0 += batch) {
      const int nf = static_cast<int>(std::min<size_t>(batch, files.size() - f0));
      algorithm.processDepthBatch(depth.data() + f0 * n, nf, k[0], k[4], k[2], k[5], all.data() + f0 * n);
    }
    const double total = us_since(t0);
    std::cout << "batch mode (" << batch << " frames per call, raw depth in): " << total / files.size() << " us per frame, FPS "
              << 1e6 * files.size() / total << "\n";
    // both modes label the last frame identically
    const bool same = std::equal(labels.begin(), labels.end(), all.end() - n);
    std::cout << "last frame: " << *std::max_element(labels.begin(), labels.end()) << " planes, modes agree: " << (same ? "yes" : "NO") << "\n";
    return same ? 0 : 1;
  } catch (const std::exception& e) {
    std::cerr << e.what() << "\n";
    return 1;
  }
}
