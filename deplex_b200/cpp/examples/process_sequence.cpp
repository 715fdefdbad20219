// process_sequence.cpp -- the reference's examples/process_sequence.cpp workflow (a directory of depth PNGs processed
// in order, min / max / mean time per frame and FPS) on the B200 library, in two modes:
//   latency   one frame per call, like the reference loop (decode + back-projection + process per frame)
//   batch     frames decoded up front, then processed B at a time from raw depth (toPointCloud runs on the device)
//   process_sequence <dir-with-png> <intrinsics.K> [config.ini] [batch]
// and, for throughput on every GPU of the box with no Python anywhere (deplex::SequenceExtractor):
//   process_sequence --clouds <file.bin> <height> <width> [config.ini|-] [batch] [lanes] [batches-per-device]
// where file.bin holds organized clouds back to back (row-major float32 [N x 3] each).  The frames are tiled to
// `batch` per device; "host" = the whole sequence from host memory sharded over the GPUs (copies included),
// "device" = batches-per-device batches resident in each GPU's memory through a `lanes`-lane pipeline.
#include <dirent.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include <cstring>
#include <fstream>
#include <numeric>

#include <deplex/deplex.h>
#include <deplex_b200.h>

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

// --clouds mode: SequenceExtractor over every visible GPU, host-pointer and device-resident
int run_clouds(int argc, char** argv) {
  if (argc < 5) {
    std::cerr << "usage: " << argv[0] << " --clouds <file.bin> <height> <width> [config.ini|-] [batch] [lanes] [batches-per-device]\n";
    return 2;
  }
  const int h = std::atoi(argv[3]), w = std::atoi(argv[4]);
  const deplex::config::Config config =
      (argc > 5 && std::strcmp(argv[5], "-") != 0) ? deplex::config::Config(std::string(argv[5])) : deplex::config::Config();
  const int batch = argc > 6 ? std::max(1, std::atoi(argv[6])) : 256;
  const int lanes = argc > 7 ? std::max(1, std::atoi(argv[7])) : 3;
  const int per_dev = argc > 8 ? std::max(1, std::atoi(argv[8])) : 16;
  const size_t np = static_cast<size_t>(h) * w;
  std::ifstream in(argv[2], std::ios::binary | std::ios::ate);
  if (!in) throw std::runtime_error(std::string("cannot open ") + argv[2]);
  const size_t unique = static_cast<size_t>(in.tellg()) / (np * 12);
  if (unique == 0) throw std::runtime_error("file holds less than one frame");
  in.seekg(0);

  deplex::SequenceExtractor seq(h, w, config, {}, batch);
  const int G = seq.deviceCount();
  // host sequence: `batch` frames per device, pinned so that the copies run at PCIe speed
  const size_t n_host = static_cast<size_t>(batch) * G;
  float* clouds = nullptr;
  int32_t* labels = nullptr;
  if (dpx_host_alloc(reinterpret_cast<void**>(&clouds), n_host * np * 12) != DPX_OK ||
      dpx_host_alloc(reinterpret_cast<void**>(&labels), n_host * np * 4) != DPX_OK)
    throw std::runtime_error(dpx_last_error(nullptr));
  const size_t n_read = std::min(unique, n_host);
  in.read(reinterpret_cast<char*>(clouds), static_cast<std::streamsize>(n_read * np * 12));
  for (size_t f = n_read; f < n_host; ++f) std::memcpy(clouds + f * np * 3, clouds + (f % n_read) * np * 3, np * 12);

  // ---- host mode: H2D + kernels + D2H, sharded over the GPUs ----
  seq.process(clouds, static_cast<int64_t>(n_host), deplex::PointLayout::RowMajor, labels);  // warm-up (allocations)
  const int host_passes = 3;
  auto t0 = std::chrono::high_resolution_clock::now();
  for (int i = 0; i < host_passes; ++i) seq.process(clouds, static_cast<int64_t>(n_host), deplex::PointLayout::RowMajor, labels);
  const double host_us = us_since(t0);
  uint64_t checksum = 0;
  for (size_t f = 0; f < n_read; ++f)
    for (size_t i = 0; i < np; i += 997) checksum += static_cast<uint64_t>(labels[f * np + i]) * (1 + (i & 1023));
  std::cout << "host mode: " << G << " GPU(s), " << n_host << " frames per pass, "
            << 1e6 * host_passes * n_host / host_us << " frames/s (copies included)\n";

  // ---- device mode: per_dev batches resident on every GPU ----
  const int64_t n_dev = static_cast<int64_t>(batch) * per_dev;
  std::vector<float const*> d_points(G);
  std::vector<int32_t*> d_labels(G);
  for (int g = 0; g < G; ++g) {
    void *dp = nullptr, *dl = nullptr;
    if (dpx_device_alloc(g, &dp, static_cast<size_t>(n_dev) * np * 12) != DPX_OK ||
        dpx_device_alloc(g, &dl, static_cast<size_t>(n_dev) * np * 4) != DPX_OK)
      throw std::runtime_error(dpx_last_error(nullptr));
    for (int b = 0; b < per_dev; ++b)  // device g holds its own slice of the host sequence, repeated
      if (dpx_memcpy_to_device(g, static_cast<char*>(dp) + static_cast<size_t>(b) * batch * np * 12,
                               clouds + static_cast<size_t>(g) * batch * np * 3, static_cast<size_t>(batch) * np * 12) != DPX_OK)
        throw std::runtime_error(dpx_last_error(nullptr));
    d_points[g] = static_cast<float const*>(dp);
    d_labels[g] = static_cast<int32_t*>(dl);
  }
  seq.processDevice(d_points, n_dev, deplex::PointLayout::RowMajor, d_labels, lanes);  // warm-up
  double best_ms = 1e30;
  for (int rep = 0; rep < 5; ++rep) {
    const std::vector<float> ms = seq.processDevice(d_points, n_dev, deplex::PointLayout::RowMajor, d_labels, lanes);
    best_ms = std::min(best_ms, static_cast<double>(*std::max_element(ms.begin(), ms.end())));
  }
  // the device-resident labels of GPU 0's first batch must equal what the host mode produced for the same frames
  std::vector<int32_t> back(static_cast<size_t>(batch) * np);
  if (dpx_memcpy_to_host(0, back.data(), d_labels[0], back.size() * 4) != DPX_OK) throw std::runtime_error(dpx_last_error(nullptr));
  const bool same = std::equal(back.begin(), back.end(), labels);
  std::cout << "device mode: " << G << " GPU(s) x " << n_dev << " resident frames, " << lanes << " lanes, "
            << G * n_dev / (best_ms * 1e-3) << " frames/s (slowest GPU's CUDA-event time, best of 5)\n";
  std::cout << "labels checksum " << checksum << ", device == host labels: " << (same ? "yes" : "NO") << "\n";
  for (int g = 0; g < G; ++g) {
    dpx_device_free(g, const_cast<float*>(d_points[g]));
    dpx_device_free(g, d_labels[g]);
  }
  dpx_host_free(clouds);
  dpx_host_free(labels);
  return same ? 0 : 1;
}
}  // namespace

int main(int argc, char** argv) {
  if (argc > 1 && std::strcmp(argv[1], "--clouds") == 0) {
    try {
      return run_clouds(argc, argv);
    } catch (const std::exception& e) {
      std::cerr << e.what() << "\n";
      return 1;
    }
  }
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
