// deplex/utils/depth_image.h -- deplex::utils::DepthImage: a 16-bit depth PNG and its back-projection.
//
// Mirrors cpp/deplex/include/deplex/utils/depth_image.h:24-56.  The reference decodes with the vendored
// stb_image; here a small PNG reader over zlib does the same job (non-interlaced PNG, 8 or 16 bits per
// sample, grey / grey+alpha / RGB / RGBA, converted to 16-bit grey like stbi_load_16(..., STBI_grey)).
// This is file I/O that feeds the hot path; it runs on the host.
#pragma once

#include <array>
#include <cstdint>
#include <string>
#include <vector>

#if !defined(DEPLEX_NO_EIGEN) && defined(__has_include)
#if __has_include(<Eigen/Core>)
#include <Eigen/Core>
#ifndef DEPLEX_HAS_EIGEN
#define DEPLEX_HAS_EIGEN 1
#endif
#endif
#endif

namespace deplex {
namespace utils {
/** Row-major 3x3 pinhole matrix [[fx, 0, cx], [0, fy, cy], [0, 0, 1]]. */
using Intrinsics = std::array<float, 9>;

class DepthImage {
 public:
  DepthImage();
  /** Throws std::runtime_error "Error: Couldn't read image <path>" (depth_image.cpp:33-35). */
  DepthImage(std::string const& image_path);

  int32_t getWidth() const;
  int32_t getHeight() const;

  /** Raw depth samples, row-major, height * width entries. */
  uint16_t const* data() const;

  /**
   * Organized cloud of height*width points, ROW-major [N x 3] (depth_image.cpp:55-78):
   * z = float(raw), x = (col - cx) * z / fx, y = (row - cy) * z / fy, each evaluated left to right in fp32.
   */
  std::vector<float> toPointCloudRowMajor(Intrinsics const& intrinsics) const;

#ifdef DEPLEX_HAS_EIGEN
  Eigen::MatrixX3f toPointCloud(Eigen::Matrix3f const& intrinsics) const {
    Intrinsics k;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) k[3 * i + j] = intrinsics(i, j);
    std::vector<float> rm = toPointCloudRowMajor(k);
    return Eigen::Map<Eigen::Matrix<float, Eigen::Dynamic, 3, Eigen::RowMajor>>(rm.data(), width_ * height_, 3);
  }
#endif

  void reset(std::string const& image_path);

 private:
  std::vector<uint16_t> image_;
  int32_t width_;
  int32_t height_;
};
}  // namespace utils
}  // namespace deplex
