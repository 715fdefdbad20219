// deplex/plane_extractor.h -- deplex::PlaneExtractor running on one B200 through libdeplex_b200.so.
//
// Drop-in for the reference class (cpp/deplex/include/deplex/plane_extractor.h:28-56): same constructor,
// same `process`, move-only PIMPL with an out-of-line destructor.  `Impl` owns nothing but a C-ABI handle
// (include/deplex_b200.h); all arithmetic runs in CUDA kernels.  The Eigen signature is compiled when
// Eigen is on the include path (it is what the reference's callers use); the raw-pointer overloads are
// always available and are what the Eigen overload forwards to.
#pragma once

#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>

#if !defined(DEPLEX_NO_EIGEN) && defined(__has_include)
#if __has_include(<Eigen/Core>)
#include <Eigen/Core>
#define DEPLEX_HAS_EIGEN 1
#endif
#endif

#include "deplex/config.h"

namespace deplex {

/** Memory order of an [N x 3] point matrix handed over as a raw pointer. */
enum class PointLayout : int32_t {
  ColMajor = 0, /**< X[N] Y[N] Z[N] -- Eigen::MatrixX3f's default order */
  RowMajor = 1  /**< x0 y0 z0 x1 ... -- numpy C order, DepthImage::toPointCloud's order */
};

/** One extracted plane (the reference computes these and discards them, plane_extractor.cpp:394-426). */
struct PlaneParams {
  float normal[3];
  float d;
  float mean[3];
  float mse, score;
  int32_t n_points;
  int32_t merge_label;
};

class PlaneExtractor {
 public:
  /**
   * @param image_height Image height in pixels.
   * @param image_width Image width in pixels.
   * @param config Parameters of plane extraction algorithm.
   * Throws std::runtime_error with the reference's text when patchSize is 0 (plane_extractor.cpp:161-164).
   */
  PlaneExtractor(int32_t image_height, int32_t image_width, config::Config config = config::Config());
  /** Same, with device scratch for `max_batch` frames per call on CUDA device `device` (-1 = current). */
  PlaneExtractor(int32_t image_height, int32_t image_width, config::Config config, int32_t max_batch, int32_t device = -1);
  ~PlaneExtractor();

#ifdef DEPLEX_HAS_EIGEN
  /**
   * Extract planes from one ORGANIZED point cloud [N x 3]; returns one label per point, 0 = non-planar.
   * Throws std::runtime_error when N != height * width (plane_extractor.cpp:188-194).
   */
  Eigen::VectorXi process(Eigen::MatrixX3f const& pcd_array) {
    Eigen::VectorXi labels(pcd_array.rows());
    process(pcd_array.data(), static_cast<int64_t>(pcd_array.rows()), PointLayout::ColMajor, labels.data());
    return labels;
  }
#endif

  /** Raw-pointer form of process(): `points` holds n_points x 3 floats in host memory. */
  std::vector<int32_t> process(float const* points, int64_t n_points, PointLayout layout);
  void process(float const* points, int64_t n_points, PointLayout layout, int32_t* labels);

  /** n_frames independent frames back to back in host memory (copies and kernels are pipelined). */
  void processBatch(float const* points, int32_t n_frames, PointLayout layout, int32_t* labels);
  /** n_frames <= max_batch frames resident in device memory; asynchronous on `cuda_stream` (a cudaStream_t). */
  void processBatchDevice(float const* d_points, int32_t n_frames, PointLayout layout, int32_t* d_labels,
                          void* cuda_stream = nullptr);

  /** Raw uint16 depth frames + pinhole intrinsics instead of points: DepthImage::toPointCloud
   *  (depth_image.cpp:55-78) is evaluated on the device, fused into the first kernel where the geometry allows. */
  void processDepthBatch(uint16_t const* depth, int32_t n_frames, float fx, float fy, float cx, float cy, int32_t* labels);

  /** Planes of frame `frame` of the last call (post-merge statistics). */
  std::vector<PlaneParams> planes(int32_t frame = 0);

  int32_t imageHeight() const;
  int32_t imageWidth() const;
  /** The C-ABI handle (dpx_extractor*), for callers that mix both layers. */
  void* handle() const;

  PlaneExtractor(PlaneExtractor&& op) noexcept;
  PlaneExtractor& operator=(PlaneExtractor&& op) noexcept;

 private:
  class Impl;
  std::unique_ptr<Impl> impl_;
};
}  // namespace deplex
