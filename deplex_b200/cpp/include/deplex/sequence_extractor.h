// deplex/sequence_extractor.h -- batches in flight on one GPU, and a frame sequence sharded over all GPUs of a box.
//
// Additions next to the reference class: its caller processes a sequence with a loop of process() calls on one thread
// (examples/process_sequence.cpp:30-43).  Frames are independent (plane_extractor.cpp:281,428), so
//   deplex::BatchPipeline      keeps several device-resident batches in flight on one GPU (dpx_pipeline_*), and
//   deplex::SequenceExtractor  splits a host frame range into contiguous ranges, one worker thread per GPU
//                              (dpx_sequence_*), with no inter-GPU traffic.
// Both are thin owners of C-ABI handles (include/deplex_b200.h); labels are identical to process() frame by frame.
#pragma once

#include <cstdint>
#include <memory>
#include <vector>

#include "deplex/config.h"
#include "deplex/plane_extractor.h"

namespace deplex {

class BatchPipeline {
 public:
  /** `lanes` extractors on CUDA device `device` (-1 = current), each sized for `max_batch` frames per submit. */
  BatchPipeline(int32_t image_height, int32_t image_width, config::Config config, int32_t max_batch, int32_t lanes = 3,
                int32_t device = -1);
  ~BatchPipeline();
  BatchPipeline(BatchPipeline&&) noexcept;
  BatchPipeline& operator=(BatchPipeline&&) noexcept;

  /** Queue n_frames <= max_batch device-resident frames; asynchronous, ordered behind `producer_stream` (cudaStream_t). */
  void submit(float const* d_points, int32_t n_frames, PointLayout layout, int32_t* d_labels, void* producer_stream = nullptr);
  void submitDepth(uint16_t const* d_depth, int32_t n_frames, float fx, float fy, float cx, float cy, int32_t* d_labels,
                   void* producer_stream = nullptr);
  /** Order `consumer_stream` behind everything submitted so far (device-side). */
  void join(void* consumer_stream = nullptr);
  /** join + wait on the host. */
  void synchronize();
  int32_t lanes() const;
  /** The C-ABI handle (dpx_pipeline*). */
  void* handle() const;

 private:
  class Impl;
  std::unique_ptr<Impl> impl_;
};

class SequenceExtractor {
 public:
  /** One extractor per device; `devices` empty = every visible CUDA device. */
  SequenceExtractor(int32_t image_height, int32_t image_width, config::Config config = config::Config(),
                    std::vector<int32_t> const& devices = {}, int32_t max_batch = 64);
  ~SequenceExtractor();
  SequenceExtractor(SequenceExtractor&&) noexcept;
  SequenceExtractor& operator=(SequenceExtractor&&) noexcept;

  /** n_frames organized clouds back to back in host memory -> n_frames * height * width labels. */
  void process(float const* points, int64_t n_frames, PointLayout layout, int32_t* labels);
  /** The same from raw uint16 depth frames (DepthImage::toPointCloud runs on the devices). */
  void processDepth(uint16_t const* depth, int64_t n_frames, float fx, float fy, float cx, float cy, int32_t* labels);
  /** Every device processes n_frames_per_device frames resident in ITS memory through a `lanes`-lane pipeline;
   *  returns the CUDA-event milliseconds each device took. */
  std::vector<float> processDevice(std::vector<float const*> const& d_points, int64_t n_frames_per_device, PointLayout layout,
                                   std::vector<int32_t*> const& d_labels, int32_t lanes = 3);

  int32_t deviceCount() const;
  /** [begin, end) of device slot `slot` for a sequence of n_frames (contiguous, sizes differ by at most one). */
  void frameRange(int64_t n_frames, int32_t slot, int64_t* begin, int64_t* end) const;
  void* handle() const;

 private:
  class Impl;
  std::unique_ptr<Impl> impl_;
};

}  // namespace deplex
