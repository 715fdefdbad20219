// sequence_extractor.cpp -- deplex::BatchPipeline and deplex::SequenceExtractor as thin owners of C-ABI handles
// (dpx_pipeline, dpx_sequence).  They replace the per-frame loop of examples/process_sequence.cpp:30-43; errors are
// std::runtime_error with the C-ABI's message (the reference's own texts for the reference's own errors).
#include "deplex/sequence_extractor.h"

#include <stdexcept>
#include <string>

#include "deplex_b200.h"

namespace deplex {

class BatchPipeline::Impl {
 public:
  dpx_pipeline* p = nullptr;
  ~Impl() { dpx_pipeline_destroy(p); }
  void check(dpx_status st) const {
    if (st != DPX_OK) throw std::runtime_error(dpx_pipeline_last_error(p));
  }
};

BatchPipeline::BatchPipeline(int32_t image_height, int32_t image_width, config::Config config, int32_t max_batch, int32_t lanes,
                             int32_t device)
    : impl_(new Impl()) {
  dpx_config c;
  config.toC(&c);
  if (dpx_pipeline_create(image_height, image_width, &c, device, max_batch, lanes, &impl_->p) != DPX_OK)
    throw std::runtime_error(dpx_pipeline_last_error(nullptr));
}
BatchPipeline::~BatchPipeline() = default;
BatchPipeline::BatchPipeline(BatchPipeline&&) noexcept = default;
BatchPipeline& BatchPipeline::operator=(BatchPipeline&&) noexcept = default;

void BatchPipeline::submit(float const* d_points, int32_t n_frames, PointLayout layout, int32_t* d_labels, void* producer_stream) {
  impl_->check(dpx_pipeline_submit_device(impl_->p, d_points, n_frames, static_cast<dpx_layout>(layout), d_labels, producer_stream));
}
void BatchPipeline::submitDepth(uint16_t const* d_depth, int32_t n_frames, float fx, float fy, float cx, float cy, int32_t* d_labels,
                                void* producer_stream) {
  dpx_intrinsics k{fx, fy, cx, cy};
  impl_->check(dpx_pipeline_submit_depth_device(impl_->p, d_depth, n_frames, &k, d_labels, producer_stream));
}
void BatchPipeline::join(void* consumer_stream) { impl_->check(dpx_pipeline_join(impl_->p, consumer_stream)); }
void BatchPipeline::synchronize() { impl_->check(dpx_pipeline_synchronize(impl_->p)); }
int32_t BatchPipeline::lanes() const { return dpx_pipeline_lanes(impl_->p); }
void* BatchPipeline::handle() const { return impl_->p; }

class SequenceExtractor::Impl {
 public:
  dpx_sequence* s = nullptr;
  ~Impl() { dpx_sequence_destroy(s); }
  void check(dpx_status st) const {
    if (st != DPX_OK) throw std::runtime_error(dpx_sequence_last_error(s));
  }
};

SequenceExtractor::SequenceExtractor(int32_t image_height, int32_t image_width, config::Config config,
                                     std::vector<int32_t> const& devices, int32_t max_batch)
    : impl_(new Impl()) {
  dpx_config c;
  config.toC(&c);
  if (dpx_sequence_create(image_height, image_width, &c, devices.empty() ? nullptr : devices.data(),
                          static_cast<int32_t>(devices.size()), max_batch, &impl_->s) != DPX_OK)
    throw std::runtime_error(dpx_sequence_last_error(nullptr));
}
SequenceExtractor::~SequenceExtractor() = default;
SequenceExtractor::SequenceExtractor(SequenceExtractor&&) noexcept = default;
SequenceExtractor& SequenceExtractor::operator=(SequenceExtractor&&) noexcept = default;

void SequenceExtractor::process(float const* points, int64_t n_frames, PointLayout layout, int32_t* labels) {
  impl_->check(dpx_sequence_process_host(impl_->s, points, n_frames, static_cast<dpx_layout>(layout), labels));
}
void SequenceExtractor::processDepth(uint16_t const* depth, int64_t n_frames, float fx, float fy, float cx, float cy,
                                     int32_t* labels) {
  dpx_intrinsics k{fx, fy, cx, cy};
  impl_->check(dpx_sequence_process_depth_host(impl_->s, depth, n_frames, &k, labels));
}
std::vector<float> SequenceExtractor::processDevice(std::vector<float const*> const& d_points, int64_t n_frames_per_device,
                                                    PointLayout layout, std::vector<int32_t*> const& d_labels, int32_t lanes) {
  const size_t G = static_cast<size_t>(deviceCount());
  if (d_points.size() != G || d_labels.size() != G) throw std::runtime_error("processDevice: one pointer per device is required");
  std::vector<float> ms(G, 0.f);
  impl_->check(dpx_sequence_process_device(impl_->s, d_points.data(), n_frames_per_device, static_cast<dpx_layout>(layout),
                                           d_labels.data(), lanes, ms.data()));
  return ms;
}
int32_t SequenceExtractor::deviceCount() const { return dpx_sequence_devices(impl_->s); }
void SequenceExtractor::frameRange(int64_t n_frames, int32_t slot, int64_t* begin, int64_t* end) const {
  dpx_sequence_range(impl_->s, n_frames, slot, begin, end);
}
void* SequenceExtractor::handle() const { return impl_->s; }

}  // namespace deplex
