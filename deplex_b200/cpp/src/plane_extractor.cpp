// plane_extractor.cpp -- deplex::PlaneExtractor as a thin owner of a C-ABI handle.
//
// Reference: cpp/deplex/src/deplex/plane_extractor.cpp:153-185 (constructor, destructor, move operations,
// process forwarding to Impl).  Error behaviour: every DPX_ERR_RUNTIME carries the reference's own
// std::runtime_error text; unsupported inputs (undefined behaviour in the reference) and CUDA failures are
// std::runtime_error too, with an explanatory text.  There is no CPU path behind this class.
#include "deplex/plane_extractor.h"

#include <stdexcept>
#include <string>

#include "deplex_b200.h"

namespace deplex {

class PlaneExtractor::Impl {
 public:
  Impl(int32_t h, int32_t w, config::Config const& cfg, int32_t max_batch, int32_t device) : height(h), width(w) {
    dpx_config c;
    cfg.toC(&c);
    const dpx_status st = dpx_create(h, w, &c, device, max_batch, &ex);
    if (st != DPX_OK) throw std::runtime_error(dpx_last_error(nullptr));
  }
  ~Impl() { dpx_destroy(ex); }
  Impl(Impl const&) = delete;
  Impl& operator=(Impl const&) = delete;

  void check(dpx_status st) const {
    if (st != DPX_OK) throw std::runtime_error(dpx_last_error(ex));
  }

  dpx_extractor* ex = nullptr;
  int32_t height, width;
};

PlaneExtractor::PlaneExtractor(int32_t image_height, int32_t image_width, config::Config config)
    : impl_(new Impl(image_height, image_width, config, 1, -1)) {}

PlaneExtractor::PlaneExtractor(int32_t image_height, int32_t image_width, config::Config config, int32_t max_batch,
                               int32_t device)
    : impl_(new Impl(image_height, image_width, config, max_batch, device)) {}

PlaneExtractor::~PlaneExtractor() = default;
PlaneExtractor::PlaneExtractor(PlaneExtractor&&) noexcept = default;
PlaneExtractor& PlaneExtractor::operator=(PlaneExtractor&& op) noexcept = default;

void PlaneExtractor::process(float const* points, int64_t n_points, PointLayout layout, int32_t* labels) {
  impl_->check(dpx_process_host(impl_->ex, points, n_points, static_cast<dpx_layout>(layout), labels));
}

std::vector<int32_t> PlaneExtractor::process(float const* points, int64_t n_points, PointLayout layout) {
  std::vector<int32_t> labels(n_points > 0 ? static_cast<size_t>(n_points) : 0);
  process(points, n_points, layout, labels.data());
  return labels;
}

void PlaneExtractor::processBatch(float const* points, int32_t n_frames, PointLayout layout, int32_t* labels) {
  impl_->check(dpx_process_batch_host(impl_->ex, points, n_frames, static_cast<dpx_layout>(layout), labels));
}

void PlaneExtractor::processBatchDevice(float const* d_points, int32_t n_frames, PointLayout layout, int32_t* d_labels,
                                        void* cuda_stream) {
  impl_->check(dpx_process_batch_device(impl_->ex, d_points, n_frames, static_cast<dpx_layout>(layout), d_labels, cuda_stream));
}

void PlaneExtractor::processDepthBatch(uint16_t const* depth, int32_t n_frames, float fx, float fy, float cx, float cy,
                                       int32_t* labels) {
  dpx_intrinsics k{fx, fy, cx, cy};
  impl_->check(dpx_process_depth_batch_host(impl_->ex, depth, n_frames, &k, labels));
}

std::vector<PlaneParams> PlaneExtractor::planes(int32_t frame) {
  dpx_info info;
  impl_->check(dpx_get_info(impl_->ex, &info));
  std::vector<dpx_plane> raw(static_cast<size_t>(info.plane_capacity));
  int32_t n = 0;
  impl_->check(dpx_get_planes(impl_->ex, frame, raw.data(), info.plane_capacity, &n));
  std::vector<PlaneParams> out(static_cast<size_t>(n));
  for (int32_t i = 0; i < n; ++i) {
    for (int a = 0; a < 3; ++a) {
      out[i].normal[a] = raw[i].normal[a];
      out[i].mean[a] = raw[i].mean[a];
    }
    out[i].d = raw[i].d;
    out[i].mse = raw[i].mse;
    out[i].score = raw[i].score;
    out[i].n_points = raw[i].n_points;
    out[i].merge_label = raw[i].merge_label;
  }
  return out;
}

int32_t PlaneExtractor::imageHeight() const { return impl_->height; }
int32_t PlaneExtractor::imageWidth() const { return impl_->width; }
void* PlaneExtractor::handle() const { return impl_->ex; }

}  // namespace deplex
