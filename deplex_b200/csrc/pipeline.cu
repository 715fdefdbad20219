// pipeline.cu -- the two orchestration layers of the C-ABI above a single extractor (include/deplex_b200.h):
//
//   dpx_pipeline   several batches in flight on ONE GPU: `lanes` extractors on their own streams, fed round-robin
//   dpx_sequence   a frame sequence sharded over the GPUs of one box: contiguous frame ranges, one worker thread per GPU
//
// Both replace the caller-side loop of the reference (examples/process_sequence.cpp:30-43: one process() per frame, one
// thread, one device).  Frames are independent (plane_extractor.cpp:281,428), so neither layer exchanges any data
// between lanes or devices; they only own handles, streams, events and threads.  Built on the public dpx_* entry
// points: nothing here touches labels.
#include "deplex_b200.h"

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "error_state.h"

using namespace dpx;

struct dpx_pipeline {
  int device = 0;
  std::vector<dpx_extractor*> lanes;
  std::vector<cudaStream_t> streams;
  std::vector<cudaEvent_t> done;  // recorded on the lane's stream after its last submitted batch
  std::vector<char> dirty;
  cudaEvent_t ready = nullptr;    // recorded on the producer's stream at submit
  int next = 0;
  std::string err;
};

struct dpx_sequence {
  int height = 0, width = 0;
  int max_batch = 0;
  dpx_config cfg{};
  std::vector<int> devices;
  std::vector<dpx_extractor*> ex;        // one per device: the host-pointer entry points
  std::vector<dpx_pipeline*> pipes;      // one per device, created on first device-resident call
  int pipe_lanes = 0;
  std::string err;
};

namespace {

struct DeviceScope {
  int prev = -1;
  bool ok = true;
  explicit DeviceScope(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceScope() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};

dpx_status pfail(dpx_pipeline* p, dpx_status st, const std::string& msg) {
  if (p) p->err = msg;
  else set_thread_error(msg);
  return st;
}

dpx_status sfail(dpx_sequence* s, dpx_status st, const std::string& msg) {
  if (s) s->err = msg;
  else set_thread_error(msg);
  return st;
}

#define PIPE_CUDA(p, call)                                                                                  \
  do {                                                                                                      \
    cudaError_t e__ = (call);                                                                               \
    if (e__ != cudaSuccess) return pfail((p), DPX_ERR_CUDA, std::string("CUDA error in " #call ": ") + cudaGetErrorString(e__)); \
  } while (0)

// the lane that takes the next batch, its stream ordered behind the producer's
dpx_status take_lane(dpx_pipeline* p, void* producer_stream, int* lane_out) {
  const int lane = p->next;
  p->next = (lane + 1) % static_cast<int>(p->lanes.size());
  PIPE_CUDA(p, cudaEventRecord(p->ready, static_cast<cudaStream_t>(producer_stream)));
  PIPE_CUDA(p, cudaStreamWaitEvent(p->streams[lane], p->ready, 0));
  *lane_out = lane;
  return DPX_OK;
}

dpx_status after_submit(dpx_pipeline* p, int lane, dpx_status st) {
  if (st != DPX_OK) return pfail(p, st, dpx_last_error(p->lanes[lane]));
  PIPE_CUDA(p, cudaEventRecord(p->done[lane], p->streams[lane]));
  p->dirty[lane] = 1;
  return DPX_OK;
}

}  // namespace

namespace {
// one worker thread per device over its contiguous range; calls of at most 4 * max_batch frames keep the per-call chunk
// pipeline long enough to hide its head and tail
template <class Call>
dpx_status run_sharded(dpx_sequence* s, int64_t n_frames, Call call) {
  const int G = static_cast<int>(s->devices.size());
  std::vector<dpx_status> status(G, DPX_OK);
  std::vector<std::thread> workers;
  for (int g = 0; g < G; ++g) {
    workers.emplace_back([&, g] {
      int64_t b = 0, e = 0;
      dpx_sequence_range(s, n_frames, g, &b, &e);
      const int64_t step = std::max<int64_t>(1, static_cast<int64_t>(s->max_batch) * 4);
      for (int64_t f = b; f < e && status[g] == DPX_OK; f += step)
        status[g] = call(s->ex[g], f, static_cast<int32_t>(std::min<int64_t>(step, e - f)));
    });
  }
  for (auto& t : workers) t.join();
  for (int g = 0; g < G; ++g)
    if (status[g] != DPX_OK)
      return sfail(s, status[g], "device " + std::to_string(s->devices[g]) + ": " + dpx_last_error(s->ex[g]));
  return DPX_OK;
}
}  // namespace

extern "C" {

// ------------------------------------------------------------------------------------------------ dpx_pipeline

dpx_status dpx_pipeline_create(int32_t height, int32_t width, const dpx_config* cfg, int32_t device, int32_t max_batch,
                               int32_t lanes, dpx_pipeline** out) {
  if (!out) return pfail(nullptr, DPX_ERR_ARGUMENT, "dpx_pipeline_create: out is NULL");
  *out = nullptr;
  if (lanes < 1 || lanes > 16) return pfail(nullptr, DPX_ERR_ARGUMENT, "dpx_pipeline_create: lanes must be in [1, 16]");
  if (device < 0) {
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess)
      return pfail(nullptr, DPX_ERR_CUDA, std::string("no usable CUDA device (this library has no CPU path): ") + cudaGetErrorString(e));
  }
  dpx_pipeline* p = new dpx_pipeline();
  p->device = device;
  for (int i = 0; i < lanes; ++i) {
    dpx_extractor* ex = nullptr;
    const dpx_status st = dpx_create(height, width, cfg, device, max_batch, &ex);
    if (st != DPX_OK) {  // the reference's constructor errors pass through unchanged (thread-local message)
      dpx_pipeline_destroy(p);
      return st;
    }
    p->lanes.push_back(ex);
  }
  DeviceScope scope(device);
  p->streams.assign(lanes, nullptr);
  p->done.assign(lanes, nullptr);
  p->dirty.assign(lanes, 0);
  cudaError_t e = cudaEventCreateWithFlags(&p->ready, cudaEventDisableTiming);
  for (int i = 0; i < lanes && e == cudaSuccess; ++i) {
    e = cudaStreamCreateWithFlags(&p->streams[i], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->done[i], cudaEventDisableTiming);
  }
  if (e != cudaSuccess) {
    dpx_pipeline_destroy(p);
    return pfail(nullptr, DPX_ERR_CUDA, std::string("dpx_pipeline_create: ") + cudaGetErrorString(e));
  }
  *out = p;
  return DPX_OK;
}

void dpx_pipeline_destroy(dpx_pipeline* p) {
  if (!p) return;
  {
    DeviceScope scope(p->device);
    for (cudaStream_t s : p->streams)
      if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
    for (cudaEvent_t e : p->done)
      if (e) cudaEventDestroy(e);
    if (p->ready) cudaEventDestroy(p->ready);
  }
  for (dpx_extractor* ex : p->lanes) dpx_destroy(ex);
  delete p;
}

const char* dpx_pipeline_last_error(const dpx_pipeline* p) { return p ? p->err.c_str() : thread_error(); }

int32_t dpx_pipeline_lanes(const dpx_pipeline* p) { return p ? static_cast<int32_t>(p->lanes.size()) : 0; }

dpx_extractor* dpx_pipeline_lane(dpx_pipeline* p, int32_t lane) {
  if (!p || lane < 0 || lane >= static_cast<int32_t>(p->lanes.size())) return nullptr;
  return p->lanes[lane];
}

dpx_status dpx_pipeline_submit_device(dpx_pipeline* p, const float* d_xyz, int32_t n_frames, dpx_layout layout, int32_t* d_labels,
                                      void* producer_stream) {
  if (!p) return DPX_ERR_ARGUMENT;
  DeviceScope scope(p->device);
  if (!scope.ok) return pfail(p, DPX_ERR_CUDA, "cudaSetDevice failed");
  int lane = 0;
  const dpx_status st = take_lane(p, producer_stream, &lane);
  if (st != DPX_OK) return st;
  return after_submit(p, lane, dpx_process_batch_device(p->lanes[lane], d_xyz, n_frames, layout, d_labels, p->streams[lane]));
}

dpx_status dpx_pipeline_submit_depth_device(dpx_pipeline* p, const uint16_t* d_depth, int32_t n_frames, const dpx_intrinsics* k,
                                            int32_t* d_labels, void* producer_stream) {
  if (!p) return DPX_ERR_ARGUMENT;
  DeviceScope scope(p->device);
  if (!scope.ok) return pfail(p, DPX_ERR_CUDA, "cudaSetDevice failed");
  int lane = 0;
  const dpx_status st = take_lane(p, producer_stream, &lane);
  if (st != DPX_OK) return st;
  return after_submit(p, lane, dpx_process_depth_batch_device(p->lanes[lane], d_depth, n_frames, k, d_labels, p->streams[lane]));
}

dpx_status dpx_pipeline_join(dpx_pipeline* p, void* consumer_stream) {
  if (!p) return DPX_ERR_ARGUMENT;
  DeviceScope scope(p->device);
  if (!scope.ok) return pfail(p, DPX_ERR_CUDA, "cudaSetDevice failed");
  for (size_t i = 0; i < p->lanes.size(); ++i) {
    if (!p->dirty[i]) continue;
    PIPE_CUDA(p, cudaStreamWaitEvent(static_cast<cudaStream_t>(consumer_stream), p->done[i], 0));
    p->dirty[i] = 0;
  }
  return DPX_OK;
}

dpx_status dpx_pipeline_synchronize(dpx_pipeline* p) {
  if (!p) return DPX_ERR_ARGUMENT;
  DeviceScope scope(p->device);
  if (!scope.ok) return pfail(p, DPX_ERR_CUDA, "cudaSetDevice failed");
  for (size_t i = 0; i < p->lanes.size(); ++i) PIPE_CUDA(p, cudaStreamSynchronize(p->streams[i]));
  return DPX_OK;
}

int64_t dpx_pipeline_kernel_launches(const dpx_pipeline* p) {
  int64_t n = 0;
  if (p)
    for (const dpx_extractor* ex : p->lanes) n += dpx_kernel_launches(ex);
  return n;
}

// ------------------------------------------------------------------------------------------------ dpx_sequence

dpx_status dpx_sequence_create(int32_t height, int32_t width, const dpx_config* cfg, const int32_t* devices, int32_t n_devices,
                               int32_t max_batch, dpx_sequence** out) {
  if (!out) return sfail(nullptr, DPX_ERR_ARGUMENT, "dpx_sequence_create: out is NULL");
  *out = nullptr;
  int visible = 0;
  {
    cudaError_t e = cudaGetDeviceCount(&visible);
    if (e != cudaSuccess || visible == 0)
      return sfail(nullptr, DPX_ERR_CUDA, std::string("no usable CUDA device (this library has no CPU path): ") + cudaGetErrorString(e));
  }
  dpx_sequence* s = new dpx_sequence();
  s->height = height;
  s->width = width;
  s->max_batch = max_batch;
  if (cfg) s->cfg = *cfg;
  else dpx_config_default(&s->cfg);
  if (devices && n_devices > 0) s->devices.assign(devices, devices + n_devices);
  else
    for (int d = 0; d < visible; ++d) s->devices.push_back(d);
  for (int d : s->devices) {
    dpx_extractor* ex = nullptr;
    const dpx_status st = dpx_create(height, width, &s->cfg, d, max_batch, &ex);
    if (st != DPX_OK) {
      dpx_sequence_destroy(s);
      return st;
    }
    s->ex.push_back(ex);
  }
  *out = s;
  return DPX_OK;
}

void dpx_sequence_destroy(dpx_sequence* s) {
  if (!s) return;
  for (dpx_pipeline* p : s->pipes) dpx_pipeline_destroy(p);
  for (dpx_extractor* ex : s->ex) dpx_destroy(ex);
  delete s;
}

const char* dpx_sequence_last_error(const dpx_sequence* s) { return s ? s->err.c_str() : thread_error(); }

int32_t dpx_sequence_devices(const dpx_sequence* s) { return s ? static_cast<int32_t>(s->devices.size()) : 0; }

void dpx_sequence_range(const dpx_sequence* s, int64_t n_frames, int32_t slot, int64_t* begin, int64_t* end) {
  const int64_t G = s ? static_cast<int64_t>(s->devices.size()) : 1;
  const int64_t base = n_frames / G, extra = n_frames % G;
  const int64_t b = slot * base + std::min<int64_t>(slot, extra);
  if (begin) *begin = b;
  if (end) *end = b + base + (slot < extra ? 1 : 0);
}

dpx_status dpx_sequence_process_host(dpx_sequence* s, const float* xyz, int64_t n_frames, dpx_layout layout, int32_t* labels) {
  if (!s) return DPX_ERR_ARGUMENT;
  if (n_frames < 0) return sfail(s, DPX_ERR_ARGUMENT, "negative n_frames");
  if (n_frames == 0) return DPX_OK;
  if (!xyz || !labels) return sfail(s, DPX_ERR_ARGUMENT, "null host pointer");
  const size_t np = static_cast<size_t>(s->height) * s->width;
  return run_sharded(s, n_frames, [&](dpx_extractor* ex, int64_t f, int32_t nf) {
    return dpx_process_batch_host(ex, xyz + static_cast<size_t>(f) * np * 3, nf, layout, labels + static_cast<size_t>(f) * np);
  });
}

dpx_status dpx_sequence_process_depth_host(dpx_sequence* s, const uint16_t* depth, int64_t n_frames, const dpx_intrinsics* k,
                                           int32_t* labels) {
  if (!s) return DPX_ERR_ARGUMENT;
  if (n_frames < 0) return sfail(s, DPX_ERR_ARGUMENT, "negative n_frames");
  if (n_frames == 0) return DPX_OK;
  if (!depth || !labels || !k) return sfail(s, DPX_ERR_ARGUMENT, "null pointer");
  const size_t np = static_cast<size_t>(s->height) * s->width;
  return run_sharded(s, n_frames, [&](dpx_extractor* ex, int64_t f, int32_t nf) {
    return dpx_process_depth_batch_host(ex, depth + static_cast<size_t>(f) * np, nf, k, labels + static_cast<size_t>(f) * np);
  });
}

dpx_status dpx_sequence_process_device(dpx_sequence* s, const float* const* d_xyz, int64_t n_frames_per_device, dpx_layout layout,
                                       int32_t* const* d_labels, int32_t lanes, float* ms_per_device) {
  if (!s) return DPX_ERR_ARGUMENT;
  if (n_frames_per_device < 0) return sfail(s, DPX_ERR_ARGUMENT, "negative n_frames");
  if (!d_xyz || !d_labels) return sfail(s, DPX_ERR_ARGUMENT, "null pointer table");
  if (lanes < 1) lanes = 3;
  const int G = static_cast<int>(s->devices.size());
  if (s->pipes.empty() || s->pipe_lanes != lanes) {
    for (dpx_pipeline* p : s->pipes) dpx_pipeline_destroy(p);
    s->pipes.clear();
    for (int g = 0; g < G; ++g) {
      dpx_pipeline* p = nullptr;
      const dpx_status st = dpx_pipeline_create(s->height, s->width, &s->cfg, s->devices[g], s->max_batch, lanes, &p);
      if (st != DPX_OK) return sfail(s, st, dpx_pipeline_last_error(nullptr));
      s->pipes.push_back(p);
    }
    s->pipe_lanes = lanes;
  }
  const size_t np = static_cast<size_t>(s->height) * s->width;
  std::vector<dpx_status> status(G, DPX_OK);
  std::vector<std::string> msg(G);
  std::vector<float> ms(G, 0.f);
  std::vector<std::thread> workers;
  for (int g = 0; g < G; ++g) {
    workers.emplace_back([&, g] {
      auto cuda_ok = [&](cudaError_t e, const char* what) {
        if (e == cudaSuccess) return true;
        status[g] = DPX_ERR_CUDA;
        msg[g] = std::string(what) + ": " + cudaGetErrorString(e);
        return false;
      };
      if (!cuda_ok(cudaSetDevice(s->devices[g]), "cudaSetDevice")) return;
      cudaStream_t st = nullptr;
      cudaEvent_t e0 = nullptr, e1 = nullptr;
      if (!cuda_ok(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking), "cudaStreamCreate")) return;
      if (cuda_ok(cudaEventCreate(&e0), "cudaEventCreate") && cuda_ok(cudaEventCreate(&e1), "cudaEventCreate") &&
          cuda_ok(cudaEventRecord(e0, st), "cudaEventRecord")) {
        dpx_pipeline* p = s->pipes[g];
        for (int64_t f = 0; f < n_frames_per_device; f += s->max_batch) {
          const int32_t nf = static_cast<int32_t>(std::min<int64_t>(s->max_batch, n_frames_per_device - f));
          const dpx_status r = dpx_pipeline_submit_device(p, d_xyz[g] + static_cast<size_t>(f) * np * 3, nf, layout,
                                                          d_labels[g] + static_cast<size_t>(f) * np, st);
          if (r != DPX_OK) {
            status[g] = r;
            msg[g] = dpx_pipeline_last_error(p);
            break;
          }
        }
        if (status[g] == DPX_OK && dpx_pipeline_join(p, st) != DPX_OK) {
          status[g] = DPX_ERR_CUDA;
          msg[g] = dpx_pipeline_last_error(p);
        }
        if (cuda_ok(cudaEventRecord(e1, st), "cudaEventRecord") && cuda_ok(cudaStreamSynchronize(st), "cudaStreamSynchronize"))
          cuda_ok(cudaEventElapsedTime(&ms[g], e0, e1), "cudaEventElapsedTime");
      }
      if (e0) cudaEventDestroy(e0);
      if (e1) cudaEventDestroy(e1);
      cudaStreamDestroy(st);
    });
  }
  for (auto& t : workers) t.join();
  for (int g = 0; g < G; ++g) {
    if (ms_per_device) ms_per_device[g] = ms[g];
    if (status[g] != DPX_OK) return sfail(s, status[g], "device " + std::to_string(s->devices[g]) + ": " + msg[g]);
  }
  return DPX_OK;
}

// ------------------------------------------------------------------------------------------------ device memory helpers

int32_t dpx_device_count(void) {
  int n = 0;
  return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

dpx_status dpx_device_alloc(int32_t device, void** ptr, size_t bytes) {
  if (!ptr) return DPX_ERR_ARGUMENT;
  DeviceScope scope(device);
  if (!scope.ok) return pfail(nullptr, DPX_ERR_CUDA, "cudaSetDevice failed");
  PIPE_CUDA(nullptr, cudaMalloc(ptr, bytes));
  return DPX_OK;
}

void dpx_device_free(int32_t device, void* ptr) {
  if (!ptr) return;
  DeviceScope scope(device);
  cudaFree(ptr);
}

dpx_status dpx_memcpy_to_device(int32_t device, void* dst, const void* src, size_t bytes) {
  DeviceScope scope(device);
  if (!scope.ok) return pfail(nullptr, DPX_ERR_CUDA, "cudaSetDevice failed");
  PIPE_CUDA(nullptr, cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
  return DPX_OK;
}

dpx_status dpx_memcpy_to_host(int32_t device, void* dst, const void* src, size_t bytes) {
  DeviceScope scope(device);
  if (!scope.ok) return pfail(nullptr, DPX_ERR_CUDA, "cudaSetDevice failed");
  PIPE_CUDA(nullptr, cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
  return DPX_OK;
}

}  // extern "C"
