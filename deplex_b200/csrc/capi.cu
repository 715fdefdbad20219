// capi.cu -- the C-ABI of libdeplex_b200.so (include/deplex_b200.h): handle management, device scratch,
// stream plumbing and the three-stage launch sequence.  No arithmetic on labels happens here.
//
// Reference behaviour mirrored (file:line under the reference tree):
//   constructor geometry, patch clamp, patchSize==0 error      plane_extractor.cpp:153-176
//   size check and its message                                  plane_extractor.cpp:188-194
//   stage order cell grid -> histogram -> growing -> merging -> labels   plane_extractor.cpp:195-283
#include "deplex_b200.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "cell_stats.cuh"
#include "depth_points.cuh"
#include "error_state.h"
#include "host_pool.h"
#include "labeling.cuh"
#include "refine.cuh"
#include "region_grow.cuh"

namespace dpx {
namespace {
thread_local std::string g_thread_error;
}
void set_thread_error(const std::string& msg) { g_thread_error = msg; }
const char* thread_error() { return g_thread_error.c_str(); }
}  // namespace dpx

using namespace dpx;

constexpr int kHostSlots = 4;     // chunks in flight on the host-pointer path
constexpr int kHostLookahead = 2; // chunks queued behind the one whose labels the host is waiting for

struct dpx_extractor {
  dpx_config cfg;
  Geometry geom;
  Thresholds thr;
  Tables tb;
  int device = 0, max_batch = 0, sm_count = 0;
  int tile_cells = 0;
  int stream_warps = 16;      // env DPX_STREAM_WARPS=8|12|16 (A/B measurement)
  int force_tile_kernel = 0;  // env DPX_CELL_KERNEL=tile (A/B measurement of the two stage-1 kernels)
  int fuse_labeling = 1;      // env DPX_FUSE_LABELING=0: keep stage 3 as its own kernel (A/B measurement)
  RegionPlan plan{};
  uint32_t* mt_init = nullptr;       // std::mt19937 default state for the refinement stage
  unsigned long long* refine_work = nullptr;  // [max_batch][2] point passes, rounds (dpx_get_refine_work)
  long long* region_prof = nullptr;  // [max_batch][kRegionProfSlots], written while profiling is on
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  // host-pointer path: kHostSlots staging buffers in flight on three streams (copy in, kernels, copy out)
  int host_chunk_xyz = 0, host_chunk_depth = 0;  // frames per pipelined chunk (sized by bytes, see ensure_host_path)
  float* d_xyz[kHostSlots] = {};
  uint16_t* d_depth[kHostSlots] = {};  // staging of the raw-depth host path
  float* d_conv = nullptr;             // points generated from depth when the fused path cannot run
  size_t d_conv_frames = 0;
  int32_t* d_lab[kHostSlots] = {};
  int lab_chunk = 0;                   // frames d_lab[] / d_nar[] / h_nar[] are sized for
  // narrow label transport: labels cross PCIe as uint16 and are widened into the caller's int32 buffer by host threads
  int label_transport = DPX_LABELS_AUTO;
  int uniform_variant = 0;    // std::uniform_int_distribution mapping of the refinement stage (dpx_set_rng_compat)
  uint16_t* d_nar[kHostSlots] = {};
  uint16_t* h_nar[kHostSlots] = {};    // pinned
  uint64_t widen_ticket[kHostSlots] = {};
  std::unique_ptr<HostPool> pool;
  cudaStream_t s_h2d = nullptr, s_run = nullptr, s_d2h = nullptr;
  cudaEvent_t e_h2d[kHostSlots] = {}, e_run[kHostSlots] = {}, e_d2h[kHostSlots] = {};
  // measurement
  bool profiling = false;
  cudaEvent_t e_stage[DPX_N_STAGES + 1] = {};
  bool stage_valid = false;
  int64_t launches = 0;
  int last_frames = 0;
  int last_first = 0;  // index, within the last call, of the first frame whose tables are still resident (host path: last chunk)
  std::string err;
};

namespace {

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};

dpx_status fail(dpx_extractor* ex, dpx_status st, const std::string& msg) {
  if (ex) ex->err = msg;
  else set_thread_error(msg);
  return st;
}

dpx_status cuda_fail(dpx_extractor* ex, cudaError_t e, const char* what) {
  return fail(ex, DPX_ERR_CUDA, std::string("CUDA error in ") + what + ": " + cudaGetErrorString(e));
}

#define DPX_CUDA(ex, call)                                        \
  do {                                                            \
    cudaError_t e__ = (call);                                     \
    if (e__ != cudaSuccess) return cuda_fail((ex), e__, #call);   \
  } while (0)

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Carve the frame-major tables out of one allocation.
size_t carve_tables(const Geometry& g, int max_batch, bool bins_in_smem, char* base, Tables* tb) {
  size_t off = 0;
  const size_t F = static_cast<size_t>(max_batch), C = static_cast<size_t>(g.n_cells), P = static_cast<size_t>(g.plane_cap);
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off = align_up(off + bytes, 256);
    return p;
  };
  tb->rec_a = reinterpret_cast<float4*>(take(F * C * 2 * sizeof(float4)));
  tb->rec_b = reinterpret_cast<float4*>(take(F * C * 3 * sizeof(float4)));
  tb->bin = reinterpret_cast<int16_t*>(take(F * C * sizeof(int16_t)));
  tb->flags = reinterpret_cast<uint8_t*>(take(F * C));
  tb->mse = reinterpret_cast<float*>(take(F * C * sizeof(float)));
  tb->edge = reinterpret_cast<uint8_t*>(take(F * C));
  tb->seg_label = reinterpret_cast<int32_t*>(take(F * C * sizeof(int32_t)));
  tb->cell_label = reinterpret_cast<int32_t*>(take(F * C * sizeof(int32_t)));
  tb->queue = reinterpret_cast<int32_t*>(take(F * C * sizeof(int32_t)));
  tb->pairs = reinterpret_cast<uint32_t*>(take(F * C * 2 * sizeof(uint32_t)));
  tb->skeys = reinterpret_cast<unsigned long long*>(take(F * C * sizeof(unsigned long long)));
  tb->bin_work = reinterpret_cast<int16_t*>(take(bins_in_smem ? 0 : F * C * sizeof(int16_t)));
  tb->cell_words = reinterpret_cast<uint32_t*>(take(F * C * sizeof(uint32_t)));
  tb->paint_state = reinterpret_cast<int32_t*>(take((F + 2) * sizeof(int32_t)));
  tb->segs = reinterpret_cast<float*>(take(F * P * kSegFloats * sizeof(float)));
  tb->merge = reinterpret_cast<int32_t*>(take(F * P * sizeof(int32_t)));
  tb->n_planes = reinterpret_cast<int32_t*>(take(F * sizeof(int32_t)));
  tb->axis_work = reinterpret_cast<int32_t*>(take((1 + kAxisWorkCap) * sizeof(int32_t)));
  return off;
}

// The three stages on one stream (plane_extractor.cpp:195-283).
// d_xyz + layout, or (layout == kLayoutDepth16) d_depth + pin
dpx_status run_stages(dpx_extractor* ex, const float* d_xyz, int n_frames, int layout, int32_t* d_labels, cudaStream_t st,
                      const uint16_t* d_depth = nullptr, const dpx_intrinsics* pin = nullptr) {
  const bool prof = ex->profiling;
  if (prof) DPX_CUDA(ex, cudaEventRecord(ex->e_stage[0], st));
  Pinhole ph{};
  if (layout == kLayoutDepth16 && ex->geom.n_cells > 0) {
    ph.fx = pin->fx; ph.fy = pin->fy; ph.cx = pin->cx; ph.cy = pin->cy;
    CellStatsArgs probe{};
    probe.depth = d_depth;
    probe.geom = ex->geom;
    probe.force_tile_kernel = ex->force_tile_kernel;
    if (ex->cfg.ransac_refinement || !cell_stats_depth_eligible(probe)) {
      // materialise the points once (row-major) and continue on the ordinary path
      if (ex->d_conv_frames < static_cast<size_t>(n_frames)) {
        // first use only, stream-ordered (cudaMallocAsync does not synchronise the device, so the call stays
        // asynchronous on the caller's stream); the buffer is kept for the life of the handle
        DPX_CUDA(ex, cudaMallocAsync(reinterpret_cast<void**>(&ex->d_conv),
                                     static_cast<size_t>(ex->max_batch) * ex->geom.n_points * 3 * sizeof(float), st));
        ex->d_conv_frames = static_cast<size_t>(ex->max_batch);
      }
      DPX_CUDA(ex, launch_depth_to_points(d_depth, n_frames, ex->geom, ph, ex->d_conv, st));
      ++ex->launches;
      d_xyz = ex->d_conv;
      layout = kLayoutRowMajor;
    }
  }
  if (ex->geom.n_cells > 0) {
    CellStatsArgs ca{};
    ca.xyz = d_xyz;
    ca.depth = d_depth;
    ca.pin = ph;
    ca.n_frames = n_frames;
    ca.layout = layout;
    ca.tile_cells = ex->tile_cells;
    ca.vec_ok = (reinterpret_cast<uintptr_t>(d_xyz) % 16 == 0) && (ex->geom.n_points % 4 == 0) && (ex->geom.width % 4 == 0);
    ca.sm_count = ex->sm_count;
    ca.force_tile_kernel = ex->force_tile_kernel;
    ca.stream_warps = ex->stream_warps;
    ca.geom = ex->geom;
    ca.thr = ex->thr;
    ca.tables = ex->tb;
    DPX_CUDA(ex, launch_cell_stats(ca, st));
    ++ex->launches;
  }
  if (prof) DPX_CUDA(ex, cudaEventRecord(ex->e_stage[1], st));
  bool labels_painted = false;
  if (ex->geom.n_cells > 0) {
    RegionArgs ra{};
    ra.n_frames = n_frames;
    ra.plan = ex->plan;
    ra.prof = ex->profiling ? ex->region_prof : nullptr;
    ra.geom = ex->geom;
    ra.thr = ex->thr;
    ra.tables = ex->tb;
    ra.labels = ex->fuse_labeling ? d_labels : nullptr;
    ra.labels_vec_ok = labels_vec_ok(ex->geom, d_labels) ? 1 : 0;
    DPX_CUDA(ex, launch_region_grow(ra, st, &labels_painted));
    ex->launches += region_grow_mode(ex->geom, ex->thr) >= 1 ? 4 : 3;  // edge masks + axis repair (+ seed sort) + region growing
  }
  if (prof) DPX_CUDA(ex, cudaEventRecord(ex->e_stage[2], st));
  {
    // stage 3; frames the region-growing kernel has already painted are skipped (their flag is set)
    LabelArgs la{};
    la.n_frames = n_frames;
    la.geom = ex->geom;
    la.cell_label = ex->tb.cell_label;
    la.labels = d_labels;
    la.todo = labels_painted ? ex->tb.paint_state + 1 : nullptr;
    la.vec_ok = labels_vec_ok(ex->geom, d_labels) ? 1 : 0;
    DPX_CUDA(ex, launch_labeling(la, st));
    if (ex->geom.n_cells > 0) ++ex->launches;
  }
  if (prof) DPX_CUDA(ex, cudaEventRecord(ex->e_stage[3], st));
  if (ex->cfg.ransac_refinement && ex->geom.n_cells > 0) {  // plane_extractor.cpp:265-267
    RefineArgs fa{};
    fa.xyz = d_xyz;
    fa.layout = layout;
    fa.n_frames = n_frames;
    fa.max_iterations = ex->cfg.ransac_max_iterations;
    fa.threshold = ex->cfg.ransac_threshold;
    fa.inliers_ratio = ex->cfg.ransac_inliers_ratio;
    fa.mt_init = ex->mt_init;
    fa.work = ex->refine_work;
    fa.uniform_variant = ex->uniform_variant;
    fa.geom = ex->geom;
    fa.tables = ex->tb;
    fa.labels = d_labels;
    DPX_CUDA(ex, launch_refine(fa, st));
    ++ex->launches;
  }
  if (prof) {
    DPX_CUDA(ex, cudaEventRecord(ex->e_stage[4], st));
    ex->stage_valid = true;
  }
  ex->last_frames = n_frames;
  ex->last_first = 0;
  return DPX_OK;
}

int env_int(const char* name, int fallback) {
  const char* e = std::getenv(name);
  return e && *e ? std::atoi(e) : fallback;
}

// Streams, events and chunk sizes of the host-pointer path (first host call only).
dpx_status ensure_host_path(dpx_extractor* ex) {
  if (ex->s_run) return DPX_OK;
  // A chunk is what one copy moves: about 64 MB, so that the copy engines work in long transfers while the kernels of
  // the previous chunk run (640x480: 16 frames of points, 34 frames of raw depth).  DPX_HOST_CHUNK overrides (A/B).
  const size_t np = std::max<size_t>(1, static_cast<size_t>(ex->geom.n_points));
  const size_t target = 64u << 20;
  const int forced = env_int("DPX_HOST_CHUNK", 0);
  auto sized = [&](size_t bytes_per_frame) {
    const int c = forced > 0 ? forced : static_cast<int>(std::min<size_t>(64, std::max<size_t>(1, target / bytes_per_frame)));
    return std::max(1, std::min(ex->max_batch, c));
  };
  ex->host_chunk_xyz = sized(np * 3 * sizeof(float));
  ex->host_chunk_depth = sized(np * (sizeof(uint16_t) + sizeof(int32_t)));  // here the labels going back are the larger copy
  for (int i = 0; i < kHostSlots; ++i) {
    DPX_CUDA(ex, cudaEventCreateWithFlags(&ex->e_h2d[i], cudaEventDisableTiming));
    DPX_CUDA(ex, cudaEventCreateWithFlags(&ex->e_run[i], cudaEventDisableTiming));
    DPX_CUDA(ex, cudaEventCreateWithFlags(&ex->e_d2h[i], cudaEventDisableTiming));
  }
  DPX_CUDA(ex, cudaStreamCreateWithFlags(&ex->s_h2d, cudaStreamNonBlocking));
  DPX_CUDA(ex, cudaStreamCreateWithFlags(&ex->s_d2h, cudaStreamNonBlocking));
  DPX_CUDA(ex, cudaStreamCreateWithFlags(&ex->s_run, cudaStreamNonBlocking));
  return DPX_OK;
}

// Device label staging for chunks of up to `chunk` frames (grown on demand: the depth path uses larger chunks).
dpx_status ensure_label_staging(dpx_extractor* ex, int chunk, bool narrow, bool widen) {
  const size_t np = static_cast<size_t>(ex->geom.n_points);
  if (chunk > ex->lab_chunk) {
    DPX_CUDA(ex, cudaDeviceSynchronize());
    for (int i = 0; i < kHostSlots; ++i) {
      if (ex->d_lab[i]) DPX_CUDA(ex, cudaFree(ex->d_lab[i]));
      if (ex->d_nar[i]) DPX_CUDA(ex, cudaFree(ex->d_nar[i]));
      if (ex->h_nar[i]) DPX_CUDA(ex, cudaFreeHost(ex->h_nar[i]));
      ex->d_lab[i] = nullptr; ex->d_nar[i] = nullptr; ex->h_nar[i] = nullptr;
      DPX_CUDA(ex, cudaMalloc(&ex->d_lab[i], std::max<size_t>(16, np * sizeof(int32_t) * chunk)));
    }
    ex->lab_chunk = chunk;
  }
  if (narrow && !ex->d_nar[0])
    for (int i = 0; i < kHostSlots; ++i)
      DPX_CUDA(ex, cudaMalloc(&ex->d_nar[i], std::max<size_t>(16, np * sizeof(uint16_t) * ex->lab_chunk)));
  if (widen && !ex->h_nar[0])
    for (int i = 0; i < kHostSlots; ++i)
      DPX_CUDA(ex, cudaHostAlloc(reinterpret_cast<void**>(&ex->h_nar[i]), std::max<size_t>(16, np * sizeof(uint16_t) * ex->lab_chunk),
                                 cudaHostAllocDefault));
  if (widen && !ex->pool) {
    // widening threads: DPX_HOST_THREADS, else the host's hardware threads shared among the visible GPUs, at most 8
    int n_dev = 1;
    cudaGetDeviceCount(&n_dev);
    const int hw = static_cast<int>(std::thread::hardware_concurrency());
    const int n = env_int("DPX_HOST_THREADS", std::max(1, std::min(8, hw / std::max(1, n_dev))));
    ex->pool.reset(new HostPool(std::max(1, n)));
  }
  return DPX_OK;
}

// Chunk sizes of one host batch: a short ramp up (the first kernels start after a small copy), full chunks, and a ramp
// down (the last labels come back after a small copy), so that the unoverlapped head and tail of the three-stage pipeline
// cost a couple of frames' worth of PCIe time instead of a full chunk's.
std::vector<int> chunk_schedule(int n_frames, int chunk) {
  std::vector<int> up, down, out;
  int ramp = 0;
  for (int c = std::max(1, chunk / 8); c < chunk; c *= 2) { up.push_back(c); ramp += c; }
  down.assign(up.rbegin(), up.rend());
  if (n_frames < 2 * ramp + chunk) {  // too short to taper: equal chunks
    for (int f = 0; f < n_frames; f += chunk) out.push_back(std::min(chunk, n_frames - f));
    return out;
  }
  out = up;
  int middle = n_frames - 2 * ramp;
  for (; middle >= chunk; middle -= chunk) out.push_back(chunk);
  if (middle > 0) out.push_back(middle);
  out.insert(out.end(), down.begin(), down.end());
  return out;
}

}  // namespace

extern "C" {

int32_t dpx_version(void) { return DPX_VERSION; }

const char* dpx_last_error(const dpx_extractor* ex) { return ex ? ex->err.c_str() : thread_error(); }

dpx_status dpx_create(int32_t height, int32_t width, const dpx_config* cfg_in, int32_t device, int32_t max_batch,
                      dpx_extractor** out) {
  if (!out) return fail(nullptr, DPX_ERR_ARGUMENT, "dpx_create: out is NULL");
  *out = nullptr;
  dpx_config cfg;
  if (cfg_in) cfg = *cfg_in;
  else dpx_config_default(&cfg);

  // plane_extractor.cpp:155-164 -- cell counts come from the UNclamped patch size, then the clamp, then the check
  if (cfg.patch_size == 0)
    return fail(nullptr, DPX_ERR_RUNTIME,
                "Error! Invalid config parameter: patchSize(" + std::to_string(cfg.patch_size) +
                    "). patchSize has to be positive.");
  if (cfg.patch_size < 0) return fail(nullptr, DPX_ERR_UNSUPPORTED, "negative patchSize is undefined behaviour in the reference");
  if (height < 0 || width < 0) return fail(nullptr, DPX_ERR_UNSUPPORTED, "negative image size");
  if (max_batch < 1) return fail(nullptr, DPX_ERR_ARGUMENT, "dpx_create: max_batch must be >= 1");

  Geometry g{};
  g.height = height;
  g.width = width;
  g.nh = width / std::max(cfg.patch_size, 1);
  g.nv = height / std::max(cfg.patch_size, 1);
  g.patch = std::min(cfg.patch_size, std::min(height, width));
  g.n_cells = g.nh * g.nv;
  g.n_points = static_cast<long long>(height) * width;

  Thresholds th{};
  if (g.n_cells > 0) {
    const int p = g.patch;
    if (p < 4 || p > 26)
      return fail(nullptr, DPX_ERR_UNSUPPORTED,
                  "patchSize " + std::to_string(p) + " is outside the supported range [4, 26] (Eigen's product kernels switch "
                  "summation order outside it; see DESIGN.md)");
    if (width % p != 0 || height % p != 0)
      return fail(nullptr, DPX_ERR_UNSUPPORTED,
                  "image size " + std::to_string(height) + " x " + std::to_string(width) + " is not divisible by patchSize " +
                      std::to_string(p) + " (out-of-bounds reads in the reference, cell_grid.cpp:71)");
    if (cfg.min_pts_per_cell == 0)
      return fail(nullptr, DPX_ERR_UNSUPPORTED, "minPtsPerCell == 0 divides by zero in the reference (cell_segment.cpp:23)");
    if (cfg.histogram_bins_per_coord < 1 || cfg.histogram_bins_per_coord > 181)
      return fail(nullptr, DPX_ERR_UNSUPPORTED, "histogramBinsPerCoord must be in [1, 181]");
    const long long mca = cfg.min_region_growing_cells_activated;
    g.plane_cap = static_cast<int>(mca >= 1 ? g.n_cells / mca : g.n_cells) + 1;
    if (g.plane_cap > 65535) return fail(nullptr, DPX_ERR_UNSUPPORTED, "more than 65535 possible plane segments per frame");
    if (g.n_cells > 131071) return fail(nullptr, DPX_ERR_UNSUPPORTED, "more than 131071 cells per frame");
    // size_t valid_pts_threshold = cell_points.size() / config.min_pts_per_cell  (signed division, then cast)
    th.valid_pts_threshold = static_cast<unsigned long long>(static_cast<long long>(3LL * p * p) / static_cast<long long>(cfg.min_pts_per_cell));
  } else {
    g.plane_cap = 1;
  }
  th.min_cos_angle_merge = cfg.min_cos_angle_merge;
  th.max_merge_dist = cfg.max_merge_dist;
  th.min_region_planarity_score = cfg.min_region_planarity_score;
  th.depth_sigma_coeff = cfg.depth_sigma_coeff;
  th.depth_sigma_margin = cfg.depth_sigma_margin;
  th.depth_discontinuity_threshold = cfg.depth_discontinuity_threshold;
  th.max_number_depth_discontinuity = cfg.max_number_depth_discontinuity;
  th.histogram_bins_per_coord = cfg.histogram_bins_per_coord;
  th.min_candidate_size = static_cast<unsigned long long>(static_cast<long long>(cfg.min_region_growing_candidate_size));
  th.min_cells_activated = static_cast<unsigned long long>(static_cast<long long>(cfg.min_region_growing_cells_activated));

  if (device < 0) {
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDevice (no usable CUDA device: this library has no CPU path)");
  }
  int n_dev = 0;
  {
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
      return fail(nullptr, DPX_ERR_CUDA,
                  std::string("no usable CUDA device (this library has no CPU path): ") + cudaGetErrorString(e));
    if (device >= n_dev) return fail(nullptr, DPX_ERR_ARGUMENT, "dpx_create: device index out of range");
  }
  DeviceGuard guard(device);
  if (!guard.ok) return fail(nullptr, DPX_ERR_CUDA, "cudaSetDevice failed");

  dpx_extractor* ex = new dpx_extractor();
  ex->cfg = cfg;
  ex->geom = g;
  ex->thr = th;
  ex->device = device;
  ex->max_batch = max_batch;
  cudaDeviceProp prop;
  {
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { delete ex; return cuda_fail(nullptr, e, "cudaGetDeviceProperties"); }
  }
  ex->sm_count = prop.multiProcessorCount;
  if (g.n_cells > 0) {
    ex->tile_cells = cell_stats_tile_cells(g.patch, g.nh);
    if (const char* e = std::getenv("DPX_STREAM_WARPS")) { const int w = std::atoi(e); ex->stream_warps = (w == 8 || w == 12) ? w : 16; }
    if (const char* e = std::getenv("DPX_CELL_KERNEL")) ex->force_tile_kernel = std::strcmp(e, "tile") == 0;
    if (const char* e = std::getenv("DPX_FUSE_LABELING")) ex->fuse_labeling = std::atoi(e) != 0;
    if (const char* e = std::getenv("DPX_RNG_COMPAT")) ex->uniform_variant = std::strcmp(e, "libstdc++10") == 0 ? 1 : 0;
    if (const char* e = std::getenv("DPX_LABEL_TRANSPORT"))
      ex->label_transport = std::strcmp(e, "u16") == 0 ? DPX_LABELS_U16 : std::strcmp(e, "i32") == 0 ? DPX_LABELS_I32 : DPX_LABELS_AUTO;
    ex->plan = region_grow_plan(g, th);
    Tables probe{};
    ex->scratch_bytes = carve_tables(g, max_batch, ex->plan.bins_smem != 0, nullptr, &probe);
    cudaError_t e = cudaMalloc(&ex->scratch, ex->scratch_bytes);
    if (e != cudaSuccess) { delete ex; return cuda_fail(nullptr, e, "cudaMalloc(scratch tables)"); }
    carve_tables(g, max_batch, ex->plan.bins_smem != 0, static_cast<char*>(ex->scratch), &ex->tb);
    e = cudaMemset(ex->tb.axis_work, 0, sizeof(int32_t));
    if (e != cudaSuccess) { dpx_destroy(ex); return cuda_fail(nullptr, e, "cudaMemset(axis work list)"); }
  }
  if (cfg.ransac_refinement) {
    uint32_t mt[kMtN];
    mt19937_default_state(mt);
    cudaError_t e = cudaMalloc(&ex->mt_init, sizeof(mt));
    if (e == cudaSuccess) e = cudaMemcpy(ex->mt_init, mt, sizeof(mt), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&ex->refine_work, static_cast<size_t>(max_batch) * 2 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(ex->refine_work, 0, static_cast<size_t>(max_batch) * 2 * sizeof(unsigned long long));
    if (e != cudaSuccess) { dpx_destroy(ex); return cuda_fail(nullptr, e, "cudaMalloc(mt19937 state)"); }
  }
  {
    cudaError_t e = cudaMalloc(&ex->region_prof, sizeof(long long) * kRegionProfSlots * max_batch);
    if (e != cudaSuccess) { dpx_destroy(ex); return cuda_fail(nullptr, e, "cudaMalloc(region profile)"); }
  }
  for (int i = 0; i <= DPX_N_STAGES; ++i) {
    cudaError_t e = cudaEventCreate(&ex->e_stage[i]);
    if (e != cudaSuccess) { dpx_destroy(ex); return cuda_fail(nullptr, e, "cudaEventCreate"); }
  }
  *out = ex;
  return DPX_OK;
}

void dpx_destroy(dpx_extractor* ex) {
  if (!ex) return;
  DeviceGuard guard(ex->device);
  cudaDeviceSynchronize();
  ex->pool.reset();
  for (int i = 0; i < kHostSlots; ++i) {
    if (ex->d_xyz[i]) cudaFree(ex->d_xyz[i]);
    if (ex->d_lab[i]) cudaFree(ex->d_lab[i]);
    if (ex->d_depth[i]) cudaFree(ex->d_depth[i]);
    if (ex->d_nar[i]) cudaFree(ex->d_nar[i]);
    if (ex->h_nar[i]) cudaFreeHost(ex->h_nar[i]);
    if (ex->e_h2d[i]) cudaEventDestroy(ex->e_h2d[i]);
    if (ex->e_run[i]) cudaEventDestroy(ex->e_run[i]);
    if (ex->e_d2h[i]) cudaEventDestroy(ex->e_d2h[i]);
  }
  if (ex->s_h2d) cudaStreamDestroy(ex->s_h2d);
  if (ex->s_d2h) cudaStreamDestroy(ex->s_d2h);
  if (ex->s_run) cudaStreamDestroy(ex->s_run);
  for (int i = 0; i <= DPX_N_STAGES; ++i)
    if (ex->e_stage[i]) cudaEventDestroy(ex->e_stage[i]);
  if (ex->scratch) cudaFree(ex->scratch);
  if (ex->region_prof) cudaFree(ex->region_prof);
  if (ex->mt_init) cudaFree(ex->mt_init);
  if (ex->refine_work) cudaFree(ex->refine_work);
  if (ex->d_conv) cudaFree(ex->d_conv);
  delete ex;
}

dpx_status dpx_get_info(const dpx_extractor* ex, dpx_info* info) {
  if (!ex || !info) return DPX_ERR_ARGUMENT;
  info->height = ex->geom.height;
  info->width = ex->geom.width;
  info->patch_size = ex->geom.patch;
  info->cells_x = ex->geom.nh;
  info->cells_y = ex->geom.nv;
  info->n_cells = ex->geom.n_cells;
  info->plane_capacity = ex->geom.plane_cap;
  info->max_batch = ex->max_batch;
  info->device = ex->device;
  info->sm_count = ex->sm_count;
  info->fused_labeling = (ex->fuse_labeling && region_grow_uses_cta(ex->geom, ex->thr)) ? 1 : 0;
  return DPX_OK;
}

dpx_status dpx_process_batch_device(dpx_extractor* ex, const float* d_xyz, int32_t n_frames, dpx_layout layout,
                                    int32_t* d_labels, void* cuda_stream) {
  if (!ex) return DPX_ERR_ARGUMENT;
  if (n_frames < 0 || n_frames > ex->max_batch)
    return fail(ex, DPX_ERR_ARGUMENT, "n_frames " + std::to_string(n_frames) + " exceeds max_batch " + std::to_string(ex->max_batch));
  if (layout != DPX_LAYOUT_COLMAJOR && layout != DPX_LAYOUT_ROWMAJOR) return fail(ex, DPX_ERR_ARGUMENT, "unknown layout");
  if (n_frames == 0 || ex->geom.n_points == 0) return DPX_OK;
  if (!d_xyz || !d_labels) return fail(ex, DPX_ERR_ARGUMENT, "null device pointer");
  DeviceGuard guard(ex->device);
  if (!guard.ok) return fail(ex, DPX_ERR_CUDA, "cudaSetDevice failed");
  return run_stages(ex, d_xyz, n_frames, layout, d_labels, static_cast<cudaStream_t>(cuda_stream));
}

namespace {
// Host pointers: tapered chunks, H2D / kernels / D2H on three streams, kHostSlots chunks in flight.  With the narrow
// label transport the labels come back as uint16 into pinned staging and are widened into `labels` by the host pool
// while the following chunks are in flight.
dpx_status process_host_impl(dpx_extractor* ex, const void* src, size_t bytes_per_frame, int32_t n_frames, int layout,
                             const dpx_intrinsics* pin, int32_t* labels, uint16_t* labels16 = nullptr) {
  DeviceGuard guard(ex->device);
  if (!guard.ok) return fail(ex, DPX_ERR_CUDA, "cudaSetDevice failed");
  dpx_status st = ensure_host_path(ex);
  if (st != DPX_OK) return st;
  const bool is_depth = layout == kLayoutDepth16;
  const size_t np = static_cast<size_t>(ex->geom.n_points);
  const int chunk = is_depth ? ex->host_chunk_depth : ex->host_chunk_xyz;
  // labels16: the caller takes uint16 labels as they are (no widening, 2 B/pixel over PCIe and in host memory)
  const bool out16 = labels16 != nullptr;
  const bool narrow = !out16 && ex->label_transport == DPX_LABELS_U16 && n_frames > 1;
  st = ensure_label_staging(ex, chunk, narrow || out16, narrow);
  if (st != DPX_OK) return st;
  if (is_depth && !ex->d_depth[0])
    for (int i = 0; i < kHostSlots; ++i) DPX_CUDA(ex, cudaMalloc(&ex->d_depth[i], std::max<size_t>(16, np * sizeof(uint16_t) * chunk)));
  if (!is_depth && !ex->d_xyz[0])
    for (int i = 0; i < kHostSlots; ++i) DPX_CUDA(ex, cudaMalloc(&ex->d_xyz[i], std::max<size_t>(16, np * 3 * sizeof(float) * chunk)));
  auto launch = [&](int slot, int nf) {
    return is_depth ? run_stages(ex, nullptr, nf, layout, ex->d_lab[slot], ex->s_run, ex->d_depth[slot], pin)
                    : run_stages(ex, ex->d_xyz[slot], nf, layout, ex->d_lab[slot], ex->s_run);
  };
  if (n_frames == 1) {
    // latency path (a single process() call): everything on one stream, no cross-stream hand-offs
    void* d_in = is_depth ? static_cast<void*>(ex->d_depth[0]) : static_cast<void*>(ex->d_xyz[0]);
    DPX_CUDA(ex, cudaMemcpyAsync(d_in, src, bytes_per_frame, cudaMemcpyHostToDevice, ex->s_run));
    st = launch(0, 1);
    if (st != DPX_OK) return st;
    if (out16) {
      DPX_CUDA(ex, launch_narrow_labels(ex->d_lab[0], ex->d_nar[0], static_cast<long long>(np), ex->s_run));
      ++ex->launches;
      DPX_CUDA(ex, cudaMemcpyAsync(labels16, ex->d_nar[0], np * sizeof(uint16_t), cudaMemcpyDeviceToHost, ex->s_run));
    } else {
      DPX_CUDA(ex, cudaMemcpyAsync(labels, ex->d_lab[0], np * sizeof(int32_t), cudaMemcpyDeviceToHost, ex->s_run));
    }
    DPX_CUDA(ex, cudaStreamSynchronize(ex->s_run));
    return DPX_OK;
  }
  const std::vector<int> sizes = chunk_schedule(n_frames, chunk);
  const int n_chunks = static_cast<int>(sizes.size());
  std::vector<int> first(n_chunks);
  for (int i = 0, f = 0; i < n_chunks; f += sizes[i], ++i) first[i] = f;
  // host side of a finished chunk (narrow transport only): wait for its D2H, hand the widening to the pool
  auto retire = [&](int i) -> dpx_status {
    const int slot = i % kHostSlots;
    DPX_CUDA(ex, cudaEventSynchronize(ex->e_d2h[slot]));
    ex->widen_ticket[slot] = ex->pool->widen(ex->h_nar[slot], labels + static_cast<size_t>(first[i]) * np, np * sizes[i]);
    return DPX_OK;
  };
  for (int i = 0; i < n_chunks; ++i) {
    const int slot = i % kHostSlots;
    const int nf = sizes[i], f0 = first[i];
    void* d_in = is_depth ? static_cast<void*>(ex->d_depth[slot]) : static_cast<void*>(ex->d_xyz[slot]);
    // H2D into slot: the kernels of the chunk that used this slot kHostSlots rounds ago must be done reading it
    if (i >= kHostSlots) DPX_CUDA(ex, cudaStreamWaitEvent(ex->s_h2d, ex->e_run[slot], 0));
    DPX_CUDA(ex, cudaMemcpyAsync(d_in, static_cast<const char*>(src) + static_cast<size_t>(f0) * bytes_per_frame,
                                 bytes_per_frame * nf, cudaMemcpyHostToDevice, ex->s_h2d));
    DPX_CUDA(ex, cudaEventRecord(ex->e_h2d[slot], ex->s_h2d));
    // kernels: need the input, and the label slot must have been drained
    DPX_CUDA(ex, cudaStreamWaitEvent(ex->s_run, ex->e_h2d[slot], 0));
    if (i >= kHostSlots) DPX_CUDA(ex, cudaStreamWaitEvent(ex->s_run, ex->e_d2h[slot], 0));
    st = launch(slot, nf);
    if (st != DPX_OK) return st;
    ex->last_first = f0;
    if (narrow || out16) {
      DPX_CUDA(ex, launch_narrow_labels(ex->d_lab[slot], ex->d_nar[slot], static_cast<long long>(np) * nf, ex->s_run));
      ++ex->launches;
    }
    DPX_CUDA(ex, cudaEventRecord(ex->e_run[slot], ex->s_run));
    // D2H
    DPX_CUDA(ex, cudaStreamWaitEvent(ex->s_d2h, ex->e_run[slot], 0));
    if (narrow) {
      ex->pool->wait(ex->widen_ticket[slot]);  // the staging buffer's previous contents have been widened
      ex->widen_ticket[slot] = 0;
      DPX_CUDA(ex, cudaMemcpyAsync(ex->h_nar[slot], ex->d_nar[slot], np * sizeof(uint16_t) * nf, cudaMemcpyDeviceToHost, ex->s_d2h));
    } else if (out16) {
      DPX_CUDA(ex, cudaMemcpyAsync(labels16 + static_cast<size_t>(f0) * np, ex->d_nar[slot], np * sizeof(uint16_t) * nf,
                                   cudaMemcpyDeviceToHost, ex->s_d2h));
    } else {
      DPX_CUDA(ex, cudaMemcpyAsync(labels + static_cast<size_t>(f0) * np, ex->d_lab[slot], np * sizeof(int32_t) * nf,
                                   cudaMemcpyDeviceToHost, ex->s_d2h));
    }
    DPX_CUDA(ex, cudaEventRecord(ex->e_d2h[slot], ex->s_d2h));
    if (narrow && i >= kHostLookahead) {
      st = retire(i - kHostLookahead);
      if (st != DPX_OK) return st;
    }
  }
  if (narrow) {
    for (int i = std::max(0, n_chunks - kHostLookahead); i < n_chunks; ++i) {
      st = retire(i);
      if (st != DPX_OK) return st;
    }
    for (int slot = 0; slot < kHostSlots; ++slot) {
      ex->pool->wait(ex->widen_ticket[slot]);
      ex->widen_ticket[slot] = 0;
    }
  }
  DPX_CUDA(ex, cudaStreamSynchronize(ex->s_d2h));
  DPX_CUDA(ex, cudaStreamSynchronize(ex->s_run));
  DPX_CUDA(ex, cudaStreamSynchronize(ex->s_h2d));
  return DPX_OK;
}
}  // namespace

dpx_status dpx_process_batch_host(dpx_extractor* ex, const float* xyz, int32_t n_frames, dpx_layout layout, int32_t* labels) {
  if (!ex) return DPX_ERR_ARGUMENT;
  if (n_frames < 0) return fail(ex, DPX_ERR_ARGUMENT, "negative n_frames");
  if (layout != DPX_LAYOUT_COLMAJOR && layout != DPX_LAYOUT_ROWMAJOR) return fail(ex, DPX_ERR_ARGUMENT, "unknown layout");
  if (n_frames == 0 || ex->geom.n_points == 0) return DPX_OK;
  if (!xyz || !labels) return fail(ex, DPX_ERR_ARGUMENT, "null host pointer");
  return process_host_impl(ex, xyz, static_cast<size_t>(ex->geom.n_points) * 3 * sizeof(float), n_frames, layout, nullptr, labels);
}

dpx_status dpx_process_depth_batch_host(dpx_extractor* ex, const uint16_t* depth, int32_t n_frames, const dpx_intrinsics* k,
                                        int32_t* labels) {
  if (!ex) return DPX_ERR_ARGUMENT;
  if (n_frames < 0) return fail(ex, DPX_ERR_ARGUMENT, "negative n_frames");
  if (n_frames == 0 || ex->geom.n_points == 0) return DPX_OK;
  if (!depth || !labels || !k) return fail(ex, DPX_ERR_ARGUMENT, "null pointer");
  return process_host_impl(ex, depth, static_cast<size_t>(ex->geom.n_points) * sizeof(uint16_t), n_frames, kLayoutDepth16, k, labels);
}

dpx_status dpx_process_batch_host_u16(dpx_extractor* ex, const float* xyz, int32_t n_frames, dpx_layout layout, uint16_t* labels) {
  if (!ex) return DPX_ERR_ARGUMENT;
  if (n_frames < 0) return fail(ex, DPX_ERR_ARGUMENT, "negative n_frames");
  if (layout != DPX_LAYOUT_COLMAJOR && layout != DPX_LAYOUT_ROWMAJOR) return fail(ex, DPX_ERR_ARGUMENT, "unknown layout");
  if (n_frames == 0 || ex->geom.n_points == 0) return DPX_OK;
  if (!xyz || !labels) return fail(ex, DPX_ERR_ARGUMENT, "null host pointer");
  return process_host_impl(ex, xyz, static_cast<size_t>(ex->geom.n_points) * 3 * sizeof(float), n_frames, layout, nullptr, nullptr, labels);
}

dpx_status dpx_process_depth_batch_host_u16(dpx_extractor* ex, const uint16_t* depth, int32_t n_frames, const dpx_intrinsics* k,
                                            uint16_t* labels) {
  if (!ex) return DPX_ERR_ARGUMENT;
  if (n_frames < 0) return fail(ex, DPX_ERR_ARGUMENT, "negative n_frames");
  if (n_frames == 0 || ex->geom.n_points == 0) return DPX_OK;
  if (!depth || !labels || !k) return fail(ex, DPX_ERR_ARGUMENT, "null pointer");
  return process_host_impl(ex, depth, static_cast<size_t>(ex->geom.n_points) * sizeof(uint16_t), n_frames, kLayoutDepth16, k, nullptr, labels);
}

dpx_status dpx_process_depth_batch_device(dpx_extractor* ex, const uint16_t* d_depth, int32_t n_frames, const dpx_intrinsics* k,
                                          int32_t* d_labels, void* cuda_stream) {
  if (!ex) return DPX_ERR_ARGUMENT;
  if (n_frames < 0 || n_frames > ex->max_batch)
    return fail(ex, DPX_ERR_ARGUMENT, "n_frames " + std::to_string(n_frames) + " exceeds max_batch " + std::to_string(ex->max_batch));
  if (n_frames == 0 || ex->geom.n_points == 0) return DPX_OK;
  if (!d_depth || !d_labels || !k) return fail(ex, DPX_ERR_ARGUMENT, "null pointer");
  DeviceGuard guard(ex->device);
  if (!guard.ok) return fail(ex, DPX_ERR_CUDA, "cudaSetDevice failed");
  return run_stages(ex, nullptr, n_frames, kLayoutDepth16, d_labels, static_cast<cudaStream_t>(cuda_stream), d_depth, k);
}

dpx_status dpx_process_host(dpx_extractor* ex, const float* xyz, int64_t n_points, dpx_layout layout, int32_t* labels) {
  if (!ex) return DPX_ERR_ARGUMENT;
  // plane_extractor.cpp:188-194
  if (n_points != ex->geom.n_points)
    return fail(ex, DPX_ERR_RUNTIME,
                "Error! Number of points doesn't match image shape: " + std::to_string(n_points) + " != " +
                    std::to_string(ex->geom.height) + " x " + std::to_string(ex->geom.width));
  return dpx_process_batch_host(ex, xyz, 1, layout, labels);
}

namespace {
// Frame index within the last call -> index into the device tables.  The batched host entry points work through a call
// in chunks, so only the frames of its last chunk are still resident afterwards.
dpx_status resident_frame(dpx_extractor* ex, int32_t frame, int* table_index) {
  const int rel = frame - ex->last_first;
  if (frame < 0 || rel >= ex->last_frames) return fail(ex, DPX_ERR_ARGUMENT, "frame index outside the last batch");
  if (rel < 0)
    return fail(ex, DPX_ERR_ARGUMENT, "frame " + std::to_string(frame) + " of the last host call was processed in an earlier chunk: only frames [" +
                                          std::to_string(ex->last_first) + ", " + std::to_string(ex->last_first + ex->last_frames) +
                                          ") are still resident (use a device-resident call to inspect a whole batch)");
  *table_index = rel;
  return DPX_OK;
}
}  // namespace

dpx_status dpx_get_cells(dpx_extractor* ex, int32_t frame, dpx_cell* out, int32_t capacity) {
  if (!ex || !out) return DPX_ERR_ARGUMENT;
  const int C = ex->geom.n_cells;
  if (dpx_status rs = resident_frame(ex, frame, &frame); rs != DPX_OK) return rs;
  if (capacity < C) return fail(ex, DPX_ERR_ARGUMENT, "dpx_get_cells: capacity < n_cells");
  if (C == 0) return DPX_OK;
  DeviceGuard guard(ex->device);
  DPX_CUDA(ex, cudaDeviceSynchronize());
  std::vector<float4> ra(2 * static_cast<size_t>(C)), rb(3 * static_cast<size_t>(C));
  std::vector<int16_t> bin(C);
  std::vector<uint8_t> flags(C);
  std::vector<int32_t> seg(C), lab(C);
  const size_t f = static_cast<size_t>(frame);
  DPX_CUDA(ex, cudaMemcpy(ra.data(), ex->tb.rec_a + f * C * 2, ra.size() * sizeof(float4), cudaMemcpyDeviceToHost));
  DPX_CUDA(ex, cudaMemcpy(rb.data(), ex->tb.rec_b + f * C * 3, rb.size() * sizeof(float4), cudaMemcpyDeviceToHost));
  DPX_CUDA(ex, cudaMemcpy(bin.data(), ex->tb.bin + f * C, C * sizeof(int16_t), cudaMemcpyDeviceToHost));
  DPX_CUDA(ex, cudaMemcpy(flags.data(), ex->tb.flags + f * C, C, cudaMemcpyDeviceToHost));
  DPX_CUDA(ex, cudaMemcpy(seg.data(), ex->tb.seg_label + f * C, C * sizeof(int32_t), cudaMemcpyDeviceToHost));
  DPX_CUDA(ex, cudaMemcpy(lab.data(), ex->tb.cell_label + f * C, C * sizeof(int32_t), cudaMemcpyDeviceToHost));
  for (int c = 0; c < C; ++c) {
    dpx_cell& o = out[c];
    std::memset(&o, 0, sizeof(o));
    o.valid = (flags[c] & kFlagValid) != 0;
    o.planar = (flags[c] & kFlagPlanar) != 0;
    o.bin = bin[c];
    o.seg_label = seg[c];
    o.final_label = lab[c];
    if (!o.valid) continue;
    const float4 a0 = ra[2 * c], a1 = ra[2 * c + 1], b0 = rb[3 * c], b1 = rb[3 * c + 1], b2 = rb[3 * c + 2];
    o.normal[0] = a0.x; o.normal[1] = a0.y; o.normal[2] = a0.z; o.d = a0.w;
    o.mean[0] = a1.x; o.mean[1] = a1.y; o.mean[2] = a1.z; o.merge_tolerance = a1.w;
    o.sum[0] = b0.x; o.sum[1] = b0.y; o.sum[2] = b0.z;
    o.var[0] = b0.w; o.var[1] = b1.x; o.var[2] = b1.y; o.var[3] = b1.z; o.var[4] = b1.w; o.var[5] = b2.x;
    o.mse = b2.y; o.score = b2.z;
  }
  return DPX_OK;
}

dpx_status dpx_get_planes(dpx_extractor* ex, int32_t frame, dpx_plane* out, int32_t capacity, int32_t* n_planes) {
  if (!ex || !n_planes) return DPX_ERR_ARGUMENT;
  *n_planes = 0;
  if (dpx_status rs = resident_frame(ex, frame, &frame); rs != DPX_OK) return rs;
  if (ex->geom.n_cells == 0) return DPX_OK;
  DeviceGuard guard(ex->device);
  DPX_CUDA(ex, cudaDeviceSynchronize());
  int32_t P = 0;
  DPX_CUDA(ex, cudaMemcpy(&P, ex->tb.n_planes + frame, sizeof(int32_t), cudaMemcpyDeviceToHost));
  *n_planes = P;
  if (!out || P == 0) return DPX_OK;
  const int n = std::min(P, capacity);
  std::vector<float> segs(static_cast<size_t>(n) * kSegFloats);
  std::vector<int32_t> merge(n);
  const size_t f = static_cast<size_t>(frame);
  DPX_CUDA(ex, cudaMemcpy(segs.data(), ex->tb.segs + f * ex->geom.plane_cap * kSegFloats, segs.size() * sizeof(float), cudaMemcpyDeviceToHost));
  DPX_CUDA(ex, cudaMemcpy(merge.data(), ex->tb.merge + f * ex->geom.plane_cap, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; ++i) {
    const float* r = &segs[static_cast<size_t>(i) * kSegFloats];
    dpx_plane& o = out[i];
    std::memcpy(o.normal, r + kSegNormal, 12);
    std::memcpy(o.mean, r + kSegMean, 12);
    o.d = r[kSegD];
    o.mse = r[kSegMse];
    o.score = r[kSegScore];
    std::memcpy(&o.n_points, r + kSegN, 4);
    o.merge_label = merge[i];
  }
  return DPX_OK;
}

dpx_status dpx_get_seed_order(dpx_extractor* ex, int32_t frame, uint64_t* out, int32_t capacity) {
  if (!ex || !out) return DPX_ERR_ARGUMENT;
  const int C = ex->geom.n_cells;
  if (dpx_status rs = resident_frame(ex, frame, &frame); rs != DPX_OK) return rs;
  if (capacity < C) return fail(ex, DPX_ERR_ARGUMENT, "dpx_get_seed_order: capacity < n_cells");
  if (C == 0) return DPX_OK;
  if (region_grow_mode(ex->geom, ex->thr) < 1)
    return fail(ex, DPX_ERR_UNSUPPORTED, "frames this small (or histograms this large) are grown without a sorted seed order");
  DeviceGuard guard(ex->device);
  DPX_CUDA(ex, cudaDeviceSynchronize());
  DPX_CUDA(ex, cudaMemcpy(out, ex->tb.skeys + static_cast<size_t>(frame) * C, static_cast<size_t>(C) * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return DPX_OK;
}

dpx_status dpx_get_refine_work(dpx_extractor* ex, int32_t frame, uint64_t* point_passes, uint64_t* rounds) {
  if (!ex || !point_passes || !rounds) return DPX_ERR_ARGUMENT;
  *point_passes = *rounds = 0;
  if (!ex->refine_work) return fail(ex, DPX_ERR_UNSUPPORTED, "dpx_get_refine_work: the handle was created with ransacRefinement = 0");
  if (dpx_status rs = resident_frame(ex, frame, &frame); rs != DPX_OK) return rs;
  DeviceGuard guard(ex->device);
  DPX_CUDA(ex, cudaDeviceSynchronize());
  unsigned long long w[2] = {0, 0};
  DPX_CUDA(ex, cudaMemcpy(w, ex->refine_work + 2 * static_cast<size_t>(frame), sizeof(w), cudaMemcpyDeviceToHost));
  *point_passes = w[0];
  *rounds = w[1];
  return DPX_OK;
}

dpx_status dpx_set_profiling(dpx_extractor* ex, int32_t enabled) {
  if (!ex) return DPX_ERR_ARGUMENT;
  ex->profiling = enabled != 0;
  ex->stage_valid = false;
  return DPX_OK;
}

dpx_status dpx_get_stage_ms(dpx_extractor* ex, float ms[DPX_N_STAGES]) {
  if (!ex || !ms) return DPX_ERR_ARGUMENT;
  if (!ex->stage_valid) return fail(ex, DPX_ERR_ARGUMENT, "no profiled batch: call dpx_set_profiling(ex, 1) first");
  DeviceGuard guard(ex->device);
  DPX_CUDA(ex, cudaEventSynchronize(ex->e_stage[DPX_N_STAGES]));
  for (int i = 0; i < DPX_N_STAGES; ++i) DPX_CUDA(ex, cudaEventElapsedTime(&ms[i], ex->e_stage[i], ex->e_stage[i + 1]));
  return DPX_OK;
}

dpx_status dpx_get_region_profile(dpx_extractor* ex, int32_t frame, int64_t out[DPX_REGION_PROFILE_SLOTS]) {
  if (!ex || !out) return DPX_ERR_ARGUMENT;
  static_assert(DPX_REGION_PROFILE_SLOTS == kRegionProfSlots, "header and kernel disagree");
  if (!ex->stage_valid) return fail(ex, DPX_ERR_ARGUMENT, "no profiled batch: call dpx_set_profiling(ex, 1) first");
  if (dpx_status rs = resident_frame(ex, frame, &frame); rs != DPX_OK) return rs;
  DeviceGuard guard(ex->device);
  DPX_CUDA(ex, cudaDeviceSynchronize());
  DPX_CUDA(ex, cudaMemcpy(out, ex->region_prof + static_cast<size_t>(frame) * kRegionProfSlots,
                          sizeof(long long) * kRegionProfSlots, cudaMemcpyDeviceToHost));
  return DPX_OK;
}

int64_t dpx_kernel_launches(const dpx_extractor* ex) { return ex ? ex->launches : 0; }

dpx_status dpx_set_rng_compat(dpx_extractor* ex, int32_t mode) {
  if (!ex) return DPX_ERR_ARGUMENT;
  if (mode != DPX_RNG_LIBSTDCXX11 && mode != DPX_RNG_LIBSTDCXX10) return fail(ex, DPX_ERR_ARGUMENT, "unknown rng compatibility mode");
  ex->uniform_variant = mode;
  return DPX_OK;
}

dpx_status dpx_set_label_transport(dpx_extractor* ex, int32_t mode) {
  if (!ex) return DPX_ERR_ARGUMENT;
  if (mode != DPX_LABELS_AUTO && mode != DPX_LABELS_I32 && mode != DPX_LABELS_U16) return fail(ex, DPX_ERR_ARGUMENT, "unknown label transport");
  ex->label_transport = mode;
  return DPX_OK;
}

dpx_status dpx_host_alloc(void** ptr, size_t bytes) {
  if (!ptr) return DPX_ERR_ARGUMENT;
  cudaError_t e = cudaHostAlloc(ptr, bytes, cudaHostAllocDefault);
  if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaHostAlloc");
  return DPX_OK;
}

void dpx_host_free(void* ptr) {
  if (ptr) cudaFreeHost(ptr);
}

}  // extern "C"
