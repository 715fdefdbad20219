/*
 * deplex_b200.h -- C-ABI of libdeplex_b200.so: the B200 (sm_100a) plane-extraction hot path.
 *
 * This is the drop-in boundary for prime-slam/deplex's
 *     deplex::PlaneExtractor(height, width, config).process(organized_pcd) -> labels
 * (reference: cpp/deplex/include/deplex/plane_extractor.h:28-56, implemented by
 * cpp/deplex/src/deplex/plane_extractor.cpp:153-283).  Plain pointers and sizes only: no C++,
 * Eigen, torch or CUDA types cross this boundary.  The C++ class shim
 * (deplex_b200/cpp/deplex/plane_extractor.h) and the pybind module `deplex.pybind`
 * (deplex_b200/pybind/) are both written against exactly these entry points.
 *
 * There is NO CPU fallback: every process call runs CUDA kernels and fails with DPX_ERR_CUDA
 * when no usable device is present.
 */
#ifndef DEPLEX_B200_H
#define DEPLEX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPX_VERSION 200

#if defined(__GNUC__)
#define DPX_API __attribute__((visibility("default")))
#else
#define DPX_API
#endif

typedef enum dpx_status {
  DPX_OK = 0,
  DPX_ERR_RUNTIME = 1,     /* the reference throws std::runtime_error here; dpx_last_error() has its exact text */
  DPX_ERR_UNSUPPORTED = 2, /* input is undefined behaviour in the reference or outside the parity domain */
  DPX_ERR_CUDA = 3,        /* CUDA runtime / driver failure (including "no device") */
  DPX_ERR_ARGUMENT = 4     /* NULL handle, batch larger than max_batch, ... */
} dpx_status;

/* Memory layout of one organized cloud of N = height*width points. */
typedef enum dpx_layout {
  DPX_LAYOUT_COLMAJOR = 0, /* X[N] Y[N] Z[N]: Eigen::MatrixX3f (plane_extractor.h:48) */
  DPX_LAYOUT_ROWMAJOR = 1  /* x0 y0 z0 x1 ...: numpy C-order (N,3); DepthImage::toPointCloud's native order */
} dpx_layout;

/* Field-for-field mirror of deplex::config::Config (cpp/deplex/include/deplex/config.h:51-81). */
typedef struct dpx_config {
  int32_t patch_size;                         /* patchSize */
  int32_t histogram_bins_per_coord;           /* histogramBinsPerCoord */
  float min_cos_angle_merge;                  /* minCosAngleForMerge */
  float max_merge_dist;                       /* maxMergeDist */
  int32_t min_region_growing_candidate_size;  /* minRegionGrowingCandidateSize */
  int32_t min_region_growing_cells_activated; /* minRegionGrowingCellsActivated */
  float min_region_planarity_score;           /* minRegionPlanarityScore */
  float depth_sigma_coeff;                    /* depthSigmaCoeff */
  float depth_sigma_margin;                   /* depthSigmaMargin */
  int32_t min_pts_per_cell;                   /* minPtsPerCell */
  float depth_discontinuity_threshold;        /* depthDiscontinuityThreshold */
  int32_t max_number_depth_discontinuity;     /* maxNumberDepthDiscontinuity */
  int32_t ransac_refinement;                  /* ransacRefinement (bool) */
  int32_t ransac_max_iterations;              /* ransacMaxIterations */
  float ransac_threshold;                     /* ransacThreshold */
  float ransac_inliers_ratio;                 /* ransacInliersRatio */
} dpx_config;

/* Pinhole intrinsics [[fx, 0, cx], [0, fy, cy], [0, 0, 1]] as DepthImage::toPointCloud reads them
 * (cpp/deplex/src/deplex/utils/depth_image.cpp:58-61). */
typedef struct dpx_intrinsics {
  float fx, fy, cx, cy;
} dpx_intrinsics;

typedef struct dpx_extractor dpx_extractor; /* opaque: PlaneExtractor::Impl's replacement */

/* Geometry and capacities fixed at creation. */
typedef struct dpx_info {
  int32_t height, width;
  int32_t patch_size;      /* after the reference's clamp min(patch, min(h, w)) (plane_extractor.cpp:160) */
  int32_t cells_x, cells_y;/* nr_horizontal_cells_, nr_vertical_cells_ (plane_extractor.cpp:155-156) */
  int32_t n_cells;
  int32_t plane_capacity;  /* upper bound on plane segments per frame */
  int32_t max_batch;
  int32_t device;
  int32_t sm_count;
  int32_t fused_labeling;  /* 1: the region-growing kernel also paints the pixels of the frames that finish early (all but
                              max(2, n_frames / 16) per batch); stage DPX_STAGE_LABELING then covers only the rest */
} dpx_info;

/* Per-cell record returned by dpx_get_cells (one per cell of one frame of the last batch). */
typedef struct dpx_cell {
  float sum[3];     /* CellSegmentStat::coord_sum_ */
  float var[6];     /* upper triangle of variance_ (X^T X): xx xy xz yy yz zz */
  float mean[3];
  float normal[3];
  float d, mse, score;
  float merge_tolerance;
  int32_t bin;      /* initial NormalsHistogram bin, -1 if the cell is not planar */
  int32_t valid;    /* passed hasValidPoints && isDepthContinuous (cell_segment.cpp:27-30) */
  int32_t planar;   /* CellSegment::isPlanar */
  int32_t seg_label;/* labels_map_ entry after region growing (1-based segment id, 0 = none) */
  int32_t final_label; /* label painted by toImageLabels for this cell */
} dpx_cell;

/* Per-plane-segment record returned by dpx_get_planes (post-merge statistics, like the reference's
 * plane_segments vector at the end of findMergedLabels). */
typedef struct dpx_plane {
  float normal[3];
  float d;
  float mean[3];
  float mse, score;
  int32_t n_points;
  int32_t merge_label; /* plane_merge_labels[i] (plane_extractor.cpp:398-425) */
} dpx_plane;

/* Stage indices for dpx_get_stage_ms. */
enum { DPX_STAGE_CELL_STATS = 0, DPX_STAGE_REGION_GROW = 1, DPX_STAGE_LABELING = 2, DPX_STAGE_REFINE = 3, DPX_N_STAGES = 4 };

/* ---- Config: replaces Config::Config() and Config::Config(std::string const&) (config.cpp:23-80) ---- */
DPX_API void dpx_config_default(dpx_config* cfg);
/* Same parsing rules as config.cpp:33-79.  DPX_ERR_RUNTIME + "Couldn't open ini file: <path>" when unreadable. */
DPX_API dpx_status dpx_config_load_ini(const char* path, dpx_config* cfg);

/* ---- Extractor: replaces PlaneExtractor::PlaneExtractor / ~PlaneExtractor (plane_extractor.cpp:153-183) ---- */
/* device < 0 selects the current CUDA device.  max_batch >= 1 frames of device scratch are allocated. */
DPX_API dpx_status dpx_create(int32_t height, int32_t width, const dpx_config* cfg, int32_t device, int32_t max_batch,
                      dpx_extractor** out);
DPX_API void dpx_destroy(dpx_extractor* ex);
/* Message of the last failure on this handle; ex == NULL returns the calling thread's last handle-less failure
 * (dpx_create, dpx_config_load_ini). */
DPX_API const char* dpx_last_error(const dpx_extractor* ex);
DPX_API dpx_status dpx_get_info(const dpx_extractor* ex, dpx_info* info);

/* ---- process: replaces PlaneExtractor::process (plane_extractor.cpp:185-283) ---- */
/* One frame, host pointers (pageable or pinned).  n_points is pcd_array.rows(); a mismatch with height*width
 * returns DPX_ERR_RUNTIME with the reference's message (plane_extractor.cpp:188-194).  labels: n_points int32. */
DPX_API dpx_status dpx_process_host(dpx_extractor* ex, const float* xyz, int64_t n_points, dpx_layout layout, int32_t* labels);
/* n_frames independent frames, host pointers, frame-major; copies are pipelined with the kernels in chunks. */
DPX_API dpx_status dpx_process_batch_host(dpx_extractor* ex, const float* xyz, int32_t n_frames, dpx_layout layout,
                                  int32_t* labels);
/* n_frames <= max_batch frames already resident in device memory; asynchronous on `cuda_stream` (a cudaStream_t,
 * NULL = the legacy default stream).  d_labels: n_frames * height * width int32 in device memory. */
DPX_API dpx_status dpx_process_batch_device(dpx_extractor* ex, const float* d_xyz, int32_t n_frames, dpx_layout layout,
                                    int32_t* d_labels, void* cuda_stream);

/* ---- raw depth in: replaces DepthImage::toPointCloud (depth_image.cpp:55-78) followed by process() ----
 * depth: n_frames * height * width uint16 samples (row-major images, raw sensor units; 0 = no measurement).
 * The points z = float(raw), x = ((col - cx) * z) / fx, y = ((row - cy) * z) / fy are generated on the device,
 * fused into the first kernel where the geometry allows, so 2 bytes per pixel cross PCIe / HBM instead of 12.
 * Labels are identical to process() on the cloud toPointCloud would have produced. */
DPX_API dpx_status dpx_process_depth_batch_host(dpx_extractor* ex, const uint16_t* depth, int32_t n_frames,
                                        const dpx_intrinsics* k, int32_t* labels);
DPX_API dpx_status dpx_process_depth_batch_device(dpx_extractor* ex, const uint16_t* d_depth, int32_t n_frames,
                                          const dpx_intrinsics* k, int32_t* d_labels, void* cuda_stream);

/* ---- introspection of the last batch (the reference computes these and discards them) ----
 * `frame` counts within the last process call.  The batched HOST entry points work through a call in chunks, so after
 * them only the frames of the last chunk are still resident (others: DPX_ERR_ARGUMENT naming the resident range);
 * a device-resident call keeps all of its frames. */
DPX_API dpx_status dpx_get_cells(dpx_extractor* ex, int32_t frame, dpx_cell* out, int32_t capacity);
DPX_API dpx_status dpx_get_planes(dpx_extractor* ex, int32_t frame, dpx_plane* out, int32_t capacity, int32_t* n_planes);
/* The seed order of one frame (n_cells entries): its cells sorted by the key [bin : 15][MSE as order-preserving
 * uint32 : 32][cell id : 17]; cells that are not planar carry the largest bin and MSE fields and come last.  createPlaneSegments' seed of a
 * bin (first strict-minimum MSE, plane_extractor.cpp:309-316) is the first entry of the bin's run that is still unassigned. */
DPX_API dpx_status dpx_get_seed_order(dpx_extractor* ex, int32_t frame, uint64_t* out, int32_t capacity);

/* Work of the refinement stage (ransacRefinement=1) on one frame of the last batch: point_passes = sum over the frame's
 * labels of (points of the label) x (rounds of 128 hypotheses scored on them) -- each point pass reads 12 bytes and evaluates
 * 128 hypotheses; rounds = the rounds themselves.  Replaces nothing in the reference (plane_extractor.cpp:472-509 does not
 * count); it is what bench.py's roofline of the stage is computed from. */
DPX_API dpx_status dpx_get_refine_work(dpx_extractor* ex, int32_t frame, uint64_t* point_passes, uint64_t* rounds);

/* ---- measurement: CUDA-event time of each stage of the last dpx_process_batch_device call ---- */
DPX_API dpx_status dpx_set_profiling(dpx_extractor* ex, int32_t enabled);
DPX_API dpx_status dpx_get_stage_ms(dpx_extractor* ex, float ms[DPX_N_STAGES]);
/* SM-cycle counters of the region-growing stage for one frame of the last profiled batch:
 * [0] whole frame, [1] histogram + grouping, [2] seed selection, [3] BFS, [4] moment accumulation (0 when the helper
 * warps do it concurrently), [5] plane fits, [6] adjacency + merging, [7] final labels (+ fused pixel painting),
 * [8] #seeds, [9] #BFS steps, [10] #regions fitted, [11] #segments. */
#define DPX_REGION_PROFILE_SLOTS 12
DPX_API dpx_status dpx_get_region_profile(dpx_extractor* ex, int32_t frame, int64_t out[DPX_REGION_PROFILE_SLOTS]);
/* Number of kernels launched by this handle since creation. */
DPX_API int64_t dpx_kernel_launches(const dpx_extractor* ex);

/* ---- which standard library the RANSAC refinement's sampling reproduces ----
 * refineLabels draws its 3-point samples with std::uniform_int_distribution<int> over a default-seeded std::mt19937
 * (libs/rtl/include/rtl/RANSAC.hpp:81-87,107-111).  The generator is fixed by the C++ standard; the distribution's
 * mapping is not, and libstdc++ changed it in GCC 11 (Lemire multiply-shift; before: scaling + rejection).  The refined
 * labels depend on it.  Default: DPX_RNG_LIBSTDCXX11, what a reference built with GCC >= 11 produces (and what the oracle
 * is checked against on this toolchain); DPX_RNG_LIBSTDCXX10 reproduces binaries built with GCC <= 10 (e.g. manylinux2014
 * wheels, .github/workflows/wheels.yml:10).  Environment DPX_RNG_COMPAT=libstdc++10 sets the default of new handles. */
enum { DPX_RNG_LIBSTDCXX11 = 0, DPX_RNG_LIBSTDCXX10 = 1 };
DPX_API dpx_status dpx_set_rng_compat(dpx_extractor* ex, int32_t mode);

/* ---- narrow labels on the host-pointer path ----
 * The reference returns int32 labels (Eigen::VectorXi, plane_extractor.h:48) and so do the entry points above.  Labels
 * never exceed plane_capacity <= 65535, and on the raw-depth path the int32 labels are two thirds of all PCIe bytes
 * (4 of 6 B/pixel), so two narrower forms exist:
 *  - dpx_process_*_host_u16: the caller takes the labels as uint16 (same values).  2 B/pixel over PCIe AND in host
 *    memory: this is the form that lifts the ceiling (640x480 raw depth: ~76 k frames/s copy-bound instead of ~45 k).
 *  - dpx_set_label_transport(ex, DPX_LABELS_U16): the int32 entry points bring the labels over PCIe as uint16 and widen
 *    them into the caller's int32 buffer with a few host threads while the next chunks are in flight.  This halves the
 *    D2H bytes but adds 6 B/pixel of host memory traffic, so it pays only where the PCIe link, not host memory bandwidth,
 *    is the limit (measured on the B200 box: no gain, profiles/r02_e2e_probe.txt); off by default (DPX_LABELS_AUTO ==
 *    DPX_LABELS_I32).  Environment DPX_LABEL_TRANSPORT=i32|u16 sets the default of new handles, DPX_HOST_THREADS the
 *    number of widening threads. */
enum { DPX_LABELS_AUTO = 0, DPX_LABELS_I32 = 1, DPX_LABELS_U16 = 2 };
DPX_API dpx_status dpx_set_label_transport(dpx_extractor* ex, int32_t mode);
DPX_API dpx_status dpx_process_batch_host_u16(dpx_extractor* ex, const float* xyz, int32_t n_frames, dpx_layout layout,
                                              uint16_t* labels);
DPX_API dpx_status dpx_process_depth_batch_host_u16(dpx_extractor* ex, const uint16_t* depth, int32_t n_frames,
                                                    const dpx_intrinsics* k, uint16_t* labels);

/* ---- batches in flight on one GPU: replaces the caller's loop over process() ----
 * (examples/process_sequence.cpp:30-43 calls process() frame after frame; here the unit is a device-resident batch.)
 * A pipeline owns `lanes` extractors, each with its own device tables and CUDA stream, and deals the submitted batches
 * to them round-robin: within one batch the HBM-bound cell-statistics kernel and the latency-bound region growing run
 * back to back, and with several batches in flight the next batch's first kernel fills the SMs the previous batch's
 * region growing has already left.  Results do not depend on `lanes`.
 * submit: asynchronous; the lane first waits (on the device) for the work queued so far on `producer_stream`
 *   (a cudaStream_t, NULL = legacy default stream), i.e. for whatever produced d_xyz.  n_frames <= max_batch.
 * join: orders `consumer_stream` behind every batch submitted so far (device-side wait, no host synchronisation).
 * synchronize: join + host wait. */
typedef struct dpx_pipeline dpx_pipeline;
DPX_API dpx_status dpx_pipeline_create(int32_t height, int32_t width, const dpx_config* cfg, int32_t device, int32_t max_batch,
                                       int32_t lanes, dpx_pipeline** out);
DPX_API void dpx_pipeline_destroy(dpx_pipeline* p);
DPX_API const char* dpx_pipeline_last_error(const dpx_pipeline* p);
DPX_API int32_t dpx_pipeline_lanes(const dpx_pipeline* p);
/* lane's extractor, for dpx_get_info / dpx_get_planes / profiling; owned by the pipeline */
DPX_API dpx_extractor* dpx_pipeline_lane(dpx_pipeline* p, int32_t lane);
DPX_API dpx_status dpx_pipeline_submit_device(dpx_pipeline* p, const float* d_xyz, int32_t n_frames, dpx_layout layout,
                                              int32_t* d_labels, void* producer_stream);
DPX_API dpx_status dpx_pipeline_submit_depth_device(dpx_pipeline* p, const uint16_t* d_depth, int32_t n_frames,
                                                    const dpx_intrinsics* k, int32_t* d_labels, void* producer_stream);
DPX_API dpx_status dpx_pipeline_join(dpx_pipeline* p, void* consumer_stream);
DPX_API dpx_status dpx_pipeline_synchronize(dpx_pipeline* p);
DPX_API int64_t dpx_pipeline_kernel_launches(const dpx_pipeline* p);

/* ---- a frame sequence sharded over the GPUs of one box: replaces the loop of examples/process_sequence.cpp:30-43 ----
 * Frames are independent (process() keeps no state across calls, plane_extractor.cpp:281,428), so a sequence of
 * n_frames host frames is split into contiguous ranges, one per device (device g of G gets frames
 * [g*n/G, (g+1)*n/G), remainder to the first devices), and one worker thread per device runs the batched host entry
 * point on its range.  No inter-GPU traffic at all: every device copies its labels into its slice of the caller's
 * buffer.  n_devices <= 0 or devices == NULL: all visible devices.  Labels are identical to calling process() on
 * every frame in order. */
typedef struct dpx_sequence dpx_sequence;
DPX_API dpx_status dpx_sequence_create(int32_t height, int32_t width, const dpx_config* cfg, const int32_t* devices,
                                       int32_t n_devices, int32_t max_batch, dpx_sequence** out);
DPX_API void dpx_sequence_destroy(dpx_sequence* s);
DPX_API const char* dpx_sequence_last_error(const dpx_sequence* s);
DPX_API int32_t dpx_sequence_devices(const dpx_sequence* s);
/* [begin, end) of device slot `slot` for a sequence of n_frames */
DPX_API void dpx_sequence_range(const dpx_sequence* s, int64_t n_frames, int32_t slot, int64_t* begin, int64_t* end);
DPX_API dpx_status dpx_sequence_process_host(dpx_sequence* s, const float* xyz, int64_t n_frames, dpx_layout layout,
                                             int32_t* labels);
DPX_API dpx_status dpx_sequence_process_depth_host(dpx_sequence* s, const uint16_t* depth, int64_t n_frames,
                                                   const dpx_intrinsics* k, int32_t* labels);
/* Device-resident form, for measurements and for callers whose frames are produced on the GPUs: every device processes
 * `n_frames_per_device` frames that are already in ITS memory (d_xyz[slot], d_labels[slot]: arrays of n_devices device
 * pointers), max_batch frames per submit through a `lanes`-lane pipeline; returns when all devices are done.
 * ms_per_device (optional, n_devices floats): CUDA-event time of each device's range. */
DPX_API dpx_status dpx_sequence_process_device(dpx_sequence* s, const float* const* d_xyz, int64_t n_frames_per_device,
                                               dpx_layout layout, int32_t* const* d_labels, int32_t lanes, float* ms_per_device);

/* ---- device memory helpers for callers without the CUDA headers (cudaMalloc / cudaFree / cudaMemcpy) ---- */
DPX_API int32_t dpx_device_count(void);
DPX_API dpx_status dpx_device_alloc(int32_t device, void** ptr, size_t bytes);
DPX_API void dpx_device_free(int32_t device, void* ptr);
DPX_API dpx_status dpx_memcpy_to_device(int32_t device, void* dst, const void* src, size_t bytes);
DPX_API dpx_status dpx_memcpy_to_host(int32_t device, void* dst, const void* src, size_t bytes);

/* ---- pinned host memory helpers (cudaHostAlloc / cudaFreeHost) ---- */
DPX_API dpx_status dpx_host_alloc(void** ptr, size_t bytes);
DPX_API void dpx_host_free(void* ptr);

DPX_API int32_t dpx_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DEPLEX_B200_H */
