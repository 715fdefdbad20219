// common.cuh -- device-side data layout shared by the three stages of the hot path.
//
// HBM layout (all arrays are frame-major, sized for max_batch frames at dpx_create):
//   rec_a   float4[F][C][2]  {nx,ny,nz,d} {mean_x,mean_y,mean_z,merge_tolerance}   32 B/cell  (read by the BFS)
//   rec_b   float4[F][C][3]  {Sx,Sy,Sz,Vxx} {Vxy,Vxz,Vyy,Vyz} {Vzz,mse,score,0}    48 B/cell  (read by the accumulation)
//   bin     int16 [F][C]     initial NormalsHistogram bin, -1 when the cell is not planar
//   flags   uint8 [F][C]     bit0 = valid (stats exist), bit1 = planar
//   mse     float [F][C]     per-cell MSE again, dense, for the seed scan (written for valid cells only)
//   seg_label / cell_label int32[F][C], queue int32[F][C], pairs uint32[F][2C], segs float[F][Pcap][24],
//   merge int32[F][Pcap], n_planes int32[F]
// rec_a / rec_b are only written for valid cells; nothing reads them for the others.
#pragma once
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

// Bounds / invariant checks of the debug build (make -C deplex_b200/csrc debug -> libdeplex_b200_dbg.so, loaded through
// DPX_LIB_PATH).  compute-sanitizer is not available on the B200 pool this was developed on, so the kernels with
// data-dependent indexing (region growing, seed sort, labeling) carry their own checks: a violated one prints the
// expression and traps, which surfaces as a CUDA error on the host.  Compiled out of the product build.
#ifdef DPX_DEBUG_CHECKS
#include <cstdio>
#define DPX_CHECK(cond)                                                                                   \
  do {                                                                                                    \
    if (!(cond)) {                                                                                        \
      printf("DPX_CHECK failed: %s (%s:%d) block %d thread %d\n", #cond, __FILE__, __LINE__, blockIdx.x, threadIdx.x); \
      __trap();                                                                                           \
    }                                                                                                     \
  } while (0)
#else
#define DPX_CHECK(cond) do { } while (0)
#endif

namespace dpx {

constexpr int kLayoutColMajor = 0;
constexpr int kLayoutRowMajor = 1;
constexpr int kLayoutDepth16 = 2;  // raw uint16 depth + pinhole intrinsics: points are generated in registers

// DepthImage::toPointCloud's inputs (depth_image.cpp:58-61)
struct Pinhole {
  float fx, fy, cx, cy;
};

constexpr uint8_t kFlagValid = 1;
constexpr uint8_t kFlagPlanar = 2;

constexpr int kSegFloats = 24;  // n(int) S3 V6 mean3 normal3 d mse score + pad
// offsets inside one segment record
constexpr int kSegN = 0, kSegS = 1, kSegV = 4, kSegMean = 10, kSegNormal = 13, kSegD = 16, kSegMse = 17, kSegScore = 18;
// while regions are being grown, the record also remembers where the region's cells sit in the cell list
constexpr int kSegOff = 19, kSegCnt = 20;

struct Geometry {
  int height, width;
  int patch;        // clamped patch size
  int nh, nv;       // cells per row / column
  int n_cells;
  int plane_cap;
  long long n_points;
};

struct Thresholds {
  // config values exactly as the reference holds them (fp32 / int32)
  float min_cos_angle_merge, max_merge_dist, min_region_planarity_score;
  float depth_sigma_coeff, depth_sigma_margin, depth_discontinuity_threshold;
  int max_number_depth_discontinuity;
  int histogram_bins_per_coord;
  unsigned long long valid_pts_threshold;       // size_t(cell_points.size() / minPtsPerCell)
  unsigned long long min_candidate_size;        // size_t(int32) conversions of the reference's comparisons
  unsigned long long min_cells_activated;
};

struct Tables {
  float4* rec_a;
  float4* rec_b;
  int16_t* bin;
  uint8_t* flags;
  float* mse;          // [F][C] dense copy of the per-cell MSE (seed selection scans it)
  uint8_t* edge;       // [F][C] precomputed growSeed edge tests, bit s = neighbour slot s (up, down, left, right)
  int32_t* seg_label;
  int32_t* cell_label;
  int32_t* queue;
  uint32_t* pairs;     // [F][2C] packed (min<<16 | max) adjacent plane pairs; scratch of the seed sort and of region growing
  unsigned long long* skeys;  // [F][C] cells sorted by (histogram bin, MSE, cell id): seed_sort.cuh
  int16_t* bin_work;   // used by region growing when the bins do not fit in shared memory
  uint32_t* cell_words; // [F][C] region-growing cell words when a frame is too large for shared memory
  int32_t* paint_state; // [2 + F] finished-frame counter, number of frames left to stage 3, then the list of those frames
  float* segs;
  int32_t* merge;
  int32_t* n_planes;
  int32_t* axis_work;  // [1 + kAxisWorkCap] per batch: count, then the cells whose bin needs the careful path (region_grow.cu)
};
constexpr int kAxisWorkCap = 65535;

// ---- programmatic dependent launch --------------------------------------------------------------------------------
// The stages of a batch are short kernels that depend on each other; between two of them the GPU drains one grid and
// dispatches the next, a few microseconds each time.  A kernel launched with the programmatic-serialization attribute is
// dispatched while its predecessor in the stream is still running and waits, at `pdl_wait()` (its first instruction
// here), until that grid has completed and its memory is visible -- the ordering is the same, the dispatch is off the
// critical path.  DPX_PDL=0 / 1 overrides the default.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
constexpr bool kPdlDefault = true;
inline bool pdl_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("DPX_PDL");
    return (e && *e) ? std::atoi(e) != 0 : kPdlDefault;
  }();
  return on;
}
#ifdef __CUDACC__
template <class... KArgs, class... Args>
inline cudaError_t launch_dependent(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

}  // namespace dpx
