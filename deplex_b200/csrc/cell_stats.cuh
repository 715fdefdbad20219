// cell_stats.cuh -- launcher interface of stage 1 (cell statistics).
#pragma once
#include "common.cuh"

namespace dpx {

constexpr int kCellStatsThreads = 128;

struct CellStatsArgs {
  const float* xyz;  // device, n_frames organized clouds (layouts 0/1)
  const uint16_t* depth;  // device, n_frames raw depth images (layout kLayoutDepth16)
  Pinhole pin;            // intrinsics for kLayoutDepth16
  int n_frames;
  int layout;
  int tile_cells;       // cells staged per CTA (from cell_stats_tile_cells)
  int tiles_per_strip;  // filled by the launcher
  int vec_ok;           // base pointer and plane strides allow 16-byte loads
  int sm_count;         // persistent grid size of the streaming kernel
  int force_tile_kernel;  // diagnostics: always take the fallback kernel
  int stream_warps;     // warps per persistent CTA of the streaming kernel (8, 12 or 16)
  Geometry geom;
  Thresholds thr;
  Tables tables;
};

int cell_stats_tile_cells(int patch, int nh);
cudaError_t launch_cell_stats(const CellStatsArgs& args, cudaStream_t stream);
// whether the fused depth -> points path of the streaming kernel can take this geometry / pointer
bool cell_stats_depth_eligible(const CellStatsArgs& args);

}  // namespace dpx
