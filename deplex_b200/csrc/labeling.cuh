// labeling.cuh -- launcher interface of stage 3 (per-pixel labels).
#pragma once
#include "common.cuh"

namespace dpx {

struct LabelArgs {
  int n_frames;
  Geometry geom;
  const int32_t* cell_label;  // [F][C]
  int32_t* labels;            // [F][H*W]
  const int32_t* todo;        // optional {count, frame ids...}: only these frames are painted (the region-growing
                              // kernel has painted the others); nullptr = all frames
};

cudaError_t launch_labeling(const LabelArgs& args, cudaStream_t stream);
// frames the fused painting of region_grow_cta_kernel leaves to stage 3: the slowest sixteenth by finishing order
__host__ __device__ inline int labeling_deferred_frames(int n_frames) { const int d = n_frames / 16 > 2 ? n_frames / 16 : 2; return d < n_frames ? d : n_frames; }

}  // namespace dpx
