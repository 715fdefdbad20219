// labeling.cuh -- launcher interface of stage 3 (per-pixel labels).
#pragma once
#include "common.cuh"

namespace dpx {

struct LabelArgs {
  int n_frames;
  Geometry geom;
  const int32_t* cell_label;  // [F][C]
  int32_t* labels;            // [F][H*W]
};

cudaError_t launch_labeling(const LabelArgs& args, cudaStream_t stream);

}  // namespace dpx
