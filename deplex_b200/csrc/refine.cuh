// refine.cuh -- launcher interface of stage 4 (optional RANSAC refinement of the labels).
#pragma once
#include "common.cuh"

namespace dpx {

struct RefineArgs {
  const float* xyz;    // [F] organized clouds, same layout as stage 1 read
  int layout;
  int n_frames;
  int max_iterations;  // ransacMaxIterations
  float threshold;     // ransacThreshold
  float inliers_ratio; // ransacInliersRatio
  int uniform_variant; // std::uniform_int_distribution<int> mapping: 0 = libstdc++ >= 11, 1 = libstdc++ <= 10 (refine.cu)
  const uint32_t* mt_init;  // [624] std::mt19937 default-seeded state (before the first twist)
  Geometry geom;
  Tables tables;       // reads cell_label / n_planes; uses queue and pairs as scratch
  int32_t* labels;     // [F][H*W], updated in place
  unsigned long long* work;  // [F][2] out (may be null): point passes and rounds of the frame (dpx_get_refine_work)
};

constexpr int kMtN = 624;
void mt19937_default_state(uint32_t out[kMtN]);  // host
cudaError_t launch_refine(const RefineArgs& args, cudaStream_t stream);

}  // namespace dpx
