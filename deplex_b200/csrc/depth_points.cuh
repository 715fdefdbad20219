// depth_points.cuh -- raw depth -> organized cloud on the device (fallback of the fused path in cell_stats.cu).
#pragma once
#include "common.cuh"

namespace dpx {
// Writes n_frames row-major [N x 3] clouds: DepthImage::toPointCloud (depth_image.cpp:55-78) per pixel.
cudaError_t launch_depth_to_points(const uint16_t* depth, int n_frames, const Geometry& g, const Pinhole& k, float* xyz,
                                   cudaStream_t stream);
}  // namespace dpx
