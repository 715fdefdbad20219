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
  int vec_ok;                 // 16-byte stores allowed: width % 4 == 0 and `labels` 16-byte aligned (labels_vec_ok)
};

// whether the painters may use st.global.v4 on this label buffer (every row then starts on a 16-byte boundary)
inline bool labels_vec_ok(const Geometry& g, const void* labels) {
  return (g.width & 3) == 0 && (reinterpret_cast<uintptr_t>(labels) & 15u) == 0;
}

// int32 -> uint16 copy of n labels (the narrow label transport of the host-pointer path); n % 8 == 0 and both pointers
// 16-byte aligned take the vector path
cudaError_t launch_narrow_labels(const int32_t* labels, uint16_t* out, long long n, cudaStream_t stream);

cudaError_t launch_labeling(const LabelArgs& args, cudaStream_t stream);
// frames the fused painting of region_grow_cta_kernel leaves to stage 3: the slowest sixteenth by finishing order
__host__ __device__ inline int labeling_deferred_frames(int n_frames) { const int d = n_frames / 16 > 2 ? n_frames / 16 : 2; return d < n_frames ? d : n_frames; }

}  // namespace dpx
