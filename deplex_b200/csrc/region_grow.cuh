// region_grow.cuh -- launcher interface of stage 2 (histogram, region growing, merging).
#pragma once
#include "common.cuh"

namespace dpx {

struct RegionArgs {
  int n_frames;
  int bins_in_smem;  // per-cell working bins fit in shared memory (else Tables::bin_work)
  Geometry geom;
  Thresholds thr;
  Tables tables;
};

size_t region_grow_smem_bytes(const Geometry& g, const Thresholds& th, bool bins_in_smem);
cudaError_t launch_region_grow(const RegionArgs& args, cudaStream_t stream);

}  // namespace dpx
