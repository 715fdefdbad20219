// region_grow.cuh -- launcher interface of stage 2 (histogram, region growing, merging).
#pragma once
#include "common.cuh"

namespace dpx {

// Where the per-frame working set lives (shared memory when it fits, else the global scratch tables).
struct RegionPlan {
  int bins_smem, list_smem, members_smem, merge_smem;
  int off_hist, off_binoff, off_cursor, off_rowbits, off_bins, off_edge, off_list, off_members, off_msem, off_merge;
  size_t bytes;  // dynamic shared memory per CTA
};

constexpr int kRegionProfSlots = 12;

struct RegionArgs {
  int n_frames;
  long long* prof;  // optional [F][kRegionProfSlots] per-phase cycle counters, or nullptr
  int32_t* labels;  // optional [F][H*W]: when set and the CTA kernel runs, it also paints the per-pixel labels (stage 3 fused)
  int labels_vec_ok;  // 16-byte stores into `labels` allowed (labeling.cuh: labels_vec_ok)
  RegionPlan plan;
  Geometry geom;
  Thresholds thr;
  Tables tables;
};

RegionPlan region_grow_plan(const Geometry& g, const Thresholds& th);
// whether launch_region_grow takes the CTA kernel (which can paint the pixel labels itself) for this geometry
bool region_grow_uses_cta(const Geometry& g, const Thresholds& th);
// storage mode of the CTA kernel for this geometry (0: all shared memory, no seed sort; 1 .. 3: sorted seeds), -1 = generic kernel
int region_grow_mode(const Geometry& g, const Thresholds& th);
// *painted (optional) tells the caller whether the per-pixel labels were written, i.e. whether stage 3 is still needed
cudaError_t launch_region_grow(const RegionArgs& args, cudaStream_t stream, bool* painted = nullptr);

}  // namespace dpx
