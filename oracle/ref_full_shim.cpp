// ref_full_shim.cpp -- test infrastructure: a C face on the UNMODIFIED reference build, for pinning the oracle.
//
// Compiled only by `make -C oracle ref_full EIGEN3_INCLUDE_DIR=<dir holding Eigen/Core>` together with the reference's
// own sources where they lie under $(REF) (nothing is copied): plane_extractor.cpp, cell_grid.cpp, cell_segment.cpp,
// cell_segment_stat.cpp, normals_histogram.cpp, config.cpp, libs/dsyev/src/*.c, and the header-only libs/rtl, with the
// reference's Release flags (-O3 -DNDEBUG, C++14, no -march).  Eigen 3.4 is not vendored by the reference
// (external/eigen3/CMakeLists.txt:1-16 fetches it) and is absent from this image, so the target is a no-op until an
// Eigen tree is supplied; tests/test_oracle_cpu.py::test_oracle_matches_reference_build then compares labels and the
// raw coord_sum_ / variance_ bits of every cell with the oracle's.
#include <cstdint>
#include <cstring>
#include <exception>
#include <string>

#include <Eigen/Core>

#include "deplex/config.h"
#include "deplex/plane_extractor.h"
// the per-cell statistics have no getters for coord_sum_ / variance_ (cell_segment_stat.h): open the class for the dump
#define private public
#include "cell_grid.h"
#undef private

namespace {
struct ShimConfig {  // == dpxo_config == dpx_config, field for field (deplex/config.h:51-81)
  int32_t patch_size, histogram_bins_per_coord;
  float min_cos_angle_merge, max_merge_dist;
  int32_t min_region_growing_candidate_size, min_region_growing_cells_activated;
  float min_region_planarity_score, depth_sigma_coeff, depth_sigma_margin;
  int32_t min_pts_per_cell;
  float depth_discontinuity_threshold;
  int32_t max_number_depth_discontinuity, ransac_refinement, ransac_max_iterations;
  float ransac_threshold, ransac_inliers_ratio;
};

deplex::config::Config to_config(const ShimConfig& c) {
  deplex::config::Config o;
  o.patch_size = c.patch_size;
  o.histogram_bins_per_coord = c.histogram_bins_per_coord;
  o.min_cos_angle_merge = c.min_cos_angle_merge;
  o.max_merge_dist = c.max_merge_dist;
  o.min_region_growing_candidate_size = c.min_region_growing_candidate_size;
  o.min_region_growing_cells_activated = c.min_region_growing_cells_activated;
  o.min_region_planarity_score = c.min_region_planarity_score;
  o.depth_sigma_coeff = c.depth_sigma_coeff;
  o.depth_sigma_margin = c.depth_sigma_margin;
  o.min_pts_per_cell = c.min_pts_per_cell;
  o.depth_discontinuity_threshold = c.depth_discontinuity_threshold;
  o.max_number_depth_discontinuity = c.max_number_depth_discontinuity;
  o.ransac_refinement = c.ransac_refinement != 0;
  o.ransac_max_iterations = c.ransac_max_iterations;
  o.ransac_threshold = c.ransac_threshold;
  o.ransac_inliers_ratio = c.ransac_inliers_ratio;
  return o;
}

int report(const std::exception& e, char* err, int errlen) {
  if (err && errlen > 0) {
    std::strncpy(err, e.what(), static_cast<size_t>(errlen) - 1);
    err[errlen - 1] = 0;
  }
  return 1;
}
}  // namespace

extern "C" {

// PlaneExtractor(h, w, cfg).process(pcd): xyz is column-major [n_points x 3] (Eigen::MatrixX3f's own order).
// Returns 0, or 1 with the std::runtime_error text in err.
int ref_process(int32_t h, int32_t w, const ShimConfig* cfg, const float* xyz_colmajor, int64_t n_points, int32_t* labels,
                char* err, int errlen) {
  try {
    deplex::PlaneExtractor extractor(h, w, to_config(*cfg));
    Eigen::Map<const Eigen::MatrixX3f> pcd(xyz_colmajor, n_points, 3);
    const Eigen::VectorXi out = extractor.process(pcd);
    std::memcpy(labels, out.data(), sizeof(int32_t) * static_cast<size_t>(out.size()));
    return 0;
  } catch (const std::exception& e) {
    return report(e, err, errlen);
  }
}

// The CellGrid of one frame exactly as process() builds it (plane_extractor.cpp:199): per cell the raw coord_sum_ [3],
// variance_ [9, row i col j at 3*i+j], normal [3], mse, score, d and the planar flag.
int ref_cell_stats(int32_t h, int32_t w, const ShimConfig* cfg, const float* xyz_colmajor, float* sum, float* var,
                   float* normal, float* mse, float* score, float* d, uint8_t* planar, char* err, int errlen) {
  try {
    const deplex::config::Config config = to_config(*cfg);
    const int32_t p = config.patch_size;
    const int32_t nh = w / p, nv = h / p;
    Eigen::Map<const Eigen::MatrixX3f> pcd(xyz_colmajor, static_cast<int64_t>(h) * w, 3);
    const Eigen::MatrixX3f owned = pcd;  // process() receives a MatrixX3f const&: the same implicit conversion follows
    deplex::CellGrid grid(owned, config, nh, nv);
    for (size_t c = 0; c < grid.size(); ++c) {
      const deplex::CellSegmentStat& st = grid[c].getStat();
      for (int i = 0; i < 3; ++i) {
        sum[3 * c + i] = st.coord_sum_(i);
        normal[3 * c + i] = st.normal_(i);
        for (int j = 0; j < 3; ++j) var[9 * c + 3 * i + j] = st.variance_(i, j);
      }
      mse[c] = st.mse_;
      score[c] = st.score_;
      d[c] = st.d_;
      planar[c] = grid[c].isPlanar() ? 1 : 0;
    }
    return 0;
  } catch (const std::exception& e) {
    return report(e, err, errlen);
  }
}

const char* ref_eigen_version() {
  static const std::string v = std::to_string(EIGEN_WORLD_VERSION) + "." + std::to_string(EIGEN_MAJOR_VERSION) + "." +
                               std::to_string(EIGEN_MINOR_VERSION);
  return v.c_str();
}

}  // extern "C"
