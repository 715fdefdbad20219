// config.cpp -- deplex::config::Config over the C-ABI's ini reader (reference: cpp/deplex/src/deplex/config.cpp:23-80).
#include "deplex/config.h"

#include <stdexcept>

#include "deplex_b200.h"

namespace deplex {
namespace config {

Config::Config() = default;

// config.cpp:25-26 in the reference is an empty TODO body: the map is ignored there too.
Config::Config(std::unordered_map<std::string, std::string> const& param_map) { (void)param_map; }

Config::Config(std::string const& config_path) {
  dpx_config c;
  const dpx_status st = dpx_config_load_ini(config_path.c_str(), &c);
  if (st != DPX_OK) throw std::runtime_error(dpx_last_error(nullptr));
  *this = fromC(c);
}

void Config::toC(dpx_config* out) const {
  out->patch_size = patch_size;
  out->histogram_bins_per_coord = histogram_bins_per_coord;
  out->min_cos_angle_merge = min_cos_angle_merge;
  out->max_merge_dist = max_merge_dist;
  out->min_region_growing_candidate_size = min_region_growing_candidate_size;
  out->min_region_growing_cells_activated = min_region_growing_cells_activated;
  out->min_region_planarity_score = min_region_planarity_score;
  out->depth_sigma_coeff = depth_sigma_coeff;
  out->depth_sigma_margin = depth_sigma_margin;
  out->min_pts_per_cell = min_pts_per_cell;
  out->depth_discontinuity_threshold = depth_discontinuity_threshold;
  out->max_number_depth_discontinuity = max_number_depth_discontinuity;
  out->ransac_refinement = ransac_refinement ? 1 : 0;
  out->ransac_max_iterations = ransac_max_iterations;
  out->ransac_threshold = ransac_threshold;
  out->ransac_inliers_ratio = ransac_inliers_ratio;
}

Config Config::fromC(dpx_config const& in) {
  Config c;
  c.patch_size = in.patch_size;
  c.histogram_bins_per_coord = in.histogram_bins_per_coord;
  c.min_cos_angle_merge = in.min_cos_angle_merge;
  c.max_merge_dist = in.max_merge_dist;
  c.min_region_growing_candidate_size = in.min_region_growing_candidate_size;
  c.min_region_growing_cells_activated = in.min_region_growing_cells_activated;
  c.min_region_planarity_score = in.min_region_planarity_score;
  c.depth_sigma_coeff = in.depth_sigma_coeff;
  c.depth_sigma_margin = in.depth_sigma_margin;
  c.min_pts_per_cell = in.min_pts_per_cell;
  c.depth_discontinuity_threshold = in.depth_discontinuity_threshold;
  c.max_number_depth_discontinuity = in.max_number_depth_discontinuity;
  c.ransac_refinement = in.ransac_refinement != 0;
  c.ransac_max_iterations = in.ransac_max_iterations;
  c.ransac_threshold = in.ransac_threshold;
  c.ransac_inliers_ratio = in.ransac_inliers_ratio;
  return c;
}

}  // namespace config
}  // namespace deplex
