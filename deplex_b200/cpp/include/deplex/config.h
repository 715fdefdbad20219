// deplex/config.h -- deplex::config::Config, the parameter block of the plane extractor.
//
// Drop-in for the reference header cpp/deplex/include/deplex/config.h:29-82: the same struct, the same
// 16 public fields with the same defaults, the same three constructors.  The ini reader itself lives
// behind the C-ABI (dpx_config_load_ini, include/deplex_b200.h) so that C++, Python and any other host
// language parse identically; this header only moves the values in and out of `dpx_config`.
#pragma once

#include <cstdint>
#include <string>
#include <unordered_map>

struct dpx_config;

namespace deplex {
namespace config {
struct Config {
 public:
  /** Default parameters (config.h:51-81). */
  Config();
  /** The reference declares this constructor and leaves it empty (config.cpp:25-26): defaults are kept. */
  Config(std::unordered_map<std::string, std::string> const& param_map);
  /** Read `key=value` lines from an .ini file (config.cpp:28-80).  Throws std::runtime_error
   *  "Couldn't open ini file: <path>" when the file cannot be opened. */
  Config(std::string const& config_path);

  int32_t patch_size = 10;
  int32_t histogram_bins_per_coord = 20;
  float min_cos_angle_merge = 0.90;
  float max_merge_dist = 500;
  int32_t min_region_growing_candidate_size = 5;
  int32_t min_region_growing_cells_activated = 4;
  float min_region_planarity_score = 0.55;
  float depth_sigma_coeff = 1.425e-6;
  float depth_sigma_margin = 10.;
  int32_t min_pts_per_cell = 3;
  float depth_discontinuity_threshold = 160;
  int32_t max_number_depth_discontinuity = 1;
  bool ransac_refinement = false;
  int32_t ransac_max_iterations = 1000;
  float ransac_threshold = 1.;
  float ransac_inliers_ratio = 0.9;

  /** Conversions to and from the C-ABI mirror (field for field). */
  void toC(dpx_config* out) const;
  static Config fromC(dpx_config const& in);
};
}  // namespace config
}  // namespace deplex
