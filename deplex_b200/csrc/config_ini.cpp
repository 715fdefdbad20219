// config_ini.cpp -- host-side Config: defaults and the .ini reader behind dpx_config_default /
// dpx_config_load_ini.  Replaces deplex::config::Config (cpp/deplex/include/deplex/config.h:51-81,
// cpp/deplex/src/deplex/config.cpp:23-80).  Same 16 keys, same parsing rules:
//   * empty lines, lines starting with '#', lines without '=' and lines starting with '=' are skipped
//     (so "[Parameters]" is ignored and ";key=value" is reported as an unknown key);
//   * no whitespace trimming, exact case-sensitive key match;
//   * integers via std::stoi, floats via std::stof (trailing junk such as '\r' tolerated);
//   * unknown keys are reported on stderr and otherwise ignored;
//   * an unreadable file is an error: "Couldn't open ini file: <path>".
#include "deplex_b200.h"
#include "error_state.h"

#include <cstring>
#include <fstream>
#include <iostream>
#include <string>

namespace {

struct Key {
  const char* name;
  enum Kind { kInt, kFloat, kBool } kind;
  size_t offset;
};

#define DPX_KEY(name, kind, field) {name, Key::kind, offsetof(dpx_config, field)}
const Key kKeys[] = {
    DPX_KEY("patchSize", kInt, patch_size),
    DPX_KEY("histogramBinsPerCoord", kInt, histogram_bins_per_coord),
    DPX_KEY("minCosAngleForMerge", kFloat, min_cos_angle_merge),
    DPX_KEY("maxMergeDist", kFloat, max_merge_dist),
    DPX_KEY("minRegionGrowingCandidateSize", kInt, min_region_growing_candidate_size),
    DPX_KEY("minRegionGrowingCellsActivated", kInt, min_region_growing_cells_activated),
    DPX_KEY("minRegionPlanarityScore", kFloat, min_region_planarity_score),
    DPX_KEY("depthSigmaCoeff", kFloat, depth_sigma_coeff),
    DPX_KEY("depthSigmaMargin", kFloat, depth_sigma_margin),
    DPX_KEY("minPtsPerCell", kInt, min_pts_per_cell),
    DPX_KEY("depthDiscontinuityThreshold", kFloat, depth_discontinuity_threshold),
    DPX_KEY("maxNumberDepthDiscontinuity", kInt, max_number_depth_discontinuity),
    DPX_KEY("ransacRefinement", kBool, ransac_refinement),
    DPX_KEY("ransacMaxIterations", kInt, ransac_max_iterations),
    DPX_KEY("ransacThreshold", kFloat, ransac_threshold),
    DPX_KEY("ransacInliersRatio", kFloat, ransac_inliers_ratio),
};
#undef DPX_KEY

}  // namespace

extern "C" void dpx_config_default(dpx_config* cfg) {
  if (!cfg) return;
  cfg->patch_size = 10;
  cfg->histogram_bins_per_coord = 20;
  cfg->min_cos_angle_merge = 0.90;
  cfg->max_merge_dist = 500;
  cfg->min_region_growing_candidate_size = 5;
  cfg->min_region_growing_cells_activated = 4;
  cfg->min_region_planarity_score = 0.55;
  cfg->depth_sigma_coeff = 1.425e-6;
  cfg->depth_sigma_margin = 10.;
  cfg->min_pts_per_cell = 3;
  cfg->depth_discontinuity_threshold = 160;
  cfg->max_number_depth_discontinuity = 1;
  cfg->ransac_refinement = 0;
  cfg->ransac_max_iterations = 1000;
  cfg->ransac_threshold = 1.;
  cfg->ransac_inliers_ratio = 0.9;
}

extern "C" dpx_status dpx_config_load_ini(const char* path, dpx_config* cfg) {
  if (!path || !cfg) {
    dpx::set_thread_error("dpx_config_load_ini: null argument");
    return DPX_ERR_ARGUMENT;
  }
  dpx_config_default(cfg);
  std::ifstream ini(path);
  if (!ini.is_open()) {
    dpx::set_thread_error(std::string("Couldn't open ini file: ") + path);
    return DPX_ERR_RUNTIME;
  }
  std::string line;
  while (ini) {
    line.clear();
    std::getline(ini, line);
    if (line.empty() || line[0] == '#') continue;
    const size_t eq = line.find('=');
    if (eq == std::string::npos || eq == 0) continue;
    const std::string key = line.substr(0, eq);
    const std::string value = line.substr(eq + 1);
    const Key* hit = nullptr;
    for (const Key& k : kKeys)
      if (key == k.name) {
        hit = &k;
        break;
      }
    if (!hit) {
      std::cerr << "Unknown parameter name: " << key << '\n';
      continue;
    }
    char* field = reinterpret_cast<char*>(cfg) + hit->offset;
    try {
      if (hit->kind == Key::kFloat) {
        const float v = std::stof(value);
        std::memcpy(field, &v, sizeof(v));
      } else {
        int32_t v = std::stoi(value);
        if (hit->kind == Key::kBool) v = v != 0;
        std::memcpy(field, &v, sizeof(v));
      }
    } catch (const std::exception& e) {
      // the reference lets std::invalid_argument / std::out_of_range escape the constructor
      dpx::set_thread_error(std::string(e.what()));
      return DPX_ERR_RUNTIME;
    }
  }
  return DPX_OK;
}
