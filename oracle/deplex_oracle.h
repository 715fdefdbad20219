/*
 * deplex_oracle.h -- C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  The oracle is a CPU restatement of the reference's
 * plane-extraction hot path (prime-slam/deplex, `PlaneExtractor::process`).  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  Nothing under deplex_b200/ includes, links or calls it.
 *
 * PARITY STATUS: "parity unpinned at the Eigen boundary" -- see deplex_oracle.cpp.
 */
#ifndef DEPLEX_ORACLE_H
#define DEPLEX_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Field-for-field mirror of deplex::config::Config (cpp/deplex/include/deplex/config.h:51-81). */
typedef struct dpxo_config {
  int32_t patch_size;
  int32_t histogram_bins_per_coord;
  float min_cos_angle_merge;
  float max_merge_dist;
  int32_t min_region_growing_candidate_size;
  int32_t min_region_growing_cells_activated;
  float min_region_planarity_score;
  float depth_sigma_coeff;
  float depth_sigma_margin;
  int32_t min_pts_per_cell;
  float depth_discontinuity_threshold;
  int32_t max_number_depth_discontinuity;
  int32_t ransac_refinement;
  int32_t ransac_max_iterations;
  float ransac_threshold;
  float ransac_inliers_ratio;
} dpxo_config;

enum { DPXO_LAYOUT_COLMAJOR = 0, /* X[N] Y[N] Z[N]   (Eigen::MatrixX3f default) */
       DPXO_LAYOUT_ROWMAJOR = 1  /* x0 y0 z0 x1 y1 z1 (numpy C-order (N,3))      */ };

/* Optional per-stage dumps.  Every pointer may be NULL.  Cell arrays hold
 * n_cells = (w/patch)*(h/patch) entries, plane arrays hold plane_capacity entries. */
typedef struct dpxo_debug {
  /* per cell */
  uint8_t* cell_valid;   /* passed valid-point + depth-continuity checks (stats exist) */
  uint8_t* cell_planar;  /* planar mask */
  float* cell_sum;       /* [n_cells][3]  S   */
  float* cell_var;       /* [n_cells][9]  X^T X, row-major 3x3 */
  float* cell_mean;      /* [n_cells][3] */
  float* cell_normal;    /* [n_cells][3] */
  float* cell_d;
  float* cell_mse;
  float* cell_score;
  float* cell_tol;       /* merge tolerance */
  double* cell_eval;     /* [n_cells][3] eigenvalues as returned by the 3x3 solver */
  int32_t* cell_bin;     /* initial histogram bin, -1 = none */
  int32_t* cell_seglabel;/* labels_map_ after region growing (1-based segment, 0 = none) */
  /* per plane segment (before merging; stats of merged-into planes are post-merge) */
  int32_t plane_capacity;
  int32_t* n_planes;     /* out: number of segments */
  float* plane_normal;   /* [cap][3] */
  float* plane_d;
  float* plane_mean;     /* [cap][3] */
  float* plane_mse;
  float* plane_score;
  int32_t* plane_npts;
  int32_t* merge_labels; /* [cap] */
  /* counters */
  int32_t* n_seeds;      /* number of growSeed calls */
  int32_t* n_ql_fallback;/* number of 3x3 solves that took the QL branch */
} dpxo_debug;

void dpxo_config_default(dpxo_config* cfg);
/* 0 ok, 1 = could not open (err holds the reference's message). Unknown keys go to stderr. */
int dpxo_config_load_ini(const char* path, dpxo_config* cfg, char* err, int errlen);

/* One PlaneExtractor(h, w, cfg).process(xyz).  Returns 0 on success; 1 when the reference
 * would throw std::runtime_error (text in err); 2 for inputs outside the supported domain. */
int dpxo_process(int32_t h, int32_t w, const dpxo_config* cfg, const float* xyz, int64_t n_points,
                 int layout, int32_t* labels, dpxo_debug* dbg, char* err, int errlen);

/* Frame-parallel batch (one extractor per thread), used only as the CPU baseline. */
int dpxo_process_batch(int32_t h, int32_t w, const dpxo_config* cfg, const float* xyz,
                       int32_t n_frames, int layout, int32_t* labels, int32_t n_threads,
                       char* err, int errlen);

/* Sensitivity switch for tools/order_sensitivity.py: 0 = restated Eigen 3.4 reduction orders (default), 1 = plain
 * left-to-right fp32, 2 = fp64 accumulation rounded once.  Process-global; parity is always judged against 0. */
void dpxo_set_sum_variant(int variant);

/* Which libstdc++ generation's std::uniform_int_distribution<int> the RANSAC refinement draws with: 0 = GCC >= 11
 * (Lemire multiply-shift, default), 1 = GCC <= 10 (scaling + rejection).  Process-global.
 * dpxo_uniform_below draws one value in [0, n) from a caller-owned std::mt19937 (tests pin it to <random>). */
void dpxo_set_uniform_int_variant(int variant);
int dpxo_uniform_below(void* mt19937, uint32_t n);
/* mismatches between variant 0 and this build host's std::uniform_int_distribution<int>(0, n-1) over `draws` draws */
int dpxo_uniform_selftest(uint32_t n, int draws);

/* The oracle's own restatement of Kopp's hybrid 3x3 symmetric eigensolver (row-major 3x3 in/out).
 * Returns 1 if the QL branch was taken, else 0. */
int dpxo_eig3(const double* A, double* Q, double* w);

/* DepthImage::toPointCloud restated (depth_image.cpp:55-78): row-major (N,3) output. */
void dpxo_depth_to_cloud(const uint16_t* depth, int32_t h, int32_t w, float fx, float fy, float cx,
                         float cy, float* xyz_rowmajor);

#ifdef __cplusplus
}
#endif
#endif
