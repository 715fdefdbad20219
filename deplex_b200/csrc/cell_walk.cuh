// cell_walk.cuh -- per-cell arithmetic of stage 1: the order-exact moment walk over one cell and everything
// computed from the moments (validity, plane fit, planarity, merge tolerance, histogram bin).
//
// Reference: CellSegment::CellSegment and its predicates (cell_segment.cpp:21-35,57-110),
// CellSegmentStat::CellSegmentStat (cell_segment_stat.cpp:29-35), NormalsHistogram's bin formula
// (normals_histogram.cpp:31-45).  Summation orders are Eigen 3.4's.
#pragma once
#include "common.cuh"
#include "plane_fit.cuh"
#include "normal_bins.cuh"

namespace dpx {

// ---- shared-memory tile addressing -------------------------------------------------------------
// Row-major input (x y z interleaved): tile[(i * tw + col) * 3 + a]
// Col-major input (three planes):      tile[(a * p + i) * tw + col]
template <int LAYOUT>
__device__ __forceinline__ int tile_index(int tw, int p, int i, int col, int a) {
  return LAYOUT == kLayoutRowMajor ? (i * tw + col) * 3 + a : (a * p + i) * tw + col;
}

// Eigen 3.4 DenseBase::sum() over one contiguous fp32 column of N entries whose first 16-byte aligned
// entry is S (Redux.h, LinearVectorizedTraversal/NoUnrolling, SSE2 Packet4f): two packet accumulators
// over [S, S + 8*floor((N-S)/8)), an optional remainder packet, predux (a0+a2)+(a1+a3), then the
// leading and trailing scalars.  `add(k, v)` is called with k = 0..N-1 in order; after full
// unrolling every index below is a compile-time constant, so all state lives in registers.
template <int N, int S0>
struct ColumnSum {
  static constexpr int kS = S0 > N ? N : S0;
  static constexpr int kAligned = ((N - kS) / 4) * 4;
  static constexpr int kAligned2 = ((N - kS) / 8) * 8;
  static constexpr int kTrail = N - kS - kAligned;
  float a0[4], a1[4], rem[4], lead[3], trail[3], seq;

  __device__ __forceinline__ void add(int k, float v) {
    if (kAligned == 0) {
      seq = (k == 0) ? v : __fadd_rn(seq, v);
      return;
    }
    if (k < kS) {
      lead[k] = v;
      return;
    }
    const int m = k - kS;
    if (m < kAligned2) {
      const int pi = m >> 2, l = m & 3;
      if (pi == 0) a0[l] = v;
      else if (pi == 1) a1[l] = v;
      else if ((pi & 1) == 0) a0[l] = __fadd_rn(a0[l], v);
      else a1[l] = __fadd_rn(a1[l], v);
    } else if (m < kAligned) {
      rem[m - kAligned2] = v;
    } else {
      trail[m - kAligned] = v;
    }
  }

  __device__ __forceinline__ float result() const {
    if (kAligned == 0) return seq;
    float r[4];
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      if (kAligned2 == 0) {
        r[l] = rem[l];
      } else {
        r[l] = __fadd_rn(a0[l], a1[l]);
        if (kAligned > kAligned2) r[l] = __fadd_rn(r[l], rem[l]);
      }
    }
    float res = __fadd_rn(__fadd_rn(r[0], r[2]), __fadd_rn(r[1], r[3]));
#pragma unroll
    for (int i = 0; i < kS; ++i) res = __fadd_rn(res, lead[i]);
#pragma unroll
    for (int i = 0; i < kTrail; ++i) res = __fadd_rn(res, trail[i]);
    return res;
  }
};

struct CellRaw {
  Moments m;
  int valid_cnt, hcnt, vcnt;
  float first[3], last[3];
};

__device__ __forceinline__ void scan_step(float cur, float& prev, int& cnt, float thr) {
  // cell_segment.cpp:68-73 / :84-88
  if (cur > 0.f && fabsf(__fsub_rn(cur, prev)) < thr)
    prev = cur;
  else if (cur > 0.f)
    ++cnt;
}

// Load the P points of row i of cell t with the widest shared-memory vector the alignment allows.
// `rows` is the number of image rows held by the staged block (col-major planes are `rows` rows apart).
template <int LAYOUT, int P>
__device__ __forceinline__ void load_cell_row(const float* tile, int tw, int rows, int i, int t, float (&x)[P],
                                              float (&y)[P], float (&z)[P]) {
  constexpr int VW = (P % 4 == 0) ? 4 : ((P % 2 == 0) ? 2 : 1);
  if (LAYOUT == kLayoutRowMajor) {
    const float* src = tile + (i * tw + t * P) * 3;
    float buf[3 * P];
    if (VW == 4) {
#pragma unroll
      for (int q = 0; q < 3 * P / 4; ++q) {
        const float4 v = reinterpret_cast<const float4*>(src)[q];
        buf[4 * q] = v.x; buf[4 * q + 1] = v.y; buf[4 * q + 2] = v.z; buf[4 * q + 3] = v.w;
      }
    } else if (VW == 2) {
#pragma unroll
      for (int q = 0; q < 3 * P / 2; ++q) {
        const float2 v = reinterpret_cast<const float2*>(src)[q];
        buf[2 * q] = v.x; buf[2 * q + 1] = v.y;
      }
    } else {
#pragma unroll
      for (int q = 0; q < 3 * P; ++q) buf[q] = src[q];
    }
#pragma unroll
    for (int j = 0; j < P; ++j) {
      x[j] = buf[3 * j]; y[j] = buf[3 * j + 1]; z[j] = buf[3 * j + 2];
    }
  } else {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float* src = tile + (a * rows + i) * tw + t * P;
      float(&dst)[P] = (a == 0) ? x : (a == 1 ? y : z);
      if (VW == 4) {
#pragma unroll
        for (int q = 0; q < P / 4; ++q) {
          const float4 v = reinterpret_cast<const float4*>(src)[q];
          dst[4 * q] = v.x; dst[4 * q + 1] = v.y; dst[4 * q + 2] = v.z; dst[4 * q + 3] = v.w;
        }
      } else if (VW == 2) {
#pragma unroll
        for (int q = 0; q < P / 2; ++q) {
          const float2 v = reinterpret_cast<const float2*>(src)[q];
          dst[2 * q] = v.x; dst[2 * q + 1] = v.y;
        }
      } else {
#pragma unroll
        for (int q = 0; q < P; ++q) dst[q] = src[q];
      }
    }
  }
}

// Stateful, fully unrolled walk over one cell (compile-time patch size), fed one image row at a time so
// that rows can be consumed as they arrive in shared memory.  `i` must be a compile-time constant after
// unrolling at the call site.
template <int P>
struct CellWalk {
  static constexpr int N = P * P;
  ColumnSum<N, (4 - (0 * N) % 4) % 4> sx;
  ColumnSum<N, (4 - (1 * N) % 4) % 4> sy;
  ColumnSum<N, (4 - (2 * N) % 4) % 4> sz;
  float vxx, vxy, vxz, vyy, vyz, vzz;
  int valid, hcnt, vcnt;
  float hprev, vprev;
  float first[3], last[3];

  __device__ __forceinline__ void reset() {
    vxx = vxy = vxz = vyy = vyz = vzz = 0.f;
    valid = hcnt = vcnt = 0;
    hprev = vprev = 0.f;
  }

  __device__ __forceinline__ void row(int i, const float (&x)[P], const float (&y)[P], const float (&z)[P], float disc_thr) {
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const int k = i * P + j;
      const float px = x[j], py = y[j], pz = z[j];
      if (k == 0) { first[0] = px; first[1] = py; first[2] = pz; }
      if (k == N - 1) { last[0] = px; last[1] = py; last[2] = pz; }
      // X^T X: one sequential chain per entry (Eigen GEBP scalar path; cell_segment_stat.cpp:32)
      vxx = __fadd_rn(__fmul_rn(px, px), vxx);
      vxy = __fadd_rn(__fmul_rn(px, py), vxy);
      vxz = __fadd_rn(__fmul_rn(px, pz), vxz);
      vyy = __fadd_rn(__fmul_rn(py, py), vyy);
      vyz = __fadd_rn(__fmul_rn(py, pz), vyz);
      vzz = __fadd_rn(__fmul_rn(pz, pz), vzz);
      // column sums (cell_segment_stat.cpp:31)
      sx.add(k, px);
      sy.add(k, py);
      sz.add(k, pz);
      // hasValidPoints (cell_segment.cpp:57-60)
      valid += (pz > 0.f) ? 1 : 0;
      // isHorizontalContinuous: indices [N/2, N/2 + P) (cell_segment.cpp:62-76)
      if (k == N / 2) hprev = pz;
      if (k >= N / 2 && k < N / 2 + P) scan_step(pz, hprev, hcnt, disc_thr);
      // isVerticalContinuous: indices P/2, P/2 + P, ... (cell_segment.cpp:78-91)
      if (j == P / 2) {
        if (i == 0) vprev = pz;
        scan_step(pz, vprev, vcnt, disc_thr);
      }
    }
  }

  __device__ __forceinline__ void finish(CellRaw& out) const {
    out.m.n = N;
    out.m.s[0] = sx.result(); out.m.s[1] = sy.result(); out.m.s[2] = sz.result();
    out.m.v[0] = vxx; out.m.v[1] = vxy; out.m.v[2] = vxz; out.m.v[3] = vyy; out.m.v[4] = vyz; out.m.v[5] = vzz;
    out.valid_cnt = valid; out.hcnt = hcnt; out.vcnt = vcnt;
#pragma unroll
    for (int a = 0; a < 3; ++a) { out.first[a] = first[a]; out.last[a] = last[a]; }
  }
};

// Runtime-patch fallback (any 4 <= patch <= 26): same orders, loops not unrolled.
template <int LAYOUT>
__device__ float column_sum_runtime(const float* tile, int tw, int p, int t, int a, int n, int s) {
  auto at = [&](int k) { return tile[tile_index<LAYOUT>(tw, p, k / p, t * p + k % p, a)]; };
  if (s > n) s = n;
  const int aligned2 = ((n - s) / 8) * 8, aligned = ((n - s) / 4) * 4;
  const int end2 = s + aligned2, end1 = s + aligned;
  float res;
  if (aligned) {
    float a0[4], a1[4];
#pragma unroll
    for (int l = 0; l < 4; ++l) a0[l] = at(s + l);
    if (aligned > 4) {
#pragma unroll
      for (int l = 0; l < 4; ++l) a1[l] = at(s + 4 + l);
      for (int i = s + 8; i < end2; i += 8) {
#pragma unroll
        for (int l = 0; l < 4; ++l) a0[l] = __fadd_rn(a0[l], at(i + l));
#pragma unroll
        for (int l = 0; l < 4; ++l) a1[l] = __fadd_rn(a1[l], at(i + 4 + l));
      }
#pragma unroll
      for (int l = 0; l < 4; ++l) a0[l] = __fadd_rn(a0[l], a1[l]);
      if (end1 > end2) {
#pragma unroll
        for (int l = 0; l < 4; ++l) a0[l] = __fadd_rn(a0[l], at(end2 + l));
      }
    }
    res = __fadd_rn(__fadd_rn(a0[0], a0[2]), __fadd_rn(a0[1], a0[3]));
    for (int i = 0; i < s; ++i) res = __fadd_rn(res, at(i));
    for (int i = end1; i < n; ++i) res = __fadd_rn(res, at(i));
  } else {
    res = at(0);
    for (int i = 1; i < n; ++i) res = __fadd_rn(res, at(i));
  }
  return res;
}

template <int LAYOUT>
__device__ void walk_cell_runtime(const float* tile, int tw, int p, int t, float disc_thr, CellRaw& out) {
  const int n = p * p;
  float vxx = 0.f, vxy = 0.f, vxz = 0.f, vyy = 0.f, vyz = 0.f, vzz = 0.f;
  int valid = 0, hcnt = 0, vcnt = 0;
  float hprev = 0.f, vprev = 0.f;
  for (int i = 0; i < p; ++i)
    for (int j = 0; j < p; ++j) {
      const int k = i * p + j;
      const float px = tile[tile_index<LAYOUT>(tw, p, i, t * p + j, 0)];
      const float py = tile[tile_index<LAYOUT>(tw, p, i, t * p + j, 1)];
      const float pz = tile[tile_index<LAYOUT>(tw, p, i, t * p + j, 2)];
      if (k == 0) { out.first[0] = px; out.first[1] = py; out.first[2] = pz; }
      if (k == n - 1) { out.last[0] = px; out.last[1] = py; out.last[2] = pz; }
      vxx = __fadd_rn(__fmul_rn(px, px), vxx);
      vxy = __fadd_rn(__fmul_rn(px, py), vxy);
      vxz = __fadd_rn(__fmul_rn(px, pz), vxz);
      vyy = __fadd_rn(__fmul_rn(py, py), vyy);
      vyz = __fadd_rn(__fmul_rn(py, pz), vyz);
      vzz = __fadd_rn(__fmul_rn(pz, pz), vzz);
      valid += (pz > 0.f) ? 1 : 0;
      if (k == n / 2) hprev = pz;
      if (k >= n / 2 && k < n / 2 + p) scan_step(pz, hprev, hcnt, disc_thr);
      if (j == p / 2) {
        if (i == 0) vprev = pz;
        scan_step(pz, vprev, vcnt, disc_thr);
      }
    }
  out.m.n = n;
  for (int a = 0; a < 3; ++a) out.m.s[a] = column_sum_runtime<LAYOUT>(tile, tw, p, t, a, n, (4 - (a * n) % 4) % 4);
  out.m.v[0] = vxx; out.m.v[1] = vxy; out.m.v[2] = vxz; out.m.v[3] = vyy; out.m.v[4] = vyz; out.m.v[5] = vzz;
  out.valid_cnt = valid; out.hcnt = hcnt; out.vcnt = vcnt;
}

// Everything after the moments: validity, plane fit, planarity, merge tolerance, histogram bin.
__device__ __forceinline__ void finish_cell(const CellRaw& raw, const Thresholds& th, const Tables& tb, long long cell) {
  const bool valid = static_cast<unsigned long long>(raw.valid_cnt) >= th.valid_pts_threshold &&
                     raw.hcnt < th.max_number_depth_discontinuity && raw.vcnt < th.max_number_depth_discontinuity;
  uint8_t flags = 0;
  float mse_out = 0.f;
  int bin = -1;
  if (valid) {
    flags = kFlagValid;
    PlaneFit fit;
    fit_plane(raw.m, fit);
    mse_out = fit.mse;

    // hasSmallPlaneError (cell_segment.cpp:99-102): fp32 threshold, compared in fp64
    const float thr = __fadd_rn(__fmul_rn(th.depth_sigma_coeff, __fmul_rn(fit.mean[2], fit.mean[2])), th.depth_sigma_margin);
    const bool planar = static_cast<double>(fit.mse) <= __dmul_rn(static_cast<double>(thr), static_cast<double>(thr));

    // calculateMergeTolerance (cell_segment.cpp:104-110), minimum 20.0 hard-coded at :34
    const float sin_merge = __fsqrt_rn(__fsub_rn(1.0f, __fmul_rn(th.min_cos_angle_merge, th.min_cos_angle_merge)));
    const float dx = __fsub_rn(raw.first[0], raw.last[0]);
    const float dy = __fsub_rn(raw.first[1], raw.last[1]);
    const float dz = __fsub_rn(raw.first[2], raw.last[2]);
    const float diam = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fadd_rn(__fmul_rn(dy, dy), __fmul_rn(dz, dz))));
    const float a = __fmul_rn(diam, sin_merge);
    const float mx = (a < 20.0f) ? 20.0f : a;                              // std::max(a, 20.0f)
    const float tr = (th.max_merge_dist < mx) ? th.max_merge_dist : mx;    // std::min(mx, max_merge_dist)
    const float tol = __fmul_rn(tr, tr);

    if (planar) {
      // (CUDA's acos / atan2 here; the few normals next to an axis, where their last ulp can decide the bin, are worked out
      // again by the next kernel: region_grow.cu edge_mask_kernel / repair_axis_cell)
      const int b = histogram_bin<false>(fit.normal[0], fit.normal[1], fit.normal[2], th.histogram_bins_per_coord);
      if (b >= 0) {
        bin = b;
        flags |= kFlagPlanar;
      }
    }
    tb.rec_a[2 * cell] = make_float4(fit.normal[0], fit.normal[1], fit.normal[2], fit.d);
    tb.rec_a[2 * cell + 1] = make_float4(fit.mean[0], fit.mean[1], fit.mean[2], tol);
    tb.rec_b[3 * cell] = make_float4(raw.m.s[0], raw.m.s[1], raw.m.s[2], raw.m.v[0]);
    tb.rec_b[3 * cell + 1] = make_float4(raw.m.v[1], raw.m.v[2], raw.m.v[3], raw.m.v[4]);
    tb.rec_b[3 * cell + 2] = make_float4(raw.m.v[5], fit.mse, fit.score, 0.f);
  }
  if (valid) tb.mse[cell] = mse_out;
  tb.flags[cell] = flags;
  tb.bin[cell] = static_cast<int16_t>(bin);
}

}  // namespace dpx
