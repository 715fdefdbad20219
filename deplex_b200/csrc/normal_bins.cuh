// normal_bins.cuh -- NormalsHistogram's bin of a unit normal (normals_histogram.cpp:33-45), shared by the cell-statistics
// kernels (cell_walk.cuh finish_cell) and the repair of the cells whose bin hangs on the sign of a zero (region_grow.cu).
#pragma once
#include "exact_math.cuh"

namespace dpx {

// acos / atan2 for the histogram bin (normals_histogram.cpp:33-45).  The bin is trunc((B - 1) * angle / range), so an angle
// one ulp off flips the bin only when it sits on a bin boundary -- which generic normals never do (probability ~1e-14) and
// axis-aligned ones do all the time: a plane whose normal lies in the y-z plane comes out of the fp64 eigen-solver with
// nx ~ 1e-16, ny < 0, and its azimuth pi - 7e-16 is the last bin's upper edge.  glibc rounds such angles correctly, CUDA's
// libm is allowed two ulp (a 720p frame in tools/soak_refine.py got a different seed order from one such cell).  Next to
// the axes the angle is (a multiple of pi/2) - (a tiny arctangent / arcsine), and adding the tiny term to the constant's
// low word first gives the correctly rounded sum; everywhere else CUDA's function is used.
__device__ __forceinline__ double small_atan(double t) {  // |t| < 2^-20: t - t^3/3, error far below 2^-53 * |t|
  return __dsub_rn(t, __ddiv_rn(__dmul_rn(__dmul_rn(t, t), t), 3.0));
}
__device__ __forceinline__ double atan2_for_bins(double s, double c) {
  constexpr double kNear = 9.5367431640625e-07;  // 2^-20
  constexpr double kPiHi = 3.141592653589793116, kPiLo = 1.2246467991473532e-16;
  constexpr double kPio2Hi = 1.5707963267948966, kPio2Lo = 6.123233995736766e-17;
  const double as = ::fabs(s), ac = ::fabs(c);
  if (c < 0.0 && as < __dmul_rn(kNear, ac)) {  // +-(pi - eps)
    const double r = __dadd_rn(kPiHi, __dsub_rn(kPiLo, small_atan(__ddiv_rn(as, ac))));
    return ::signbit(s) ? -r : r;
  }
  if (as > 0.0 && ac < __dmul_rn(kNear, as)) {  // +-(pi/2 - eps), eps of either sign
    const double r = __dadd_rn(kPio2Hi, __dsub_rn(kPio2Lo, small_atan(__ddiv_rn(c, as))));
    return s < 0.0 ? -r : r;
  }
  return ::atan2(s, c);
}
__device__ __forceinline__ double acos_for_bins(double x) {
  constexpr double kNear = 9.5367431640625e-07;
  constexpr double kPio2Hi = 1.5707963267948966, kPio2Lo = 6.123233995736766e-17;
  if (::fabs(x) < kNear) {  // pi/2 - asin(x), asin(x) = x + x^3/6
    const double a = __dadd_rn(x, __ddiv_rn(__dmul_rn(__dmul_rn(x, x), x), 6.0));
    return __dadd_rn(kPio2Hi, __dsub_rn(kPio2Lo, a));
  }
  return ::acos(x);
}

// NormalsHistogram's bin of a unit normal (normals_histogram.cpp:33-45; isZero() precision 1e-5; fp64 trigonometry), or -1:
// a zero normal is skipped by the reference, and outside [0, B*B) it writes out of bounds -- such a cell is dropped here.
// CAREFUL = false: CUDA's acos / atan2 (the streaming kernel); true: the axis-aware versions above (the repair path).
template <bool CAREFUL>
__device__ __forceinline__ int histogram_bin(float nx, float ny, float nz, int B) {
  if (fabsf(nx) <= 1e-5f && fabsf(ny) <= 1e-5f && fabsf(nz) <= 1e-5f) return -1;
  const f64 dnx = static_cast<double>(nx), dny = static_cast<double>(ny);
  const f64 proj = sqrt(dnx * dnx + dny * dny);
  const double polar = CAREFUL ? acos_for_bins(static_cast<double>(-nz)) : ::acos(static_cast<double>(-nz));
  const double azimuth = CAREFUL ? atan2_for_bins((dnx / proj).v, (dny / proj).v) : ::atan2((dnx / proj).v, (dny / proj).v);
  const double kPi = 3.14159265358979323846;
  const int xq = __double2int_rz(((f64(static_cast<double>(B - 1)) * (f64(polar) - f64(0.0))) / f64(kPi)).v);
  int yq = 0;
  if (xq > 0)
    yq = __double2int_rz(((f64(static_cast<double>(B - 1)) * (f64(azimuth) - f64(-kPi))) / (f64(kPi) - f64(-kPi))).v);
  const int b = yq * B + xq;
  return (b >= 0 && b < B * B) ? b : -1;
}

// The normals whose bin hangs on the last ulp of the trigonometry: within ~1e-12 of an axis of the azimuth or of the polar
// angle.  A normal fitted to noisy or quantised depth is never that close; the x component of an exactly axis-aligned
// noise-free plane is zero or rounding noise of the fp64 solver (1e-16), and always is.
__device__ __forceinline__ bool bin_needs_care(float nx, float ny, float nz) {
  constexpr float kNear = 9.094947017729282e-13f;  // 2^-40
  const float ax = fabsf(nx), ay = fabsf(ny);
  return (ny < 0.f && ax < kNear * ay) || ay < kNear * ax || fabsf(nz) < kNear;
}
// ... and, of those, the ones on the azimuth's wrap (+-pi): the sign of a zero or noise-level x component decides between
// the first and the last bin, and it comes out of the solver's sin / cos (cr_math.cuh)
__device__ __forceinline__ bool normal_on_wrap(float nx, float ny) { return ny < 0.f && fabsf(nx) < 9.094947017729282e-13f * fabsf(ny); }

}  // namespace dpx
