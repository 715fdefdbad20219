// cr_math.cuh -- correctly rounded fp64 sin / cos / atan2 in double-double arithmetic, for the few cells where the last
// ulp of the eigen-solver's trigonometry is visible in the labels.
//
// The reference's solver (libs/dsyev/src/dsyevc3.c:46-77) calls libm's atan2, cos and sin; glibc returns correctly
// rounded values (its error bound is 0.55 ulp, so it can differ from the correctly rounded value only when that lies
// within 0.05 ulp of a rounding boundary), CUDA's are allowed one to two ulp.  For almost every cell the difference
// disappears in the cast to fp32.  It does not for a cell whose normal has an exact zero component: the smallest
// eigenvalue of such a cell's covariance is zero up to rounding, its computed sign (+-1e-17) depends on the last ulp of
// cos / sin, that sign ends up as the sign of the zero (dsyevh3.c:88-92: Q[0][0] = Q[0][1] + A[0][2] * w[0]), and
// atan2(+-0, ny < 0) = +-pi puts the cell in the first or the last azimuth bin (normals_histogram.cpp:33-45).
// region_grow.cu repair_axis_cell fits exactly those cells again with the functions below (plane_fit.cuh fit_plane<true>).
//
// Everything is `__host__ __device__` so that tests/test_cr_math_cpu.py can check the same code against glibc on the CPU.
#pragma once
#include <cmath>

#ifdef __CUDACC__
#define CRM_HD __host__ __device__ __forceinline__
#else
#define CRM_HD inline
#endif

namespace dpx {
namespace crm {

struct dd {
  double hi, lo;
};

CRM_HD double fma_(double a, double b, double c) {
#ifdef __CUDA_ARCH__
  return __fma_rn(a, b, c);
#else
  return std::fma(a, b, c);
#endif
}
// error-free transformations (every operation below rounds once, to nearest; compiled with contraction off)
CRM_HD dd two_sum(double a, double b) {
  const double s = a + b;
  const double bb = s - a;
  return {s, (a - (s - bb)) + (b - bb)};
}
CRM_HD dd quick_two_sum(double a, double b) {  // |a| >= |b|
  const double s = a + b;
  return {s, b - (s - a)};
}
CRM_HD dd two_prod(double a, double b) {
  const double p = a * b;
  return {p, fma_(a, b, -p)};
}
CRM_HD dd add(dd a, dd b) {
  dd s = two_sum(a.hi, b.hi);
  const dd t = two_sum(a.lo, b.lo);
  s.lo += t.hi;
  s = quick_two_sum(s.hi, s.lo);
  s.lo += t.lo;
  return quick_two_sum(s.hi, s.lo);
}
CRM_HD dd neg(dd a) { return {-a.hi, -a.lo}; }
CRM_HD dd sub(dd a, dd b) { return add(a, neg(b)); }
CRM_HD dd mul(dd a, dd b) {
  dd p = two_prod(a.hi, b.hi);
  p.lo += a.hi * b.lo + a.lo * b.hi;
  return quick_two_sum(p.hi, p.lo);
}
CRM_HD dd mul_d(dd a, double b) {
  dd p = two_prod(a.hi, b);
  p.lo += a.lo * b;
  return quick_two_sum(p.hi, p.lo);
}
CRM_HD dd div(dd a, dd b) {  // long division, three quotient digits
  const double q1 = a.hi / b.hi;
  dd r = sub(a, mul_d(b, q1));
  const double q2 = r.hi / b.hi;
  r = sub(r, mul_d(b, q2));
  const double q3 = r.hi / b.hi;
  dd q = quick_two_sum(q1, q2);
  return add(q, {q3, 0.0});
}
CRM_HD double to_double(dd a) { return a.hi + a.lo; }  // (hi, lo) is normalised: this is the rounding of the sum

// sin and cos of a double in [0, 3.2], as double-doubles: Taylor series of a / 32, five angle doublings.
//   sin x = x (1 - x^2/(2*3) (1 - x^2/(4*5) (...))),  cos x = 1 - x^2/(1*2) (1 - x^2/(3*4) (...)),
// nine factors each (x <= 0.1: x^20 / 20! < 1e-38), the reciprocals 1/((2k)(2k+1)) and 1/((2k-1)(2k)) as constants.
CRM_HD void sincos_dd(double a, dd& s, dd& c) {
  const dd rs[9] = {{0.16666666666666666, 9.25185853854297e-18},  {0.05, -2.7755575615628915e-18},
                    {0.023809523809523808, 1.32169407693471e-18}, {0.013888888888888888, 7.709882115452476e-19},
                    {0.00909090909090909, 4.415659757031872e-19}, {0.00641025641025641, 2.2240044563805217e-19},
                    {0.004761904761904762, -4.295505750037808e-19}, {0.003676470588235294, 5.102127870520021e-20},
                    {0.0029239766081871343, 1.6231330769373633e-19}};
  const dd rc[9] = {{0.5, 0.0},                                   {0.08333333333333333, 4.625929269271485e-18},
                    {0.03333333333333333, 4.625929269271486e-19}, {0.017857142857142856, 9.912705577010326e-19},
                    {0.011111111111111112, -4.2404351634988616e-19}, {0.007575757575757576, -2.1026951223961299e-19},
                    {0.005494505494505495, -4.2891514515910067e-19}, {0.004166666666666667, 5.782411586589357e-20},
                    {0.0032679738562091504, -9.920804192677818e-20}};
  const double x = a * 0.03125;  // exact
  const dd x2 = two_prod(x, x);
  dd ps = {1.0, 0.0}, pc = {1.0, 0.0};
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int k = 8; k >= 0; --k) {
    ps = sub({1.0, 0.0}, mul(mul(x2, rs[k]), ps));
    pc = sub({1.0, 0.0}, mul(mul(x2, rc[k]), pc));
  }
  s = mul_d(ps, x);
  c = pc;
  for (int i = 0; i < 5; ++i) {  // sin 2t = 2 sin t cos t, cos 2t = 1 - 2 sin^2 t
    const dd sc = mul(s, c);
    const dd ss = mul(s, s);
    s = {2.0 * sc.hi, 2.0 * sc.lo};
    c = sub({1.0, 0.0}, {2.0 * ss.hi, 2.0 * ss.lo});
  }
}

// Correctly rounded sin and cos of a in [0, 3.2] (the solver's phi is in [0, pi/3]).
CRM_HD void sincos_cr(double a, double& sn, double& cs) {
  dd s, c;
  sincos_dd(a, s, c);
  sn = to_double(s);
  cs = to_double(c);
}

// Correctly rounded atan2(y, x) for y >= 0, (x, y) != (0, 0): a libm value refined by one Newton step in double-double,
//   a1 = a0 + (y cos a0 - x sin a0) / (x cos a0 + y sin a0).
CRM_HD double atan2_cr(double y, double x, double a0) {  // a0 = the platform's atan2(y, x), within a few ulp
  if (y == 0.0) return a0;                             // 0 or pi exactly (the sign of zero is the caller's)
  dd s, c;
  sincos_dd(a0, s, c);
  const dd num = sub(mul_d(c, y), mul_d(s, x));
  const dd den = add(mul_d(c, x), mul_d(s, y));
  const dd delta = div(num, den);
  return to_double(add({a0, 0.0}, delta));
}

}  // namespace crm
}  // namespace dpx
