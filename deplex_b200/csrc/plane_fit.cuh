// plane_fit.cuh -- device-side PCA plane fit: fp32 covariance -> fp64 3x3 symmetric eigensolve ->
// normal / d / mse / score, with the planarity test fused by the callers.
//
// Replaces CellSegmentStat::fitPlane (cell_segment_stat.cpp:55-81) and the vendored C eigensolver it
// calls (libs/dsyev/src/dsyevh3.c:31-134, dsyevc3.c:31-80, dsyevq3.c:29-134, dsytrd3.h:27-102).
// The solver is Kopp's "hybrid" method: Cardano eigenvalues + cross-product eigenvectors, falling
// back to Householder + implicit-shift QL when the cross products are ill-conditioned.  Every
// operation rounds once (exact_math.cuh) in the reference's order, so results match the CPU bit for
// bit except where CUDA's sin/cos/atan2 differ from glibc's in the last ulp.
//
// Everything lives in registers: the QL iteration is unrolled over its (at most three) static
// index patterns so that no array is dynamically indexed.
#pragma once
#include "cr_math.cuh"
#include "exact_math.cuh"

namespace dpx {

struct Sym3 {
  f64 a00, a01, a02, a11, a12, a22;  // upper triangle; the reference never reads the lower one
};

struct Eig3 {
  f64 w[3];
  f64 q[9];  // row-major; column j is the eigenvector of w[j]
  bool ql;
};

// Cardano's closed form (dsyevc3.c:46-77).  Order of the results: w[0] >= w[2] >= w[1].
// EXACT: atan2 / sin / cos correctly rounded (cr_math.cuh) instead of CUDA's, for the cells where the last ulp shows.
template <bool EXACT>
__device__ __forceinline__ void eig3_values(const Sym3& A, f64 (&w)[3]) {
  const f64 de = A.a01 * A.a12;
  const f64 dd = sq(A.a01);
  const f64 ee = sq(A.a12);
  const f64 ff = sq(A.a02);
  const f64 m = A.a00 + A.a11 + A.a22;
  const f64 c1 = (A.a00 * A.a11 + A.a00 * A.a22 + A.a11 * A.a22) - (dd + ee + ff);
  const f64 c0 = A.a22 * dd + A.a00 * ee + A.a11 * ff - A.a00 * A.a11 * A.a22 - f64(2.0) * A.a02 * de;

  const f64 p = sq(m) - f64(3.0) * c1;
  const f64 q = m * (p - f64(1.5) * c1) - f64(13.5) * c0;
  const f64 sqrt_p = sqrt(abs(p));

  f64 phi = f64(27.0) * (f64(0.25) * sq(c1) * (p - c1) + c0 * (q + f64(6.75) * c0));
  {
    const double ay = sqrt(abs(phi)).v;
    double at = ::atan2(ay, q.v);
    if (EXACT && ay == ay && q.v == q.v && !(ay == 0.0 && q.v == 0.0)) at = crm::atan2_cr(ay, q.v, at);
    phi = f64(1.0 / 3.0) * f64(at);
  }

  double sn, cs;
  if (EXACT && phi.v >= 0.0 && phi.v <= 3.2)
    crm::sincos_cr(phi.v, sn, cs);
  else
    ::sincos(phi.v, &sn, &cs);
  const f64 c = sqrt_p * f64(cs);
  const f64 s = f64(1.0 / 1.73205080756887729352744634151) * sqrt_p * f64(sn);

  w[1] = f64(1.0 / 3.0) * (m - c);
  w[2] = w[1] + s;
  w[0] = w[1] + c;
  w[1] -= s;
}

// One Givens step of the QL sweep at static position I (dsyevq3.c:95-127).
template <int I>
__device__ __forceinline__ void ql_rotate(f64 (&w)[3], f64 (&e)[3], f64 (&Q)[9], f64& s, f64& c, f64& p, f64& g) {
  const f64 f = s * e[I];
  const f64 b = c * e[I];
  f64 r;
  if (abs(f) > abs(g)) {
    c = g / f;
    r = sqrt(sq(c) + f64(1.0));
    e[I + 1] = f * r;
    s = f64(1.0) / r;
    c *= s;
  } else {
    s = f / g;
    r = sqrt(sq(s) + f64(1.0));
    e[I + 1] = g * r;
    c = f64(1.0) / r;
    s *= c;
  }
  g = w[I + 1] - p;
  r = (w[I] - g) * s + f64(2.0) * c * b;
  p = s * r;
  w[I + 1] = g + p;
  g = c * r - b;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const f64 t = Q[3 * k + I + 1];
    Q[3 * k + I + 1] = s * Q[3 * k + I] + c * t;
    Q[3 * k + I] = c * Q[3 * k + I] - s * t;
  }
}

// QL sweeps for off-diagonal element L (dsyevq3.c:56-131).  Returns false on non-convergence
// (the reference returns -1 there, which its caller ignores).
template <int L>
__device__ __forceinline__ bool ql_converge(f64 (&w)[3], f64 (&e)[3], f64 (&Q)[9]) {
  for (int n_iter = 0;; ++n_iter) {
    int m = L;
    {
      f64 g = abs(w[L]) + abs(w[L + 1]);
      if (!(abs(e[L]) + g == g)) {
        m = L + 1;
        if (L == 0) {
          g = abs(w[1]) + abs(w[2]);
          if (!(abs(e[1]) + g == g)) m = 2;
        }
      }
    }
    if (m == L) return true;
    if (n_iter >= 30) return false;

    f64 g = (w[L + 1] - w[L]) / (e[L] + e[L]);
    const f64 r = sqrt(sq(g) + f64(1.0));
    const f64 wm = (L == 0 && m == 1) ? w[1] : w[2];
    if (g > f64(0.0))
      g = wm - w[L] + e[L] / (g + r);
    else
      g = wm - w[L] + e[L] / (g - r);

    f64 s = 1.0, c = 1.0, p = 0.0;
    if (L == 0) {
      if (m == 2) ql_rotate<1>(w, e, Q, s, c, p, g);
      ql_rotate<0>(w, e, Q, s, c, p, g);
    } else {
      ql_rotate<1>(w, e, Q, s, c, p, g);
    }
    w[L] -= p;
    e[L] = g;
    if (L == 0 && m == 1)
      e[1] = 0.0;
    else
      e[2] = 0.0;
  }
}

// Householder tridiagonalisation + QL (dsytrd3.h:27-102, dsyevq3.c:29-134).
static __device__ __noinline__ void eig3_ql(const Sym3& A, Eig3& out) {
  f64(&Q)[9] = out.q;
  f64(&w)[3] = out.w;
  f64 e[3];
#pragma unroll
  for (int i = 0; i < 9; ++i) Q[i] = 0.0;
  Q[0] = Q[4] = Q[8] = 1.0;

  const f64 h = sq(A.a01) + sq(A.a02);
  const f64 g = (A.a01 > f64(0.0)) ? -sqrt(h) : sqrt(h);
  e[0] = g;
  f64 f = g * A.a01;
  const f64 u1 = A.a01 - g;
  const f64 u2 = A.a02;
  f64 omega = h - f;
  if (omega > f64(0.0)) {
    omega = f64(1.0) / omega;
    f64 K = 0.0;
    f = A.a11 * u1 + A.a12 * u2;
    f64 q1 = omega * f;
    K += u1 * f;
    f = A.a12 * u1 + A.a22 * u2;
    f64 q2 = omega * f;
    K += u2 * f;
    K *= f64(0.5) * sq(omega);
    q1 = q1 - K * u1;
    q2 = q2 - K * u2;
    w[0] = A.a00;
    w[1] = A.a11 - f64(2.0) * q1 * u1;
    w[2] = A.a22 - f64(2.0) * q2 * u2;
    f = omega * u1;
    Q[4] = Q[4] - f * u1;
    Q[7] = Q[7] - f * u2;
    f = omega * u2;
    Q[5] = Q[5] - f * u1;
    Q[8] = Q[8] - f * u2;
    e[1] = A.a12 - q1 * u2 - u1 * q2;
  } else {
    w[0] = A.a00;
    w[1] = A.a11;
    w[2] = A.a22;
    e[1] = A.a12;
  }
  e[2] = 0.0;
  if (!ql_converge<0>(w, e, Q)) return;
  ql_converge<1>(w, e, Q);
}

// dsyevh3.c:64-133.
template <bool EXACT = false>
__device__ __forceinline__ void eig3_hybrid(const Sym3& A, Eig3& out) {
  f64(&w)[3] = out.w;
  eig3_values<EXACT>(A, w);

  f64 t = abs(w[0]);
  f64 u = abs(w[1]);
  if (u > t) t = u;
  u = abs(w[2]);
  if (u > t) t = u;
  u = (t < f64(1.0)) ? t : sq(t);
  const f64 error = f64(256.0 * 2.2204460492503131e-16) * sq(u);

  f64 q01 = A.a01 * A.a12 - A.a02 * A.a11;
  f64 q11 = A.a02 * A.a01 - A.a12 * A.a00;
  f64 q21 = sq(A.a01);

  f64 q00 = q01 + A.a02 * w[0];
  f64 q10 = q11 + A.a12 * w[0];
  f64 q20 = (A.a00 - w[0]) * (A.a11 - w[0]) - q21;
  f64 norm = sq(q00) + sq(q10) + sq(q20);
  bool fallback = norm <= error;
  if (!fallback) {
    norm = sqrt(f64(1.0) / norm);
    q00 = q00 * norm;
    q10 = q10 * norm;
    q20 = q20 * norm;

    q01 = q01 + A.a02 * w[1];
    q11 = q11 + A.a12 * w[1];
    q21 = (A.a00 - w[1]) * (A.a11 - w[1]) - q21;
    norm = sq(q01) + sq(q11) + sq(q21);
    fallback = norm <= error;
    if (!fallback) {
      norm = sqrt(f64(1.0) / norm);
      q01 = q01 * norm;
      q11 = q11 * norm;
      q21 = q21 * norm;
      out.q[0] = q00; out.q[3] = q10; out.q[6] = q20;
      out.q[1] = q01; out.q[4] = q11; out.q[7] = q21;
      out.q[2] = q10 * q21 - q20 * q11;
      out.q[5] = q20 * q01 - q00 * q21;
      out.q[8] = q00 * q11 - q10 * q01;
    }
  }
  out.ql = fallback;
  if (fallback) eig3_ql(A, out);
}

// Moments of a cell or of a grown region: CellSegmentStat's n, coord_sum_, variance_ (upper triangle).
struct Moments {
  int n;
  float s[3];
  float v[6];  // xx xy xz yy yz zz
};

struct PlaneFit {
  float mean[3];
  float normal[3];
  float d, mse, score;
};

// CellSegmentStat::fitPlane (cell_segment_stat.cpp:55-81) with mean = coord_sum_/nr_pts_ (:33,41).
template <bool EXACT = false>
__device__ __forceinline__ void fit_plane(const Moments& m, PlaneFit& out) {
  const float fn = static_cast<float>(m.n);
  // cov(i,j) = V(i,j) - (S_i*S_j)/n, each operation rounded to fp32 (cell_segment_stat.cpp:56)
  Sym3 A;
  A.a00 = static_cast<double>(__fsub_rn(m.v[0], __fdiv_rn(__fmul_rn(m.s[0], m.s[0]), fn)));
  A.a01 = static_cast<double>(__fsub_rn(m.v[1], __fdiv_rn(__fmul_rn(m.s[0], m.s[1]), fn)));
  A.a02 = static_cast<double>(__fsub_rn(m.v[2], __fdiv_rn(__fmul_rn(m.s[0], m.s[2]), fn)));
  A.a11 = static_cast<double>(__fsub_rn(m.v[3], __fdiv_rn(__fmul_rn(m.s[1], m.s[1]), fn)));
  A.a12 = static_cast<double>(__fsub_rn(m.v[4], __fdiv_rn(__fmul_rn(m.s[1], m.s[2]), fn)));
  A.a22 = static_cast<double>(__fsub_rn(m.v[5], __fdiv_rn(__fmul_rn(m.s[2], m.s[2]), fn)));

  Eig3 eg;
  eig3_hybrid<EXACT>(A, eg);

  // std::min_element / std::max_element: first of equal elements (cell_segment_stat.cpp:67-68)
  int imin = 0, imax = 0;
  if (eg.w[1] < eg.w[imin]) imin = 1;
  if (eg.w[2] < ((imin == 0) ? eg.w[0] : eg.w[1])) imin = 2;
  if (eg.w[0] < eg.w[1]) imax = 1;
  if (((imax == 0) ? eg.w[0] : eg.w[1]) < eg.w[2]) imax = 2;
  const f64 wmin = (imin == 0) ? eg.w[0] : (imin == 1 ? eg.w[1] : eg.w[2]);
  const f64 wmax = (imax == 0) ? eg.w[0] : (imax == 1 ? eg.w[1] : eg.w[2]);

  float v[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const f64 qi = (imin == 0) ? eg.q[3 * i] : (imin == 1 ? eg.q[3 * i + 1] : eg.q[3 * i + 2]);
    v[i] = __double2float_rn(qi.v);
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) out.mean[i] = __fdiv_rn(m.s[i], fn);

  float d = -dot3(out.mean[0], out.mean[1], out.mean[2], v[0], v[1], v[2]);
  const bool keep = d > 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) out.normal[i] = keep ? v[i] : -v[i];
  out.d = keep ? d : -d;
  out.mse = __double2float_rn((wmin / f64(static_cast<double>(m.n))).v);
  out.score = __double2float_rn((wmax / ((eg.w[0] + eg.w[1]) + eg.w[2])).v);
}

}  // namespace dpx
