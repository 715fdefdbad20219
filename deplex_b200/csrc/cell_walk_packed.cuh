// cell_walk_packed.cuh -- the order-exact moment walk of cell_walk.cuh restated on Blackwell's packed fp32
// pipe (add.rn.f32x2 -> SASS FADD2: two IEEE round-to-nearest additions per issue slot).
//
// Same arithmetic as CellWalk<P> (reference: cell_segment.cpp:57-91, cell_segment_stat.cpp:29-35; Eigen 3.4
// summation orders); only the instruction selection differs:
//   * the six X^T X chains advance as three packed additions per point: (xx,xy) (xz,yy) (yz,zz).  The
//     products stay scalar mul.rn.f32: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under
//     --fmad=false, which would round once instead of twice;
//   * the column sums keep Eigen's 2 x 4 packet accumulators, i.e. one accumulator per (k mod 8), laid out so
//     that the 8-byte pairs coming out of shared memory are added as they are (row-major input: (x_k,y_k)
//     (z_k,x_k+1) (y_k+1,z_k+1); column-major input: (x_k,x_k+1) ...) -- no register shuffling;
//   * the valid-point count is a packed fp32 sum of FSET results (exact: integers below 2^24).
// Accumulators that Eigen initialises by assignment start at -0.0f here: (-0) + v == v for every v,
// including both zeros, so "add" and "assign" coincide bit for bit.
// Requires an even patch size (so that P*P % 4 == 0 and every column is 16-byte aligned in Eigen's
// cell-contiguous temporary, S = 0) -- 4, 6, 8, 10.
#pragma once
#include "cell_walk.cuh"

namespace dpx {

typedef unsigned long long u64;

__device__ __forceinline__ u64 pk2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ float lo2(u64 v) { float a, b; upk2(v, a, b); return a; }
__device__ __forceinline__ float hi2(u64 v) { float a, b; upk2(v, a, b); return b; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float valid_f(float z) { return z > 0.f ? 1.0f : 0.0f; }  // FSET.BF.GT

constexpr u64 kNegZero2 = 0x8000000080000000ull;

template <int LAYOUT, int P>
struct CellWalkPacked {
  static_assert(P % 2 == 0, "packed walk needs an even patch size");
  static constexpr int N = P * P;
  static constexpr int kA2 = (N / 8) * 8;       // entries covered by the two packet accumulators
  static constexpr bool kRem = (N % 8) == 4;    // a trailing 4-wide remainder packet exists
  static_assert(kA2 >= 8, "patch too small");

  u64 vA, vB, vC;  // (xx,xy) (xz,yy) (yz,zz)
  // column-sum accumulators, pair index m = (k mod 8) / 2 for even k:
  //   row-major:    s0[m] = (sx[2m], sy[2m])    s1[m] = (sz[2m], sx[2m+1])   s2[m] = (sy[2m+1], sz[2m+1])
  //   column-major: s0[m] = (sx[2m], sx[2m+1])  s1[m] = (sy[2m], sy[2m+1])   s2[m] = (sz[2m],   sz[2m+1])
  u64 s0[4], s1[4], s2[4];
  u64 r0[2], r1[2], r2[2];  // remainder packet, same pairing
  u64 vv;                   // valid counts of even / odd points
  int hcnt, vcnt;
  float hprev, vprev;
  float first[3], last[3];

  __device__ __forceinline__ void reset() {
    vA = vB = vC = 0ull;
    vv = 0ull;
#pragma unroll
    for (int m = 0; m < 4; ++m) s0[m] = s1[m] = s2[m] = kNegZero2;
    r0[0] = r0[1] = r1[0] = r1[1] = r2[0] = r2[1] = kNegZero2;
    hcnt = vcnt = 0;
    hprev = vprev = 0.f;
  }

  // One pair of consecutive points (k, k + 1) of image row i: everything order-sensitive happens here.
  // l0/l1/l2 are the same six values in the pairing of the accumulators (see s0/s1/s2 above).
  __device__ __forceinline__ void pair(int i, int jj, float xe, float ye, float ze, float xo, float yo, float zo, u64 l0,
                                       u64 l1, u64 l2, float disc_thr) {
    const int k = i * P + 2 * jj;  // even point of the pair; k + 1 is in the same row and the same region
    if (k == 0) { first[0] = xe; first[1] = ye; first[2] = ze; }
    if (k + 1 == N - 1) { last[0] = xo; last[1] = yo; last[2] = zo; }
    // X^T X chains, point k then point k + 1 (cell_segment_stat.cpp:32)
    vA = add2(pk2(__fmul_rn(xe, xe), __fmul_rn(xe, ye)), vA);
    vB = add2(pk2(__fmul_rn(xe, ze), __fmul_rn(ye, ye)), vB);
    vC = add2(pk2(__fmul_rn(ye, ze), __fmul_rn(ze, ze)), vC);
    vA = add2(pk2(__fmul_rn(xo, xo), __fmul_rn(xo, yo)), vA);
    vB = add2(pk2(__fmul_rn(xo, zo), __fmul_rn(yo, yo)), vB);
    vC = add2(pk2(__fmul_rn(yo, zo), __fmul_rn(zo, zo)), vC);
    // column sums (cell_segment_stat.cpp:31)
    if (k < kA2) {
      const int m = (k & 7) >> 1;
      s0[m] = add2(s0[m], l0); s1[m] = add2(s1[m], l1); s2[m] = add2(s2[m], l2);
    } else {
      const int m = (k - kA2) >> 1;
      r0[m] = l0; r1[m] = l1; r2[m] = l2;
    }
    // hasValidPoints (cell_segment.cpp:57-60)
    vv = add2(vv, pk2(valid_f(ze), valid_f(zo)));
    // isHorizontalContinuous: indices [N/2, N/2 + P) = row P/2 (cell_segment.cpp:62-76)
    if (i == P / 2) {
      if (jj == 0) hprev = ze;
      scan_step(ze, hprev, hcnt, disc_thr);
      scan_step(zo, hprev, hcnt, disc_thr);
    }
    // isVerticalContinuous: column P/2 of every row (cell_segment.cpp:78-91)
    if (2 * jj == P / 2 || 2 * jj + 1 == P / 2) {
      const float zc = (2 * jj == P / 2) ? ze : zo;
      if (i == 0) vprev = zc;
      scan_step(zc, vprev, vcnt, disc_thr);
    }
  }

  // Consume image row i (compile-time after unrolling) of cell `t` from a staged block.
  // Row-major block: blk[(rr * tw + col) * 3 + a]; column-major block: blk[(a * rows + rr) * tw + col].
  __device__ __forceinline__ void row(int i, const float* blk, int tw, int rows, int rr, int t, float disc_thr) {
    const u64* p0;
    const u64* p1 = nullptr;
    const u64* p2 = nullptr;
    if (LAYOUT == kLayoutRowMajor) {
      p0 = reinterpret_cast<const u64*>(blk + (rr * tw + t * P) * 3);
    } else {
      p0 = reinterpret_cast<const u64*>(blk + (0 * rows + rr) * tw + t * P);
      p1 = reinterpret_cast<const u64*>(blk + (1 * rows + rr) * tw + t * P);
      p2 = reinterpret_cast<const u64*>(blk + (2 * rows + rr) * tw + t * P);
    }
#pragma unroll
    for (int jj = 0; jj < P / 2; ++jj) {
      u64 l0, l1, l2;
      float xe, ye, ze, xo, yo, zo;
      if (LAYOUT == kLayoutRowMajor) {
        l0 = p0[3 * jj]; l1 = p0[3 * jj + 1]; l2 = p0[3 * jj + 2];
        upk2(l0, xe, ye); upk2(l1, ze, xo); upk2(l2, yo, zo);
      } else {
        l0 = p0[jj]; l1 = p1[jj]; l2 = p2[jj];
        upk2(l0, xe, xo); upk2(l1, ye, yo); upk2(l2, ze, zo);
      }
      pair(i, jj, xe, ye, ze, xo, yo, zo, l0, l1, l2, disc_thr);
    }
  }

  // Same, from raw depth: the staged block holds uint16 samples, blk16[rr * tw + col], and the points are
  // DepthImage::toPointCloud's (depth_image.cpp:64-73): z = float(raw), x = ((col - cx) * z) / fx,
  // y = ((row - cy) * z) / fy, every step rounded to fp32.  Uses the column-major pairing (LAYOUT must be
  // kLayoutColMajor).  `colf` = image column of the cell's first pixel, `rowf` = image row, both exact in fp32.
  __device__ __forceinline__ void row_depth(int i, const uint16_t* blk16, int tw, int rr, int t, float disc_thr, float colf,
                                            float rowf, const Pinhole& k) {
    const uint32_t* pz = reinterpret_cast<const uint32_t*>(blk16 + rr * tw + t * P);
    const float ry = __fsub_rn(rowf, k.cy);
#pragma unroll
    for (int jj = 0; jj < P / 2; ++jj) {
      const uint32_t w = pz[jj];
      const float ze = static_cast<float>(w & 0xffffu), zo = static_cast<float>(w >> 16);
      const float xe = __fdiv_rn(__fmul_rn(__fsub_rn(__fadd_rn(colf, static_cast<float>(2 * jj)), k.cx), ze), k.fx);
      const float xo = __fdiv_rn(__fmul_rn(__fsub_rn(__fadd_rn(colf, static_cast<float>(2 * jj + 1)), k.cx), zo), k.fx);
      const float ye = __fdiv_rn(__fmul_rn(ry, ze), k.fy);
      const float yo = __fdiv_rn(__fmul_rn(ry, zo), k.fy);
      pair(i, jj, xe, ye, ze, xo, yo, zo, pk2(xe, xo), pk2(ye, yo), pk2(ze, zo), disc_thr);
    }
  }

  __device__ __forceinline__ void finish(CellRaw& out) const {
    float sx[8], sy[8], sz[8], rx[4], ry[4], rz[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      if (LAYOUT == kLayoutRowMajor) {
        upk2(s0[m], sx[2 * m], sy[2 * m]); upk2(s1[m], sz[2 * m], sx[2 * m + 1]); upk2(s2[m], sy[2 * m + 1], sz[2 * m + 1]);
      } else {
        upk2(s0[m], sx[2 * m], sx[2 * m + 1]); upk2(s1[m], sy[2 * m], sy[2 * m + 1]); upk2(s2[m], sz[2 * m], sz[2 * m + 1]);
      }
    }
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      if (LAYOUT == kLayoutRowMajor) {
        upk2(r0[m], rx[2 * m], ry[2 * m]); upk2(r1[m], rz[2 * m], rx[2 * m + 1]); upk2(r2[m], ry[2 * m + 1], rz[2 * m + 1]);
      } else {
        upk2(r0[m], rx[2 * m], rx[2 * m + 1]); upk2(r1[m], ry[2 * m], ry[2 * m + 1]); upk2(r2[m], rz[2 * m], rz[2 * m + 1]);
      }
    }
    // Eigen redux tail: a0 += a1; a0 += remainder packet; predux = (a0[0] + a0[2]) + (a0[1] + a0[3])
    auto reduce = [](const float (&s)[8], const float (&r)[4]) {
      float q[4];
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        q[l] = __fadd_rn(s[l], s[l + 4]);
        if (kRem) q[l] = __fadd_rn(q[l], r[l]);
      }
      return __fadd_rn(__fadd_rn(q[0], q[2]), __fadd_rn(q[1], q[3]));
    };
    out.m.n = N;
    out.m.s[0] = reduce(sx, rx); out.m.s[1] = reduce(sy, ry); out.m.s[2] = reduce(sz, rz);
    upk2(vA, out.m.v[0], out.m.v[1]); upk2(vB, out.m.v[2], out.m.v[3]); upk2(vC, out.m.v[4], out.m.v[5]);
    out.valid_cnt = __float2int_rn(__fadd_rn(lo2(vv), hi2(vv)));
    out.hcnt = hcnt; out.vcnt = vcnt;
#pragma unroll
    for (int a = 0; a < 3; ++a) { out.first[a] = first[a]; out.last[a] = last[a]; }
  }
};

}  // namespace dpx
