// exact_math.cuh -- round-once arithmetic wrappers.
//
// Label parity with the reference hinges on reproducing its fp32/fp64 results bit for bit
// (SURVEY.md section 7, H1): the reference's default x86-64 build has no FMA, so every + - * /
// rounds exactly once.  nvcc contracts a*b+c into FMA by default; the wrappers below route every
// operator through the *_rn intrinsics, which are never contracted, so parity does not depend on
// a compiler flag (the build passes -fmad=false as well).
#pragma once
#include <cuda_runtime.h>

namespace dpx {

struct f32 {
  float v;
  __host__ __device__ f32() {}
  __host__ __device__ f32(float x) : v(x) {}
};
__device__ __forceinline__ f32 operator+(f32 a, f32 b) { return __fadd_rn(a.v, b.v); }
__device__ __forceinline__ f32 operator-(f32 a, f32 b) { return __fsub_rn(a.v, b.v); }
__device__ __forceinline__ f32 operator*(f32 a, f32 b) { return __fmul_rn(a.v, b.v); }
__device__ __forceinline__ f32 operator/(f32 a, f32 b) { return __fdiv_rn(a.v, b.v); }
__device__ __forceinline__ f32 operator-(f32 a) { return -a.v; }
__device__ __forceinline__ f32 sqrt(f32 a) { return __fsqrt_rn(a.v); }

struct f64 {
  double v;
  __host__ __device__ f64() {}
  __host__ __device__ f64(double x) : v(x) {}
};
__device__ __forceinline__ f64 operator+(f64 a, f64 b) { return __dadd_rn(a.v, b.v); }
__device__ __forceinline__ f64 operator-(f64 a, f64 b) { return __dsub_rn(a.v, b.v); }
__device__ __forceinline__ f64 operator*(f64 a, f64 b) { return __dmul_rn(a.v, b.v); }
__device__ __forceinline__ f64 operator/(f64 a, f64 b) { return __ddiv_rn(a.v, b.v); }
__device__ __forceinline__ f64 operator-(f64 a) { return -a.v; }
__device__ __forceinline__ f64& operator+=(f64& a, f64 b) { a = a + b; return a; }
__device__ __forceinline__ f64& operator-=(f64& a, f64 b) { a = a - b; return a; }
__device__ __forceinline__ f64& operator*=(f64& a, f64 b) { a = a * b; return a; }
__device__ __forceinline__ bool operator>(f64 a, f64 b) { return a.v > b.v; }
__device__ __forceinline__ bool operator<(f64 a, f64 b) { return a.v < b.v; }
__device__ __forceinline__ bool operator<=(f64 a, f64 b) { return a.v <= b.v; }
__device__ __forceinline__ bool operator>=(f64 a, f64 b) { return a.v >= b.v; }
__device__ __forceinline__ bool operator==(f64 a, f64 b) { return a.v == b.v; }
__device__ __forceinline__ f64 sqrt(f64 a) { return __dsqrt_rn(a.v); }
__device__ __forceinline__ f64 abs(f64 a) { return ::fabs(a.v); }
__device__ __forceinline__ f64 sq(f64 a) { return a * a; }

// Fixed-size-3 dot product in Eigen's unrolled order a0*b0 + (a1*b1 + a2*b2)
// (Eigen 3.4 redux_novec_unroller<0,3>; call sites plane_extractor.cpp:380-381,407-410,
// cell_segment_stat.cpp:74).
__device__ __forceinline__ float dot3(float a0, float a1, float a2, float b0, float b1, float b2) {
  return __fadd_rn(__fmul_rn(a0, b0), __fadd_rn(__fmul_rn(a1, b1), __fmul_rn(a2, b2)));
}

}  // namespace dpx
