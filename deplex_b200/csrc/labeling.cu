// labeling.cu -- stage 3 of the hot path: paint the merged cell labels onto the pixels.
//
// Replaces PlaneExtractor::Impl::toImageLabels (plane_extractor.cpp:455-470) and the all-zero early
// exit (plane_extractor.cpp:230-232: with no segments every cell label is 0).  Pure write stream:
// 4 B/pixel out, the cell-label table (4 B/cell) in.  One thread owns a group of 4 consecutive pixels
// of a cell row: the group's labels are identical for all `patch` image rows of the cell, so they are
// looked up once and stored `patch` times as 16-byte streaming stores.  Pixels right of / below the
// cell grid (width or height not a multiple of patch is rejected at create) do not exist.
#include "labeling.cuh"

namespace dpx {
namespace {

__device__ __forceinline__ void st_stream_i4(int4* p, int4 v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

constexpr int kLabelThreads = 128;

// grid: x = column groups, y = cell row, z = frame
__global__ void __launch_bounds__(kLabelThreads) labeling_kernel(const LabelArgs args) {
  pdl_wait();
  const Geometry& g = args.geom;
  const int p = g.patch;
  const int col = (blockIdx.x * kLabelThreads + threadIdx.x) * 4;
  if (col >= g.width) return;
  const int cell_row = blockIdx.y;
  int frame = blockIdx.z;
  if (args.todo != nullptr) {
    if (frame >= args.todo[0]) return;
    frame = args.todo[1 + frame];
  }
  const int32_t* cl = args.cell_label + static_cast<long long>(frame) * g.n_cells + static_cast<long long>(cell_row) * g.nh;
  int32_t* out = args.labels + static_cast<long long>(frame) * g.n_points +
                 static_cast<long long>(cell_row) * p * g.width + col;

  DPX_CHECK(frame >= 0 && frame < args.n_frames && cell_row < g.nv);
  int cq = col / p;
  int rem = col - cq * p;
  int lab[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    lab[i] = (col + i < g.width) ? __ldg(cl + cq) : 0;
    if (++rem == p) { rem = 0; ++cq; }
  }
  if (args.vec_ok) {
    const int4 v = make_int4(lab[0], lab[1], lab[2], lab[3]);
    for (int r = 0; r < p; ++r) st_stream_i4(reinterpret_cast<int4*>(out + static_cast<long long>(r) * g.width), v);
  } else {
    for (int r = 0; r < p; ++r)
      for (int i = 0; i < 4; ++i)
        if (col + i < g.width) out[static_cast<long long>(r) * g.width + i] = lab[i];
  }
}

// 8 labels per thread: two 16-byte loads, one 16-byte store
__global__ void __launch_bounds__(256) narrow_labels_kernel(const int32_t* __restrict__ in, uint16_t* __restrict__ out, long long n8,
                                                            long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i < n8) {
    const int4 a = __ldcs(reinterpret_cast<const int4*>(in) + 2 * i), b = __ldcs(reinterpret_cast<const int4*>(in) + 2 * i + 1);
    uint4 o;
    o.x = (static_cast<unsigned>(a.x) & 0xffffu) | (static_cast<unsigned>(a.y) << 16);
    o.y = (static_cast<unsigned>(a.z) & 0xffffu) | (static_cast<unsigned>(a.w) << 16);
    o.z = (static_cast<unsigned>(b.x) & 0xffffu) | (static_cast<unsigned>(b.y) << 16);
    o.w = (static_cast<unsigned>(b.z) & 0xffffu) | (static_cast<unsigned>(b.w) << 16);
    __stcs(reinterpret_cast<uint4*>(out) + i, o);
  }
  // scalar tail (n not a multiple of 8): the last block's first threads
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x < n - 8 * n8) {
    const long long j = 8 * n8 + threadIdx.x;
    out[j] = static_cast<uint16_t>(in[j]);
  }
}

}  // namespace

cudaError_t launch_narrow_labels(const int32_t* labels, uint16_t* out, long long n, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  const bool vec = (reinterpret_cast<uintptr_t>(labels) & 15u) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
  const long long n8 = vec ? n / 8 : 0;
  if (n - 8 * n8 > 256) {  // unaligned buffers: plain element-wise copy through the tail path is not enough
    return cudaErrorInvalidValue;
  }
  const long long blocks = (n8 + 255) / 256 > 0 ? (n8 + 255) / 256 : 1;
  narrow_labels_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(labels, out, n8, n);
  return cudaGetLastError();
}

cudaError_t launch_labeling(const LabelArgs& args, cudaStream_t stream) {
  const Geometry& g = args.geom;
  if (args.n_frames == 0 || g.n_points == 0) return cudaSuccess;
  if (g.n_cells == 0)  // enormous patch: no cells, all labels zero (plane_extractor.cpp:230-232)
    return cudaMemsetAsync(args.labels, 0, sizeof(int32_t) * g.n_points * args.n_frames, stream);
  const int groups = (g.width + 3) / 4;
  dim3 grid((groups + kLabelThreads - 1) / kLabelThreads, g.nv, args.todo ? labeling_deferred_frames(args.n_frames) : args.n_frames);
  if (const cudaError_t e = launch_dependent(labeling_kernel, dim3(grid), dim3(kLabelThreads), 0, stream, args); e != cudaSuccess) return e;
  return cudaGetLastError();
}

}  // namespace dpx
