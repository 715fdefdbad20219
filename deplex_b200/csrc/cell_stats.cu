// cell_stats.cu -- stage 1 of the hot path: cell-grid partitioning + per-cell statistics + plane fit
// + planarity test + histogram bin, fused into one bandwidth-bound pass over the organized cloud.
//
// Replaces, per frame (reference file:line):
//   CellGrid::CellGrid / cellContinuousOrganize      cell_grid.cpp:21-44,69-83   (no re-tiled copy is made:
//                                                     rows of a tile of cells are staged in shared memory)
//   CellSegment::CellSegment and its predicates      cell_segment.cpp:21-35,57-110
//   CellSegmentStat::CellSegmentStat / fitPlane      cell_segment_stat.cpp:29-35,55-81 (+ libs/dsyev)
//   NormalsHistogram's per-cell bin                  normals_histogram.cpp:31-45
//
// Two kernels share the per-cell arithmetic in cell_walk.cuh:
//
//  * cell_stats_stream_kernel (the fast path; compile-time patch size, 16-byte aligned input):
//    persistent CTAs, one per SM, 8 warps each.  A warp owns a tile of 32 horizontally adjacent cells
//    (one cell per lane) and streams the tile's image rows through a private ring of shared-memory
//    slots filled by cp.async.bulk (TMA bulk copies, one per contiguous row segment) that complete on
//    per-slot mbarriers.  The warp is its own producer: after the lanes have consumed a slot, lane 0
//    re-arms its mbarrier and issues the copy of the row RING slots ahead, so RING-1 slots of loads are
//    always in flight per warp while the lanes run the order-exact walk on the current one.  No
//    __syncthreads on the data path.
//
//  * cell_stats_tile_kernel (fallback: any supported patch size, unaligned input): one CTA stages a whole
//    tile of cells with plain loads, then one thread per cell walks it.
//
// The walk is sequential per cell on purpose: the reference's X^T X entries are sequential fp32 chains
// and its column sums an 8-lane strided tree (Eigen 3.4); labels only match if those orders are
// reproduced exactly.  A lane carries 6 + 24 independent accumulators, so the chains
// overlap in the FMA pipe and the kernel is bound by HBM, not by the dependency chains.
//
// Algorithmic bytes: 12 B/pixel read once; <= 87 B/cell written.
#include "cell_stats.cuh"

#include <type_traits>

#include "cell_walk.cuh"
#include "cell_walk_packed.cuh"

namespace dpx {
namespace {

__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// ---------------------------------------------------------------------------------------------------
// Fallback: one CTA = one tile of up to `tile_cells` cells of one cell-row of one frame.
// ---------------------------------------------------------------------------------------------------
template <int LAYOUT, int PATCH>
__global__ void __launch_bounds__(kCellStatsThreads) cell_stats_tile_kernel(const CellStatsArgs args) {
  extern __shared__ float4 smem_f4[];
  float* tile = reinterpret_cast<float*>(smem_f4);

  const Geometry& g = args.geom;
  const int p = PATCH ? PATCH : g.patch;
  const int tc = args.tile_cells;
  const int tw = tc * p;  // shared-memory row pitch in points

  int b = blockIdx.x;
  const int tix = b % args.tiles_per_strip;
  b /= args.tiles_per_strip;
  const int strip = b % g.nv;
  const int frame = b / g.nv;
  const int c0 = tix * tc;
  const int cnt = min(tc, g.nh - c0);
  const int seg = cnt * p;  // points per row segment

  const float* src = args.xyz + static_cast<long long>(frame) * 3 * g.n_points;
  const long long row0 = static_cast<long long>(strip) * p * g.width + static_cast<long long>(c0) * p;

  if (LAYOUT == kLayoutRowMajor) {
    const int row_floats = seg * 3;
    const bool vec = args.vec_ok && (row_floats % 4 == 0) && ((row0 * 3) % 4 == 0);
    if (vec) {
      const int nv4 = row_floats / 4;
      for (int i = 0; i < p; ++i) {
        const float4* s4 = reinterpret_cast<const float4*>(src + (row0 + static_cast<long long>(i) * g.width) * 3);
        float4* d4 = reinterpret_cast<float4*>(tile + i * tw * 3);
        for (int v = threadIdx.x; v < nv4; v += blockDim.x) d4[v] = ld_stream_f4(s4 + v);
      }
    } else {
      for (int i = 0; i < p; ++i) {
        const float* s1 = src + (row0 + static_cast<long long>(i) * g.width) * 3;
        float* d1 = tile + i * tw * 3;
        for (int v = threadIdx.x; v < row_floats; v += blockDim.x) d1[v] = __ldg(s1 + v);
      }
    }
  } else {
    const bool vec = args.vec_ok && (seg % 4 == 0) && (row0 % 4 == 0);
    for (int a = 0; a < 3; ++a)
      for (int i = 0; i < p; ++i) {
        const float* s1 = src + a * g.n_points + row0 + static_cast<long long>(i) * g.width;
        float* d1 = tile + (a * p + i) * tw;
        if (vec) {
          const float4* s4 = reinterpret_cast<const float4*>(s1);
          float4* d4 = reinterpret_cast<float4*>(d1);
          for (int v = threadIdx.x; v < seg / 4; v += blockDim.x) d4[v] = ld_stream_f4(s4 + v);
        } else {
          for (int v = threadIdx.x; v < seg; v += blockDim.x) d1[v] = __ldg(s1 + v);
        }
      }
  }
  __syncthreads();

  const int t = threadIdx.x;
  if (t < cnt) {
    CellRaw raw;
    if (PATCH) {
      constexpr int P = PATCH ? PATCH : 4;
      CellWalk<P> walk;
      walk.reset();
#pragma unroll
      for (int i = 0; i < P; ++i) {
        float x[P], y[P], z[P];
        load_cell_row<LAYOUT, P>(tile, tw, P, i, t, x, y, z);
        walk.row(i, x, y, z, args.thr.depth_discontinuity_threshold);
      }
      walk.finish(raw);
    } else {
      walk_cell_runtime<LAYOUT>(tile, tw, p, t, args.thr.depth_discontinuity_threshold, raw);
    }
    const long long cell = static_cast<long long>(frame) * g.n_cells + static_cast<long long>(strip) * g.nh + c0 + t;
    finish_cell(raw, args.thr, args.tables, cell);
  }
}

// ---------------------------------------------------------------------------------------------------
// Fast path: persistent, warp-private TMA-bulk row pipeline.
// ---------------------------------------------------------------------------------------------------

// rows staged per slot: the largest divisor of P whose slot stays <= 4 KB (else one row)
__host__ __device__ constexpr int rows_per_slot(int P) {
  int best = 1;
  for (int r = 1; r <= P; ++r)
    if (P % r == 0 && r * 32 * P * 12 <= 4096) best = r;
  return best;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
// The input is read exactly once: an L2 evict-first policy keeps it from displacing the per-cell tables this
// kernel writes, which the next stage reads back from L2.
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

// scalar walk behind the same row-at-a-time interface (odd patch sizes)
template <int LAYOUT, int P>
struct CellWalkScalar {
  CellWalk<P> w;
  __device__ __forceinline__ void reset() { w.reset(); }
  __device__ __forceinline__ void row(int i, const float* blk, int tw, int rows, int rr, int t, float disc_thr) {
    float x[P], y[P], z[P];
    load_cell_row<LAYOUT, P>(blk, tw, rows, rr, t, x, y, z);
    w.row(i, x, y, z, disc_thr);
  }
  __device__ __forceinline__ void finish(CellRaw& out) const { w.finish(out); }
};

template <int LAYOUT, int P>
struct WalkFor {
  // raw depth generates its points in registers and feeds them in the column-major pairing
  static constexpr int kWalkLayout = LAYOUT == kLayoutDepth16 ? kLayoutColMajor : LAYOUT;
  using type = typename std::conditional<P % 2 == 0, CellWalkPacked<kWalkLayout, P>, CellWalkScalar<kWalkLayout, P>>::type;
};

// where a warp's tile starts in the input, and where its cells go
struct TileDesc {
  const char* src;      // first sample of the tile's first row (component 0 for column-major input)
  long long cell0;      // frame * n_cells + strip * nh + c0
  unsigned seg_bytes;   // bytes of one component of one row segment
  int cnt;              // cells in the tile (32, fewer at the right edge); 0 = past the end
  int col0, row0;       // image column / row of the tile's first pixel
};

template <int LAYOUT, int P, int WARPS, int RING>
__global__ void __launch_bounds__(WARPS * 32, 1) cell_stats_stream_kernel(const CellStatsArgs args) {
  static_assert(LAYOUT != kLayoutDepth16 || P % 2 == 0, "the fused depth path needs the packed walk");
  constexpr int RPS = rows_per_slot(P);
  constexpr int SPT = P / RPS;               // stages (slots) per tile
  constexpr int TW = 32 * P;                 // points per staged row
  constexpr int SLOT_FLOATS = RPS * TW * 3;
  constexpr int ELEM = LAYOUT == kLayoutDepth16 ? 2 : 4;  // bytes per staged sample
  extern __shared__ float4 smem_f4[];
  __shared__ __align__(8) uint64_t bars[WARPS * RING];

  const Geometry& g = args.geom;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* ring = reinterpret_cast<float*>(smem_f4) + static_cast<size_t>(warp) * RING * SLOT_FLOATS;
  uint64_t* bar = bars + warp * RING;

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < RING; ++s) mbar_init(bar + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  const uint64_t policy = l2_evict_first_policy();
  const long long total_tiles = static_cast<long long>(args.n_frames) * g.nv * args.tiles_per_strip;
  const long long gw = static_cast<long long>(blockIdx.x) * WARPS + warp;
  const long long total_warps = static_cast<long long>(gridDim.x) * WARPS;
  if (gw >= total_tiles) return;
  const long long my_tiles = (total_tiles - gw + total_warps - 1) / total_warps;
  // bytes between image rows / between the component planes of a column-major cloud
  const long long row_bytes = static_cast<long long>(g.width) * (LAYOUT == kLayoutRowMajor ? 12 : ELEM);
  const long long plane_bytes = g.n_points * 4;
  const char* base = LAYOUT == kLayoutDepth16 ? reinterpret_cast<const char*>(args.depth) : reinterpret_cast<const char*>(args.xyz);
  const long long frame_bytes = g.n_points * (LAYOUT == kLayoutDepth16 ? 2 : 12);

  // tile n of this warp (the divisions run once per tile, not once per stage)
  auto locate = [&](long long n, TileDesc& d) {
    if (n >= my_tiles) {
      d.cnt = 0;
      return;
    }
    long long tile = gw + n * total_warps;
    const int tix = static_cast<int>(tile % args.tiles_per_strip);
    tile /= args.tiles_per_strip;
    const int strip = static_cast<int>(tile % g.nv);
    const long long frame = tile / g.nv;
    const int c0 = tix * 32;
    d.cnt = min(32, g.nh - c0);
    d.seg_bytes = static_cast<unsigned>(d.cnt) * P * ELEM;
    d.cell0 = frame * g.n_cells + static_cast<long long>(strip) * g.nh + c0;
    d.col0 = c0 * P;
    d.row0 = strip * P;
    const long long px0 = static_cast<long long>(strip) * P * g.width + static_cast<long long>(c0) * P;
    d.src = base + frame * frame_bytes + px0 * (LAYOUT == kLayoutRowMajor ? 12 : ELEM);
  };
  // lane 0: arm the slot's mbarrier and issue the bulk copies of stage `st` of tile `d`
  auto issue = [&](const TileDesc& d, int st, int slot) {
    if (d.cnt == 0 || lane != 0) return;
    char* dst = reinterpret_cast<char*>(ring + slot * SLOT_FLOATS);
    const char* src = d.src + static_cast<long long>(st) * RPS * row_bytes;
    if (LAYOUT == kLayoutRowMajor) {
      mbar_expect_tx(bar + slot, d.seg_bytes * 3 * RPS);
#pragma unroll
      for (int rr = 0; rr < RPS; ++rr) bulk_g2s(dst + rr * TW * 12, src + rr * row_bytes, d.seg_bytes * 3, bar + slot, policy);
    } else if (LAYOUT == kLayoutColMajor) {
      mbar_expect_tx(bar + slot, d.seg_bytes * 3 * RPS);
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int rr = 0; rr < RPS; ++rr)
          bulk_g2s(dst + (a * RPS + rr) * TW * 4, src + a * plane_bytes + rr * row_bytes, d.seg_bytes, bar + slot, policy);
    } else {
      mbar_expect_tx(bar + slot, d.seg_bytes * RPS);
#pragma unroll
      for (int rr = 0; rr < RPS; ++rr) bulk_g2s(dst + rr * TW * 2, src + rr * row_bytes, d.seg_bytes, bar + slot, policy);
    }
  };

  // the issue cursor runs RING stages ahead of the consumer
  TileDesc cur, ahead;
  locate(0, cur);
  ahead = cur;
  long long ahead_n = 0;
  int ahead_st = 0;
  auto issue_next = [&](int slot) {
    issue(ahead, ahead_st, slot);
    if (++ahead_st == SPT) {
      ahead_st = 0;
      locate(++ahead_n, ahead);
    }
  };
#pragma unroll
  for (int s = 0; s < RING; ++s) issue_next(s);

  int slot = 0;
  unsigned phase = 0;
  for (long long n = 0; n < my_tiles; ++n) {
    typename WalkFor<LAYOUT, P>::type walk;
    walk.reset();
#pragma unroll
    for (int st = 0; st < SPT; ++st) {
      mbar_wait(bar + slot, phase);
      const float* blk = ring + slot * SLOT_FLOATS;
      if (lane < cur.cnt) {
#pragma unroll
        for (int rr = 0; rr < RPS; ++rr) {
          if constexpr (LAYOUT == kLayoutDepth16) {
            walk.row_depth(st * RPS + rr, reinterpret_cast<const uint16_t*>(blk), TW, rr, lane,
                           args.thr.depth_discontinuity_threshold, static_cast<float>(cur.col0 + lane * P),
                           static_cast<float>(cur.row0 + st * RPS + rr), args.pin);
          } else {
            walk.row(st * RPS + rr, blk, TW, RPS, rr, lane, args.thr.depth_discontinuity_threshold);
          }
        }
      }
      __syncwarp();        // every lane is done reading the slot ...
      issue_next(slot);    // ... before the async proxy refills it
      if (++slot == RING) {
        slot = 0;
        phase ^= 1u;
      }
    }
    if (lane < cur.cnt) {
      CellRaw raw;
      walk.finish(raw);
      finish_cell(raw, args.thr, args.tables, cur.cell0 + lane);
    }
    locate(n + 1, cur);
  }
}

template <int LAYOUT>
cudaError_t launch_tile(const CellStatsArgs& a, int grid, size_t smem, cudaStream_t st) {
#define DPX_CASE(P)                                                                                               \
  case P:                                                                                                         \
    cudaFuncSetAttribute(cell_stats_tile_kernel<LAYOUT, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    cell_stats_tile_kernel<LAYOUT, P><<<grid, kCellStatsThreads, smem, st>>>(a);                                  \
    break;
  switch (a.geom.patch) {
    DPX_CASE(4)
    DPX_CASE(5)
    DPX_CASE(6)
    DPX_CASE(8)
    DPX_CASE(10)
    default:
      cudaFuncSetAttribute(cell_stats_tile_kernel<LAYOUT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cell_stats_tile_kernel<LAYOUT, 0><<<grid, kCellStatsThreads, smem, st>>>(a);
      break;
  }
#undef DPX_CASE
  return cudaGetLastError();
}

template <int LAYOUT, int P, int WARPS, int RING>
cudaError_t launch_stream_pw(const CellStatsArgs& a, cudaStream_t st) {
  constexpr size_t smem = static_cast<size_t>(WARPS) * RING * rows_per_slot(P) * 32 * P * 12;
  static_assert(smem <= 200 * 1024, "ring does not fit");
  cudaFuncSetAttribute(cell_stats_stream_kernel<LAYOUT, P, WARPS, RING>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const long long total_tiles = static_cast<long long>(a.n_frames) * a.geom.nv * a.tiles_per_strip;
  const long long ctas = (total_tiles + WARPS - 1) / WARPS;
  const int grid = static_cast<int>(ctas < a.sm_count ? ctas : a.sm_count);
  cell_stats_stream_kernel<LAYOUT, P, WARPS, RING><<<grid, WARPS * 32, smem, st>>>(a);
  return cudaGetLastError();
}

template <int LAYOUT, int P>
cudaError_t launch_stream_p(const CellStatsArgs& a, cudaStream_t st) {
  switch (a.stream_warps) {
    case 8: return launch_stream_pw<LAYOUT, P, 8, 4>(a, st);
    case 16: return launch_stream_pw<LAYOUT, P, 16, 3>(a, st);
    case 12: return launch_stream_pw<LAYOUT, P, 12, 4>(a, st);
    default: return launch_stream_pw<LAYOUT, P, 16, 3>(a, st);
  }
}

template <int LAYOUT>
bool launch_stream(const CellStatsArgs& a, cudaStream_t st, cudaError_t* err) {
  switch (a.geom.patch) {
    case 4: *err = launch_stream_p<LAYOUT, 4>(a, st); return true;
    case 6: *err = launch_stream_p<LAYOUT, 6>(a, st); return true;
    case 8: *err = launch_stream_p<LAYOUT, 8>(a, st); return true;
    case 10: *err = launch_stream_p<LAYOUT, 10>(a, st); return true;
    case 5:
      if constexpr (LAYOUT != kLayoutDepth16) {
        *err = launch_stream_p<LAYOUT, 5>(a, st);
        return true;
      }
      return false;
    default: return false;
  }
}

}  // namespace

int cell_stats_tile_cells(int patch, int nh) {
  // fallback kernel: largest power-of-two tile (<= one thread per cell per CTA) whose staging buffer stays
  // <= 40 KB, so that several CTAs are resident per SM and loads of one overlap the walk of another
  const size_t per_cell = static_cast<size_t>(patch) * patch * 12;
  int tc = kCellStatsThreads;
  while (tc > 4 && tc * per_cell > 40 * 1024) tc >>= 1;
  while (tc > 4 && tc / 2 >= nh) tc >>= 1;
  return tc;
}

bool cell_stats_stream_eligible(const CellStatsArgs& a) {
  const Geometry& g = a.geom;
  if (!a.vec_ok || a.force_tile_kernel) return false;
  const int p = g.patch;
  if (!(p == 4 || p == 5 || p == 6 || p == 8 || p == 10)) return false;
  // every bulk copy must start on a 16-byte boundary and move a multiple of 16 bytes
  if ((32 * p) % 4 != 0) return false;
  const int tail = g.nh % 32;
  if ((tail * p) % 4 != 0) return false;
  return true;
}

bool cell_stats_depth_eligible(const CellStatsArgs& a) {
  const Geometry& g = a.geom;
  const int p = g.patch;
  if (a.force_tile_kernel || !(p == 4 || p == 6 || p == 8 || p == 10)) return false;
  // 2-byte samples: rows, frames and row segments must start on 16-byte boundaries and be multiples of 16 bytes
  if (reinterpret_cast<uintptr_t>(a.depth) % 16 != 0 || g.width % 8 != 0 || g.n_points % 8 != 0) return false;
  const int tail = g.nh % 32;
  if ((tail * p) % 8 != 0) return false;
  return true;
}

cudaError_t launch_cell_stats(const CellStatsArgs& args_in, cudaStream_t stream) {
  CellStatsArgs a = args_in;
  const Geometry& g = a.geom;
  if (g.n_cells == 0 || a.n_frames == 0) return cudaSuccess;
  if (a.layout == kLayoutDepth16) {
    if (!cell_stats_depth_eligible(a)) return cudaErrorInvalidValue;  // the caller converts to points instead
    a.tiles_per_strip = (g.nh + 31) / 32;
    cudaError_t err = cudaSuccess;
    return launch_stream<kLayoutDepth16>(a, stream, &err) ? err : cudaErrorInvalidValue;
  }
  if (cell_stats_stream_eligible(a)) {
    a.tiles_per_strip = (g.nh + 31) / 32;
    cudaError_t err = cudaSuccess;
    const bool ok = a.layout == kLayoutRowMajor ? launch_stream<kLayoutRowMajor>(a, stream, &err)
                                                : launch_stream<kLayoutColMajor>(a, stream, &err);
    if (ok) return err;
  }
  a.tiles_per_strip = (g.nh + a.tile_cells - 1) / a.tile_cells;
  const size_t smem = static_cast<size_t>(a.tile_cells) * g.patch * g.patch * 12;
  const long long grid = static_cast<long long>(a.n_frames) * g.nv * a.tiles_per_strip;
  if (grid > 0x7fffffffLL) return cudaErrorInvalidValue;
  return a.layout == kLayoutRowMajor ? launch_tile<kLayoutRowMajor>(a, static_cast<int>(grid), smem, stream)
                                     : launch_tile<kLayoutColMajor>(a, static_cast<int>(grid), smem, stream);
}

}  // namespace dpx
