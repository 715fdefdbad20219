// seed_sort.cuh -- stage 2a': the seed order of every frame, computed up front.
//
// createPlaneSegments (plane_extractor.cpp:302-316) picks, seed after seed, the most frequent histogram bin and, among
// the still-unassigned cells of that bin, the one with the smallest MSE (first strict minimum in ascending cell order,
// i.e. ties go to the smallest cell id).  Cells only ever LEAVE the unassigned set, so if a bin's cells are listed once
// in (MSE, cell id) order, the seed of that bin is always "the first entry that is still unassigned", and a cursor per
// bin that never moves backwards finds it in amortised O(1) -- instead of a linear minimum scan over the bin's members
// for every seed, which is what made noisy fine grids slow (thousands of one-cell regions out of bins with thousands of
// members: 80 % of the region-growing time at 1920x1080 / patch 5, profiles/r02a_frame_profiles.txt).
//
// This kernel sorts the cells of each frame by the 64-bit key
//     [ histogram bin : 15 ][ MSE as an order-preserving uint32 : 32 ][ cell id : 17 ]
// in two stages, so that a single large frame is not sorted by a single SM:
//   1. seed_chunk_sort_kernel: one CTA of 1024 threads per chunk of 4096 cells, least-significant-digit radix sort in
//      shared memory (8-bit digits over bits 17..63; the cell id needs no pass: the input is in cell order and every pass
//      is stable);
//   2. seed_rank_merge_kernel: keys are unique, so a key's final position is its index in its own chunk plus its lower
//      bound in every other chunk of the frame -- one thread per key, binary searches in L2-resident memory.
// Cells that are not planar get the largest bin / MSE fields and end up behind everything.  Only frames too large for the
// all-shared-memory mode of region growing are sorted (a few thousand cells and up): below that a minimum scan over the
// bin's members is cheaper than any sort.  Integer work on 8 bytes per cell: bit-exact by construction.
#pragma once
#include <cstring>

#include "common.cuh"

namespace dpx {

constexpr int kSeedCellBits = 17;                       // n_cells <= 131071 (checked at dpx_create)
constexpr unsigned long long kSeedCellMask = (1ull << kSeedCellBits) - 1ull;
constexpr int kSeedBinShift = kSeedCellBits + 32;       // bin id in bits 49..63 (bins per frame <= 32761 < 2^15)

// float -> uint32 whose unsigned order is the float order (negative values first); NaN sorts last and is never a seed
__host__ __device__ inline uint32_t seed_mse_order(float m) {
  if (m != m) return 0xffffffffu;
  uint32_t fb;
#ifdef __CUDA_ARCH__
  fb = __float_as_uint(m);
#else
  static_assert(sizeof(float) == 4, "float");
  memcpy(&fb, &m, 4);
#endif
  return fb ^ ((fb >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float seed_mse_from_order(uint32_t ord) {
  return __uint_as_float(ord ^ ((ord >> 31) ? 0x80000000u : 0xffffffffu));
}

namespace {

constexpr int kSortThreads = 1024;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortItems = 4;         // keys per thread: a chunk is kSortItems sub-tiles of kSortThreads keys
constexpr int kSortChunk = kSortThreads * kSortItems;  // 4096 cells sorted by one CTA in shared memory
constexpr int kSortPasses = 6;        // digits of 8 bits over key bits 17..64 (the last digit has 7 bits)
constexpr int kSortRows = kSortItems * kSortWarps;  // (sub-tile, warp) pairs of a chunk, in key order

struct SortShared {
  unsigned long long keys[2][kSortChunk];  // ping-pong
  uint32_t hist[kSortPasses][256];    // digit histograms of all passes (taken once: the key multiset does not change)
  uint32_t offs[256];                 // where every digit's keys start in the output
  uint32_t sub_offs[kSortItems][256]; // ... and where those of sub-tile j start
  uint32_t warp_tot[kSortWarps];
  int skip;
  uint16_t wrank[kSortRows][256];     // keys with this digit in earlier warps of the same sub-tile
};

#ifdef DPX_SORT_PROBE
__device__ long long g_sort_probe[16];
#define SORT_PROBE(slot) do { if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) { const long long t__ = clock64(); g_sort_probe[slot] += t__ - probe_t; probe_t = t__; } } while (0)
#else
#define SORT_PROBE(slot) do { } while (0)
#endif

// lanes of this warp holding the same 8-bit digit (restricted to `valid` lanes)
__device__ __forceinline__ unsigned match_digit(unsigned digit, unsigned valid) {
  return __match_any_sync(0xffffffffu, digit) & valid;
}

// one histogram update per warp and digit value: a single atomic when the whole warp holds the same digit (the usual case
// for the high bytes), else one conflict-free atomic per lane
__device__ __forceinline__ void hist_add(uint32_t* hist, unsigned digit, bool valid, unsigned vm, int lane) {
  const unsigned first = __shfl_sync(0xffffffffu, digit, __ffs(vm | 0x80000000u) - 1);
  if (__all_sync(0xffffffffu, !valid || digit == first)) {
    if (vm != 0 && lane == __ffs(vm) - 1) atomicAdd(&hist[digit], __popc(vm));
  } else if (valid) {
    atomicAdd(&hist[digit], 1u);
  }
}

// Stage 1: CTA (chunk, frame) sorts cells [chunk * kSortChunk, ...) of its frame by key, entirely in shared memory, and
// writes the sorted chunk to keys_tmp at the same positions.
__global__ void __launch_bounds__(kSortThreads) seed_chunk_sort_kernel(const int16_t* __restrict__ bin, const float* __restrict__ mse,
                                                                       unsigned long long* keys_tmp, int n_cells) {
  extern __shared__ unsigned long long sort_smem[];
  SortShared& s = *reinterpret_cast<SortShared*>(sort_smem);
#ifdef DPX_SORT_PROBE
  long long probe_t = clock64();
#endif
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const long long fc = static_cast<long long>(blockIdx.y) * n_cells;
  const int first = blockIdx.x * kSortChunk;
  const int n = min(kSortChunk, n_cells - first);
  bin += fc + first;
  mse += fc + first;
  keys_tmp += fc + first;
  unsigned long long* cur = s.keys[0];
  unsigned long long* nxt = s.keys[1];

  for (int i = tid; i < kSortPasses * 256; i += kSortThreads) (&s.hist[0][0])[i] = 0;
  __syncthreads();
  // ---- keys in cell order + the digit histograms of all passes ------------------------------------------------
  {
    int b[kSortItems];
    float m[kSortItems];
#pragma unroll
    for (int j = 0; j < kSortItems; ++j) {
      const int i = j * kSortThreads + tid;
      b[j] = i < n ? static_cast<int>(bin[i]) : -1;
      m[j] = i < n ? mse[i] : 0.f;  // (written for valid cells only: ignored below when the cell is not planar)
    }
#pragma unroll
    for (int j = 0; j < kSortItems; ++j) {
      const int i = j * kSortThreads + tid;
      const bool valid = i < n;
      const unsigned long long cell = static_cast<unsigned long long>(first + i);
      // cells that are not planar: the largest bin value and MSE order, so that they end up behind everything, still unique
      unsigned long long key = (0x7fffull << kSeedBinShift) | (0xffffffffull << kSeedCellBits) | cell;
      if (b[j] >= 0)
        key = (static_cast<unsigned long long>(b[j]) << kSeedBinShift) |
              (static_cast<unsigned long long>(seed_mse_order(m[j])) << kSeedCellBits) | cell;
      if (valid) cur[i] = key;
      const unsigned vm = __ballot_sync(0xffffffffu, valid);
#pragma unroll
      for (int p = 0; p < kSortPasses; ++p)
        hist_add(s.hist[p], static_cast<unsigned>(key >> (kSeedCellBits + 8 * p)) & 255u, valid, vm, lane);
    }
  }
  __syncthreads();
  SORT_PROBE(0);  // key building

  for (int pass = 0; pass < kSortPasses; ++pass) {
    const int shift = kSeedCellBits + 8 * pass;
    uint32_t* hist_cur = s.hist[pass];
    // exclusive prefix of the digit histogram -> offs; a pass whose keys all share one digit changes nothing
    if (tid == 0) s.skip = 0;
    __syncthreads();
    if (tid < 256) {
      const uint32_t h = hist_cur[tid];
      if (h == static_cast<uint32_t>(n)) s.skip = 1;
      uint32_t incl = h;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane == 31) s.warp_tot[warp] = incl;
      s.offs[tid] = incl - h;  // exclusive within the warp; the warp bases are added below
    }
    {
      uint32_t* z = reinterpret_cast<uint32_t*>(s.wrank);
#pragma unroll
      for (int k = 0; k < kSortRows * 256 / 2 / kSortThreads; ++k) z[k * kSortThreads + tid] = 0;
    }
    __syncthreads();
    if (s.skip != 0) continue;
    if (tid < 256) {
      uint32_t base = 0;
      for (int w = 0; w < warp; ++w) base += s.warp_tot[w];
      s.offs[tid] += base;
    }
    SORT_PROBE(1);  // pass setup
    // sub-tile j holds positions j * kSortThreads + tid
    unsigned long long key[kSortItems];
    bool valid[kSortItems];
    unsigned dg[kSortItems], rank[kSortItems];
#pragma unroll
    for (int j = 0; j < kSortItems; ++j) {
      valid[j] = j * kSortThreads + tid < n;
      key[j] = valid[j] ? cur[j * kSortThreads + tid] : ~0ull;
    }
#pragma unroll
    for (int j = 0; j < kSortItems; ++j) {
      const unsigned vm = __ballot_sync(0xffffffffu, valid[j]);
      dg[j] = static_cast<unsigned>(key[j] >> shift) & 255u;
      const unsigned peers = match_digit(dg[j], vm);
      rank[j] = __popc(peers & lt_mask);
      if (valid[j] && rank[j] == 0) s.wrank[j * kSortWarps + warp][dg[j]] = static_cast<uint16_t>(__popc(peers));
    }
    SORT_PROBE(2);  // matching
    __syncthreads();
    {
      // thread (q, d): the 32 warp rows of sub-tile q for digit d -- counts -> keys in earlier warps of the sub-tile
      const int q = tid >> 8, d = tid & 255;
      uint16_t c[kSortWarps];
#pragma unroll
      for (int w = 0; w < kSortWarps; ++w) c[w] = s.wrank[q * kSortWarps + w][d];
      uint32_t run = 0;
#pragma unroll
      for (int w = 0; w < kSortWarps; ++w) {
        s.wrank[q * kSortWarps + w][d] = static_cast<uint16_t>(run);
        run += c[w];
      }
      s.sub_offs[q][d] = run;  // the sub-tile's total for now
    }
    __syncthreads();
    if (tid < 256) {
      uint32_t o = s.offs[tid];
#pragma unroll
      for (int q = 0; q < kSortItems; ++q) {
        const uint32_t t = s.sub_offs[q][tid];
        s.sub_offs[q][tid] = o;
        o += t;
      }
    }
    SORT_PROBE(3);  // prefix over the rows
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kSortItems; ++j)
      if (valid[j]) {
        const uint32_t o = s.sub_offs[j][dg[j]] + s.wrank[j * kSortWarps + warp][dg[j]] + rank[j];
        DPX_CHECK(o < static_cast<uint32_t>(n));
        nxt[o] = key[j];
      }
    SORT_PROBE(4);  // scatter
    __syncthreads();
    unsigned long long* t = cur;
    cur = nxt;
    nxt = t;
  }
  for (int i = tid; i < n; i += kSortThreads) keys_tmp[i] = cur[i];
  SORT_PROBE(5);  // write back
}

// Stage 2: the sorted chunks of a frame are merged by ranking.  Keys are unique, so the final position of a key is the
// number of keys of its frame below it: its index in its own chunk plus a lower bound in every other chunk (binary
// searches in L2-resident memory, every key independent of the others).
constexpr int kMergeThreads = 256;
__global__ void __launch_bounds__(kMergeThreads) seed_rank_merge_kernel(const unsigned long long* __restrict__ keys_tmp,
                                                                        unsigned long long* __restrict__ keys_out, int n_cells) {
  const long long fc = static_cast<long long>(blockIdx.y) * n_cells;
  const int e = blockIdx.x * kMergeThreads + threadIdx.x;
  if (e >= n_cells) return;
  keys_tmp += fc;
  const unsigned long long key = keys_tmp[e];
  const int own = e / kSortChunk;
  int pos = e - own * kSortChunk;
  const int n_chunks = (n_cells + kSortChunk - 1) / kSortChunk;
  for (int c = 0; c < n_chunks; ++c) {
    if (c == own) continue;
    const unsigned long long* base = keys_tmp + c * kSortChunk;
    int lo = 0, hi = min(kSortChunk, n_cells - c * kSortChunk);
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(base + mid) < key) lo = mid + 1;
      else hi = mid;
    }
    pos += lo;
  }
  DPX_CHECK(pos >= 0 && pos < n_cells);
  keys_out[fc + pos] = key;
}

}  // namespace

// keys_out / keys_tmp: [n_frames][n_cells] each; keys_tmp is scratch
inline cudaError_t launch_seed_sort(const int16_t* bin, const float* mse, unsigned long long* keys_out, unsigned long long* keys_tmp,
                                    int n_frames, int n_cells, cudaStream_t stream) {
  if (n_frames == 0 || n_cells == 0) return cudaSuccess;
  const int n_chunks = (n_cells + kSortChunk - 1) / kSortChunk;
  const size_t dyn = sizeof(SortShared);
  cudaFuncSetAttribute(seed_chunk_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(dyn));
  // a frame of a single chunk needs no merge: its chunk is sorted straight into keys_out
  seed_chunk_sort_kernel<<<dim3(n_chunks, n_frames), kSortThreads, dyn, stream>>>(bin, mse, n_chunks == 1 ? keys_out : keys_tmp, n_cells);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess || n_chunks == 1) return e;
  seed_rank_merge_kernel<<<dim3((n_cells + kMergeThreads - 1) / kMergeThreads, n_frames), kMergeThreads, 0, stream>>>(keys_tmp, keys_out, n_cells);
  return cudaGetLastError();
}

}  // namespace dpx
