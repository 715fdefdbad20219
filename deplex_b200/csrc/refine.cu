// refine.cu -- stage 4 of the hot path (opt-in, ransacRefinement=1): per-plane RANSAC refinement of the labels.
//
// Replaces, per frame (reference file:line):
//   PlaneExtractor::Impl::refineLabels                          plane_extractor.cpp:472-509
//   RTL::PlaneRANSAC::FindBest / FindInliers / IsContinued /
//     GenerateModel / EvaluateModel                             libs/rtl/include/rtl/RANSAC.hpp:25-98
//   PlaneEstimator::ComputeModel / ComputeError                 libs/rtl/include/rtl/Plane.hpp:13-49
//   std::mt19937 (default seed 5489, one generator per process() call, shared across the labels) and
//   libstdc++'s std::uniform_int_distribution<int> (GCC >= 11: Lemire's multiply-shift with rejection)
//
// One thread-block cluster (8 CTAs of 512 threads, distributed shared memory) per frame; CTA 0 of the cluster leads.
// The reference's loop is sequential twice over -- labels share one random stream, and each
// label's iterations stop as soon as a hypothesis reaches the target inlier ratio -- so the CTA walks the labels in
// order and evaluates the next 32 hypotheses of the current label speculatively and at once:
//   warp 0     (leader) draws the sample ranks of 32 hypotheses: the next generator outputs are tempered in parallel
//              into a tape and lane h takes outputs 3h..3h+2; a rejection or a repeated sample sends the round down the
//              exact sequential path.  The draws each hypothesis took are remembered;
//   96 threads turn (hypothesis, sample) ranks into pixels: the k-th pixel of a label in image order follows
//              from the label's cells sorted by cell id (cells are painted whole), no per-pixel index lists;
//   32 threads build the plane models (fp32, the reference's expression order);
//   all warps  of all 8 CTAs score: a warp stages 32 points in shared memory, then lane g scores hypothesis g on each
//              of them (the loss is a count, so the order of the points is free); per-warp counts go to the leader's
//              shared memory with distributed-shared-memory atomics;
//   one warp   replays the reference's sequential loop over the 32 losses as a prefix minimum (best-so-far,
//              IsContinued) and reports how many hypotheses were really consumed; warp 0 leaves the generator at
//              exactly that point.
// FindInliers + the relabelling loop (which stops at the last inlier, plane_extractor.cpp:500-507) become two
// passes: the largest inlier pixel index, then "non-inlier before it -> 0".
#include "refine.cuh"

#include <cooperative_groups.h>

#include "exact_math.cuh"

namespace dpx {
namespace {

constexpr int kRefThreads = 512;
constexpr int kRefWarps = kRefThreads / 32;
constexpr int kSub = 4;          // groups of 32 hypotheses per round
constexpr int kHyp = 32 * kSub;  // hypotheses evaluated per round (one cluster-wide pass over the label's points)
constexpr int kRefCluster = 8;  // CTAs per frame (portable cluster size limit)
constexpr int kTape = 128;      // generator outputs prepared per round (32 hypotheses x 3 draws + slack)
constexpr int kMaxRows = 1023;  // cell rows the per-label row table can hold
constexpr int kCellCache = 4096;  // cells of one label cached in shared memory (larger labels read the global list)
namespace cg = cooperative_groups;
constexpr unsigned kFullMask = 0xffffffffu;

struct RefShared {
  uint32_t mt[kMtN];
  uint32_t mt_bak[kMtN];
  alignas(16) float model[kHyp][4];  // the round's hypotheses: computed by the leader, pushed into every CTA's copy
  unsigned loss[kHyp];        // leader: the cluster's totals
  unsigned loss_cta[kHyp];    // this CTA's share of a round
  int draws_cum[kHyp];        // generator draws used up to and including hypothesis g
  int rank[kHyp][3];          // sample ranks, ascending (std::set order)
  long long pix[kHyp][3];     // their pixels
  float best[4];
  double bestloss;            // HUGE_VAL until a hypothesis has been accepted
  int iteration, consumed, go_on;
  int max_inlier_pix;
  float4 stage[kRefWarps][32];
  uint32_t tape[kTape];       // tempered generator outputs of this round, in draw order
  int rowstart[kMaxRows + 1]; // per label: index of the first of its cells in each cell row (cells are sorted)
  int32_t cells[kCellCache];  // per label, every CTA: a copy of the label's sorted cell list when it fits
};

// ---- std::mt19937, executed by warp 0 (all lanes compute the same values; lane 0 owns the stores) ----------
__device__ __forceinline__ void mt_twist_warp(uint32_t* mt, int lane) {
  // new[i] = old[(i + 397) % 624] ^ f(old[i], old[i + 1]); entries i >= 227 read already updated entries,
  // so the update runs in waves of 227 (dependency distance) with a warp barrier in between.
  auto f = [](uint32_t a, uint32_t b) {
    const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
  };
  for (int base = 0; base < kMtN; base += 224) {  // 224 <= 227, multiple of 32
    const int end = min(base + 224, kMtN);
    uint32_t v[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const int i = base + k * 32 + lane;
      if (i < end) {
        const uint32_t nxt = mt[(i + 1) % kMtN];  // for i = 623 this is the already updated mt[0], as in the reference
        v[k] = mt[(i + 397) % kMtN] ^ f(mt[i], nxt);
      }
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const int i = base + k * 32 + lane;
      if (i < end) mt[i] = v[k];
    }
    __syncwarp();
  }
}

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

// `idx` is the generator's position, held in a register by every lane of warp 0 (all lanes agree)
__device__ __forceinline__ uint32_t mt_next_warp(RefShared& s, int lane, int& idx) {
  if (idx >= kMtN) {
    mt_twist_warp(s.mt, lane);
    idx = 0;
  }
  return mt_temper(s.mt[idx++]);
}

// std::uniform_int_distribution<int>(0, n - 1)(gen) over std::mt19937.  The mapping is implementation-defined and
// libstdc++ changed it (bits/uniform_int_dist.h), so both generations are here, selected by RefineArgs::uniform_variant:
//   0  libstdc++ >= 11 (_S_nd): Lemire's nearly divisionless method -- product = u * n, the low 32 bits below
//      (2^32 - n) % n are rejected, result = product >> 32;
//   1  libstdc++ <= 10: scaling = (2^32 - 1) / n, draws >= n * scaling are rejected, result = u / scaling.
// One generator output either yields a value or is rejected; `accept_draw` is that step for both variants.
struct UniformMap {
  uint32_t n, threshold, scaling, past_lo;  // past = n * scaling fits 32 bits (<= 2^32 - 1)
  int variant;
};
__device__ __forceinline__ UniformMap make_uniform_map(uint32_t n, int variant) {
  UniformMap m;
  m.n = n;
  m.variant = variant;
  m.threshold = (0u - n) % n;
  m.scaling = 0xffffffffu / n;
  m.past_lo = n * m.scaling;
  return m;
}
__device__ __forceinline__ bool accept_draw(const UniformMap& m, uint32_t u, int& value) {
  if (m.variant == 1) {
    value = static_cast<int>(u / m.scaling);
    return u < m.past_lo;
  }
  const unsigned long long product = static_cast<unsigned long long>(u) * m.n;
  value = static_cast<int>(product >> 32);
  return static_cast<uint32_t>(product) >= m.threshold;  // (low < n && low < threshold) rejects; threshold < n always
}
// `draws` counts generator calls.
__device__ __forceinline__ int uniform_below_warp(RefShared& s, int lane, int& idx, const UniformMap& m, int& draws) {
  int value;
  do {
    ++draws;
  } while (!accept_draw(m, mt_next_warp(s, lane, idx), value));
  return value;
}

struct LabelCells {
  const int32_t* cells;  // the label's cells, ascending cell id
  int count;             // number of cells
};

// The k-th pixel (image order) among the pixels of a label whose cells are `lc` (plane_extractor.cpp:473-478 builds
// this list explicitly).  A cell row holding m of the label's cells contributes p image rows of m*p pixels each.
// rowstart[r] = index of the label's first cell in cell row r (rowstart[nv] = count), or nullptr to search.
__device__ __forceinline__ long long kth_pixel(const LabelCells& lc, const int* rowstart, int k, int p, int nh, int width) {
  const int p2 = p * p;
  const int t = k / p2;
  const int r = lc.cells[t] / nh;  // cell row containing rank k
  int s, e;
  if (rowstart) {
    s = rowstart[r];
    e = rowstart[r + 1];
  } else {
    s = t;
    e = t + 1;
    while (s > 0 && lc.cells[s - 1] / nh == r) --s;
    while (e < lc.count && lc.cells[e] / nh == r) ++e;
  }
  const int m = e - s;
  const int kk = k - s * p2;
  const int i = kk / (p * m), rem = kk - i * (p * m);
  const int j = rem / p, xo = rem - j * p;
  const int cell = lc.cells[s + j];
  const int q = cell - r * nh;
  return static_cast<long long>(r * p + i) * width + q * p + xo;
}

template <int LAYOUT>
__device__ __forceinline__ void load_point(const float* xyz, long long n_points, long long pix, float& x, float& y, float& z) {
  if (LAYOUT == kLayoutRowMajor) {
    x = __ldg(xyz + 3 * pix); y = __ldg(xyz + 3 * pix + 1); z = __ldg(xyz + 3 * pix + 2);
  } else {
    x = __ldg(xyz + pix); y = __ldg(xyz + n_points + pix); z = __ldg(xyz + 2 * n_points + pix);
  }
}

// PlaneEstimator::ComputeError (Plane.hpp:45-48): fp32, left to right
__device__ __forceinline__ float plane_error(const float (&m)[4], float x, float y, float z) {
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[0], x), __fmul_rn(m[1], y)), __fmul_rn(m[2], z)), m[3]);
}

#ifdef DPX_REFINE_PROBE
// probe builds only (make NVFLAGS_EXTRA=-DDPX_REFINE_PROBE, tools/refine_probe.py): cycles of the leader's thread 0 per phase
__device__ long long g_refine_probe[8];
#define REF_PROBE(slot) do { if (leader && tid == 0 && blockIdx.x < kRefCluster) { const long long t__ = clock64(); g_refine_probe[slot] += t__ - rp_t; rp_t = t__; } } while (0)
#else
#define REF_PROBE(slot) do { } while (0)
#endif

template <int LAYOUT>
__global__ void __launch_bounds__(kRefThreads) refine_kernel(const RefineArgs args) {
  __shared__ RefShared s;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned crank = cluster.block_rank();
  const bool leader = crank == 0;
  RefShared* lead = cluster.map_shared_rank(&s, 0);  // the leader CTA's copy (distributed shared memory)
  const Geometry& g = args.geom;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int frame = blockIdx.x / kRefCluster;
#ifdef DPX_REFINE_PROBE
  long long rp_t = clock64();
#endif
  const int C = g.n_cells, p = g.patch, p2 = p * p, nh = g.nh;
  const long long fc = static_cast<long long>(frame) * C;
  const int nseg = args.tables.n_planes[frame];
  if (nseg <= 0) return;  // no plane: nothing is labelled (plane_extractor.cpp:230-232); the whole cluster leaves

  const int32_t* cell_label = args.tables.cell_label + fc;
  int32_t* lab_cells = args.tables.queue + fc;                               // [C] cells sorted by (label, cell id)
  int* lab_end = reinterpret_cast<int*>(args.tables.pairs + 2 * fc);          // [nseg] end of each label's run
  const float* xyz = args.xyz + static_cast<long long>(frame) * 3 * g.n_points;
  int32_t* labels = args.labels + static_cast<long long>(frame) * g.n_points;
  // SetParamThreshold(double) <- float config value: the threshold is a float widened to double, and errors are floats
  // widened to double, so every comparison against it is exact in fp32
  const float thr_f = args.threshold;
  const double ratio = static_cast<double>(args.inliers_ratio);
  // this thread's share of a label's points: chunks of 32 points, dealt round-robin over the cluster's warps
  const int gwarp = static_cast<int>(crank) * kRefWarps + warp;
  constexpr int kStride = kRefCluster * kRefThreads;

  int mt_idx = kMtN, mt_idx_bak = kMtN;  // generator position; meaningful in warp 0 of the leader only
  if (tid < kHyp) s.loss_cta[tid] = 0;
  __syncthreads();
  if (leader) {
    // ---- labels_indices (plane_extractor.cpp:473-478), per cell instead of per pixel ---------------------
    for (int i = tid; i < nseg; i += kRefThreads) lab_end[i] = 0;
    for (int i = tid; i < kMtN; i += kRefThreads) s.mt[i] = args.mt_init[i];
    __syncthreads();
    for (int c = tid; c < C; c += kRefThreads) {
      const int l = cell_label[c];
      if (l > 0) atomicAdd(&lab_end[l - 1], 1);
    }
    __syncthreads();
    if (warp == 0) {
      // exclusive scan of the counts -> run starts (kept as running cursors), then a stable fill in ascending cell id
      int run = 0;
      for (int b0 = 0; b0 < nseg; b0 += 32) {
        const int i = b0 + lane;
        const int cnt = i < nseg ? lab_end[i] : 0;
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(kFullMask, incl, o);
          if (lane >= o) incl += t;
        }
        if (i < nseg) lab_end[i] = run + incl - cnt;
        run += __shfl_sync(kFullMask, incl, 31);
      }
      __syncwarp();
      for (int c0 = 0; c0 < C; c0 += 32) {
        const int c = c0 + lane;
        const int l = c < C ? cell_label[c] : 0;
        const unsigned act = __ballot_sync(kFullMask, l > 0);
        if (l > 0) {
          const unsigned grp = __match_any_sync(act, l);
          const int ldr = __ffs(grp) - 1;
          int base = 0;
          if (lane == ldr) {
            base = lab_end[l - 1];
            lab_end[l - 1] = base + __popc(grp);
          }
          base = __shfl_sync(grp, base, ldr);
          lab_cells[base + __popc(grp & ((1u << lane) - 1u))] = c;
        }
        __syncwarp();
      }
    }
    __threadfence();
  }
  cluster.sync();  // the label runs (global memory) are visible to the whole cluster

  // ---- one label after the other (plane_extractor.cpp:486-508) ----------------------------------------------
  for (int L = 0; L < nseg; ++L) {
    const int start = L ? lab_end[L - 1] : 0;
    LabelCells lc;
    lc.cells = lab_cells + start;
    lc.count = lab_end[L] - start;
    if (lc.count == 0) continue;  // labels_indices[label].size() == 0 (the same decision in every CTA)
    const int n = lc.count * p2;
    // the label's cell list is read by every rank -> pixel lookup and by every scored point: keep it in shared memory
    __syncthreads();  // (the previous label's last readers of the cache are done)
    if (lc.count <= kCellCache) {
      for (int i = tid; i < lc.count; i += kRefThreads) s.cells[i] = lc.cells[i];
      lc.cells = s.cells;
    }
    __syncthreads();

    const bool rows_ok = g.nv <= kMaxRows;
    if (leader && rows_ok) {
      // first cell of the label in every cell row (lower bound over the sorted cell list): ranks -> pixels become
      // two table reads instead of a walk along the row
      for (int r = tid; r <= g.nv; r += kRefThreads) {
        int lo = 0, hi = lc.count;
        const int key = r * nh;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (lc.cells[mid] < key) lo = mid + 1;
          else hi = mid;
        }
        s.rowstart[r] = lo;
      }
    }
    if (leader && tid == 0) {
      s.best[0] = s.best[1] = s.best[2] = s.best[3] = 0.f;  // Eigen::Vector4f::Zero()
      s.bestloss = HUGE_VAL;
      s.iteration = 0;
      s.max_inlier_pix = -1;
      // IsContinued(0, N - HUGE_VAL, N): int(-inf) is INT_MIN on x86-64
      s.go_on = (0 < args.max_iterations) && (static_cast<double>(INT_MIN) < ratio * n);
    }
    cluster.sync();
    int go_on = lead->go_on;

    // ---- FindBest (RANSAC.hpp:25-51), kHyp hypotheses per round ------------------------------------------------
    while (go_on) {
      REF_PROBE(6);
      if (leader) {
        if (warp == 0) {
          // The samples of the next kHyp iterations (RANSAC.hpp:81-87), kSub groups of 32 one after the other.  The
          // generator is saved first: when the search stops inside the round it is rewound and advanced by exactly the
          // draws the reference would have made.
          for (int i = lane; i < kMtN; i += 32) s.mt_bak[i] = s.mt[i];
          mt_idx_bak = mt_idx;
          __syncwarp();
          const UniformMap umap = make_uniform_map(static_cast<uint32_t>(n), args.uniform_variant);
          int draws_base = 0;          // draws of the groups before this one
          bool twisted_round = false;  // the generator has left the block saved in mt_bak
          for (int sub = 0; sub < kSub; ++sub) {
            // Almost always every iteration takes exactly three draws (a rejection in the distribution or a repeated
            // sample has probability ~3/n), so the next kTape generator outputs are tempered in parallel and lane h takes
            // outputs from 3h on.  Only while the tape stays inside the generator's current block; the group that
            // crosses into the next block (one in six or seven) draws one value at a time instead.
            const int idx0 = mt_idx;
            const int avail = max(0, min(kTape, kMtN - idx0));  // tape entries the current block still holds
            int off = 0, extra = 0, a = 0, b = 0, c = 0;
            bool ok = false, twisted = false;
            if (avail == kTape || !twisted_round) {
              for (int j = lane; j < avail; j += 32) s.tape[j] = mt_temper(s.mt[idx0 + j]);
              if (avail < kTape) {
                // the tape runs into the generator's next block (mt_bak still holds the current one: first twist of
                // the round)
                __syncwarp();
                mt_twist_warp(s.mt, lane);
                twisted = true;
                for (int j = avail + lane; j < kTape; j += 32) s.tape[j] = mt_temper(s.mt[j - avail]);
              }
              __syncwarp();
              // Lane h consumes tape entries from 3h + (extra draws of the hypotheses before it) until it holds three
              // distinct accepted values.  The extra draws shift everything behind them, so the offsets are iterated to
              // a fixed point: a prefix sum of the lanes' extra draws per pass, normally one pass.
              for (int pass = 0; pass < 6; ++pass) {
                int pos = 3 * lane + off, cnt = 0;
                a = b = c = -1;
                while (cnt < 3 && pos < kTape) {
                  int vv;
                  if (!accept_draw(umap, s.tape[pos++], vv)) continue;  // the distribution draws again
                  if (vv == a || vv == b || vv == c) continue;          // std::set already holds it
                  if (cnt == 0) a = vv;
                  else if (cnt == 1) { if (vv < a) { b = a; a = vv; } else b = vv; }
                  else {
                    if (vv < a) { c = b; b = a; a = vv; }
                    else if (vv < b) { c = b; b = vv; }
                    else c = vv;
                  }
                  ++cnt;
                }
                const bool complete = cnt == 3;
                extra = pos - (3 * lane + off) - 3;
                int incl = extra;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                  const int t = __shfl_up_sync(kFullMask, incl, o);
                  if (lane >= o) incl += t;
                }
                const int new_off = incl - extra;
                const bool stable = __all_sync(kFullMask, complete && new_off == off);
                off = new_off;
                if (stable) {
                  ok = true;
                  break;
                }
                if (!__all_sync(kFullMask, complete)) break;  // ran off the tape: take the sequential path
              }
            }
            if (ok) {
              const int h = 32 * sub + lane;
              s.rank[h][0] = a; s.rank[h][1] = b; s.rank[h][2] = c;
              const int cum = 3 * (lane + 1) + off + extra;  // draws of this group up to and including hypothesis `lane`
              s.draws_cum[h] = draws_base + cum;
              const int used = __shfl_sync(kFullMask, cum, 31);
              if (!twisted) {
                mt_idx = idx0 + used;
              } else if (used > avail) {
                mt_idx = used - avail;  // inside the block the tape has already generated
                twisted_round = true;
              } else {
                // the draws end inside the old block after all: back to it
                for (int i = lane; i < kMtN; i += 32) s.mt[i] = s.mt_bak[i];
                __syncwarp();
                mt_idx = idx0 + used;
              }
              draws_base += used;
            } else {
              if (twisted) {
                for (int i = lane; i < kMtN; i += 32) s.mt[i] = s.mt_bak[i];
                __syncwarp();
              }
              int draws = 0;
              for (int hh = 0; hh < 32; ++hh) {
                int sa = -1, sb = -1, sc = -1, cnt = 0;  // the std::set<int>, kept sorted
                while (cnt < 3) {
                  const int vv = uniform_below_warp(s, lane, mt_idx, umap, draws);
                  if (vv == sa || vv == sb || vv == sc) continue;
                  if (cnt == 0) sa = vv;
                  else if (cnt == 1) { if (vv < sa) { sb = sa; sa = vv; } else sb = vv; }
                  else {
                    if (vv < sa) { sc = sb; sb = sa; sa = vv; }
                    else if (vv < sb) { sc = sb; sb = vv; }
                    else sc = vv;
                  }
                  ++cnt;
                }
                if (lane == 0) {
                  const int h = 32 * sub + hh;
                  s.rank[h][0] = sa; s.rank[h][1] = sb; s.rank[h][2] = sc;
                  s.draws_cum[h] = draws_base + draws;
                }
              }
              draws_base += draws;
              if (mt_idx < idx0) twisted_round = true;  // the generator moved on to its next block
            }
            __syncwarp();
          }
        }
        if (tid < kHyp) s.loss[tid] = 0;
        __syncthreads();
        REF_PROBE(0);  // sampling
        if (tid < kHyp * 3) s.pix[tid / 3][tid % 3] = kth_pixel(lc, rows_ok ? s.rowstart : nullptr, s.rank[tid / 3][tid % 3], p, nh, g.width);
        __syncthreads();
        if (tid < kHyp) {
          // PlaneEstimator::ComputeModel (Plane.hpp:13-43), fp32 in the reference's expression order
          float x0, y0, z0, x1, y1, z1, x2, y2, z2;
          load_point<LAYOUT>(xyz, g.n_points, s.pix[tid][0], x0, y0, z0);
          load_point<LAYOUT>(xyz, g.n_points, s.pix[tid][1], x1, y1, z1);
          load_point<LAYOUT>(xyz, g.n_points, s.pix[tid][2], x2, y2, z2);
          const f32 X0(x0), X1(x1), X2(x2), Y0(y0), Y1(y1), Y2(y2), Z0(z0), Z1(z1), Z2(z2);
          const f32 D = X0 * Y1 - X1 * Y0 - X0 * Y2 + X2 * Y0 + X1 * Y2 - X2 * Y1;
          const f32 a = (Z0 * (Y1 - Y2)) / D - (Z1 * (Y0 - Y2)) / D + (Z2 * (Y0 - Y1)) / D;
          const f32 b = (Z1 * (X0 - X2)) / D - (Z0 * (X1 - X2)) / D - (Z2 * (X0 - X1)) / D;
          const f32 d = (Z2 * (X0 * Y1 - X1 * Y0)) / D - (Z1 * (X0 * Y2 - X2 * Y0)) / D + (Z0 * (X1 * Y2 - X2 * Y1)) / D;
          const f32 c(-1.0f);
          // `sqrt(float)` in Plane.hpp:36 resolves to ::sqrt(double): double square root, rounded to float on assignment
          const f32 l(__double2float_rn(__dsqrt_rn(static_cast<double>((a * a + b * b + c * c).v))));
          // pushed into every CTA's shared memory (fire-and-forget stores, visible after the cluster barrier): 1024
          // threads per CTA fetching them from the leader afterwards queue up on its distributed-shared-memory port
          const float4 mv = make_float4((a / l).v, (b / l).v, (c / l).v, (d / l).v);
#pragma unroll
          for (int r = 0; r < kRefCluster; ++r) *reinterpret_cast<float4*>(cluster.map_shared_rank(&s.model[tid][0], r)) = mv;
        }
      }
      REF_PROBE(1);  // ranks -> pixels, models
      cluster.sync();  // (1) the kHyp models and the zeroed losses are in the leader's shared memory
      REF_PROBE(2);
      {
        // EvaluateModel (RANSAC.hpp:89-98): lane g scores hypotheses g, g + 32, ...; loss += (fabs(error) >= threshold)
        float m[kSub][4];
#pragma unroll
        for (int j = 0; j < kSub; ++j)
#pragma unroll
          for (int k = 0; k < 4; ++k) m[j][k] = s.model[32 * j + lane][k];
        unsigned loss[kSub];
#pragma unroll
        for (int j = 0; j < kSub; ++j) loss[j] = 0;
        for (int e0 = gwarp * 32; e0 < n; e0 += kStride) {
          const int e = e0 + lane;
          float x = 0.f, y = 0.f, z = 0.f;
          if (e < n) {
            const int t = e / p2, in = e - t * p2;
            const int cell = lc.cells[t];
            const int r = cell / nh, q = cell - r * nh;
            const int i = in / p, j = in - i * p;
            load_point<LAYOUT>(xyz, g.n_points, static_cast<long long>(r * p + i) * g.width + q * p + j, x, y, z);
          }
          s.stage[warp][lane] = make_float4(x, y, z, 0.f);
          __syncwarp();
          const int cnt = min(32, n - e0);
          for (int k = 0; k < cnt; ++k) {
            const float4 pt = s.stage[warp][k];
#pragma unroll
            for (int j = 0; j < kSub; ++j) {
              // the reference compares fabs(double(error)) with double(float threshold): the same predicate in fp32
              loss[j] += (::fabsf(plane_error(m[j], pt.x, pt.y, pt.z)) >= thr_f) ? 1u : 0u;
            }
          }
          __syncwarp();
        }
        // per-warp counts -> this CTA's totals -> one distributed-shared-memory atomic per hypothesis and CTA
        // (128 warps adding straight into the leader's counters serialise there)
#pragma unroll
        for (int j = 0; j < kSub; ++j)
          if (loss[j]) atomicAdd(&s.loss_cta[32 * j + lane], loss[j]);
        __syncthreads();
        if (tid < kHyp) {
          const unsigned v = s.loss_cta[tid];
          s.loss_cta[tid] = 0;
          if (v) atomicAdd(&lead->loss[tid], v);
        }
      }
      REF_PROBE(3);  // scoring
      cluster.sync();  // (2) every CTA's counts are in
      REF_PROBE(4);
      if (leader) {
        if (warp == 1) {
          // The reference's sequential loop over these hypotheses (RANSAC.hpp:33-46), evaluated by one warp, 32 at a time:
          // lane h owns hypothesis h of the group.  The best-so-far loss before hypothesis h is a prefix minimum; the loop
          // runs while IsContinued holds, which is monotone (the best loss only falls, the iteration count only grows), so
          // the number of iterations really run is the number of hypotheses whose check passes.
          const double inf = HUGE_VAL;
          double best0 = s.bestloss;
          int iter0 = s.iteration;
          int consumed_all = 0;
          for (int sub = 0; sub < kSub; ++sub) {
            const double mine = static_cast<double>(s.loss[32 * sub + lane]);
            double incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const double t = __shfl_up_sync(kFullMask, incl, o);
              if (lane >= o) incl = ::fmin(incl, t);
            }
            double before = __shfl_up_sync(kFullMask, incl, 1);
            if (lane == 0) before = inf;
            before = ::fmin(before, best0);
            const int inl = ::isinf(before) ? INT_MIN : static_cast<int>(n - before);
            const bool go_h = (iter0 + lane < args.max_iterations) && (static_cast<double>(inl) < ratio * n);
            const int consumed = __popc(__ballot_sync(kFullMask, go_h));
            // best loss over the iterations really run, and the first of them that reaches it (strict '<' updates)
            double best = lane < consumed ? mine : inf;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) best = ::fmin(best, __shfl_xor_sync(kFullMask, best, o));
            const unsigned hit = __ballot_sync(kFullMask, lane < consumed && mine == best);
            if (hit && best < best0) {
              if (lane == 0) {
                const int winner = 32 * sub + __ffs(hit) - 1;
                s.best[0] = s.model[winner][0]; s.best[1] = s.model[winner][1]; s.best[2] = s.model[winner][2]; s.best[3] = s.model[winner][3];
              }
              best0 = best;
            }
            iter0 += consumed;
            consumed_all += consumed;
            if (consumed < 32) break;
          }
          if (lane == 0) {
            s.bestloss = best0;
            s.iteration = iter0;
            bool go = consumed_all == kHyp;
            if (go) {
              const int inl2 = ::isinf(best0) ? INT_MIN : static_cast<int>(n - best0);
              go = (s.iteration < args.max_iterations) && (static_cast<double>(inl2) < ratio * n);
            }
            s.consumed = consumed_all;
            for (int r = 0; r < kRefCluster; ++r) *cluster.map_shared_rank(&s.go_on, r) = go ? 1 : 0;
          }
        }
        __syncthreads();
        if (warp == 0 && s.consumed < kHyp) {
          // the search stopped inside the round: leave the generator just after the last iteration the reference ran
          for (int i = lane; i < kMtN; i += 32) s.mt[i] = s.mt_bak[i];
          mt_idx = mt_idx_bak;
          __syncwarp();
          int redo = s.consumed > 0 ? s.draws_cum[s.consumed - 1] : 0;
          while (redo > 0) {  // advancing the generator is moving its position, block by block
            if (mt_idx >= kMtN) {
              mt_twist_warp(s.mt, lane);
              __syncwarp();
              mt_idx = 0;
            }
            const int step = min(redo, kMtN - mt_idx);
            mt_idx += step;
            redo -= step;
          }
        }
      }
      REF_PROBE(5);  // evaluate + generator
      cluster.sync();  // (3) the decision is visible
      go_on = s.go_on;
    }

    // ---- FindInliers + relabelling (RANSAC.hpp:53-62, plane_extractor.cpp:498-507) ---------------------------
    float m[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) m[k] = lead->best[k];
    for (int pass = 0; pass < 2; ++pass) {
      const int last = pass ? lead->max_inlier_pix : -1;
      if (pass == 1 && last < 0) break;  // no inlier at all: the relabelling loop never runs (same in every CTA)
      int local_max = -1;
      for (int e = gwarp * 32 + lane; e < n; e += kStride - 0) {
        const int t = e / p2, in = e - t * p2;
        const int cell = lc.cells[t];
        const int r = cell / nh, q = cell - r * nh;
        const int i = in / p, j = in - i * p;
        const long long pix = static_cast<long long>(r * p + i) * g.width + q * p + j;
        float x, y, z;
        load_point<LAYOUT>(xyz, g.n_points, pix, x, y, z);
        const bool inlier = ::fabsf(plane_error(m, x, y, z)) < thr_f;
        if (pass == 0) {
          if (inlier) local_max = max(local_max, static_cast<int>(pix));
        } else if (!inlier && pix < last) {
          labels[pix] = 0;  // points after the last inlier keep their label, as in the reference's loop
        }
      }
      if (pass == 0) {
        local_max = __reduce_max_sync(kFullMask, local_max);
        if (lane == 0 && local_max >= 0) atomicMax(&lead->max_inlier_pix, local_max);
        cluster.sync();  // the largest inlier pixel of the whole label is known
      }
    }
    cluster.sync();  // nobody still reads this label's state when the leader resets it for the next one
  }
  cluster.sync();  // no CTA leaves while another may still touch its shared memory
}

}  // namespace

#ifdef DPX_REFINE_PROBE
}  // namespace dpx
extern "C" __attribute__((visibility("default"))) int dpx_debug_refine_probe(long long* out, int reset) {
  cudaDeviceSynchronize();
  if (out) cudaMemcpyFromSymbol(out, dpx::g_refine_probe, sizeof(long long) * 8);
  if (reset) {
    long long zero[8] = {};
    cudaMemcpyToSymbol(dpx::g_refine_probe, zero, sizeof(zero));
  }
  return 0;
}
namespace dpx {
#endif

void mt19937_default_state(uint32_t out[kMtN]) {
  out[0] = 5489u;
  for (int i = 1; i < kMtN; ++i) out[i] = 1812433253u * (out[i - 1] ^ (out[i - 1] >> 30)) + static_cast<uint32_t>(i);
}

cudaError_t launch_refine(const RefineArgs& args, cudaStream_t stream) {
  if (args.n_frames == 0 || args.geom.n_cells == 0) return cudaSuccess;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(args.n_frames) * kRefCluster, 1, 1);
  cfg.blockDim = dim3(kRefThreads, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kRefCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (args.layout == kLayoutRowMajor) return cudaLaunchKernelEx(&cfg, refine_kernel<kLayoutRowMajor>, args);
  return cudaLaunchKernelEx(&cfg, refine_kernel<kLayoutColMajor>, args);
}

}  // namespace dpx
