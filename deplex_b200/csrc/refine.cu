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
// One thread-block cluster (8 CTAs of 512 threads, 16 when the launch is a few frames; distributed shared memory) per
// frame; CTA 0 of the cluster leads.
// The reference's loop is sequential twice over -- labels share one random stream, and each label's iterations stop as
// soon as a hypothesis reaches the target inlier ratio -- so the cluster walks the labels in order and evaluates the
// next kHyp = 128 hypotheses of the current label speculatively and at once.  A round is a two-stage pipeline: while
// the cluster scores round r, four producer warps of the leader prepare round r + 1 as if round r ran to its end.
//   producers  (leader, warps 0-3), one warp per group of 32 hypotheses.  Sampling: in the common case each hypothesis'
//              three generator outputs are accepted by the distribution and distinct, read straight from the generator's
//              block (96 draws per group); otherwise the group's next 128 outputs go to a tape, lane h takes outputs from
//              3h on, and rejections and repeated samples shift the lanes behind them (a prefix sum, iterated to a fixed
//              point).  Group w starts as if the groups before it took 96 draws each; a small table and one named barrier
//              per pass tell the warps whether that held, and the groups behind one with extra draws go again.  A round
//              that does not settle takes the exact path (warp 0, group by group, draw by draw where it must).  The draws
//              each hypothesis took are remembered.
//              The generator's blocks of 624 words live in a ring of three, each made out of place from the one before,
//              so any position of the last ~1200 draws can be returned to by resetting a counter.
//              Then thread h turns its three ranks into pixels -- the k-th pixel of a label in image order follows from
//              the label's cells sorted by cell id (cells are painted whole), no per-pixel index lists; the divisions
//              are multiply-shifts -- builds the plane model (fp32, the reference's expression order) and pushes it into
//              every CTA's shared memory;
//   scorers    (every other warp of the cluster; in a 16-CTA cluster the leader does not score).  An equal share of the
//              label's points per warp; a warp stages 32 points in shared memory, then lane g scores hypotheses g,
//              g + 32, g + 64, g + 96 on each of them (the loss is a count, so the order of the points is free); per-warp
//              counts are summed per CTA and go to the leader with one distributed-shared-memory atomic per hypothesis;
//   warp 0     replays the reference's sequential loop over the 128 losses as a prefix minimum (best-so-far,
//              IsContinued), finds how many hypotheses were really run, and leaves the generator just after the last
//              of them when the search stops (the round prepared ahead is dropped); it pushes the continue flag and,
//              at the end of a label, the label's model to every CTA.
// FindInliers + the relabelling loop (which stops at the last inlier, plane_extractor.cpp:500-507) become two passes --
// the largest inlier pixel index, then "non-inlier before it -> 0" -- run one label late by the scoring warps, behind the
// producers' preparation of the next label's first round.  launch_refine (end of file) picks cluster size and register
// budget by batch size.
#include "refine.cuh"

#include <cstdio>
#include <mutex>

#include <cooperative_groups.h>

#include "exact_math.cuh"

#if defined(DPX_DEBUG_CHECKS) && !defined(DPX_REFINE_FORCE_RESEED)
#define DPX_REFINE_FORCE_RESEED  // the debug build also walks the generator's regeneration path (never taken otherwise)
#endif

namespace dpx {
namespace {

constexpr int kRefThreads = 512;
constexpr int kRefWarps = kRefThreads / 32;
// groups of 32 hypotheses per round: KSUB, a template parameter; 4 -> 128 hypotheses per cluster-wide pass over the label's points
constexpr int kRefCluster = 8;   // CTAs per frame (portable cluster size limit) ...
constexpr int kRefClusterWide = 16;  // ... or 16 (non-portable, opt-in) when the launch is a few frames only
constexpr int kTape = 128;      // generator outputs prepared per group of 32 hypotheses (3 draws each + slack)
constexpr int kMaxRows = 1023;  // cell rows the per-label row table can hold
constexpr int kSortLabels = 256;  // labels the shared-memory label sort handles (kRefWarps * kSortLabels <= kCellCache)
constexpr int kCellCache = 4096;  // cells of one label cached in shared memory (larger labels read the global list)
namespace cg = cooperative_groups;
constexpr unsigned kFullMask = 0xffffffffu;

constexpr unsigned kNoLoss = 0xffffffffu;  // "no hypothesis accepted yet" (the reference's HUGE_VAL best loss)

template <int KSUB>
struct RefShared {
  static constexpr int kHyp = 32 * KSUB;
  static constexpr int kRing = KSUB > 4 ? 5 : 3;  // generator blocks kept: a round and the round prepared ahead of it always fit
  uint32_t mtb[kRing][kMtN];  // leader: generator block b (the state after b twists of the seeded state) in slot b % kRing
  alignas(16) float model[2][kHyp][4];  // the hypotheses of the round being scored / being prepared, in every CTA
  alignas(16) unsigned loss[kHyp];      // leader: the cluster's totals
  unsigned loss_cta[kHyp];    // this CTA's share of a round
  int draws_cum[2][kHyp];     // generator draws of the round up to and including hypothesis g
  int rank[kHyp][3];          // sample ranks, ascending (std::set order): hand-over of the exact sequential path
  alignas(8) int2 grp[2][KSUB];  // per group of 32 hypotheses: (assumed offset into the round's draws, draws taken)
  alignas(16) float best[4];   // leader: best model of the label being searched
  alignas(16) float dbest[4];  // every CTA: best model of the label whose inlier passes are pending
  unsigned bestloss;          // kNoLoss until a hypothesis has been accepted
  int iteration, go_on;
  int prod_gp0;               // leader: stream position the round being prepared starts at
  int max_inlier_pix;
  float4 stage[kRefWarps][32];
  uint32_t tape[KSUB][kTape];  // per producer warp: tempered generator outputs of a group of 32 hypotheses, in draw order
  int rowstart[kMaxRows + 1]; // per label: index of the first of its cells in each cell row (cells are sorted)
  int32_t cells[kCellCache];  // per label, every CTA: a copy of the label's sorted cell list when it fits
};

#ifdef DPX_REFINE_PROBE
__device__ long long g_refine_probe_cnt[4];  // frame 0: generator blocks made, reseeds, rounds sampled draw by draw / in more than one pass
#endif
// ---- std::mt19937, executed by warp 0 of the leader (all lanes compute the same values) -----------------------
// The next block of 624 words from the current one, out of place:
//   next[i] = (i < 227 ? cur[i + 397] : next[i - 227]) ^ f(cur[i], i < 623 ? cur[i + 1] : next[0]);
// entries i >= 227 read entries of the new block, so the update runs in waves of 224 (<= 227, a multiple of 32).
template <int BASE>
__device__ __forceinline__ void mt_wave(uint32_t* next, const uint32_t* cur, int lane) {
  constexpr int END = BASE + 224 < kMtN ? BASE + 224 : kMtN;
  uint32_t v[7];
  // branch-free: the operand addresses are selected, every load is unconditional (lanes past the end compute a value they
  // do not store), and all the loads of a wave come before its stores
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    if (BASE + k * 32 >= END) continue;
    const int i = min(BASE + k * 32 + lane, kMtN - 1);
    const uint32_t* far = BASE + k * 32 + 31 < 227 ? cur + i + 397 : BASE + k * 32 >= 227 ? next + i - 227 : (i < 227 ? cur + i + 397 : next + i - 227);
    const uint32_t* nxt = BASE + k * 32 + 31 < kMtN - 1 ? cur + i + 1 : (i < kMtN - 1 ? cur + i + 1 : next);
    const uint32_t y = (cur[i] & 0x80000000u) | (*nxt & 0x7fffffffu);
    v[k] = *far ^ (y >> 1) ^ ((0u - (y & 1u)) & 0x9908b0dfu);
  }
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    if (BASE + k * 32 >= END) continue;
    const int i = BASE + k * 32 + lane;
    if (i < END) next[i] = v[k];
  }
  __syncwarp();
}
__device__ __forceinline__ void mt_next_block_warp(uint32_t* next, const uint32_t* cur, int lane) {
  mt_wave<0>(next, cur, lane);
  mt_wave<224>(next, cur, lane);
  mt_wave<448>(next, cur, lane);
}

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

// The generator as a stream: draw number t (counted from the seeded state, whose own 624 words are "used up", so the
// first real draw is t = 624) is word t % 624 of block t / 624.  Blocks gen_hi - 2 .. gen_hi are in the ring; a block
// ahead is generated on demand, a block that has left the ring is regenerated from the seed (never in practice: a round
// would have to take several hundred draws more than its 384).  `gen_hi` lives in a register of every lane of warp 0.
template <class S>
__device__ __forceinline__ void gen_cover(S& s, const uint32_t* mt_init, int lane, int& gen_hi, int t_lo, int t_hi) {
  const int b_lo = t_lo / kMtN, b_hi = t_hi / kMtN;
#ifdef DPX_REFINE_FORCE_RESEED  // test builds: take the regeneration path on every step back (tools/gpu_debug_checks.sh)
  if (b_lo < gen_hi) {
#else
  if (b_lo < gen_hi - (S::kRing - 1)) {
#endif
    for (int i = lane; i < kMtN; i += 32) s.mtb[0][i] = mt_init[i];
    __syncwarp();
    gen_hi = 0;
#ifdef DPX_REFINE_PROBE
    if (threadIdx.x == 0 && blockIdx.x == 0) g_refine_probe_cnt[1] += 1;
#endif
  }
  while (gen_hi < b_hi) {
    mt_next_block_warp(s.mtb[(gen_hi + 1) % S::kRing], s.mtb[gen_hi % S::kRing], lane);
    ++gen_hi;
#ifdef DPX_REFINE_PROBE
    if (threadIdx.x == 0 && blockIdx.x == 0) g_refine_probe_cnt[0] += 1;
#endif
  }
}
template <class S>
__device__ __forceinline__ uint32_t gen_word(const S& s, int t) {
  const int b = t / kMtN;
  return mt_temper(s.mtb[b % S::kRing][t - b * kMtN]);
}
// The same for the next 624 draws from a fixed position on, without a division per word: the position's block and the
// one after it.
struct GenWindow {
  const uint32_t* cur;   // block holding draw t0, already offset to it: cur[j] is draw t0 + j while j < left
  const uint32_t* next;  // the block after, offset so that next[j] is draw t0 + j for j >= left
  int left;              // draws left in the first block
};
template <class S>
__device__ __forceinline__ GenWindow gen_window(const S& s, int t0) {
  const int b = t0 / kMtN, off = t0 - b * kMtN, slot = b % S::kRing;
  GenWindow w;
  w.left = kMtN - off;
  w.cur = s.mtb[slot] + off;
  w.next = s.mtb[slot == S::kRing - 1 ? 0 : slot + 1] - w.left;
  return w;
}
__device__ __forceinline__ uint32_t gen_word(const GenWindow& w, int j) { return mt_temper(j < w.left ? w.cur[j] : w.next[j]); }

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
// one value of the distribution, drawn word by word from position t on (t is advanced past the words used)
template <class S>
__device__ __forceinline__ int uniform_below_warp(S& s, const uint32_t* mt_init, int lane, int& gen_hi, int& t, const UniformMap& m) {
  int value;
  uint32_t u;
  do {
    if (t / kMtN > gen_hi || t / kMtN < gen_hi - (S::kRing - 1)) gen_cover(s, mt_init, lane, gen_hi, t, t);
    u = gen_word(s, t);
    ++t;
  } while (!accept_draw(m, u, value));
  return value;
}

struct LabelCells {
  const int32_t* cells;  // the label's cells, ascending cell id
  int count;             // number of cells
};

// Division by a number fixed for the whole kernel (patch size, cells per row): multiply-shift with a round-up multiplier,
// exact for dividends in [0, 2^31) (Granlund-Montgomery: m = ceil(2^(31+l) / d), l = ceil(log2 d), fits 32 bits).
// The index arithmetic of this kernel is divisions, and an integer division by a run-time value is ~25 instructions.
struct FastDiv {
  uint32_t mul;
  int shift;
};
__device__ __forceinline__ FastDiv make_fastdiv(int d) {
  const int l = 32 - __clz(static_cast<unsigned>(d) - 1u);
  FastDiv f;
  f.mul = static_cast<uint32_t>(((1ull << (31 + l)) + static_cast<unsigned>(d) - 1ull) / static_cast<unsigned>(d));
  f.shift = 31 + l;
  return f;
}
__device__ __forceinline__ int fdiv(int x, const FastDiv& f) {
  return static_cast<int>((static_cast<unsigned long long>(static_cast<unsigned>(x)) * f.mul) >> f.shift);
}
// u / m for 0 <= u, 0 < m, both below 2^24 when `small`: the fp32 quotient is off by at most one, fixed up exactly
__device__ __forceinline__ int div_small(int u, int m, bool small) {
  if (!small) return u / m;
  int q = static_cast<int>(__fmul_rn(static_cast<float>(u), __frcp_rn(static_cast<float>(m))));
  const int rem = u - q * m;
  if (rem < 0) --q;
  else if (rem >= m) ++q;
  return q;
}
struct IndexMath {
  FastDiv by_p2, by_nh, by_p;
  int p, p2, nh, width;
  bool narrow;  // width < 2^24
};

// The k-th pixel (image order) among the pixels of a label whose cells are `lc` (plane_extractor.cpp:473-478 builds
// this list explicitly).  A cell row holding m of the label's cells contributes p image rows of m*p pixels each.
// rowstart[r] = index of the label's first cell in cell row r (rowstart[nv] = count), or nullptr to search.
__device__ __forceinline__ long long kth_pixel(const LabelCells& lc, const int* rowstart, int k, const IndexMath& im) {
  const int t = fdiv(k, im.by_p2);
  const int r = fdiv(lc.cells[t], im.by_nh);  // cell row containing rank k
  int s, e;
  if (rowstart) {
    s = rowstart[r];
    e = rowstart[r + 1];
  } else {
    s = t;
    e = t + 1;
    while (s > 0 && fdiv(lc.cells[s - 1], im.by_nh) == r) --s;
    while (e < lc.count && fdiv(lc.cells[e], im.by_nh) == r) ++e;
  }
  const int m = e - s;
  const int kk = k - s * im.p2;          // rank inside the row: kk = (i * m + j) * p + xo
  const int u = fdiv(kk, im.by_p), xo = kk - u * im.p;
  const int i = div_small(u, m, im.narrow), j = u - i * m;  // u < m * p <= width
  const int cell = lc.cells[s + j];
  const int q = cell - r * im.nh;
  return static_cast<long long>(r * im.p + i) * im.width + q * im.p + xo;
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
__device__ long long g_refine_probe[24];  // [0, 12): the leader's thread 0 (a producer); [12, 24): thread 0 of CTA 1 (a scorer)
// (accumulated in registers and written once at the end: a global read-modify-write per probe would cost more than most phases)
#define REF_PROBE(slot) do { const long long t__ = clock64(); rp_acc[slot] += t__ - rp_t; rp_t = t__; } while (0)
#else
#define REF_PROBE(slot) do { } while (0)
#endif

// MINB = CTAs per SM the register budget is set for: 1 (~125 registers, nothing spilled) when the launch has no more CTAs
// than the GPU has SMs -- the latency case -- and 2 (64 registers) for batches, where a second resident cluster per SM
// fills the first one's barriers and serial phases.
template <int LAYOUT, int MINB, int CL, int KSUB>
__global__ void __launch_bounds__(kRefThreads, MINB) refine_kernel(const RefineArgs args) {
  constexpr int kRefCluster = CL;  // CTAs of this frame's cluster
  constexpr int kSub = KSUB, kHyp = 32 * KSUB, kProdWarps = KSUB;  // one producer warp per group of 32 hypotheses
  using Shared = RefShared<KSUB>;
  extern __shared__ __align__(16) unsigned char ref_smem[];
  Shared& s = *reinterpret_cast<Shared*>(ref_smem);
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned crank = cluster.block_rank();
  const bool leader = crank == 0;
  Shared* lead = cluster.map_shared_rank(&s, 0);  // the leader CTA's copy (distributed shared memory)
  const Geometry& g = args.geom;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int frame = blockIdx.x / kRefCluster;
#ifdef DPX_REFINE_PROBE
  long long rp_t = clock64();
  long long rp_acc[12] = {};
#endif
  const int C = g.n_cells, p = g.patch, p2 = p * p, nh = g.nh;
  const long long fc = static_cast<long long>(frame) * C;
  IndexMath im;
  im.by_p2 = make_fastdiv(p2); im.by_nh = make_fastdiv(nh); im.by_p = make_fastdiv(p);
  im.p = p; im.p2 = p2; im.nh = nh; im.width = g.width;
  im.narrow = g.width < (1 << 24);
  const int nseg = args.tables.n_planes[frame];
  if (nseg <= 0) {  // no plane: nothing is labelled (plane_extractor.cpp:230-232); the whole cluster leaves
    if (leader && tid == 0 && args.work) args.work[2 * frame] = args.work[2 * frame + 1] = 0;
    return;
  }

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
  // during the search the leader's producer warps do not score: the other warps of the cluster share the points
  const bool producer = leader && warp < kProdWarps;
  // leader warps [0, kLeaderIdle) do not score: the producers, and in a 16-CTA cluster the whole leader (its scoring warps
  // are a twentieth of the cluster's and slow the producers down, which is the longer leg there)
  constexpr int kLeaderIdle = CL > 8 ? kRefWarps : kProdWarps;
  constexpr int kScoreWarps = kRefCluster * kRefWarps - kLeaderIdle;
  const int swarp = gwarp - kLeaderIdle;

  // the generator (warp 0 of the leader): position of the next draw in the stream, newest block in the ring
  int gp = kMtN, gen_hi = 0;
  unsigned long long work_points = 0, work_rounds = 0;  // (leader, thread 0) what the frame's search scored
  if (tid < kHyp) s.loss_cta[tid] = 0;
  __syncthreads();
  if (leader) {
    // ---- labels_indices (plane_extractor.cpp:473-478), per cell instead of per pixel ---------------------
    // lab_cells = the cells sorted by (label, cell id), lab_end[l] = end of label l's run: a stable counting sort.
    for (int i = tid; i < kMtN; i += kRefThreads) s.mtb[0][i] = args.mt_init[i];
    if (nseg <= kSortLabels) {
      // Every warp takes a contiguous sixteenth of the cells: per-(warp, label) counts in shared memory, a scan in
      // (label, warp) order, then each warp places its cells in order behind its own cursors.
      int* mat = s.cells;        // [kRefWarps][nseg] counts, then cursors relative to the label's run
      int* lstart = s.rowstart;  // [nseg] totals, then run starts
      for (int i = tid; i < kRefWarps * nseg; i += kRefThreads) mat[i] = 0;
      __syncthreads();
      const int chunk = ((C + kRefWarps - 1) / kRefWarps + 31) & ~31;
      const int c_begin = min(C, warp * chunk), c_end = min(C, c_begin + chunk);
      for (int c = c_begin + lane; c < c_end; c += 32) {
        const int l = cell_label[c];
        if (l > 0) atomicAdd(&mat[warp * nseg + l - 1], 1);
      }
      __syncthreads();
      if (tid < nseg) {
        int run = 0;
        for (int w = 0; w < kRefWarps; ++w) {
          const int t = mat[w * nseg + tid];
          mat[w * nseg + tid] = run;
          run += t;
        }
        lstart[tid] = run;
      }
      __syncthreads();
      if (warp == 0) {
        int run = 0;
        for (int b0 = 0; b0 < nseg; b0 += 32) {
          const int i = b0 + lane;
          const int cnt = i < nseg ? lstart[i] : 0;
          int incl = cnt;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(kFullMask, incl, o);
            if (lane >= o) incl += t;
          }
          if (i < nseg) {
            lstart[i] = run + incl - cnt;
            lab_end[i] = run + incl;
          }
          run += __shfl_sync(kFullMask, incl, 31);
        }
      }
      __syncthreads();
      int l_next = c_begin + lane < c_end ? cell_label[c_begin + lane] : 0;
      for (int c0 = c_begin; c0 < c_end; c0 += 32) {
        const int c = c0 + lane;
        const int l = l_next;
        l_next = c + 32 < c_end ? cell_label[c + 32] : 0;  // in flight while this batch is placed
        const unsigned act = __ballot_sync(kFullMask, l > 0);
        if (l > 0) {
          const unsigned grp = __match_any_sync(act, l);
          const int ldr = __ffs(grp) - 1;
          int base = 0;
          if (lane == ldr) {
            base = mat[warp * nseg + l - 1];
            mat[warp * nseg + l - 1] = base + __popc(grp);
          }
          base = __shfl_sync(grp, base, ldr);
          lab_cells[lstart[l - 1] + base + __popc(grp & ((1u << lane) - 1u))] = c;
        }
        __syncwarp();
      }
    } else {
    for (int i = tid; i < nseg; i += kRefThreads) lab_end[i] = 0;
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
    }
    __threadfence();
  }
  cluster.sync();  // the label runs (global memory) are visible to the whole cluster

  // ---- one label after the other (plane_extractor.cpp:486-508) ----------------------------------------------
  // FindInliers + relabelling of a label (RANSAC.hpp:53-62, plane_extractor.cpp:498-507) run one label late, by the
  // scoring warps, while the producers prepare the first round of the next label: pass 0 (largest inlier pixel) before
  // that round's barrier, pass 1 (non-inliers before it -> 0; points after the last inlier keep their label, as in the
  // reference's loop) right after it.  The pending label's cells are read from the global list (the shared-memory
  // copy already holds the next label's), its best model from `dbest`.
  bool pending = false;
  LabelCells plc = {nullptr, 0};
  int pn = 0;
  auto inlier_pass = [&](const int pass, const int last) {
    if (swarp < 0 || (pass == 1 && last < 0)) return;  // no inlier at all: the relabelling loop never runs
    const float m[4] = {s.dbest[0], s.dbest[1], s.dbest[2], s.dbest[3]};
    const int e_lo = static_cast<int>(static_cast<long long>(swarp) * pn / kScoreWarps);
    const int e_hi = static_cast<int>(static_cast<long long>(swarp + 1) * pn / kScoreWarps);
    int local_max = -1;
    for (int e = e_lo + lane; e < e_hi; e += 32) {
      const int t = fdiv(e, im.by_p2), in = e - t * p2;
      const int cell = __ldg(plc.cells + t);
      const int r = fdiv(cell, im.by_nh), q = cell - r * nh;
      const int i = fdiv(in, im.by_p), j = in - i * p;
      const long long pix = static_cast<long long>(r * p + i) * g.width + q * p + j;
      float x, y, z;
      load_point<LAYOUT>(xyz, g.n_points, pix, x, y, z);
      const bool inlier = ::fabsf(plane_error(m, x, y, z)) < thr_f;
      if (pass == 0) {
        if (inlier) local_max = max(local_max, static_cast<int>(pix));
      } else if (!inlier && pix < last) {
        labels[pix] = 0;
      }
    }
    if (pass == 0) {
      local_max = __reduce_max_sync(kFullMask, local_max);
      if (lane == 0 && local_max >= 0) atomicMax(&lead->max_inlier_pix, local_max);
    }
  };
  auto last_inlier = [&]() -> int {  // after the cluster barrier that follows pass 0
    int v = 0;
    if (lane == 0) v = lead->max_inlier_pix;
    return __shfl_sync(kFullMask, v, 0);
  };

  for (int L = 0; L < nseg; ++L) {
    const int start = L ? lab_end[L - 1] : 0;
    LabelCells lc;
    lc.cells = lab_cells + start;
    lc.count = lab_end[L] - start;
    if (lc.count == 0) continue;  // labels_indices[label].size() == 0 (the same decision in every CTA)
    const LabelCells glc = lc;    // the global list
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
      s.bestloss = kNoLoss;
      s.iteration = 0;
      s.max_inlier_pix = -1;  // of the pending label: its pass 0 runs after the next barrier
      // IsContinued(0, N - HUGE_VAL, N): int(-inf) is INT_MIN on x86-64
      s.go_on = (0 < args.max_iterations) && (static_cast<double>(INT_MIN) < ratio * n);
    }
    cluster.sync();
    int go_on = lead->go_on;

    // ---- FindBest (RANSAC.hpp:25-51), kHyp hypotheses per round ------------------------------------------------
    // The producers' work for one round: samples -> pixels -> models of hypotheses [0, kHyp) drawn from position gp0 of
    // the generator's stream, into buffer `buf`.  Returns the draws taken (warp 0).
    auto produce = [&](const int buf, const int gp_w0) -> int {  // gp_w0: the stream position, known to warp 0
      const UniformMap umap = make_uniform_map(static_cast<uint32_t>(n), args.uniform_variant);
      // One group of 32 hypotheses from the stream position t_first (RANSAC.hpp:81-87).  Almost always every
      // iteration takes exactly three draws (a rejection in the distribution or a repeated sample has probability ~3/n),
      // so the next kTape generator outputs are tempered in parallel and lane h takes outputs from 3h on.  Lane h
      // consumes tape entries from 3h + (extra draws of the hypotheses before it) until it holds three distinct accepted
      // values; the extra draws shift everything behind them, so the offsets are iterated to a fixed point: a prefix sum
      // of the lanes' extra draws per pass, normally one pass.  Returns the draws the group took, or -1 if it did not
      // settle inside the tape; `cum` = draws up to and including this lane's hypothesis.
      auto sample_group = [&](const int t_first, int& a, int& b, int& c, int& cum) -> int {
        const GenWindow win = gen_window(s, t_first);
        {
          // the common case first, straight from the generator's block: every lane's three outputs are accepted and
          // distinct, so the group takes exactly 96 draws
          int v0, v1, v2;
          const bool k0 = accept_draw(umap, gen_word(win, 3 * lane), v0);
          const bool k1 = accept_draw(umap, gen_word(win, 3 * lane + 1), v1);
          const bool k2 = accept_draw(umap, gen_word(win, 3 * lane + 2), v2);
          if (__all_sync(kFullMask, k0 && k1 && k2 && v0 != v1 && v0 != v2 && v1 != v2)) {
            const int lo = min(v0, v1), hi = max(v0, v1);
            a = min(lo, v2);
            c = max(hi, v2);
            b = max(lo, min(hi, v2));
            cum = 3 * (lane + 1);
            return 96;
          }
        }
        uint32_t* tape = s.tape[warp];
        for (int j = lane; j < kTape; j += 32) tape[j] = gen_word(win, j);
        __syncwarp();
        int off = 0, extra = 0;
        bool ok = false;
        for (int pass = 0; pass < 6; ++pass) {
          int pos = 3 * lane + off, cnt = 0;
          a = b = c = -1;
          while (cnt < 3 && pos < kTape) {
            int vv;
            if (!accept_draw(umap, tape[pos++], vv)) continue;  // the distribution draws again
            if (vv == a || vv == b || vv == c) continue;        // std::set already holds it
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
          if (!__all_sync(kFullMask, complete)) break;  // ran off the tape
        }
        __syncwarp();
        cum = 3 * (lane + 1) + off + extra;
        return ok ? __shfl_sync(kFullMask, cum, 31) : -1;
      };
      auto prod_bar = [] { asm volatile("bar.sync 1, %0;" ::"n"(kProdWarps * 32) : "memory"); };

      // The four groups of the round are sampled side by side, one per producer warp, group w assuming that the groups
      // before it took exactly 96 draws each.  Where that turns out wrong (a group with extra draws) the groups behind
      // it go again from their true offsets; group w is certainly right after pass w.
      REF_PROBE(11);
      if (warp == 0) {
        gen_cover(s, args.mt_init, lane, gen_hi, gp_w0, gp_w0 + kSub * kTape - 1);
        if (lane == 0) s.prod_gp0 = gp_w0;
      }
      REF_PROBE(9);  // generator blocks
      prod_bar();
      const int gp0 = s.prod_gp0;
      int start = 96 * warp, used = 0, a = -1, b = -1, c = -1, cum = 0;
      int grp_start = -1, used_own = 0;  // the offset this warp's current result was computed at, the draws it took
      bool settled = false, serial = false;
#ifdef DPX_REFINE_PROBE
      bool grp_multi = false;
#endif
      for (int it = 0; it < kSub && !settled; ++it) {
#ifdef DPX_REFINE_PROBE
        grp_multi = it > 0;
#endif
        if (it == 0 || grp_start != start) {
          used_own = sample_group(gp0 + start, a, b, c, cum);
          grp_start = start;
        }
        if (lane == 0) s.grp[it & 1][warp] = make_int2(grp_start, used_own);  // (two tables: one barrier per pass)
        prod_bar();
        REF_PROBE(10);  // group sampling
        int sum = 0, mine_start = 0;
        settled = true;
#pragma unroll
        for (int w = 0; w < kSub; ++w) {
          const int2 gw = s.grp[it & 1][w];
          if (w == warp) mine_start = sum;
          if (gw.y < 0) serial = true;
          if (gw.x != sum) settled = false;
          sum += gw.y;
        }
        if (serial) break;
        start = mine_start;
        used = sum;  // total of the round (final once settled)
      }
#ifdef DPX_REFINE_PROBE
      if (tid == 0 && blockIdx.x == 0 && grp_multi) g_refine_probe_cnt[3] += 1;
      if (tid == 0 && blockIdx.x == 0 && (serial || !settled)) g_refine_probe_cnt[2] += 1;
#endif
      if (serial || !settled) {
        // exact path: warp 0 walks the groups in order, a group that does not settle draw by draw
        if (warp == 0) {
          int t0 = gp0;
          for (int sub = 0; sub < kSub; ++sub) {
            gen_cover(s, args.mt_init, lane, gen_hi, t0, t0 + kTape - 1);
            int ga, gb, gc, gcum;
            const int gused = sample_group(t0, ga, gb, gc, gcum);
            if (gused >= 0) {
              const int h = 32 * sub + lane;
              s.rank[h][0] = ga; s.rank[h][1] = gb; s.rank[h][2] = gc;
              s.draws_cum[buf][h] = t0 - gp0 + gcum;
              t0 += gused;
            } else {
              for (int hh = 0; hh < 32; ++hh) {
                int sa = -1, sb = -1, sc = -1, cnt = 0;  // the std::set<int>, kept sorted
                while (cnt < 3) {
                  const int vv = uniform_below_warp(s, args.mt_init, lane, gen_hi, t0, umap);
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
                  s.draws_cum[buf][h] = t0 - gp0;
                }
              }
            }
            __syncwarp();
          }
          used = t0 - gp0;
        }
        prod_bar();
        a = s.rank[tid][0]; b = s.rank[tid][1]; c = s.rank[tid][2];
      } else {
        s.draws_cum[buf][tid] = start + cum;
      }
      REF_PROBE(0);  // sampling
      // this thread's hypothesis: ranks -> pixels -> model
      const int* rs = rows_ok ? s.rowstart : nullptr;
      const long long pa = kth_pixel(lc, rs, a, im), pb = kth_pixel(lc, rs, b, im), pc = kth_pixel(lc, rs, c, im);
      REF_PROBE(2);
      {
        // PlaneEstimator::ComputeModel (Plane.hpp:13-43), fp32 in the reference's expression order
        float x0, y0, z0, x1, y1, z1, x2, y2, z2;
        load_point<LAYOUT>(xyz, g.n_points, pa, x0, y0, z0);
        load_point<LAYOUT>(xyz, g.n_points, pb, x1, y1, z1);
        load_point<LAYOUT>(xyz, g.n_points, pc, x2, y2, z2);
#ifdef DPX_REFINE_PROBE
        if (x0 + x1 + x2 + y0 + y1 + y2 + z0 + z1 + z2 == 12345.678f) rp_t += 1;  // wait for the loads
        REF_PROBE(4);
#endif
        const f32 X0(x0), X1(x1), X2(x2), Y0(y0), Y1(y1), Y2(y2), Z0(z0), Z1(z1), Z2(z2);
        const f32 D = X0 * Y1 - X1 * Y0 - X0 * Y2 + X2 * Y0 + X1 * Y2 - X2 * Y1;
        const f32 a = (Z0 * (Y1 - Y2)) / D - (Z1 * (Y0 - Y2)) / D + (Z2 * (Y0 - Y1)) / D;
        const f32 b = (Z1 * (X0 - X2)) / D - (Z0 * (X1 - X2)) / D - (Z2 * (X0 - X1)) / D;
        const f32 d = (Z2 * (X0 * Y1 - X1 * Y0)) / D - (Z1 * (X0 * Y2 - X2 * Y0)) / D + (Z0 * (X1 * Y2 - X2 * Y1)) / D;
        const f32 c(-1.0f);
        // `sqrt(float)` in Plane.hpp:36 resolves to ::sqrt(double): double square root, rounded to float on assignment
        const f32 l(__double2float_rn(__dsqrt_rn(static_cast<double>((a * a + b * b + c * c).v))));
        // pushed into every CTA's shared memory (fire-and-forget stores, visible after the next cluster barrier): 512
        // threads per CTA fetching them from the leader afterwards queue up on its distributed-shared-memory port
        const float4 mv = make_float4((a / l).v, (b / l).v, (c / l).v, (d / l).v);
#pragma unroll
        for (int r = 0; r < kRefCluster; ++r) *reinterpret_cast<float4*>(cluster.map_shared_rank(&s.model[buf][tid][0], r)) = mv;
      }
      REF_PROBE(1);  // ranks -> pixels, models
      return used;
    };

    if (go_on) {
      int buf = 0, draws_cur = 0, draws_next = 0;
      REF_PROBE(6);
      if (producer) draws_cur = produce(0, gp);  // the first round of a label hides behind the previous label's pass 0
      else if (pending) inlier_pass(0, -1);
      if (leader && tid < kHyp) s.loss[tid] = 0;
      cluster.sync();  // round 0's models are in every CTA, the leader's counters are zero, the pending label's last inlier is known
      REF_PROBE(7);
      if (pending && !producer) inlier_pass(1, last_inlier());
      pending = false;
      for (;;) {
        if (producer) {
          // the next round, as if this one ran all its iterations (it does, except the last round of a label)
          draws_next = produce(buf ^ 1, gp + draws_cur);
        } else if (swarp >= 0) {
          // EvaluateModel (RANSAC.hpp:89-98): lane g scores hypotheses g, g + 32, ...; loss += (fabs(error) >= threshold)
          float m[kSub][4];
#pragma unroll
          for (int j = 0; j < kSub; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) m[j][k] = s.model[buf][32 * j + lane][k];
          unsigned loss[kSub];
#pragma unroll
          for (int j = 0; j < kSub; ++j) loss[j] = 0;
          // an equal share of the label's points for every scoring warp (dealing whole chunks of 32 leaves some warps a
          // chunk more than others: a quarter of the round idle on a 13 000-point label)
          const int e_lo = static_cast<int>(static_cast<long long>(swarp) * n / kScoreWarps);
          const int e_hi = static_cast<int>(static_cast<long long>(swarp + 1) * n / kScoreWarps);
          for (int e0 = e_lo; e0 < e_hi; e0 += 32) {
            const int e = e0 + lane;
            float x = 0.f, y = 0.f, z = 0.f;
            if (e < e_hi) {
              const int t = fdiv(e, im.by_p2), in = e - t * p2;
              const int cell = lc.cells[t];
              const int r = fdiv(cell, im.by_nh), q = cell - r * nh;
              const int i = fdiv(in, im.by_p), j = in - i * p;
              load_point<LAYOUT>(xyz, g.n_points, static_cast<long long>(r * p + i) * g.width + q * p + j, x, y, z);
            }
            s.stage[warp][lane] = make_float4(x, y, z, 0.f);
            __syncwarp();
            const int cnt = min(32, e_hi - e0);
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
          // (all warps adding straight into the leader's counters serialise there)
#pragma unroll
          for (int j = 0; j < kSub; ++j)
            if (loss[j]) atomicAdd(&s.loss_cta[32 * j + lane], loss[j]);
          REF_PROBE(8);  // scoring (the leader's first scoring warp)
        }
        __syncthreads();
        REF_PROBE(2);  // producers: waiting for the scorers; scorers: waiting for the CTA
        if (tid < kHyp) {
          const unsigned v = s.loss_cta[tid];
          s.loss_cta[tid] = 0;
          if (v) atomicAdd(&lead->loss[tid], v);
        }
        cluster.sync();  // (1) every CTA's counts are in
        REF_PROBE(3);
        if (leader && warp == 0) {
          // The reference's sequential loop over these hypotheses (RANSAC.hpp:33-46), evaluated by one warp: lane l owns
          // hypotheses kSub * l .. kSub * l + kSub - 1.  The best-so-far loss before a hypothesis is a prefix minimum; the loop runs while
          // IsContinued holds, which is monotone (the best loss only falls, the iteration count only grows), so the number
          // of iterations really run is the number of hypotheses whose check passes.  Losses are counts: the reference's
          // doubles hold the same integers, HUGE_VAL is kNoLoss here.
          work_points += static_cast<unsigned long long>(n);
          work_rounds += 1;
          const unsigned best0 = s.bestloss;
          const int iter0 = s.iteration;
          unsigned mine[kSub], pre[kSub];  // this lane's losses; their running minimum
#pragma unroll
          for (int q4 = 0; q4 < kSub / 4; ++q4) {
            const uint4 v4 = *reinterpret_cast<const uint4*>(&s.loss[kSub * lane + 4 * q4]);
            *reinterpret_cast<uint4*>(&s.loss[kSub * lane + 4 * q4]) = make_uint4(0, 0, 0, 0);  // for the next round
            mine[4 * q4] = v4.x; mine[4 * q4 + 1] = v4.y; mine[4 * q4 + 2] = v4.z; mine[4 * q4 + 3] = v4.w;
          }
          pre[0] = mine[0];
#pragma unroll
          for (int j = 1; j < kSub; ++j) pre[j] = min(pre[j - 1], mine[j]);
          unsigned incl = pre[kSub - 1];
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(kFullMask, incl, o);
            if (lane >= o) incl = min(incl, t);
          }
          unsigned ex = __shfl_up_sync(kFullMask, incl, 1);
          if (lane == 0) ex = kNoLoss;
          ex = min(ex, best0);
          const double target = ratio * n;
          int consumed = 0;
#pragma unroll
          for (int j = 0; j < kSub; ++j) {
            // IsContinued(iteration, N - best loss, N); int(N - HUGE_VAL) is INT_MIN on x86-64
            const unsigned before = j ? min(ex, pre[j - 1]) : ex;
            const int inl = before == kNoLoss ? INT_MIN : n - static_cast<int>(before);
            const bool go_h = (iter0 + kSub * lane + j < args.max_iterations) && (static_cast<double>(inl) < target);
            consumed += __popc(__ballot_sync(kFullMask, go_h));
          }
          // best loss over the iterations really run, and the first of them that reaches it (strict '<' updates)
          unsigned cand = kNoLoss;
#pragma unroll
          for (int j = 0; j < kSub; ++j)
            if (kSub * lane + j < consumed) cand = min(cand, mine[j]);
          const unsigned best = __reduce_min_sync(kFullMask, cand);
          unsigned bl = best0;
          if (best < best0) {
            unsigned first = kNoLoss;
#pragma unroll
            for (int j = kSub - 1; j >= 0; --j)
              if (kSub * lane + j < consumed && mine[j] == best) first = kSub * lane + j;
            const unsigned winner = __reduce_min_sync(kFullMask, first);
            if (lane < 4) s.best[lane] = s.model[buf][winner][lane];
            bl = best;
          }
          bool go = consumed == kHyp;
          if (go) {
            const int inl2 = bl == kNoLoss ? INT_MIN : n - static_cast<int>(bl);
            go = (iter0 + consumed < args.max_iterations) && (static_cast<double>(inl2) < target);
          }
          if (lane == 0) {
            s.bestloss = bl;
            s.iteration = iter0 + consumed;
          }
          // the generator stops just after the last iteration the reference ran; the round prepared ahead is dropped
          gp += consumed > 0 ? s.draws_cum[buf][consumed - 1] : 0;
          if (lane < kRefCluster) *cluster.map_shared_rank(&s.go_on, lane) = go ? 1 : 0;
          if (!go) {
            // the label's model goes to every CTA for the inlier passes
            __syncwarp();
            const float4 bm = *reinterpret_cast<const float4*>(s.best);
            if (lane < kRefCluster) *reinterpret_cast<float4*>(cluster.map_shared_rank(&s.dbest[0], lane)) = bm;
          }
        }
        REF_PROBE(4);  // evaluate
        cluster.sync();  // (2) the decision is visible
        REF_PROBE(5);
        if (!s.go_on) break;
        buf ^= 1;
        draws_cur = draws_next;
      }
    }

    else {
      // no search (ransacMaxIterations = 0): the model stays Eigen::Vector4f::Zero()
      if (pending) {
        inlier_pass(0, -1);
        cluster.sync();
        inlier_pass(1, last_inlier());
      }
      __syncthreads();
      if (tid < 4) s.dbest[tid] = 0.f;
      __syncthreads();
    }
    pending = true;
    plc = glc;
    pn = n;
  }
  if (pending) {
    if (leader && tid == 0) s.max_inlier_pix = -1;
    cluster.sync();
    inlier_pass(0, -1);
    cluster.sync();  // the largest inlier pixel of the whole label is known
    inlier_pass(1, last_inlier());
  }
  cluster.sync();  // no CTA leaves while another may still touch its shared memory
  if (leader && tid == 0 && args.work) {
    args.work[2 * frame] = work_points;
    args.work[2 * frame + 1] = work_rounds;
  }
#ifdef DPX_REFINE_PROBE
  if (tid == 0 && blockIdx.x < 2)  // frame 0: the leader's thread 0 (a producer) and thread 0 of the next CTA (a scorer)
    for (int i = 0; i < 12; ++i) g_refine_probe[(blockIdx.x ? 12 : 0) + i] += rp_acc[i];
#endif
}

}  // namespace

#ifdef DPX_REFINE_PROBE
}  // namespace dpx
extern "C" __attribute__((visibility("default"))) int dpx_debug_refine_probe(long long* out, int reset) {
  cudaDeviceSynchronize();
  if (out) cudaMemcpyFromSymbol(out, dpx::g_refine_probe, sizeof(long long) * 24);
  {
    long long cnt[4];
    cudaMemcpyFromSymbol(cnt, dpx::g_refine_probe_cnt, sizeof(cnt));
    fprintf(stderr, "[refine probe] since load, frame 0: generator blocks %lld, reseeds %lld, rounds sampled draw by draw %lld, in several passes %lld\n", cnt[0], cnt[1], cnt[2], cnt[3]);
  }
  if (reset) {
    long long zero[24] = {};
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

template <int LAYOUT, int MINB, int CL>
cudaError_t launch_refine_as(const RefineArgs& args, cudaStream_t stream, bool probe_only) {
  // 256-hypothesis rounds (KSUB = 8) were measured with 16-CTA clusters: 5 % faster on the TUM frame, 40 % slower on the ICL
  // frame (many small labels: more rounds that do not settle at once, twice the work thrown away at the end of each label)
  constexpr int KSUB = 4;
  // (function attributes belong to the current device: set on every launch, like the other stages do -- one process may
  // drive several GPUs, dpx_sequence)
  cudaError_t e = cudaFuncSetAttribute(refine_kernel<LAYOUT, MINB, CL, KSUB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(sizeof(RefShared<KSUB>)));
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(args.n_frames) * CL, 1, 1);
  cfg.blockDim = dim3(kRefThreads, 1, 1);
  cfg.dynamicSmemBytes = sizeof(RefShared<KSUB>);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (CL > 8) {
    e = cudaFuncSetAttribute(refine_kernel<LAYOUT, MINB, CL, KSUB>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
  }
  if (probe_only) {
    // can this device place the clusters of the launch side by side at all?
    int n_clusters = 0;
    e = cudaOccupancyMaxActiveClusters(&n_clusters, refine_kernel<LAYOUT, MINB, CL, KSUB>, &cfg);
    if (e != cudaSuccess) return e;
    return n_clusters >= args.n_frames ? cudaSuccess : cudaErrorInvalidConfiguration;
  }
  return cudaLaunchKernelEx(&cfg, refine_kernel<LAYOUT, MINB, CL, KSUB>, args);
}

cudaError_t launch_refine(const RefineArgs& args, cudaStream_t stream) {
  if (args.n_frames == 0 || args.geom.n_cells == 0) return cudaSuccess;
  // per device (one process may drive several): SM count, and whether n 16-CTA clusters can be resident side by side
  struct DeviceFacts {
    int n_sm = 0;
    signed char wide_ok[2][16] = {};  // [layout][frames]: 0 unknown, 1 yes, -1 no
  };
  static DeviceFacts facts[64];
  static std::mutex mu;
  int dev = 0;
  cudaGetDevice(&dev);
  const bool row = args.layout == kLayoutRowMajor;
  bool wide = false;
  int n_sm = 0;
  {
    std::lock_guard<std::mutex> lock(mu);
    DeviceFacts& f = facts[dev & 63];
    if (f.n_sm == 0) cudaDeviceGetAttribute(&f.n_sm, cudaDevAttrMultiProcessorCount, dev);
    n_sm = f.n_sm;
    // A handful of frames: 16 CTAs per frame halve the scoring time of a round (the clusters must all be resident at once
    // for that to be a gain; asked of the occupancy calculator once per device and frame count).
    if (args.n_frames * kRefClusterWide <= n_sm && args.n_frames < 16) {
      signed char& ok = f.wide_ok[row ? 0 : 1][args.n_frames];
      if (ok == 0) {
        const cudaError_t e = row ? launch_refine_as<kLayoutRowMajor, 1, kRefClusterWide>(args, stream, true)
                                  : launch_refine_as<kLayoutColMajor, 1, kRefClusterWide>(args, stream, true);
        ok = e == cudaSuccess ? 1 : -1;
        (void)cudaGetLastError();
      }
      wide = ok > 0;
    }
  }
  if (wide)
    return row ? launch_refine_as<kLayoutRowMajor, 1, kRefClusterWide>(args, stream, false)
               : launch_refine_as<kLayoutColMajor, 1, kRefClusterWide>(args, stream, false);
  if (args.n_frames * kRefCluster <= n_sm)
    return row ? launch_refine_as<kLayoutRowMajor, 1, kRefCluster>(args, stream, false) : launch_refine_as<kLayoutColMajor, 1, kRefCluster>(args, stream, false);
  return row ? launch_refine_as<kLayoutRowMajor, 2, kRefCluster>(args, stream, false) : launch_refine_as<kLayoutColMajor, 2, kRefCluster>(args, stream, false);
}

}  // namespace dpx
