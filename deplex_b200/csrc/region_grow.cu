// region_grow.cu -- stage 2 of the hot path: normals histogram, seeded region growing over the cell
// adjacency graph, plane merging and the per-cell label table.  One warp owns one frame; a batch runs
// one warp per frame concurrently (frames are independent, seeds within a frame are not).
//
// Replaces, per frame (reference file:line):
//   NormalsHistogram ctor / getPointsFromMostFrequentBin / removePoint   normals_histogram.cpp:21-72
//   PlaneExtractor::Impl::createPlaneSegments                            plane_extractor.cpp:297-347
//   PlaneExtractor::Impl::growSeed                                       plane_extractor.cpp:349-392
//   PlaneExtractor::Impl::getConnectedComponents / findMergedLabels      plane_extractor.cpp:394-453
//   the labels_map_ -> merged-label lookup of toImageLabels              plane_extractor.cpp:464-465
//
// Order-sensitive parts and how they are kept exact:
//   * first-max bin / first strict-min MSE seed: warp reductions on (value, index) pairs.
//   * FIFO BFS: the warp pops up to 8 queue entries per step, lanes = entry*4 + neighbour slot in the
//     reference's push order (up, down, left, right).  A cell reached by several lanes in one step is
//     claimed by the lowest lane (__match_any_sync), and winners are appended in lane order, which is
//     exactly the order the sequential queue would have produced.
//   * region moments: fp32 chains over the FIFO order, one lane per component (the seed is counted
//     twice, as in the reference).
// What is NOT order-sensitive is taken off the sequential path: whether a grown region is accepted
// (its plane fit, fp64) influences neither the histogram nor the unassigned mask, so all regions are
// grown first and then fitted 32 at a time, one lane per region; segment ids are assigned by an
// ordered compaction, which reproduces the reference's numbering.
//
// Working set per frame (bins, cell list, per-bin member runs with their MSE) lives in shared memory when
// it fits (always at 640x480 / patch 10: 36 KB), else in the global scratch tables.
#include "region_grow.cuh"

#include <cstdlib>
#include <cstring>

#include "normal_bins.cuh"
#include "plane_fit.cuh"
#include "seed_sort.cuh"

namespace dpx {
namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kNoSeed = 0x7fffffff;

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

__device__ __forceinline__ void load_moments(const float* rec, Moments& m) {
  m.n = __float_as_int(__ldcg(rec + kSegN));
#pragma unroll
  for (int i = 0; i < 3; ++i) m.s[i] = __ldcg(rec + kSegS + i);
#pragma unroll
  for (int i = 0; i < 6; ++i) m.v[i] = __ldcg(rec + kSegV + i);
}

__device__ __forceinline__ void load_seg(const float* rec, Moments& m, PlaneFit& f) {
  load_moments(rec, m);
#pragma unroll
  for (int i = 0; i < 3; ++i) f.mean[i] = __ldcg(rec + kSegMean + i);
#pragma unroll
  for (int i = 0; i < 3; ++i) f.normal[i] = __ldcg(rec + kSegNormal + i);
  f.d = __ldcg(rec + kSegD);
  f.mse = __ldcg(rec + kSegMse);
  f.score = __ldcg(rec + kSegScore);
}

__device__ __forceinline__ void store_seg(float* rec, const Moments& m, const PlaneFit& f) {
  __stcg(rec + kSegN, __int_as_float(m.n));
#pragma unroll
  for (int i = 0; i < 3; ++i) __stcg(rec + kSegS + i, m.s[i]);
#pragma unroll
  for (int i = 0; i < 6; ++i) __stcg(rec + kSegV + i, m.v[i]);
#pragma unroll
  for (int i = 0; i < 3; ++i) __stcg(rec + kSegMean + i, f.mean[i]);
#pragma unroll
  for (int i = 0; i < 3; ++i) __stcg(rec + kSegNormal + i, f.normal[i]);
  __stcg(rec + kSegD, f.d);
  __stcg(rec + kSegMse, f.mse);
  __stcg(rec + kSegScore, f.score);
}

// Directed edge tests of growSeed (plane_extractor.cpp:369-383) for every cell and neighbour slot.  The test
// between a frontier cell u and its neighbour v reads only per-cell constants (n_u, d_u, n_v, mean_v, tol_v),
// never BFS state, so all 4*C tests of a frame are evaluated up front, fully in parallel over the batch;
// the sequential BFS then only consults one byte per popped cell.  Bit s of edge[u] = "u may activate its
// neighbour in slot s" (slots in the reference's push order: up, down, left, right).
// A planar cell whose normal is next to an axis of the histogram (normal_bins.cuh bin_needs_care).  There the last ulp of
// acos / atan2 can decide the bin -- the azimuth of (6e-16, -0.89, -0.45) is pi - 7e-16, the upper edge of the last bin --
// so the bin is worked out again with the axis-aware functions; and when the x component is zero or rounding noise with
// y < 0, its sign decides between the first and the last azimuth bin and hangs on the last ulp of the solver's sin / cos:
// that cell is first fitted again with correctly rounded trigonometry (cr_math.cuh).
// The moments come back from rec_b; the point count is the whole cell.  Out of line and rare: axis-aligned planes of
// rendered or synthetic depth (167 of the ICL frame's 19 200 cells, none of the TUM frame's).
__device__ __noinline__ void repair_axis_cell(const Tables& tb, int patch, int bins_per_coord, long long cell, float4 na) {
  if (normal_on_wrap(na.x, na.y)) {
    const float4 r0 = tb.rec_b[3 * cell], r1 = tb.rec_b[3 * cell + 1], r2 = tb.rec_b[3 * cell + 2];
    Moments m;
    m.n = patch * patch;
    m.s[0] = r0.x; m.s[1] = r0.y; m.s[2] = r0.z;
    m.v[0] = r0.w; m.v[1] = r1.x; m.v[2] = r1.y; m.v[3] = r1.z; m.v[4] = r1.w; m.v[5] = r2.x;
    PlaneFit fit;
    fit_plane<true>(m, fit);
    na = make_float4(fit.normal[0], fit.normal[1], fit.normal[2], fit.d);
    tb.rec_a[2 * cell] = na;
  }
  const int b = histogram_bin<true>(na.x, na.y, na.z, bins_per_coord);
  if (b >= 0) tb.bin[cell] = static_cast<int16_t>(b);  // (always: the components that are not zero do not move)
}

// The cells edge_mask_kernel listed, a few threads for a short list (a frame of sensor data lists none: this kernel is one
// word read).  The bins are first read after it (seed sort, region growing).  A list that overflowed is replaced by a scan.
__global__ void __launch_bounds__(128) axis_repair_kernel(const RegionArgs args) {
  pdl_wait();
  const int count = args.tables.axis_work[0];
  if (count == 0) return;
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
  const Tables& tb = args.tables;
  if (count <= kAxisWorkCap) {
    for (int i = gtid; i < count; i += gsize) {
      const long long cell = tb.axis_work[1 + i];
      repair_axis_cell(tb, args.geom.patch, args.thr.histogram_bins_per_coord, cell, tb.rec_a[2 * cell]);
    }
  } else {
    const long long total = static_cast<long long>(args.n_frames) * args.geom.n_cells;
    for (long long cell = gtid; cell < total; cell += gsize) {
      if (tb.bin[cell] < 0) continue;
      const float4 na = tb.rec_a[2 * cell];
      if (bin_needs_care(na.x, na.y, na.z)) repair_axis_cell(tb, args.geom.patch, args.thr.histogram_bins_per_coord, cell, na);
    }
  }
}

__global__ void __launch_bounds__(256) edge_mask_kernel(const RegionArgs args) {
  pdl_wait();
  const Geometry& g = args.geom;
  const long long total = static_cast<long long>(args.n_frames) * g.n_cells;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % g.n_cells);
  const long long base = idx - c;  // first cell of this frame
  if (idx == 0) {
    // reset the fused-painting bookkeeping of this batch (region_grow_cta_kernel runs after this kernel)
    args.tables.paint_state[0] = 0;
    args.tables.paint_state[1] = 0;
  }
  const int r = c / g.nh, q = c - r * g.nh;
  unsigned mask = 0;
  if (args.tables.bin[idx] >= 0) {
    const float4 nu = __ldg(args.tables.rec_a + 2 * idx);
    // a normal next to an axis of the histogram: its bin is worked out again by axis_repair_kernel (a call in here costs this
    // kernel half its occupancy and a stack frame; an entry in a list costs nothing)
    if (bin_needs_care(nu.x, nu.y, nu.z)) {
      const int pos = atomicAdd(args.tables.axis_work, 1);
      if (pos < kAxisWorkCap) args.tables.axis_work[1 + pos] = static_cast<int>(idx);
    }
    const double min_cos = static_cast<double>(args.thr.min_cos_angle_merge);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      int v = -1;
      if (s == 0) v = (r >= 1) ? c - g.nh : -1;
      else if (s == 1) v = (r + 1 < g.nv) ? c + g.nh : -1;
      else if (s == 2) v = (q >= 1) ? c - 1 : -1;
      else v = (q + 1 < g.nh) ? c + 1 : -1;
      if (v < 0 || args.tables.bin[base + v] < 0) continue;
      const float4 nv = __ldg(args.tables.rec_a + 2 * (base + v));
      const float4 mv = __ldg(args.tables.rec_a + 2 * (base + v) + 1);
      const double cos_angle = static_cast<double>(dot3(nu.x, nu.y, nu.z, nv.x, nv.y, nv.z));
      const double t = __dadd_rn(static_cast<double>(dot3(nu.x, nu.y, nu.z, mv.x, mv.y, mv.z)), static_cast<double>(nu.w));
      const double merge_dist = __dmul_rn(t, t);
      if (cos_angle >= min_cos && merge_dist <= static_cast<double>(mv.w)) mask |= 1u << s;
    }
  }
  args.tables.edge[idx] = static_cast<uint8_t>(mask);
}

}  // namespace
}  // namespace dpx

#include "region_grow_cta.cuh"

namespace dpx {
namespace {

template <bool SMEM>
__global__ void __launch_bounds__(32) region_grow_kernel(const RegionArgs args) {
  extern __shared__ float4 smem_f4[];
  const Geometry& g = args.geom;
  const Thresholds& th = args.thr;
  const int lane = threadIdx.x;
  const int frame = blockIdx.x;
  if (blockIdx.x == 0 && threadIdx.x == 0) args.tables.axis_work[0] = 0;  // the list of this batch has been worked off
  const int C = g.n_cells, nh = g.nh, nv = g.nv;
  const int B2 = th.histogram_bins_per_coord * th.histogram_bins_per_coord;
  const long long fc = static_cast<long long>(frame) * C;

  // ---- shared-memory carve-up (offsets computed by region_grow_plan on the host) ------------------
  char* smem = reinterpret_cast<char*>(smem_f4);
  float* stage = reinterpret_cast<float*>(smem);                  // [32][12] staging for the accumulation
  int* hist = reinterpret_cast<int*>(smem + args.plan.off_hist);
  int* bin_off = reinterpret_cast<int*>(smem + args.plan.off_binoff);   // [B2 + 1] start of each bin's member run
  int* run_end = reinterpret_cast<int*>(smem + args.plan.off_cursor);   // [B2] end of the still-unassigned members
  unsigned* rowbits = reinterpret_cast<unsigned*>(smem + args.plan.off_rowbits);
  // bins: working bin per cell during growing (-1 = assigned / not planar); afterwards reused for the segment labels
  // SMEM = the whole working set is in shared memory: pointers are then provably shared and the compiler
  // emits LDS/STS instead of generic accesses
  int16_t* bins = (SMEM || args.plan.bins_smem) ? reinterpret_cast<int16_t*>(smem + args.plan.off_bins) : args.tables.bin_work + fc;
  uint8_t* edge = (SMEM || args.plan.bins_smem) ? reinterpret_cast<uint8_t*>(smem + args.plan.off_edge) : args.tables.edge + fc;
  int32_t* list = (SMEM || args.plan.list_smem) ? reinterpret_cast<int32_t*>(smem + args.plan.off_list) : args.tables.queue + fc;
  // members of each bin (cell ids, grouped by initial bin) and their MSE, in the same order
  int32_t* members = (SMEM || args.plan.members_smem) ? reinterpret_cast<int32_t*>(smem + args.plan.off_members)
                                            : reinterpret_cast<int32_t*>(args.tables.pairs + 2 * fc);
  float* msem = (SMEM || args.plan.members_smem) ? reinterpret_cast<float*>(smem + args.plan.off_msem)
                                       : reinterpret_cast<float*>(args.tables.pairs + 2 * fc + C);
  int32_t* merge = (SMEM || args.plan.merge_smem) ? reinterpret_cast<int32_t*>(smem + args.plan.off_merge)
                                        : args.tables.merge + static_cast<long long>(frame) * g.plane_cap;

  const float4* rec_b4 = args.tables.rec_b + 3 * fc;
  const int16_t* bin_in = args.tables.bin + fc;
  const uint8_t* edge_in = args.tables.edge + fc;
  const float* mse_g = args.tables.mse + fc;
  int32_t* seg_label = args.tables.seg_label + fc;
  int32_t* cell_label = args.tables.cell_label + fc;
  float* segs = args.tables.segs + static_cast<long long>(frame) * g.plane_cap * kSegFloats;
  int32_t* merge_out = args.tables.merge + static_cast<long long>(frame) * g.plane_cap;

  const bool prof = args.prof != nullptr;
  const long long t_kernel0 = prof ? clock64() : 0;

  // ---- histogram of planar-cell bins (normals_histogram.cpp:21-49; bins come from stage 1) --------
  for (int i = lane; i < B2; i += 32) hist[i] = 0;
  __syncwarp();
  int remaining = 0;
  float* mse_tmp = reinterpret_cast<float*>(list);  // per-cell MSE parked in the (still unused) cell list
  for (int c0 = 0; c0 < C; c0 += 256) {
    int b[8];
    float m[8];
    unsigned ed[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = c0 + 32 * k + lane;
      const bool in = c < C;
      b[k] = in ? static_cast<int>(__ldg(bin_in + c)) : -2;
      m[k] = in ? __ldg(mse_g + c) : 0.f;
      ed[k] = in ? static_cast<unsigned>(__ldg(edge_in + c)) : 0u;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = c0 + 32 * k + lane;
      if (b[k] == -2) continue;
      bins[c] = static_cast<int16_t>(b[k]);
      if (SMEM || args.plan.bins_smem) edge[c] = static_cast<uint8_t>(ed[k]);
      mse_tmp[c] = m[k];
      seg_label[c] = 0;
      if (b[k] >= 0) {
        atomicAdd(&hist[b[k]], 1);
        ++remaining;
      }
    }
  }
  remaining = warp_sum(remaining);
  __syncwarp();

  // group the planar cells by bin (exclusive scan of the histogram, then an unordered fill: the seed
  // search below breaks ties by cell id itself), so that a seed search touches only its bin's cells
  {
    const int per = (B2 + 31) / 32;
    int local = 0;
    for (int i = 0; i < per; ++i) {
      const int b = lane * per + i;
      if (b < B2) local += hist[b];
    }
    int incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += t;
    }
    int run = incl - local;
    for (int i = 0; i < per; ++i) {
      const int b = lane * per + i;
      if (b < B2) {
        bin_off[b] = run;
        run_end[b] = run;
        run += hist[b];
      }
    }
    if (lane == 31) bin_off[B2] = incl;
  }
  __syncwarp();
  for (int c = lane; c < C; c += 32) {
    const int b = bins[c];
    if (b >= 0) {
      const int pos = atomicAdd(&run_end[b], 1);
      members[pos] = c;
      msem[pos] = mse_tmp[c];
    }
  }
  __syncwarp();

  const double min_cos = static_cast<double>(th.min_cos_angle_merge);
  const int cell_pts = g.patch * g.patch;
  int n_regions = 0;  // grown regions with enough cells, in seed order
  int list_off = 0;   // their cells are stored back to back in `list`
  const int slot = lane & 3;
  const int delta = (slot == 0) ? -nh : (slot == 1) ? nh : (slot == 2) ? -1 : 1;

  long long t_seed = 0, t_bfs = 0, t_acc = 0, t_mark = 0;
  int n_seeds = 0, n_steps = 0;
  const long long t_init = prof ? clock64() - t_kernel0 : 0;

  // ---- createPlaneSegments, sequential part (plane_extractor.cpp:302-331) -------------------------
  while (remaining > 0) {
    if (prof) t_mark = clock64();
    // most frequent bin, first maximum (normals_histogram.cpp:54-56)
    // key = count << 15 | (0x7fff - bin): the maximum key is the largest count with the smallest bin id
    // (count <= n_cells < 2^17 and bin < 2^15 are enforced at create)
    unsigned key = 0;
#pragma unroll 4
    for (int i = lane; i < B2; i += 32) key = max(key, (static_cast<unsigned>(hist[i]) << 15) | (0x7fffu - i));
    key = __reduce_max_sync(kFull, key);
    const int bc = static_cast<int>(key >> 15), bi = static_cast<int>(0x7fffu - (key & 0x7fffu));
    const unsigned long long n_cand = bc > 0 ? static_cast<unsigned long long>(bc) : 0ull;
    if (n_cand < th.min_candidate_size) break;  // plane_extractor.cpp:305-307

    // seed = first strict minimum of the MSE among the bin's cells (plane_extractor.cpp:309-316);
    // float -> double is monotonic, so the comparison against (double)INT_MAX is done once at the end.
    // The scan also compacts the bin's member run down to the cells that are still unassigned.
    float lm = __int_as_float(0x7f800000);  // +inf
    int seed = kNoSeed;
    {
      const int start = bin_off[bi], end = run_end[bi];
      int w = start;
      for (int i0 = start; i0 < end; i0 += 32) {
        const int i = i0 + lane;
        const bool in = i < end;
        const int c = in ? members[i] : 0;
        const float m = in ? msem[i] : 0.f;
        const bool alive = in && bins[c] == bi;
        if (alive && (m < lm || (m == lm && c < seed))) { lm = m; seed = c; }
        const unsigned am = __ballot_sync(kFull, alive);
        if (alive) {
          const int pos = w + __popc(am & ((1u << lane) - 1u));
          members[pos] = c;
          msem[pos] = m;
        }
        w += __popc(am);
        __syncwarp();
      }
      if (lane == 0) run_end[bi] = w;
    }
    {
      // lexicographic minimum of (mse, cell id) over the lanes: two integer warp reductions on the
      // order-preserving integer image of the float
      const unsigned fb = __float_as_uint(lm);
      const unsigned ord = fb ^ ((fb >> 31) ? 0xffffffffu : 0x80000000u);
      const unsigned best = __reduce_min_sync(kFull, seed == kNoSeed ? 0xffffffffu : ord);
      const unsigned cand = (seed != kNoSeed && ord == best) ? static_cast<unsigned>(seed) : static_cast<unsigned>(kNoSeed);
      seed = static_cast<int>(__reduce_min_sync(kFull, cand));
      const unsigned bb = best ^ ((best >> 31) ? 0x80000000u : 0xffffffffu);
      lm = __uint_as_float(bb);
    }
    // no candidate with mse < INT_MAX: the reference reads an uninitialised seed id here
    if (seed == kNoSeed || !(static_cast<double>(lm) < 2147483647.0)) break;

    if (prof) { const long long t = clock64(); t_seed += t - t_mark; t_mark = t; ++n_seeds; }
    // growSeed (plane_extractor.cpp:349-392): batched FIFO BFS into list[list_off ...).  Edge tests were
    // precomputed (edge_mask_kernel); a neighbour is taken if its edge bit is set and it is still unassigned.
    // Queue entries carry the cell id in the low 24 bits and the cell's edge mask in the high 8, so a pop
    // costs one shared-memory read; a lane's neighbour is v = u + delta(slot) whenever its edge bit is set.
    int32_t* q = list + list_off;
    if (lane == 0) {
      q[0] = seed | (static_cast<int>(edge[seed]) << 24);
      bins[seed] = -1;
      hist[bi] -= 1;
    }
    __syncwarp();
    int head = 0, tail = 1;
    while (head < tail) {
      const int nb = min(8, tail - head);
      int v = 0, vb = -1;
      unsigned ev = 0;
      if ((lane >> 2) < nb) {
        const unsigned pk = static_cast<unsigned>(q[head + (lane >> 2)]);
        if ((pk >> (24 + slot)) & 1u) {
          v = static_cast<int>(pk & 0xffffffu) + delta;
          vb = bins[v];
          ev = edge[v];
        }
      }
      const bool pass = vb >= 0;  // edge test passed, still unassigned, not yet activated
      const unsigned pm = __ballot_sync(kFull, pass);
      unsigned wm = 0;
      if (pm) {
        // a cell reached by several lanes goes to the lowest one (= the earliest in FIFO order); conflicts
        // need at least two passing lanes from different queue entries
        bool win = pass;
        // (lanes of one queue entry have distinct targets: only passing lanes of different entries can clash)
        const unsigned same_entry = 0xfu << ((__ffs(pm) - 1) & ~3);  // the 4 lanes of the lowest passing entry
        if (pass && (pm & ~same_entry)) {
          const unsigned grp = __match_any_sync(pm, v);
          win = (__ffs(grp) - 1) == lane;
        }
        wm = __ballot_sync(kFull, win);
        if (win) {
          q[tail + __popc(wm & ((1u << lane) - 1u))] = v | static_cast<int>(ev << 24);
          bins[v] = -1;             // removePoint + unassigned_mask[v] = false (plane_extractor.cpp:324-325)
          atomicSub(&hist[vb], 1);
        }
      }
      tail += __popc(wm);
      head += nb;
      ++n_steps;
      __syncwarp();
    }
    // strip the edge masks: the list keeps plain cell ids for the later phases
    for (int i = lane; i < tail; i += 32) q[i] &= 0xffffff;
    __syncwarp();
    remaining -= tail;
    if (prof) { const long long t = clock64(); t_bfs += t - t_mark; t_mark = t; }
    if (static_cast<unsigned long long>(tail) < th.min_cells_activated) continue;  // plane_extractor.cpp:329-331

    // merge the activated cells into the candidate, seed first and twice (plane_extractor.cpp:318-323):
    // 32 cell records per round trip, staged through shared memory, then 9 sequential fp32 chains
    float acc = 0.f;
    if (lane < 9) acc = reinterpret_cast<const float*>(rec_b4)[12 * seed + lane];
    // records are fetched two chunks (64 cells) ahead of the additions
    float4 ra[3], rb[3];
    auto fetch = [&](int i0, float4 (&r)[3]) {
      if (i0 < tail) {
        const int c = q[min(i0 + lane, tail - 1)];
        r[0] = __ldg(rec_b4 + 3 * c); r[1] = __ldg(rec_b4 + 3 * c + 1); r[2] = __ldg(rec_b4 + 3 * c + 2);
      }
    };
    fetch(0, ra);
    fetch(32, rb);
    for (int i0 = 0; i0 < tail; i0 += 32) {
      const int cnt = min(32, tail - i0);
      float4* st4 = reinterpret_cast<float4*>(stage + 12 * lane);
      st4[0] = ra[0]; st4[1] = ra[1]; st4[2] = ra[2];
      ra[0] = rb[0]; ra[1] = rb[1]; ra[2] = rb[2];
      fetch(i0 + 64, rb);
      __syncwarp();
      if (lane < 9) {
        if (cnt == 32) {
#pragma unroll
          for (int k = 0; k < 32; ++k) acc = __fadd_rn(acc, stage[12 * k + lane]);
        } else {
          for (int k = 0; k < cnt; ++k) acc = __fadd_rn(acc, stage[12 * k + lane]);
        }
      }
      __syncwarp();
    }
    if (prof) t_acc += clock64() - t_mark;
    if (n_regions < g.plane_cap) {
      float* rec = segs + static_cast<long long>(n_regions) * kSegFloats;
      if (lane < 9) __stcg(rec + kSegS + lane, acc);
      if (lane == 9) __stcg(rec + kSegN, __int_as_float(cell_pts * (tail + 1)));
      if (lane == 10) __stcg(rec + kSegOff, __int_as_float(list_off));
      if (lane == 11) __stcg(rec + kSegCnt, __int_as_float(tail));
      ++n_regions;
      list_off += tail;
    }
    __syncwarp();
  }
  __syncwarp();

  // growing is over: the bins array now becomes the per-cell segment label (labels_map_), all zero
  uint16_t* segl = reinterpret_cast<uint16_t*>(bins);
  for (int c = lane; c < C; c += 32) segl[c] = 0;
  __syncwarp();

  const long long t_grow_end = prof ? clock64() : 0;
  // ---- plane fit of every grown region, one lane per region (plane_extractor.cpp:333-343) ---------
  int nseg = 0;
  for (int base = 0; base < n_regions; base += 32) {
    const int r = base + lane;
    Moments mom;
    PlaneFit fit;
    int off = 0, cnt = 0;
    bool accept = false;
    if (r < n_regions) {
      const float* rec = segs + static_cast<long long>(r) * kSegFloats;
      load_moments(rec, mom);
      off = __float_as_int(__ldcg(rec + kSegOff));
      cnt = __float_as_int(__ldcg(rec + kSegCnt));
      fit_plane(mom, fit);
      accept = fit.score > th.min_region_planarity_score;  // strict (plane_extractor.cpp:336)
    }
    __syncwarp();
    const unsigned am = __ballot_sync(kFull, accept);
    const int id = nseg + __popc(am & ((1u << lane) - 1u));  // 0-based segment index, in seed order
    if (accept) store_seg(segs + static_cast<long long>(id) * kSegFloats, mom, fit);
    // paint labels_map_ (plane_extractor.cpp:339-342)
    unsigned todo = am;
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const int o = __shfl_sync(kFull, off, src), n = __shfl_sync(kFull, cnt, src);
      const int label = __shfl_sync(kFull, id, src) + 1;
      for (int i = lane; i < n; i += 32) {
        const int c = list[o + i];
        segl[c] = static_cast<uint16_t>(label);
        seg_label[c] = label;
      }
    }
    nseg += __popc(am);
    __syncwarp();
  }
  if (lane == 0) args.tables.n_planes[frame] = nseg;
  const long long t_fit_end = prof ? clock64() : 0;
  __syncwarp();

  // ---- getConnectedComponents (plane_extractor.cpp:430-453): boundary pairs, last row/column skipped ----
  // (the pair list reuses the member runs' storage: 2*C words, free once growing is over)
  uint32_t* pairs = reinterpret_cast<uint32_t*>(members);
  int n_pairs = 0;
  if (nseg > 1) {
    const int limit = (nv - 1) * nh;
    for (int c0 = 0; c0 < limit; c0 += 32) {
      const int c = c0 + lane;
      unsigned p0 = 0xffffffffu, p1 = 0xffffffffu;
      if (c < limit) {
        const int qc = c % nh;
        const int id = segl[c];
        if (qc < nh - 1 && id > 0) {
          const int right = segl[c + 1], down = segl[c + nh];
          if (right > 0 && right != id) p0 = (static_cast<unsigned>(min(id, right) - 1) << 16) | static_cast<unsigned>(max(id, right) - 1);
          if (down > 0 && down != id) p1 = (static_cast<unsigned>(min(id, down) - 1) << 16) | static_cast<unsigned>(max(id, down) - 1);
        }
      }
      const unsigned m0 = __ballot_sync(kFull, p0 != 0xffffffffu);
      if (p0 != 0xffffffffu) pairs[n_pairs + __popc(m0 & ((1u << lane) - 1u))] = p0;
      n_pairs += __popc(m0);
      const unsigned m1 = __ballot_sync(kFull, p1 != 0xffffffffu);
      if (p1 != 0xffffffffu) pairs[n_pairs + __popc(m1 & ((1u << lane) - 1u))] = p1;
      n_pairs += __popc(m1);
    }
  }
  for (int i = lane; i < nseg; i += 32) merge[i] = i;
  __syncwarp();

  // ---- findMergedLabels (plane_extractor.cpp:402-423) ---------------------------------------------
  const int words = (nseg + 31) / 32;
  for (int r = 0; r < nseg && n_pairs > 0; ++r) {
    for (int i = lane; i < words; i += 32) rowbits[i] = 0;
    __syncwarp();
    bool any = false;
    for (int i = lane; i < n_pairs; i += 32) {
      const unsigned pr = pairs[i];
      if (static_cast<int>(pr >> 16) == r) {
        const unsigned t = pr & 0xffffu;
        atomicOr(&rowbits[t >> 5], 1u << (t & 31));
        any = true;
      }
    }
    any = __any_sync(kFull, any);
    __syncwarp();
    if (!any) continue;

    const int a = merge[r];
    Moments ma;
    PlaneFit fa;
    load_seg(segs + static_cast<long long>(a) * kSegFloats, ma, fa);
    bool expanded = false;
    for (int w = 0; w < words; ++w) {
      unsigned bits = rowbits[w];
      while (bits) {
        const int t = w * 32 + __ffs(bits) - 1;
        bits &= bits - 1;
        Moments mt;
        PlaneFit ft;
        load_seg(segs + static_cast<long long>(t) * kSegFloats, mt, ft);
        // normal/d of `a` are the ones it had when the row started (stats are refit after the row)
        const double cos_angle = static_cast<double>(dot3(fa.normal[0], fa.normal[1], fa.normal[2], ft.normal[0], ft.normal[1], ft.normal[2]));
        const float df = __fadd_rn(dot3(fa.normal[0], fa.normal[1], fa.normal[2], ft.mean[0], ft.mean[1], ft.mean[2]), fa.d);
        const double distance = __dmul_rn(static_cast<double>(df), static_cast<double>(df));
        if (cos_angle > min_cos && distance < static_cast<double>(th.max_merge_dist)) {
          ma.n += mt.n;
#pragma unroll
          for (int i = 0; i < 3; ++i) ma.s[i] = __fadd_rn(ma.s[i], mt.s[i]);
#pragma unroll
          for (int i = 0; i < 6; ++i) ma.v[i] = __fadd_rn(ma.v[i], mt.v[i]);
          if (lane == 0) merge[t] = a;
          expanded = true;
        }
      }
    }
    if (expanded) {
      fit_plane(ma, fa);
      if (lane == 0) store_seg(segs + static_cast<long long>(a) * kSegFloats, ma, fa);
    }
    __syncwarp();
  }
  __syncwarp();

  const long long t_merge_end = prof ? clock64() : 0;
  // ---- per-cell final labels (plane_extractor.cpp:464-465) -----------------------------------------
  for (int c = lane; c < C; c += 32) {
    const int l = segl[c];
    cell_label[c] = (l == 0) ? 0 : merge[l - 1] + 1;
  }
  if (SMEM || args.plan.merge_smem)
    for (int i = lane; i < nseg; i += 32) merge_out[i] = merge[i];
  if (prof && lane == 0) {
    long long* o = args.prof + static_cast<long long>(frame) * kRegionProfSlots;
    const long long t_end = clock64();
    o[0] = t_end - t_kernel0;          // whole frame
    o[1] = t_init;                     // histogram + grouping + prefetch
    o[2] = t_seed;                     // bin argmax + seed search (all seeds)
    o[3] = t_bfs;                      // BFS (all seeds)
    o[4] = t_acc;                      // moment accumulation (regions with enough cells)
    o[5] = t_fit_end - t_grow_end;     // plane fits + label painting
    o[6] = t_merge_end - t_fit_end;    // adjacency pairs + merging
    o[7] = t_end - t_merge_end;        // final labels
    o[8] = n_seeds;
    o[9] = n_steps;
    o[10] = n_regions;
    o[11] = nseg;
  }
}

}  // namespace

RegionPlan region_grow_plan(const Geometry& g, const Thresholds& th) {
  RegionPlan p{};
  const size_t budget = 160 * 1024;
  auto align16 = [](size_t v) { return (v + 15) & ~static_cast<size_t>(15); };
  const size_t B2 = static_cast<size_t>(th.histogram_bins_per_coord) * th.histogram_bins_per_coord;
  size_t off = 32 * 12 * sizeof(float);  // stage
  p.off_hist = static_cast<int>(off);
  off = align16(off + B2 * 4);
  p.off_binoff = static_cast<int>(off);
  off = align16(off + (B2 + 1) * 4);
  p.off_cursor = static_cast<int>(off);
  off = align16(off + B2 * 4);
  p.off_rowbits = static_cast<int>(off);
  off = align16(off + static_cast<size_t>((g.plane_cap + 31) / 32) * 4);
  const size_t C = static_cast<size_t>(g.n_cells);
  if (off + C * 3 <= budget) {  // working bins (int16) + edge masks (uint8)
    p.bins_smem = 1;
    p.off_bins = static_cast<int>(off);
    off = align16(off + C * 2);
    p.off_edge = static_cast<int>(off);
    off = align16(off + C);
  }
  if (off + C * 4 <= budget) {
    p.list_smem = 1;
    p.off_list = static_cast<int>(off);
    off = align16(off + C * 4);
  }
  if (off + C * 8 <= budget) {
    p.members_smem = 1;
    p.off_members = static_cast<int>(off);
    off = align16(off + C * 4);
    p.off_msem = static_cast<int>(off);
    off = align16(off + C * 4);
  }
  if (off + static_cast<size_t>(g.plane_cap) * 4 <= budget) {
    p.merge_smem = 1;
    p.off_merge = static_cast<int>(off);
    off = align16(off + static_cast<size_t>(g.plane_cap) * 4);
  }
  p.bytes = off;
  return p;
}

namespace {
bool force_warp_kernel_env() {
  static const bool v = std::getenv("DPX_REGION_KERNEL") && std::strcmp(std::getenv("DPX_REGION_KERNEL"), "warp") == 0;
  return v;
}
}  // namespace

namespace {
// 0 .. 3 = storage mode of region_grow_cta_kernel, -1 = the generic single-warp kernels
int cta_mode(const Geometry& g, const Thresholds& th, CtaPlan* plan) {
  if (g.n_cells == 0 || force_warp_kernel_env()) return -1;
  for (int mode = 0; mode <= 3; ++mode) {
    const CtaPlan p = region_grow_cta_plan(g, th, mode);
    if (p.bytes > 0) {
      if (plan) *plan = p;
      return mode;
    }
  }
  return -1;
}
template <int MODE>
cudaError_t launch_cta(const RegionArgs& args, const CtaPlan& plan, cudaStream_t stream) {
  cudaFuncSetAttribute(region_grow_cta_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(plan.bytes));
  return launch_dependent(region_grow_cta_kernel<MODE>, dim3(args.n_frames), dim3(kCtaThreads), plan.bytes, stream, args, plan);
}
}  // namespace

bool region_grow_uses_cta(const Geometry& g, const Thresholds& th) { return cta_mode(g, th, nullptr) >= 0; }
int region_grow_mode(const Geometry& g, const Thresholds& th) { return cta_mode(g, th, nullptr); }

cudaError_t launch_region_grow(const RegionArgs& args, cudaStream_t stream, bool* painted) {
  if (painted) *painted = false;
  if (args.n_frames == 0) return cudaSuccess;
  const long long cells = static_cast<long long>(args.n_frames) * args.geom.n_cells;
  cudaError_t e = launch_dependent(edge_mask_kernel, dim3(static_cast<unsigned>((cells + 255) / 256)), dim3(256), 0, stream, args);
  if (e != cudaSuccess) return e;
  e = launch_dependent(axis_repair_kernel, dim3(16), dim3(128), 0, stream, args);
  if (e != cudaSuccess) return e;
  CtaPlan cta{};
  const int mode = cta_mode(args.geom, args.thr, &cta);
  if (mode >= 0) {
    if (painted) *painted = args.labels != nullptr;
    // frames beyond mode 0: the seed order, cells sorted by (bin, MSE, cell id); `pairs` is its scratch
    if (mode >= 1) e = launch_seed_sort(args.tables.bin, args.tables.mse, args.tables.skeys, reinterpret_cast<unsigned long long*>(args.tables.pairs),
                         args.n_frames, args.geom.n_cells, stream);
    if (e != cudaSuccess) return e;
    return mode == 0   ? launch_cta<0>(args, cta, stream)
           : mode == 1 ? launch_cta<1>(args, cta, stream)
           : mode == 2 ? launch_cta<2>(args, cta, stream)
                       : launch_cta<3>(args, cta, stream);
  }
  const bool all_smem = args.plan.bins_smem && args.plan.list_smem && args.plan.members_smem && args.plan.merge_smem;
  if (all_smem) {
    cudaFuncSetAttribute(region_grow_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(args.plan.bytes));
    region_grow_kernel<true><<<args.n_frames, 32, args.plan.bytes, stream>>>(args);
  } else {
    cudaFuncSetAttribute(region_grow_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(args.plan.bytes));
    region_grow_kernel<false><<<args.n_frames, 32, args.plan.bytes, stream>>>(args);
  }
  return cudaGetLastError();
}

}  // namespace dpx

#ifdef DPX_BFS_PROBE
// probe builds only (make NVFLAGS+=-DDPX_BFS_PROBE): cycles of lane 0 in the phases of bfs_wide_step
extern "C" __attribute__((visibility("default"))) int dpx_debug_wide_probe(long long* out, int reset) {
  cudaDeviceSynchronize();
  if (out) cudaMemcpyFromSymbol(out, dpx::g_wide_probe, sizeof(long long) * 8);
  if (reset) {
    long long zero[8] = {};
    cudaMemcpyToSymbol(dpx::g_wide_probe, zero, sizeof(zero));
  }
  return 0;
}
#endif
