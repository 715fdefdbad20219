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

#include "plane_fit.cuh"

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

__global__ void __launch_bounds__(32) region_grow_kernel(const RegionArgs args) {
  extern __shared__ float4 smem_f4[];
  const Geometry& g = args.geom;
  const Thresholds& th = args.thr;
  const int lane = threadIdx.x;
  const int frame = blockIdx.x;
  const int C = g.n_cells, nh = g.nh, nv = g.nv;
  const int B2 = th.histogram_bins_per_coord * th.histogram_bins_per_coord;

  // ---- shared-memory carve-up (offsets computed by region_grow_plan on the host) ------------------
  char* smem = reinterpret_cast<char*>(smem_f4);
  float* stage = reinterpret_cast<float*>(smem);                  // [32][12] staging for the accumulation
  int* hist = reinterpret_cast<int*>(smem + args.plan.off_hist);
  unsigned* rowbits = reinterpret_cast<unsigned*>(smem + args.plan.off_rowbits);
  int16_t* bins = args.plan.bins_smem ? reinterpret_cast<int16_t*>(smem + args.plan.off_bins)
                                      : args.tables.bin_work + static_cast<long long>(frame) * C;
  int32_t* list = args.plan.list_smem ? reinterpret_cast<int32_t*>(smem + args.plan.off_list)
                                      : args.tables.queue + static_cast<long long>(frame) * C;
  const float* mse_g = args.tables.mse + static_cast<long long>(frame) * C;
  int* bin_off = reinterpret_cast<int*>(smem + args.plan.off_binoff);   // [B2 + 1] start of each bin's member run
  int* cursor = reinterpret_cast<int*>(smem + args.plan.off_cursor);    // [B2] fill cursors
  // members of each bin (cell ids, grouped by initial bin) and their MSE, in the same order
  int32_t* members = args.plan.members_smem
                         ? reinterpret_cast<int32_t*>(smem + args.plan.off_members)
                         : reinterpret_cast<int32_t*>(args.tables.pairs + 2LL * frame * C);
  float* msem = args.plan.members_smem ? reinterpret_cast<float*>(smem + args.plan.off_msem)
                                       : reinterpret_cast<float*>(args.tables.pairs + 2LL * frame * C + C);

  const float4* rec_a = args.tables.rec_a + 2LL * frame * C;
  const float4* rec_b4 = args.tables.rec_b + 3LL * frame * C;
  const int16_t* bin_in = args.tables.bin + static_cast<long long>(frame) * C;
  int32_t* seg_label = args.tables.seg_label + static_cast<long long>(frame) * C;
  int32_t* cell_label = args.tables.cell_label + static_cast<long long>(frame) * C;
  uint32_t* pairs = args.tables.pairs + 2LL * frame * C;
  float* segs = args.tables.segs + static_cast<long long>(frame) * g.plane_cap * kSegFloats;
  int32_t* merge = args.tables.merge + static_cast<long long>(frame) * g.plane_cap;

  // ---- histogram of planar-cell bins (normals_histogram.cpp:21-49; bins come from stage 1) --------
  for (int i = lane; i < B2; i += 32) hist[i] = 0;
  __syncwarp();
  int remaining = 0;
  for (int c0 = 0; c0 < C; c0 += 128) {
    int b[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = c0 + 32 * k + lane;
      b[k] = (c < C) ? static_cast<int>(__ldg(bin_in + c)) : -2;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = c0 + 32 * k + lane;
      if (b[k] == -2) continue;
      bins[c] = static_cast<int16_t>(b[k]);
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
        cursor[b] = run;
        run += hist[b];
      }
    }
    if (lane == 31) bin_off[B2] = incl;
  }
  __syncwarp();
  for (int c0 = 0; c0 < C; c0 += 128) {
    float m[4];
    int b[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = c0 + 32 * k + lane;
      b[k] = (c < C) ? static_cast<int>(bins[c]) : -1;
      m[k] = (b[k] >= 0) ? __ldg(mse_g + c) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (b[k] < 0) continue;
      const int pos = atomicAdd(&cursor[b[k]], 1);
      members[pos] = c0 + 32 * k + lane;
      msem[pos] = m[k];
    }
  }
  // warm L1 with this frame's BFS records (32 B per cell, read-only in this kernel)
  {
    const char* base = reinterpret_cast<const char*>(rec_a);
    const int lines = (C * 32 + 127) / 128;
    for (int i = lane; i < lines; i += 32) asm volatile("prefetch.global.L1 [%0];" ::"l"(base + 128LL * i));
  }
  __syncwarp();

  const double min_cos = static_cast<double>(th.min_cos_angle_merge);
  // exact u / nh for u * nh < 2^32 (u < n_cells)
  const unsigned nh_magic = static_cast<unsigned>((0x100000000ull + nh - 1) / nh);
  const int cell_pts = g.patch * g.patch;
  int n_regions = 0;  // grown regions with enough cells, in seed order
  int list_off = 0;   // their cells are stored back to back in `list`

  // ---- createPlaneSegments, sequential part (plane_extractor.cpp:302-331) -------------------------
  while (remaining > 0) {
    // most frequent bin, first maximum (normals_histogram.cpp:54-56)
    int bc = -1, bi = kNoSeed;
    for (int i = lane; i < B2; i += 32) {
      const int h = hist[i];
      if (h > bc) { bc = h; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const int oc = __shfl_xor_sync(kFull, bc, o), oi = __shfl_xor_sync(kFull, bi, o);
      if (oc > bc || (oc == bc && oi < bi)) { bc = oc; bi = oi; }
    }
    const unsigned long long n_cand = bc > 0 ? static_cast<unsigned long long>(bc) : 0ull;
    if (n_cand < th.min_candidate_size) break;  // plane_extractor.cpp:305-307

    // seed = first strict minimum of the MSE among the bin's cells (plane_extractor.cpp:309-316);
    // float -> double is monotonic, so the comparison against (double)INT_MAX is done once at the end
    float lm = __int_as_float(0x7f800000);  // +inf
    int seed = kNoSeed;
    {
      const int end = bin_off[bi + 1];
      for (int i = bin_off[bi] + lane; i < end; i += 32) {
        const int c = members[i];
        if (bins[c] == bi) {  // still unassigned
          const float m = msem[i];
          if (m < lm || (m == lm && c < seed)) { lm = m; seed = c; }
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(kFull, lm, o);
      const int os = __shfl_xor_sync(kFull, seed, o);
      if (os != kNoSeed && (seed == kNoSeed || om < lm || (om == lm && os < seed))) { lm = om; seed = os; }
    }
    // no candidate with mse < INT_MAX: the reference reads an uninitialised seed id here
    if (seed == kNoSeed || !(static_cast<double>(lm) < 2147483647.0)) break;

    // growSeed (plane_extractor.cpp:349-392): batched FIFO BFS into list[list_off ...)
    int32_t* q = list + list_off;
    if (lane == 0) {
      q[0] = seed;
      bins[seed] = -1;
      hist[bi] -= 1;
    }
    __syncwarp();
    int head = 0, tail = 1;
    while (head < tail) {
      const int nb = min(8, tail - head);
      const int e = lane >> 2, s = lane & 3;
      int v = -1, u = -1;
      if (e < nb) {
        u = q[head + e];
        const int r = static_cast<int>(__umulhi(static_cast<unsigned>(u), nh_magic)), qc = u - r * nh;
        if (s == 0) v = (r >= 1) ? u - nh : -1;
        else if (s == 1) v = (r + 1 < nv) ? u + nh : -1;
        else if (s == 2) v = (qc >= 1) ? u - 1 : -1;
        else v = (qc + 1 < nh) ? u + 1 : -1;
      }
      bool pass = false;
      int vb = -1;
      if (v >= 0) {
        vb = bins[v];
        if (vb >= 0) {  // unassigned and not yet activated
          const float4 nu = __ldg(rec_a + 2 * u);
          const float4 nvv = __ldg(rec_a + 2 * v);
          const float4 mv = __ldg(rec_a + 2 * v + 1);
          const double cos_angle = static_cast<double>(dot3(nu.x, nu.y, nu.z, nvv.x, nvv.y, nvv.z));
          const double t = __dadd_rn(static_cast<double>(dot3(nu.x, nu.y, nu.z, mv.x, mv.y, mv.z)),
                                     static_cast<double>(nu.w));
          const double merge_dist = __dmul_rn(t, t);
          pass = cos_angle >= min_cos && merge_dist <= static_cast<double>(mv.w);
        }
      }
      const unsigned pm = __ballot_sync(kFull, pass);
      if (pm) {
        bool win = false;
        if (pass) {
          const unsigned grp = __match_any_sync(pm, v);
          win = (__ffs(grp) - 1) == lane;
        }
        const unsigned wm = __ballot_sync(kFull, win);
        if (win) {
          q[tail + __popc(wm & ((1u << lane) - 1u))] = v;
          bins[v] = -1;             // removePoint + unassigned_mask[v] = false (plane_extractor.cpp:324-325)
          atomicSub(&hist[vb], 1);
        }
        tail += __popc(wm);
      }
      head += nb;
      __syncwarp();
    }
    remaining -= tail;
    if (static_cast<unsigned long long>(tail) < th.min_cells_activated) continue;  // plane_extractor.cpp:329-331

    // merge the activated cells into the candidate, seed first and twice (plane_extractor.cpp:318-323):
    // 32 cell records per round trip, staged through shared memory, then 9 sequential fp32 chains
    float acc = 0.f;
    if (lane < 9) acc = reinterpret_cast<const float*>(rec_b4)[12 * seed + lane];
    float4 r0, r1, r2;
    {
      const int c = q[min(lane, tail - 1)];
      r0 = __ldg(rec_b4 + 3 * c); r1 = __ldg(rec_b4 + 3 * c + 1); r2 = __ldg(rec_b4 + 3 * c + 2);
    }
    for (int i0 = 0; i0 < tail; i0 += 32) {
      const int cnt = min(32, tail - i0);
      float4* st4 = reinterpret_cast<float4*>(stage + 12 * lane);
      st4[0] = r0; st4[1] = r1; st4[2] = r2;
      if (i0 + 32 < tail) {  // prefetch the next 32 records while this chunk is summed
        const int c = q[min(i0 + 32 + lane, tail - 1)];
        r0 = __ldg(rec_b4 + 3 * c); r1 = __ldg(rec_b4 + 3 * c + 1); r2 = __ldg(rec_b4 + 3 * c + 2);
      }
      __syncwarp();
      if (lane < 9)
        for (int k = 0; k < cnt; ++k) acc = __fadd_rn(acc, stage[12 * k + lane]);
      __syncwarp();
    }
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
      for (int i = lane; i < n; i += 32) seg_label[list[o + i]] = label;
    }
    nseg += __popc(am);
    __syncwarp();
  }
  if (lane == 0) args.tables.n_planes[frame] = nseg;
  __syncwarp();

  // ---- getConnectedComponents (plane_extractor.cpp:430-453): boundary pairs, last row/column skipped ----
  int n_pairs = 0;
  if (nseg > 1) {
    const int limit = (nv - 1) * nh;
    for (int c0 = 0; c0 < limit; c0 += 32) {
      const int c = c0 + lane;
      unsigned p0 = 0xffffffffu, p1 = 0xffffffffu;
      if (c < limit) {
        const int qc = c % nh;
        const int id = seg_label[c];
        if (qc < nh - 1 && id > 0) {
          const int right = seg_label[c + 1], down = seg_label[c + nh];
          if (right > 0 && right != id) p0 = (static_cast<unsigned>(min(id, right) - 1) << 16) | static_cast<unsigned>(max(id, right) - 1);
          if (down > 0 && down != id) p1 = (static_cast<unsigned>(min(id, down) - 1) << 16) | static_cast<unsigned>(max(id, down) - 1);
        }
      }
      const unsigned m0 = __ballot_sync(kFull, p0 != 0xffffffffu);
      if (p0 != 0xffffffffu) __stcg(pairs + n_pairs + __popc(m0 & ((1u << lane) - 1u)), p0);
      n_pairs += __popc(m0);
      const unsigned m1 = __ballot_sync(kFull, p1 != 0xffffffffu);
      if (p1 != 0xffffffffu) __stcg(pairs + n_pairs + __popc(m1 & ((1u << lane) - 1u)), p1);
      n_pairs += __popc(m1);
    }
  }
  for (int i = lane; i < nseg; i += 32) __stcg(merge + i, i);
  __syncwarp();

  // ---- findMergedLabels (plane_extractor.cpp:402-423) ---------------------------------------------
  const int words = (nseg + 31) / 32;
  for (int r = 0; r < nseg && n_pairs > 0; ++r) {
    for (int i = lane; i < words; i += 32) rowbits[i] = 0;
    __syncwarp();
    bool any = false;
    for (int i = lane; i < n_pairs; i += 32) {
      const unsigned pr = __ldcg(pairs + i);
      if (static_cast<int>(pr >> 16) == r) {
        const unsigned t = pr & 0xffffu;
        atomicOr(&rowbits[t >> 5], 1u << (t & 31));
        any = true;
      }
    }
    any = __any_sync(kFull, any);
    __syncwarp();
    if (!any) continue;

    const int a = __ldcg(merge + r);
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
          if (lane == 0) __stcg(merge + t, a);
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

  // ---- per-cell final labels (plane_extractor.cpp:464-465) -----------------------------------------
  for (int c = lane; c < C; c += 32) {
    const int l = seg_label[c];
    cell_label[c] = (l == 0) ? 0 : __ldcg(merge + l - 1) + 1;
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
  if (off + C * 2 <= budget) {
    p.bins_smem = 1;
    p.off_bins = static_cast<int>(off);
    off = align16(off + C * 2);
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
  p.bytes = off;
  return p;
}

cudaError_t launch_region_grow(const RegionArgs& args, cudaStream_t stream) {
  if (args.n_frames == 0) return cudaSuccess;
  cudaFuncSetAttribute(region_grow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(args.plan.bytes));
  region_grow_kernel<<<args.n_frames, 32, args.plan.bytes, stream>>>(args);
  return cudaGetLastError();
}

}  // namespace dpx
