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
#include "region_grow.cuh"

#include "plane_fit.cuh"

namespace dpx {
namespace {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

__device__ __forceinline__ void load_seg(const float* rec, Moments& m, PlaneFit& f) {
  m.n = __float_as_int(__ldcg(rec + kSegN));
#pragma unroll
  for (int i = 0; i < 3; ++i) m.s[i] = __ldcg(rec + kSegS + i);
#pragma unroll
  for (int i = 0; i < 6; ++i) m.v[i] = __ldcg(rec + kSegV + i);
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
  extern __shared__ int smem_i[];
  const Geometry& g = args.geom;
  const Thresholds& th = args.thr;
  const int lane = threadIdx.x;
  const int frame = blockIdx.x;
  const int C = g.n_cells, nh = g.nh, nv = g.nv;
  const int B2 = th.histogram_bins_per_coord * th.histogram_bins_per_coord;
  const int row_words = (g.plane_cap + 31) / 32;

  int* hist = smem_i;
  unsigned* rowbits = reinterpret_cast<unsigned*>(hist + B2);
  int16_t* bins = args.bins_in_smem ? reinterpret_cast<int16_t*>(rowbits + row_words)
                                    : args.tables.bin_work + static_cast<long long>(frame) * C;

  const float4* rec_a = args.tables.rec_a + 2LL * frame * C;
  const float* rec_b = reinterpret_cast<const float*>(args.tables.rec_b + 3LL * frame * C);
  const int16_t* bin_in = args.tables.bin + static_cast<long long>(frame) * C;
  int32_t* seg_label = args.tables.seg_label + static_cast<long long>(frame) * C;
  int32_t* cell_label = args.tables.cell_label + static_cast<long long>(frame) * C;
  int32_t* queue = args.tables.queue + static_cast<long long>(frame) * C;
  uint32_t* pairs = args.tables.pairs + 2LL * frame * C;
  float* segs = args.tables.segs + static_cast<long long>(frame) * g.plane_cap * kSegFloats;
  int32_t* merge = args.tables.merge + static_cast<long long>(frame) * g.plane_cap;

  // ---- histogram of planar-cell bins (normals_histogram.cpp:21-49; bins come from stage 1) --------
  for (int i = lane; i < B2; i += 32) hist[i] = 0;
  __syncwarp();
  int remaining = 0;
  for (int c = lane; c < C; c += 32) {
    const int b = bin_in[c];
    bins[c] = static_cast<int16_t>(b);
    seg_label[c] = 0;
    if (b >= 0) {
      atomicAdd(&hist[b], 1);
      ++remaining;
    }
  }
  remaining = warp_sum(remaining);
  __syncwarp();

  const double min_cos = static_cast<double>(th.min_cos_angle_merge);
  int nseg = 0;

  // ---- createPlaneSegments (plane_extractor.cpp:302-344) ------------------------------------------
  while (remaining > 0) {
    // most frequent bin, first maximum (normals_histogram.cpp:54-56)
    int bc = -1, bi = 0x7fffffff;
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

    // seed = first strict minimum of the MSE among the bin's cells (plane_extractor.cpp:309-316)
    double lm = 2147483647.0;  // INT_MAX
    int seed = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {
      if (bins[c] == bi) {
        const double m = static_cast<double>(rec_b[12 * c + 9]);
        if (m < lm) { lm = m; seed = c; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double om = __shfl_xor_sync(kFull, lm, o);
      const int os = __shfl_xor_sync(kFull, seed, o);
      if (os != 0x7fffffff && (seed == 0x7fffffff || om < lm || (om == lm && os < seed))) { lm = om; seed = os; }
    }
    if (seed == 0x7fffffff) break;  // no candidate below INT_MAX: uninitialised read in the reference

    // growSeed (plane_extractor.cpp:349-392): batched FIFO BFS
    if (lane == 0) {
      queue[0] = seed;
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
        u = __ldcg(queue + head + e);
        const int r = u / nh, q = u - r * nh;
        if (s == 0) v = (r >= 1) ? u - nh : -1;
        else if (s == 1) v = (r + 1 < nv) ? u + nh : -1;
        else if (s == 2) v = (q >= 1) ? u - 1 : -1;
        else v = (q + 1 < nh) ? u + 1 : -1;
      }
      bool pass = false;
      int vb = -1;
      if (v >= 0) {
        vb = bins[v];
        if (vb >= 0) {  // unassigned and not yet activated
          const float4 nu = rec_a[2 * u];
          const float4 nvv = rec_a[2 * v];
          const float4 mv = rec_a[2 * v + 1];
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
          __stcg(queue + tail + __popc(wm & ((1u << lane) - 1u)), v);
          bins[v] = -1;             // removePoint + unassigned_mask[v] = false (plane_extractor.cpp:324-325)
          atomicSub(&hist[vb], 1);
        }
        tail += __popc(wm);
      }
      head += nb;
      __syncwarp();
    }
    remaining -= tail;

    // merge the activated cells into the candidate, seed first and twice (plane_extractor.cpp:318-323)
    float acc = 0.f;
    if (lane < 9) acc = rec_b[12 * seed + lane];
    for (int i0 = 0; i0 < tail; i0 += 32) {
      const int mine = (i0 + lane < tail) ? __ldcg(queue + i0 + lane) : 0;
      const int cnt = min(32, tail - i0);
      for (int k0 = 0; k0 < cnt; k0 += 8) {
        float vals[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int c = __shfl_sync(kFull, mine, (k0 + k) & 31);
          vals[k] = (lane < 9 && k0 + k < cnt) ? rec_b[12 * c + lane] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k0 + k < cnt) acc = __fadd_rn(acc, vals[k]);
      }
    }
    if (static_cast<unsigned long long>(tail) < th.min_cells_activated) continue;  // plane_extractor.cpp:329-331

    Moments mom;
    mom.n = g.patch * g.patch * (tail + 1);
#pragma unroll
    for (int i = 0; i < 3; ++i) mom.s[i] = __shfl_sync(kFull, acc, i);
#pragma unroll
    for (int i = 0; i < 6; ++i) mom.v[i] = __shfl_sync(kFull, acc, 3 + i);
    PlaneFit fit;
    fit_plane(mom, fit);  // calculateStats (plane_extractor.cpp:333)

    if (fit.score > th.min_region_planarity_score && nseg < g.plane_cap) {  // plane_extractor.cpp:336
      if (lane == 0) store_seg(segs + static_cast<long long>(nseg) * kSegFloats, mom, fit);
      ++nseg;
      for (int i = lane; i < tail; i += 32) seg_label[__ldcg(queue + i)] = nseg;
    }
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
        const int q = c % nh;
        const int id = seg_label[c];
        if (q < nh - 1 && id > 0) {
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

size_t region_grow_smem_bytes(const Geometry& g, const Thresholds& th, bool bins_in_smem) {
  const size_t B2 = static_cast<size_t>(th.histogram_bins_per_coord) * th.histogram_bins_per_coord;
  size_t bytes = B2 * 4 + static_cast<size_t>((g.plane_cap + 31) / 32) * 4;
  if (bins_in_smem) bytes += static_cast<size_t>(g.n_cells) * 2;
  return (bytes + 15) & ~static_cast<size_t>(15);
}

cudaError_t launch_region_grow(const RegionArgs& args, cudaStream_t stream) {
  if (args.n_frames == 0) return cudaSuccess;
  const size_t smem = region_grow_smem_bytes(args.geom, args.thr, args.bins_in_smem != 0);
  cudaFuncSetAttribute(region_grow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  region_grow_kernel<<<args.n_frames, 32, smem, stream>>>(args);
  return cudaGetLastError();
}

}  // namespace dpx
