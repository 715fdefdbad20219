// region_grow_cta.cuh -- stage 2 for frames of up to ~27 000 cells (every 640x480 configuration, 1280x720 and
// 1920x1080 at patch 10): one CTA of eight warps per frame; a batch runs all its frames at once and lasts as long
// as its slowest frame.
//
// Same algorithm and the same exactness arguments as region_grow_kernel (region_grow.cu); what changes is who does
// what, so that only the truly sequential part sits on the critical path:
//
//   warp 0      the sequential chain of createPlaneSegments (plane_extractor.cpp:302-331): most frequent bin,
//               seed, FIFO BFS.  One 32-bit word per cell {histogram slot, edge mask, alive, claim field} makes a
//               BFS probe a single shared-memory read; the histogram is a compacted array of keys
//               count << 15 | (0x7fff - slot) over the non-empty bins (slots ascend with the bin ids, so one integer
//               max is the reference's first-max bin), decremented once per region for the seed's bin and by rare
//               atomics for the others.  Two queue entries that reach the same cell in one step are ordered by a
//               shared-memory atomicMin on the claim field (lowest lane = earliest in FIFO order).
//               This chain is written branch-free where it can be -- predicated loads, an unrolled key scan, and
//               per-lane dummy sinks for the stores and atomics of lanes that have nothing to write: a single
//               warp pays every taken branch and every dependent instruction in full.
//   warps 1..   consume finished regions from a shared-memory ring while warp 0 keeps growing: the fp32
//               moment chains over the region's FIFO list (plane_extractor.cpp:318-327).
//   all warps   setup (histogram, grouping by bin), then after growing: plane fits one region per thread
//               (:333-337), ordered compaction into segment ids, labels_map_ painting (:339-342), adjacency
//               bit matrix (:430-453), final per-cell labels (:464-465).  findMergedLabels (:402-423) is a
//               short sequential loop run by warp 0 on shared-memory plane records.
//
// Seeds: for frames beyond mode 0 the cells arrive sorted by (histogram bin, MSE, cell id) from seed_sort_kernel
// (seed_sort.cuh), so "first strict-minimum MSE among the bin's unassigned cells" (plane_extractor.cpp:309-316) is the
// first entry of the bin's run that is still alive, found by a per-bin cursor that only moves forward.
//
// Storage modes (template parameter, chosen by the frame size): see region_grow_cta_kernel.
#pragma once
#include "labeling.cuh"

namespace dpx {
namespace {

#ifndef DPX_CTA_THREADS
#define DPX_CTA_THREADS 128  // A/B at build time: 256 (twice the helper warps, two CTAs per SM instead of three)
#endif
constexpr int kCtaThreads = DPX_CTA_THREADS;
constexpr int kCtaWarps = kCtaThreads / 32;
// cell word: bits 0-15 slot of the cell's initial bin in the compacted histogram, 16-19 edge mask, 20 alive (planar and
// unassigned), 21-31 claim field (all ones while idle; see the BFS step)
constexpr unsigned kAlive = 1u << 20;
constexpr unsigned kClaimIdle = 0x7ffu << 21;
constexpr int kRecFloats = kSegFloats;  // region / plane records use the segs layout
constexpr int kSegDirty = 21;           // merge phase: the plane's moments have grown since its last fit (see findMergedLabels)

constexpr int kRing = 1024;    // modes 1/2: most recent queue entries mirrored in shared memory (power of two)

struct CtaPlan {
  int off_stage, off_hkey, off_binslot, off_cursor, off_binend, off_cw, off_list, off_keys, off_ring, off_win, off_wpos, off_wend,
      off_recs, off_merge, off_misc;
  int rec_cap;       // region / plane records held in shared memory (the rest spill to the global segs table)
  int adj_bytes;     // bytes available to the adjacency bit matrix
  int merge_smem;    // merge labels in shared memory (else in the global merge table)
  int win;           // sorted-key entries cached per bin in shared memory (32, 16 or 8: what fits)
  size_t bytes;
};

__device__ __noinline__ void fit_plane_call(const Moments& m, PlaneFit& f) { fit_plane(m, f); }

// Cell words.  32-bit form (modes 0 and 1): bits 0-15 slot of the cell's initial bin in the compacted histogram, 16-19
// edge mask, 20 alive (planar and unassigned), 21-31 claim field (all ones while idle; see the BFS step).
// 16-bit form (mode 2, frames too large for 4 bytes per cell of shared memory): bits 0-9 slot, 10-13 edge mask, 14
// alive; no claim field -- clashes between queue entries are settled with shuffles instead.
template <bool CW16>
struct CellWordT {
  using type = unsigned;
  static constexpr unsigned slot_mask = 0xffffu, alive = kAlive;
  static constexpr int edge_shift = 16;
};
template <>
struct CellWordT<true> {
  using type = uint16_t;
  static constexpr unsigned slot_mask = 0x3ffu, alive = 1u << 14;
  static constexpr int edge_shift = 10;
};

// Queue storage of the BFS: the cell list of a frame.  Mode 0 keeps it in shared memory; the larger modes keep it in the
// global `queue` table (its reads and writes are sequential) and mirror the most recent kRing entries in shared memory,
// which is where the BFS reads them back from unless the frontier is longer than the ring.
template <bool RING>
struct QueueT {
  int32_t* list;     // mode 0: shared memory; else global, frame base
  unsigned* ring;    // [kRing] or nullptr
  __device__ __forceinline__ unsigned read(int abs_pos, bool in_ring) const {
    if (RING) return in_ring ? ring[abs_pos & (kRing - 1)] : static_cast<unsigned>(__ldcg(list + abs_pos));
    return static_cast<unsigned>(list[abs_pos]);
  }
};

// Wide BFS step (32-bit cell words only): one lane per queue entry (up to 32), each lane probes its four neighbours.
// FIFO order is (entry, slot) lexicographic; a cell reached more than once goes to the smallest entry * 4 + slot, posted
// into the cell word's claim field with a shared-memory atomicMin.  Kept out of line so that the narrow step's register
// allocation is not disturbed.  `base` = absolute list position of the region's first entry.  Returns {cells appended,
// appended cells that belong to the seed's bin}.
constexpr int kWideThreshold = 12;
#ifdef DPX_BFS_PROBE
__device__ long long g_wide_probe[8];
#define WIDE_PROBE(slot) do { if (lane == 0) { const long long t__ = clock64(); g_wide_probe[slot] += t__ - wp_t; wp_t = t__; } } while (0)
#else
#define WIDE_PROBE(slot) do { } while (0)
#endif
template <bool RING>
__device__ __noinline__ int2 bfs_wide_step(const QueueT<RING> qs, int base, unsigned* cw, unsigned* hkey, int32_t* sink, int head,
                                           int tail, int lane, int nh, int bslot, int n_cells) {
#ifdef DPX_BFS_PROBE
  long long wp_t = clock64();
#endif
  const int nbw = min(32, tail - head);
  const bool in_ring = tail - head <= kRing - 128;
  // (the list pointer reaches this out-of-line function as a generic pointer: when the list is in shared memory, address it
  // as such -- a generic load / store is slower and, for a lone warp, that is latency on the chain)
  const unsigned list_s = RING ? 0u : static_cast<unsigned>(__cvta_generic_to_shared(qs.list + base));
  const unsigned ring_s = RING ? static_cast<unsigned>(__cvta_generic_to_shared(qs.ring)) : 0u;
  unsigned pk = 0;
  if (RING) {
    if (lane < nbw) {
      if (in_ring) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(pk) : "r"(ring_s + 4u * static_cast<unsigned>((base + head + lane) & (kRing - 1))));
      else pk = static_cast<unsigned>(__ldcg(qs.list + base + head + lane));
    }
  } else {
    if (lane < nbw) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(pk) : "r"(list_s + 4u * static_cast<unsigned>(head + lane)));
  }
  const int u = static_cast<int>(pk & 0xffffffu);
  int vv[4];
  unsigned ww[4];
  unsigned passm = 0;
#pragma unroll
  for (int sl4 = 0; sl4 < 4; ++sl4) {
    const int dl = (sl4 == 0) ? -nh : (sl4 == 1) ? nh : (sl4 == 2) ? -1 : 1;
    vv[sl4] = u + dl;
    DPX_CHECK(!((pk >> (24 + sl4)) & 1u) || (vv[sl4] >= 0 && vv[sl4] < n_cells));
    ww[sl4] = ((pk >> (24 + sl4)) & 1u) ? cw[vv[sl4]] : 0u;
    if (ww[sl4] & kAlive) passm |= 1u << sl4;
  }
  const unsigned anyp = __ballot_sync(kFull, passm != 0);
  WIDE_PROBE(0);
  unsigned winm = passm;
  if (anyp & (anyp - 1u)) {  // at least two entries have candidates: they may clash
    // (branch-free: a slot without a candidate posts into the lane's private sink)
#pragma unroll
    for (int sl4 = 0; sl4 < 4; ++sl4)
      atomicMin((passm & (1u << sl4)) ? &cw[vv[sl4]] : reinterpret_cast<unsigned*>(sink) + lane,
                (ww[sl4] & ~kClaimIdle) | (static_cast<unsigned>(lane * 4 + sl4) << 21));
    __syncwarp();
#pragma unroll
    for (int sl4 = 0; sl4 < 4; ++sl4)
      if ((passm & (1u << sl4)) && (cw[vv[sl4]] >> 21) != static_cast<unsigned>(lane * 4 + sl4)) winm &= ~(1u << sl4);
  }
  WIDE_PROBE(1);
  // append position of (lane, slot) = winners of lower lanes + own lower slots.  A lane has 0..4 winners: the prefix
  // over the lanes comes from three ballots on the bits of that count instead of a five-round shuffle scan.
  const unsigned nwin = __popc(winm);
  const unsigned lt_mask = (1u << lane) - 1u;
  const unsigned b0 = __ballot_sync(kFull, nwin & 1u), b1 = __ballot_sync(kFull, nwin & 2u), b2 = __ballot_sync(kFull, nwin & 4u);
  int pos = tail + __popc(b0 & lt_mask) + 2 * __popc(b1 & lt_mask) + 4 * __popc(b2 & lt_mask);
  const int total = __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
  DPX_CHECK(base + tail + total <= n_cells);
  WIDE_PROBE(2);
  int same = 0;
  // branch-free: slots that did not win write to a private sink (see the narrow step)
#pragma unroll
  for (int sl4 = 0; sl4 < 4; ++sl4) {
    const bool won = (winm >> sl4) & 1u;
    const unsigned slt = ww[sl4] & 0xffffu;
    const bool other_bin = won && slt != static_cast<unsigned>(bslot);
    const int entry = vv[sl4] | static_cast<int>(((ww[sl4] >> 16) & 0xfu) << 24);
    unsigned* cdst = won ? cw + vv[sl4] : reinterpret_cast<unsigned*>(sink) + lane;
    if (RING) {
      // large frames: the step writes the shared-memory ring only; the caller copies the ring to the global list in bulk
      const unsigned rdst = won ? ring_s + 4u * static_cast<unsigned>((base + pos) & (kRing - 1))
                                : static_cast<unsigned>(__cvta_generic_to_shared(sink + lane));
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(rdst), "r"(entry) : "memory");
    } else {
      const unsigned qdst = won ? list_s + 4u * static_cast<unsigned>(pos)
                                : static_cast<unsigned>(__cvta_generic_to_shared(sink + lane));
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(qdst), "r"(entry) : "memory");
    }
    *cdst = ww[sl4] & ~kAlive;
    pos += won ? 1 : 0;
    same += (won && !other_bin) ? 1 : 0;
    atomicSub(other_bin ? hkey + slt : reinterpret_cast<unsigned*>(sink) + lane, 1u << 15);
    WIDE_PROBE(3 + 0 * sl4);
  }
  __syncwarp();
  WIDE_PROBE(4);
  return make_int2(total, same);
}


// MODE 0: everything in shared memory -- cell words, cell list, member runs grouped by bin (frames of up to ~5 000
//         cells, several CTAs per SM).  No sort: at this size the minimum scan over a bin's run is cheaper than sorting.
// MODE 1: sorted seeds (keys in global memory / L2 with a small window per bin in shared memory); 32-bit cell words and the
//         cell list in shared memory (up to ~21 000 cells: 640x480 at patch 4, 1920x1080 at patch 10).
// MODE 2: as mode 1 with the cell list in the global `queue` table and a shared-memory ring of its most recent entries
//         (up to ~42 000 cells); merge labels in global memory when they do not fit.
// MODE 3: as mode 2 with 16-bit cell words (up to ~85 000 cells: 1920x1080 at patch 5, 1280x720 at patch 4); queue
//         clashes are settled with shuffles and only the narrow BFS step is used.
template <int MODE>
__global__ void __launch_bounds__(kCtaThreads) region_grow_cta_kernel(const RegionArgs args, const CtaPlan plan) {
  constexpr bool ALL_SMEM = MODE == 0;   // no sort: member runs in shared memory, minimum scan per seed
  constexpr bool RING = MODE >= 2;       // cell list in global memory + shared-memory ring (else the list is in shared memory)
  constexpr bool CW16 = MODE == 3;
  using CW = CellWordT<CW16>;
  using word_t = typename CW::type;
  constexpr unsigned kAliveW = CW::alive, kSlotMask = CW::slot_mask;
  constexpr int kEdgeShift = CW::edge_shift;
  extern __shared__ float4 smem_f4[];
  const Geometry& g = args.geom;
  const Thresholds& th = args.thr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_wait();
  const int frame = blockIdx.x;
  if (blockIdx.x == 0 && threadIdx.x == 0) args.tables.axis_work[0] = 0;  // the list of this batch has been worked off
  const int C = g.n_cells, nh = g.nh, nv = g.nv;
  const int B2 = th.histogram_bins_per_coord * th.histogram_bins_per_coord;
  const long long fc = static_cast<long long>(frame) * C;

  char* smem = reinterpret_cast<char*>(smem_f4);
  float* stage_all = reinterpret_cast<float*>(smem + plan.off_stage);        // [kCtaWarps - 1][32][12]
  unsigned* hkey = reinterpret_cast<unsigned*>(smem + plan.off_hkey);        // [K] count << 15 | (0x7fff - slot), non-empty bins
  int16_t* binslot = reinterpret_cast<int16_t*>(smem + plan.off_binslot);    // [B2] bin -> slot in hkey (or -1)
  int* cursor = reinterpret_cast<int*>(smem + plan.off_cursor);              // [K] first entry of the bin's run that may be alive
  int* bin_end = reinterpret_cast<int*>(smem + plan.off_binend);             // [K] end of the bin's run in the sorted keys
  word_t* cw = reinterpret_cast<word_t*>(smem + plan.off_cw);                // [C] cell words (later the segment labels)
  // [C] BFS queues = region cell lists
  QueueT<RING> qs;
  qs.list = RING ? args.tables.queue + fc : reinterpret_cast<int32_t*>(smem + plan.off_list);
  qs.ring = RING ? reinterpret_cast<unsigned*>(smem + plan.off_ring) : nullptr;
  int32_t* list = qs.list;
  // the frame's cells sorted by (bin, MSE, cell id): seed_sort_kernel's output
  const unsigned long long* skeys_g = args.tables.skeys + fc;
  // mode 0 (small frames, where a sort would cost more than it saves): [C] cell ids grouped by initial bin and [C] their
  // MSE in the same order; the seed search is a minimum scan over the bin's run that compacts the run as it goes
  int32_t* members = reinterpret_cast<int32_t*>(smem + plan.off_keys);
  float* msem = reinterpret_cast<float*>(smem + plan.off_keys) + C;
  int* bin_off = bin_end;  // mode 0: start of the bin's member run
  int* run_end = cursor;   // mode 0: end of its still-unassigned members
  const float* mse_g = args.tables.mse + fc;
  unsigned long long* win = reinterpret_cast<unsigned long long*>(smem + plan.off_win);       // sorted modes: [K][kw], kw below
  int* wpos = reinterpret_cast<int*>(smem + plan.off_wpos);                  // modes 1/2: [K] sorted position of win[slot][0]
  int* wend = reinterpret_cast<int*>(smem + plan.off_wend);                  // modes 1/2: [K] end of the window's valid entries
  float* recs = reinterpret_cast<float*>(smem + plan.off_recs);              // [rec_cap][24]
  int32_t* merge_out = args.tables.merge + static_cast<long long>(frame) * g.plane_cap;
  int32_t* merge = plan.merge_smem ? reinterpret_cast<int32_t*>(smem + plan.off_merge) : merge_out;  // [plane_cap]
  volatile int* misc = reinterpret_cast<volatile int*>(smem + plan.off_misc);
  // misc: [0] regions published, [1] growing finished, [2] scratch counter, [3] remaining planar cells, [4] K,
  //       [8..8+kCtaWarps) per-warp counts for the ordered compaction
  // [B2] raw histogram during setup, in storage that is not in use yet (the list / the key windows)
  int* hist_tmp = ALL_SMEM ? reinterpret_cast<int*>(list) : reinterpret_cast<int*>(win);
  int32_t* dummy = reinterpret_cast<int32_t*>(smem + plan.off_misc) + 32;  // [32] per-lane sink of the branch-free BFS tail

  const float4* rec_b4 = args.tables.rec_b + 3 * fc;
  const int16_t* bin_in = args.tables.bin + fc;
  const uint8_t* edge_in = args.tables.edge + fc;
  int32_t* seg_label = args.tables.seg_label + fc;
  int32_t* cell_label = args.tables.cell_label + fc;
  float* segs_g = args.tables.segs + static_cast<long long>(frame) * g.plane_cap * kSegFloats;
  auto rec_ptr = [&](int r) -> float* {
    return r < plan.rec_cap ? recs + r * kRecFloats : segs_g + static_cast<long long>(r) * kSegFloats;
  };

  const bool prof = args.prof != nullptr;
  const long long t_kernel0 = prof ? clock64() : 0;

  // ---- setup: cell words, histogram (normals_histogram.cpp:21-49; bins come from stage 1) ----------
  for (int i = tid; i < B2; i += kCtaThreads) {
    hist_tmp[i] = 0;
    binslot[i] = -1;
    hkey[i] = 0;
  }
  if (tid < 4) hkey[B2 + tid] = 0;  // the vectorised scan may read up to three entries past the last bin
  if (tid < 8 + kCtaWarps) misc[tid] = 0;
  __syncthreads();
  // (eight consecutive cells per thread and round when the tables allow 16-byte loads: with 128 threads the loop is
  // bound by the latency of its global loads, one round trip per round)
  const bool vec8 = (C & 7) == 0;
  auto init_cell = [&](int c, int b, unsigned e) {
    const unsigned w0 = b >= 0 ? (static_cast<unsigned>(b) | (e << kEdgeShift) | kAliveW | (CW16 ? 0u : kClaimIdle)) : 0u;
    cw[c] = static_cast<word_t>(w0);
    if (b >= 0) atomicAdd(&hist_tmp[b], 1);
  };
  if (vec8) {
    for (int c0 = tid * 8; c0 < C; c0 += kCtaThreads * 8) {
      const uint4 b8 = __ldg(reinterpret_cast<const uint4*>(bin_in + c0));
      const uint2 e8 = __ldg(reinterpret_cast<const uint2*>(edge_in + c0));
      reinterpret_cast<int4*>(seg_label + c0)[0] = make_int4(0, 0, 0, 0);
      reinterpret_cast<int4*>(seg_label + c0)[1] = make_int4(0, 0, 0, 0);
      const unsigned bw[4] = {b8.x, b8.y, b8.z, b8.w};
      const unsigned ew[2] = {e8.x, e8.y};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int b = static_cast<int>(static_cast<int16_t>((bw[k >> 1] >> (16 * (k & 1))) & 0xffffu));
        init_cell(c0 + k, b, (ew[k >> 2] >> (8 * (k & 3))) & 0xffu);
      }
    }
  } else {
    for (int c = tid; c < C; c += kCtaThreads) {
      seg_label[c] = 0;
      init_cell(c, static_cast<int>(__ldg(bin_in + c)), static_cast<unsigned>(__ldg(edge_in + c)));
    }
  }
  __syncthreads();
  // compact the non-empty bins; their runs in the sorted keys follow each other in bin order (warp 0)
  if (warp == 0) {
    int K = 0, run = 0;
    for (int b0 = 0; b0 < B2; b0 += 32) {
      const int b = b0 + lane;
      const int cnt = b < B2 ? hist_tmp[b] : 0;
      const unsigned nz = __ballot_sync(kFull, cnt > 0);
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += t;
      }
      __syncwarp();  // (hist_tmp aliases the key windows: all lanes have read their counts before wpos / wend are written)
      if (cnt > 0) {
        const int slot = K + __popc(nz & ((1u << lane) - 1u));
        binslot[b] = static_cast<int16_t>(slot);
        hkey[slot] = (static_cast<unsigned>(cnt) << 15) | (0x7fffu - static_cast<unsigned>(slot));
        const int pos = run + incl - cnt;
        cursor[slot] = pos;                           // mode 0: end of the run's live members, filled by the scatter below
        bin_end[slot] = ALL_SMEM ? pos : pos + cnt;   // mode 0: start of the bin's member run
        if (!ALL_SMEM) {
          wpos[slot] = pos;
          wend[slot] = pos;  // empty window
        }
      }
      K += __popc(nz);
      run += __shfl_sync(kFull, incl, 31);
    }
    if (lane == 0) {
      misc[3] = run;
      misc[4] = K;
    }
  }
  __syncthreads();
  auto slot_cell = [&](int c, float mse) {
    const unsigned w = cw[c];
    if (w & kAliveW) {
      const unsigned slot = static_cast<unsigned>(binslot[w & kSlotMask]);
      cw[c] = static_cast<word_t>((w & ~kSlotMask) | slot);  // from here on the word carries the histogram slot, not the bin id
      if (ALL_SMEM) {
        // small frames: the bin's members in arrival order; the seed search scans (and compacts) the run
        const int pos = atomicAdd(&cursor[slot], 1);
        members[pos] = c;
        msem[pos] = mse;
      }
    }
  };
  if (ALL_SMEM && vec8) {
    for (int c0 = tid * 8; c0 < C; c0 += kCtaThreads * 8) {
      // (the MSE table is written for valid cells only; entries of the others are never used)
      const float4 m0 = __ldg(reinterpret_cast<const float4*>(mse_g + c0)), m1 = __ldg(reinterpret_cast<const float4*>(mse_g + c0) + 1);
      const float mv[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) slot_cell(c0 + k, mv[k]);
    }
  } else {
    for (int c = tid; c < C; c += kCtaThreads) slot_cell(c, ALL_SMEM ? __ldg(mse_g + c) : 0.f);
  }
  __syncthreads();  // (hist_tmp aliases the list / the windows: nobody touches those before this barrier)

  const int cell_pts = g.patch * g.patch;
  long long t_seed = 0, t_bfs = 0, t_mark = 0, t_wide = 0;
#ifdef DPX_SEED_PROBE
  long long t_pick_probe = 0;
  int n_refills_probe = 0;
#endif
  int n_seeds = 0, n_steps = 0;
  const long long t_init = prof ? clock64() - t_kernel0 : 0;

  if (warp == 0) {
    // ================= the sequential chain: createPlaneSegments (plane_extractor.cpp:302-331) ==========
    int remaining = misc[3];
    const int K4 = (misc[4] + 3) / 4;
    // entries per key window: the window storage is sized for B2 bins of plan.win entries, but only the K non-empty bins
    // need one -- they share it (a power of two, at most 32 = one warp-wide probe)
    int kw = plan.win;
    if (!ALL_SMEM) {
      const int room = (B2 * plan.win) / max(misc[4], 1);
      while (kw < 32 && 2 * kw <= room) kw *= 2;
    }
    int n_regions = 0, list_off = 0;
    const int slot4 = lane & 3;
    const int delta = (slot4 == 0) ? -nh : (slot4 == 1) ? nh : (slot4 == 2) ? -1 : 1;
    while (remaining > 0) {
      if (prof) t_mark = clock64();
      // most frequent bin, first maximum (normals_histogram.cpp:54-56): max key = largest count, smallest slot
      // (slots ascend with the bin ids, so the smallest slot is the smallest bin id)
      unsigned key = 0;
      // four keys per lane and load, the first 512 slots without a branch (entries past the last bin are zero)
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = lane + 32 * u;
        uint4 k4 = make_uint4(0u, 0u, 0u, 0u);
        if (i < K4) k4 = reinterpret_cast<const uint4*>(hkey)[i];
        key = max(max(key, max(k4.x, k4.y)), max(k4.z, k4.w));
      }
      for (int i = lane + 128; i < K4; i += 32) {
        const uint4 k4 = reinterpret_cast<const uint4*>(hkey)[i];
        key = max(max(key, max(k4.x, k4.y)), max(k4.z, k4.w));
      }
      key = __reduce_max_sync(kFull, key);
      const int bc = static_cast<int>(key >> 15), bslot = static_cast<int>(0x7fffu - (key & 0x7fffu));
#ifdef DPX_SEED_PROBE
      if (prof) t_wide += clock64() - t_mark;
#endif
      const unsigned long long n_cand = bc > 0 ? static_cast<unsigned long long>(bc) : 0ull;
      if (n_cand < th.min_candidate_size) break;  // plane_extractor.cpp:305-307

      // seed = first strict minimum of the MSE among the bin's unassigned cells (plane_extractor.cpp:309-316)
      float lm = __int_as_float(0x7f800000);
      int seed = kNoSeed;
      if (ALL_SMEM) {
        // small frames: first strict minimum of the MSE among the bin's unassigned cells (plane_extractor.cpp:309-316) by a
        // scan over the bin's member run that also compacts the run down to the cells that are still unassigned
        {
          const int start = bin_off[bslot], end = run_end[bslot];
          int w = start;
          int32_t* mp = members;
          float* sp = msem;
          // four 32-member chunks per round: their loads are independent, so a round costs one memory round trip
          if (end - start <= 32) {
            // short run (the usual case for the left-over bins that produce one-cell regions): one chunk, straight-line
            const int i = start + lane;
            const bool in = i < end;
            const int c = in ? mp[i] : 0;
            const float m = in ? sp[i] : 0.f;
            const unsigned wv = in ? cw[c] : 0u;
            const bool alive = (wv & kAliveW) != 0;
            const unsigned am = __ballot_sync(kFull, alive);
            // stable compaction; lanes without a live member write to their private sink
            const int pos = start + __popc(am & ((1u << lane) - 1u));
            int32_t* mdst = alive ? mp + pos : dummy + lane;
            float* sdst = alive ? sp + pos : reinterpret_cast<float*>(dummy) + lane;
            *mdst = c;
            *sdst = m;
            if (alive) { lm = m; seed = c; }
            w = start + __popc(am);
            __syncwarp();
          } else {
            // long run: four 32-member chunks per round; the loads of the next round are issued before this round is
            // processed, so that a round costs its processing and not a memory round trip (member runs of large
            // frames live in global memory)
            int cn[4];
            float mn[4];
            auto fetch = [&](int i0) {
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int i = i0 + 32 * u + lane;
                const bool in = i < end;
                cn[u] = in ? mp[i] : -1;
                mn[u] = in ? sp[i] : 0.f;
              }
            };
            fetch(start);
            for (int i0 = start; i0 < end; i0 += 128) {
              int c[4];
              float m[4];
              bool alive[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                c[u] = cn[u];
                m[u] = mn[u];
              }
              if (i0 + 128 < end) fetch(i0 + 128);  // reads entries the compaction below cannot reach
#pragma unroll
              for (int u = 0; u < 4; ++u) alive[u] = c[u] >= 0 && (cw[c[u]] & kAliveW);
              __syncwarp();  // all reads of this round precede its (possibly overlapping) compaction writes
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                if (alive[u] && (m[u] < lm || (m[u] == lm && c[u] < seed))) { lm = m[u]; seed = c[u]; }
                const unsigned am = __ballot_sync(kFull, alive[u]);
                if (alive[u] && (am != kFull || w != i0 + 32 * u)) {
                  const int pos = w + __popc(am & ((1u << lane) - 1u));
                  mp[pos] = c[u];
                  sp[pos] = m[u];
                }
                w += __popc(am);
              }
              __syncwarp();
            }
          }
          if (lane == 0) run_end[bslot] = w;
        }
        {
          const unsigned fb = __float_as_uint(lm);
          const unsigned ord = fb ^ ((fb >> 31) ? 0xffffffffu : 0x80000000u);
          const unsigned best = __reduce_min_sync(kFull, seed == kNoSeed ? 0xffffffffu : ord);
          const unsigned cand = (seed != kNoSeed && ord == best) ? static_cast<unsigned>(seed) : static_cast<unsigned>(kNoSeed);
          seed = static_cast<int>(__reduce_min_sync(kFull, cand));
          const unsigned bb = best ^ ((best >> 31) ? 0x80000000u : 0xffffffffu);
          lm = __uint_as_float(bb);
        }
      } else {
        // larger frames: the first entry of the bin's (MSE, cell id)-sorted run that is still alive (seed_sort.cuh);
        // everything before the cursor is dead for good
        unsigned seed_ord = 0xffffffffu;
        {
          int pos = cursor[bslot];
          const int end = bin_end[bslot];
          {
            int wb = wpos[bslot], we = wend[bslot];
            while (pos < end) {
              if (pos >= wb && pos < we) {
                // the bin's window in shared memory holds sorted entries [wb, we)
                const int i = pos + lane;
                const bool in = i < we;
                DPX_CHECK(!in || (i - wb >= 0 && i - wb < kw && i < C));
              const unsigned long long k = in ? win[bslot * kw + (i - wb)] : 0ull;
                const int c = static_cast<int>(k & kSeedCellMask);
                const bool alive = in && (cw[c] & kAliveW) != 0;
                const unsigned am = __ballot_sync(kFull, alive);
                if (am) {
                  const int f = __ffs(am) - 1;
                  seed = __shfl_sync(kFull, c, f);
                  seed_ord = __shfl_sync(kFull, static_cast<unsigned>(k >> kSeedCellBits), f);
                  pos += f + 1;
                  break;
                }
                pos = we;
                continue;
              }
  #ifdef DPX_SEED_PROBE
            ++n_refills_probe;
#endif
            // 128 entries from the L2-resident sorted keys; the entries behind the seed refill the window
              unsigned long long k[4];
              unsigned am[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int i = pos + 32 * u + lane;
                DPX_CHECK(end <= C);
              k[u] = i < end ? __ldg(skeys_g + i) : 0ull;
              }
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int i = pos + 32 * u + lane;
                const bool alive = i < end && (cw[static_cast<int>(k[u] & kSeedCellMask)] & kAliveW) != 0;
                am[u] = __ballot_sync(kFull, alive);
              }
              const int fu = am[0] ? 0 : am[1] ? 1 : am[2] ? 2 : am[3] ? 3 : -1;
              if (fu < 0) {
                pos += 128;
                continue;
              }
              const unsigned amf = fu == 0 ? am[0] : fu == 1 ? am[1] : fu == 2 ? am[2] : am[3];
              const unsigned long long kf = fu == 0 ? k[0] : fu == 1 ? k[1] : fu == 2 ? k[2] : k[3];
              const int f = __ffs(amf) - 1;
              seed = __shfl_sync(kFull, static_cast<int>(kf & kSeedCellMask), f);
              seed_ord = __shfl_sync(kFull, static_cast<unsigned>(kf >> kSeedCellBits), f);
              const int fpos = pos + 32 * fu + f;
              const int loaded_end = min(end, pos + 128);
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int i = pos + 32 * u + lane;
                if (i > fpos && i <= fpos + kw && i < loaded_end) win[bslot * kw + (i - fpos - 1)] = k[u];
              }
              wb = fpos + 1;
              we = min(fpos + 1 + kw, loaded_end);
              if (lane == 0) {
                wpos[bslot] = wb;
                wend[bslot] = we;
              }
              pos = fpos + 1;
              break;
            }
          }
          if (lane == 0) cursor[bslot] = pos;
        }
        lm = seed_mse_from_order(seed_ord);
      }
      // no candidate with mse < INT_MAX: the reference reads an uninitialised seed id here.  (double)lm < 2147483647.0
      // is lm < 2^31 for a float: the largest float below 2^31 is 2^31 - 128.
#ifdef DPX_SEED_PROBE
      if (prof) t_pick_probe += clock64() - t_mark;
#endif
      if (seed == kNoSeed || !(lm < 2147483648.0f)) break;
      DPX_CHECK(seed >= 0 && seed < C && (cw[seed] & kAliveW) != 0 && bslot < misc[4] && list_off < C);
      if (prof) { const long long t = clock64(); t_seed += t - t_mark; t_mark = t; ++n_seeds; }

      const unsigned seed_w = cw[seed];
      if (!ALL_SMEM && ((seed_w >> kEdgeShift) & 0xfu) == 0 && th.min_cells_activated > 1ull) {
        // Isolated seed (no passing edge): its region is the seed alone and is rejected (plane_extractor.cpp:329-331), so
        // all that happens is removePoint + unassigned_mask[seed] = false.  On noisy fine grids this is the common case
        // (tens of thousands per frame): skip the queue, the BFS set-up and the histogram reduction.
        __syncwarp();
        if (lane == 0) {
          cw[seed] = static_cast<word_t>(seed_w & ~kAliveW);
          hkey[bslot] -= 1u << 15;
        }
        --remaining;
        __syncwarp();
        continue;
      }

      // growSeed (plane_extractor.cpp:349-392): batched FIFO BFS into list[list_off ...)
      int same = 0;  // cells of the seed's bin activated by this lane (histogram is settled after the BFS)
      __syncwarp();
      {
        const int entry = seed | static_cast<int>(((seed_w >> kEdgeShift) & 0xfu) << 24);
        word_t* cd = lane == 0 ? cw + seed : reinterpret_cast<word_t*>(dummy + lane);
        if (RING) {
          unsigned* rd = lane == 0 ? qs.ring + (list_off & (kRing - 1)) : reinterpret_cast<unsigned*>(dummy) + lane;
          *rd = static_cast<unsigned>(entry);
        } else {
          int32_t* qd = lane == 0 ? list + list_off : dummy + lane;
          *qd = entry;
        }
        *cd = static_cast<word_t>(seed_w & ~kAliveW);
        same = lane == 0 ? 1 : 0;
      }
      __syncwarp();
      // Large frames: the steps below append to the shared-memory ring only (a global store inside the step would put its
      // latency on the chain through the warp barrier that follows it); the ring is copied to the global list in blocks of
      // kFlush entries while the region grows -- always before the ring wraps over them -- and the rest when the region is
      // published.  Regions that are discarded never reach global memory at all.
      constexpr int kFlush = 512;
      int flush_pos = list_off;  // absolute list position up to which the global list is complete
      auto flush_to = [&](int upto) {
        for (int p0 = flush_pos; p0 < upto; p0 += 32) {
          const int p1 = p0 + lane;
          if (p1 < upto) list[p1] = static_cast<int32_t>(qs.ring[p1 & (kRing - 1)]);
        }
        flush_pos = upto;
      };
      // a seed without a single passing edge is a region of one cell: no BFS step needed
      int head = ((seed_w >> kEdgeShift) & 0xfu) ? 0 : 1, tail = 1;
      const int ent = lane >> 2;
      const unsigned lt_mask = (1u << lane) - 1u;
      const unsigned my_claim = static_cast<unsigned>(lane) << 21;
      while (head < tail) {
        // (32-bit words of large frames only: the instantiation for small frames keeps the narrow step's tighter code)
        if ((MODE == 1 || MODE == 2) && tail - head > kWideThreshold) {
          const long long tw0 = prof ? clock64() : 0;
          const int2 r = bfs_wide_step<RING>(qs, list_off, reinterpret_cast<unsigned*>(cw), hkey, dummy, head, tail, lane, nh, bslot, C);
          head += min(32, tail - head);
          tail += r.x;
          same += r.y;
          ++n_steps;
          if (list_off + tail - flush_pos >= kFlush) flush_to(flush_pos + kFlush);
          if (prof) t_wide += clock64() - tw0;
          continue;
        }
        // straight-line step (a single warp pays every branch and every dependent instruction in full)
        const int nb = min(8, tail - head);
        const bool in_ring = tail - head <= kRing - 64;  // the step's entries are still mirrored in shared memory
        unsigned pk = 0;
        if (ent < nb) pk = qs.read(list_off + head + ent, in_ring);
        const bool edge_ok = ((pk >> (24 + slot4)) & 1u) != 0;
        const int v = static_cast<int>(pk & 0xffffffu) + delta;
        unsigned w = 0;
        DPX_CHECK(!edge_ok || (v >= 0 && v < C));
        if (edge_ok) w = cw[v];
        const bool pass = (w & kAliveW) != 0;  // edge test passed, still unassigned, not yet activated
        const unsigned pm = __ballot_sync(kFull, pass);
        // lanes of one queue entry have distinct targets: a clash needs passing lanes in two different nibbles.
        const unsigned nib = (pm | (pm >> 1) | (pm >> 2) | (pm >> 3)) & 0x11111111u;
        bool win = pass;
        if (nib & (nib - 1u)) {
          if (!CW16) {
            // every candidate posts its lane number into the cell word's claim field with a shared-memory atomicMin (the
            // rest of the word is identical for all of them) and the lowest lane -- the earliest in FIFO order -- finds
            // itself there.  (match.any would do the same but costs ~250 cycles on this path.)
            unsigned* cw32 = reinterpret_cast<unsigned*>(cw);
            if (pass) atomicMin(&cw32[v], (w & ~kClaimIdle) | my_claim);
            __syncwarp();
            if (pass) win = (cw32[v] >> 21) == static_cast<unsigned>(lane);
          } else {
            // no room for a claim field in 16-bit words: an earlier entry e2 reaches the same cell v through its slot s2
            // exactly when its own cell is v's neighbour on the opposite side and that lane passed too
            const int ucell = static_cast<int>(pk & 0xffffffu);
            bool lose = false;
#pragma unroll
            for (int e2 = 0; e2 < 7; ++e2) {
              const int u2 = __shfl_sync(kFull, ucell, 4 * e2);
              const int d = u2 - v;
              const unsigned p4 = (pm >> (4 * e2)) & 0xfu;
              const bool hit = (d == nh && (p4 & 1u)) || (d == -nh && (p4 & 2u)) || (d == 1 && (p4 & 4u)) || (d == -1 && (p4 & 8u));
              lose |= (e2 < ent) && hit;
            }
            win = pass && !lose;
          }
        }
        const unsigned wm = __ballot_sync(kFull, win);
        // branch-free tail: lanes that did not win write to a private dummy word instead of skipping
        const unsigned sl = w & kSlotMask;
        const bool other_bin = win && sl != static_cast<unsigned>(bslot);
        const int wpos_q = list_off + tail + __popc(wm & lt_mask);
        DPX_CHECK(!win || (wpos_q < C && (w & kSlotMask) < static_cast<unsigned>(misc[4])));
        const int entry = v | static_cast<int>(((w >> kEdgeShift) & 0xfu) << 24);
        word_t* cdst = win ? cw + v : reinterpret_cast<word_t*>(dummy + lane);
        if (RING) {
          unsigned* rdst = win ? qs.ring + (wpos_q & (kRing - 1)) : reinterpret_cast<unsigned*>(dummy) + lane;
          *rdst = static_cast<unsigned>(entry);
        } else {
          int32_t* qdst = win ? list + wpos_q : dummy + lane;
          *qdst = entry;
        }
        *cdst = static_cast<word_t>(w & ~kAliveW);  // removePoint + unassigned_mask[v] = false (plane_extractor.cpp:324-325); claim idle again
        same += (win && !other_bin) ? 1 : 0;
        atomicSub(other_bin ? hkey + sl : reinterpret_cast<unsigned*>(dummy) + lane, 1u << 15);
        tail += __popc(wm);
        head += nb;
        ++n_steps;
        __syncwarp();
        if (RING && list_off + tail - flush_pos >= kFlush) flush_to(flush_pos + kFlush);
      }
      same = __reduce_add_sync(kFull, same);
      if (lane == 0) hkey[bslot] -= static_cast<unsigned>(same) << 15;
      remaining -= tail;
      if (prof) { const long long t = clock64(); t_bfs += t - t_mark; t_mark = t; }
      if (static_cast<unsigned long long>(tail) < th.min_cells_activated) { __syncwarp(); continue; }  // :329-331
      DPX_CHECK(list_off + tail <= C && remaining >= 0);
      if (n_regions < g.plane_cap) {
        if (RING) {
          flush_to(list_off + tail);
          __threadfence_block();
          __syncwarp();
        }
        // publish the region to the accumulating warps
        if (lane == 0) {
          float* rec = rec_ptr(n_regions);
          rec[kSegOff] = __int_as_float(list_off);
          rec[kSegCnt] = __int_as_float(tail);
          rec[kSegN] = __int_as_float(seed);  // the consumer replaces it by the point count
          __threadfence_block();
          misc[0] = n_regions + 1;
        }
        ++n_regions;
        list_off += tail;
      }
      __syncwarp();
    }
    if (lane == 0) {
      __threadfence_block();
      misc[1] = 1;
    }
  } else {
    // ================= consumers: region moments in FIFO order (plane_extractor.cpp:318-327) ================
    float* stage = stage_all + (warp - 1) * 32 * 12;
    for (int r = warp - 1;; r += kCtaWarps - 1) {
      while (misc[0] <= r && misc[1] == 0) __nanosleep(2000);
      if (misc[0] <= r) {
        // growing may have finished between the two reads: look once more
        __threadfence_block();
        if (misc[0] <= r) break;
      }
      __threadfence_block();
      float* rec = rec_ptr(r);
      const int off = __float_as_int(*reinterpret_cast<volatile float*>(rec + kSegOff));
      const int cnt = __float_as_int(*reinterpret_cast<volatile float*>(rec + kSegCnt));
      const int seed = __float_as_int(*reinterpret_cast<volatile float*>(rec + kSegN));
      const volatile int32_t* q = list + off;
      // seed first and twice: the candidate starts as a copy of the seed's stats, then every activated cell,
      // the seed included, is added (plane_extractor.cpp:318-323)
      float acc = 0.f;
      if (lane < 9) acc = __ldg(reinterpret_cast<const float*>(rec_b4) + 12 * seed + lane);
      float4 ra[3], rb[3];
      auto fetch = [&](int i0, float4 (&rr)[3]) {
        if (i0 < cnt) {
          const int c = q[min(i0 + lane, cnt - 1)] & 0xffffff;
          rr[0] = __ldg(rec_b4 + 3 * c); rr[1] = __ldg(rec_b4 + 3 * c + 1); rr[2] = __ldg(rec_b4 + 3 * c + 2);
        }
      };
      fetch(0, ra);
      fetch(32, rb);
      for (int i0 = 0; i0 < cnt; i0 += 32) {
        const int n = min(32, cnt - i0);
        float4* st4 = reinterpret_cast<float4*>(stage + 12 * lane);
        st4[0] = ra[0]; st4[1] = ra[1]; st4[2] = ra[2];
        ra[0] = rb[0]; ra[1] = rb[1]; ra[2] = rb[2];
        fetch(i0 + 64, rb);
        __syncwarp();
        if (lane < 9) {
          if (n == 32) {
#pragma unroll
            for (int k = 0; k < 32; ++k) acc = __fadd_rn(acc, stage[12 * k + lane]);
          } else {
            for (int k = 0; k < n; ++k) acc = __fadd_rn(acc, stage[12 * k + lane]);
          }
        }
        __syncwarp();
      }
      if (lane < 9) rec[kSegS + lane] = acc;
      if (lane == 9) rec[kSegN] = __int_as_float(cell_pts * (cnt + 1));
    }
  }
  __syncthreads();
  const long long t_grow_end = prof ? clock64() : 0;
  const int n_regions = misc[0];

  // ---- labels_map_ starts at zero; cw becomes the per-cell segment label (<= 65535, fits either word size) ------
  for (int c = tid; c < C; c += kCtaThreads) cw[c] = 0;

  // ---- plane fit of every grown region, one thread per region (plane_extractor.cpp:333-343) -----------
  int nseg = 0;
  for (int base = 0; base < n_regions; base += kCtaThreads) {
    const int r = base + tid;
    Moments mom;
    PlaneFit fit;
    int off = 0, cnt = 0;
    bool accept = false;
    if (r < n_regions) {
      const float* rec = rec_ptr(r);
      mom.n = __float_as_int(rec[kSegN]);
#pragma unroll
      for (int i = 0; i < 3; ++i) mom.s[i] = rec[kSegS + i];
#pragma unroll
      for (int i = 0; i < 6; ++i) mom.v[i] = rec[kSegV + i];
      off = __float_as_int(rec[kSegOff]);
      cnt = __float_as_int(rec[kSegCnt]);
      fit_plane_call(mom, fit);
      accept = fit.score > th.min_region_planarity_score;  // strict (plane_extractor.cpp:336)
    }
    const unsigned am = __ballot_sync(kFull, accept);
    if (lane == 0) misc[8 + warp] = __popc(am);
    __syncthreads();  // every record of this chunk has been read; warp counts are visible
    int before = nseg;
    for (int w = 0; w < warp; ++w) before += misc[8 + w];
    int total = 0;
    for (int w = 0; w < kCtaWarps; ++w) total += misc[8 + w];
    if (accept) {
      const int id = before + __popc(am & ((1u << lane) - 1u));  // 0-based segment index, in seed order
      float* rec = rec_ptr(id);
      rec[kSegN] = __int_as_float(mom.n);
#pragma unroll
      for (int i = 0; i < 3; ++i) rec[kSegS + i] = mom.s[i];
#pragma unroll
      for (int i = 0; i < 6; ++i) rec[kSegV + i] = mom.v[i];
#pragma unroll
      for (int i = 0; i < 3; ++i) rec[kSegMean + i] = fit.mean[i];
#pragma unroll
      for (int i = 0; i < 3; ++i) rec[kSegNormal + i] = fit.normal[i];
      rec[kSegD] = fit.d;
      rec[kSegMse] = fit.mse;
      rec[kSegScore] = fit.score;
      rec[kSegOff] = __int_as_float(off);
      rec[kSegCnt] = __int_as_float(cnt);
      rec[kSegDirty] = 0.f;
      merge[id] = id;
    }
    nseg += total;
    __syncthreads();
  }
  if (tid == 0) args.tables.n_planes[frame] = nseg;
  __threadfence_block();
  __syncthreads();
  // paint labels_map_ (plane_extractor.cpp:339-342): one warp per segment
  for (int s = warp; s < nseg; s += kCtaWarps) {
    const float* rec = rec_ptr(s);
    const int off = __float_as_int(rec[kSegOff]), cnt = __float_as_int(rec[kSegCnt]);
    // (the list of a large frame is in global memory: four independent loads per round)
    for (int i = lane; i < cnt; i += 128) {
      int c4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) c4[u] = i + 32 * u < cnt ? (list[off + i + 32 * u] & 0xffffff) : -1;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (c4[u] >= 0) {
          cw[c4[u]] = static_cast<word_t>(s + 1);
          seg_label[c4[u]] = s + 1;
        }
    }
  }
  __syncthreads();
  const long long t_fit_end = prof ? clock64() : 0;

  // ---- getConnectedComponents (plane_extractor.cpp:430-453): upper-triangle adjacency bits --------------
  const int words = (nseg + 31) / 32;
  // [nseg][words]: reuses the member runs (mode 0) or the `pairs` scratch of this frame (the sort is long over)
  unsigned* adj = ALL_SMEM ? reinterpret_cast<unsigned*>(members) : args.tables.pairs + 2 * fc;
  const bool adj_fits = static_cast<long long>(nseg) * words * 4 <= plan.adj_bytes;
  unsigned* rowbits = adj;  // fallback: one row at a time
  const int limit = (nv - 1) * nh;
  if (nseg > 1 && adj_fits) {
    for (int i = tid; i < nseg * words; i += kCtaThreads) adj[i] = 0;
    __syncthreads();
    for (int c = tid; c < limit; c += kCtaThreads) {
      const int qc = c % nh;
      const int id = static_cast<int>(cw[c]);
      if (qc < nh - 1 && id > 0) {
        const int right = static_cast<int>(cw[c + 1]), down = static_cast<int>(cw[c + nh]);
        if (right > 0 && right != id) {
          const int a = min(id, right) - 1, b = max(id, right) - 1;
          atomicOr(&adj[a * words + (b >> 5)], 1u << (b & 31));
        }
        if (down > 0 && down != id) {
          const int a = min(id, down) - 1, b = max(id, down) - 1;
          atomicOr(&adj[a * words + (b >> 5)], 1u << (b & 31));
        }
      }
    }
    __syncthreads();
  }

  // ---- findMergedLabels (plane_extractor.cpp:402-423) ---------------------------------------------------
  const double min_cos = static_cast<double>(th.min_cos_angle_merge);
  // The merge loop is sequential over the rows, but almost all of its work is not: row r tests plane a = merge[r] against
  // its adjacent planes t > r with a's normal as it was when the row started and with t's record, which nothing has
  // touched yet (t > r has not been anybody's target).  While a == r -- plane r has not been merged into an earlier one --
  // that normal is r's original one too, so those tests are evaluated for all rows up front by all threads (pass0);
  // the loop then only walks the bits that passed.  Rows whose plane was merged away earlier (a < r: a's record has been
  // refit since) are tested live, as the reference does.
  const bool pre_ok = nseg > 1 && adj_fits && 2ll * nseg * words * 4 <= plan.adj_bytes;
  unsigned* pass0 = adj + nseg * words;
  if (pre_ok) {
    for (int i = tid; i < nseg * words; i += kCtaThreads) {
      const int r = i / words, w = i - r * words;
      unsigned bits = adj[i], out = 0;
      if (bits) {
        const float* rr = rec_ptr(r);
        const float an0 = rr[kSegNormal], an1 = rr[kSegNormal + 1], an2 = rr[kSegNormal + 2], ad = rr[kSegD];
        while (bits) {
          const int b = __ffs(bits) - 1;
          bits &= bits - 1;
          const float* rt = rec_ptr(w * 32 + b);
          const double cos_angle = static_cast<double>(dot3(an0, an1, an2, rt[kSegNormal], rt[kSegNormal + 1], rt[kSegNormal + 2]));
          const float df = __fadd_rn(dot3(an0, an1, an2, rt[kSegMean], rt[kSegMean + 1], rt[kSegMean + 2]), ad);
          const double distance = __dmul_rn(static_cast<double>(df), static_cast<double>(df));
          if (cos_angle > min_cos && distance < static_cast<double>(th.max_merge_dist)) out |= 1u << b;
        }
      }
      pass0[i] = out;
    }
    __syncthreads();
  }
  if (nseg > 1) {
    for (int r = 0; r < nseg; ++r) {
      if (!adj_fits) {
        // pathological segment counts: build row r alone from the label map
        __syncthreads();
        for (int i = tid; i < words; i += kCtaThreads) rowbits[i] = 0;
        __syncthreads();
        for (int c = tid; c < limit; c += kCtaThreads) {
          const int qc = c % nh;
          const int id = static_cast<int>(cw[c]);
          if (qc < nh - 1 && id > 0) {
            const int right = static_cast<int>(cw[c + 1]), down = static_cast<int>(cw[c + nh]);
            if (right > 0 && right != id && min(id, right) - 1 == r) {
              const int b = max(id, right) - 1;
              atomicOr(&rowbits[b >> 5], 1u << (b & 31));
            }
            if (down > 0 && down != id && min(id, down) - 1 == r) {
              const int b = max(id, down) - 1;
              atomicOr(&rowbits[b >> 5], 1u << (b & 31));
            }
          }
        }
        __syncthreads();
      }
      if (warp != 0) continue;
      const int a = merge[r];
      const bool pretested = pre_ok && a == r;
      const unsigned* row = pretested ? pass0 + r * words : adj_fits ? adj + r * words : rowbits;
      bool any = false;
      for (int w = lane; w < words; w += 32) any |= row[w] != 0;
      if (!__any_sync(kFull, any)) continue;

      float* ra = rec_ptr(a);
      Moments ma;
      ma.n = __float_as_int(ra[kSegN]);
#pragma unroll
      for (int i = 0; i < 3; ++i) ma.s[i] = ra[kSegS + i];
#pragma unroll
      for (int i = 0; i < 6; ++i) ma.v[i] = ra[kSegV + i];
      // The reference refits plane a after every row that expanded it (:422).  The fit (an fp64 eigensolve on one lane) is
      // only ever looked at again if a later row is tested against a's normal -- i.e. if a later plane was merged into a
      // and has neighbours of its own -- so it is deferred: an expanded plane is marked dirty, refit here when its normal
      // is needed, and all planes still dirty after the loop are refit in parallel, one thread each.  Same final records.
      if (ra[kSegDirty] != 0.f) {
        PlaneFit fa;
        fit_plane_call(ma, fa);
        __syncwarp();
        if (lane == 0) {
#pragma unroll
          for (int i = 0; i < 3; ++i) ra[kSegMean + i] = fa.mean[i];
#pragma unroll
          for (int i = 0; i < 3; ++i) ra[kSegNormal + i] = fa.normal[i];
          ra[kSegD] = fa.d;
          ra[kSegMse] = fa.mse;
          ra[kSegScore] = fa.score;
          ra[kSegDirty] = 0.f;
        }
        __syncwarp();
      }
      // normal/d of `a` are the ones it had when the row started (stats are refit after the row)
      const float an0 = ra[kSegNormal], an1 = ra[kSegNormal + 1], an2 = ra[kSegNormal + 2], ad = ra[kSegD];
      bool expanded = false;
      for (int w = 0; w < words; ++w) {
        unsigned bits = row[w];
        while (bits) {
          const int t = w * 32 + __ffs(bits) - 1;
          bits &= bits - 1;
          const float* rt = rec_ptr(t);
          bool pass = pretested;
          if (!pretested) {
            const double cos_angle = static_cast<double>(dot3(an0, an1, an2, rt[kSegNormal], rt[kSegNormal + 1], rt[kSegNormal + 2]));
            const float df = __fadd_rn(dot3(an0, an1, an2, rt[kSegMean], rt[kSegMean + 1], rt[kSegMean + 2]), ad);
            const double distance = __dmul_rn(static_cast<double>(df), static_cast<double>(df));
            pass = cos_angle > min_cos && distance < static_cast<double>(th.max_merge_dist);
          }
          if (pass) {
            ma.n += __float_as_int(rt[kSegN]);
#pragma unroll
            for (int i = 0; i < 3; ++i) ma.s[i] = __fadd_rn(ma.s[i], rt[kSegS + i]);
#pragma unroll
            for (int i = 0; i < 6; ++i) ma.v[i] = __fadd_rn(ma.v[i], rt[kSegV + i]);
            if (lane == 0) merge[t] = a;
            expanded = true;
          }
        }
      }
      if (expanded && lane == 0) {
        ra[kSegN] = __int_as_float(ma.n);
#pragma unroll
        for (int i = 0; i < 3; ++i) ra[kSegS + i] = ma.s[i];
#pragma unroll
        for (int i = 0; i < 6; ++i) ra[kSegV + i] = ma.v[i];
        ra[kSegDirty] = 1.f;
      }
      __syncwarp();
    }
  }
  __threadfence_block();
  __syncthreads();
  for (int i = tid; i < nseg; i += kCtaThreads) {
    float* ri = rec_ptr(i);
    if (ri[kSegDirty] != 0.f) {
      Moments mi;
      mi.n = __float_as_int(ri[kSegN]);
#pragma unroll
      for (int k = 0; k < 3; ++k) mi.s[k] = ri[kSegS + k];
#pragma unroll
      for (int k = 0; k < 6; ++k) mi.v[k] = ri[kSegV + k];
      PlaneFit fi;
      fit_plane_call(mi, fi);
#pragma unroll
      for (int k = 0; k < 3; ++k) ri[kSegMean + k] = fi.mean[k];
#pragma unroll
      for (int k = 0; k < 3; ++k) ri[kSegNormal + k] = fi.normal[k];
      ri[kSegD] = fi.d;
      ri[kSegMse] = fi.mse;
      ri[kSegScore] = fi.score;
      ri[kSegDirty] = 0.f;
    }
  }
  __syncthreads();
  const long long t_merge_end = prof ? clock64() : 0;

  // ---- per-cell final labels (plane_extractor.cpp:464-465), plane records back to the global tables ---------
  int32_t* final_lab = list;  // [C] the queue storage is free now
  for (int c = tid; c < C; c += kCtaThreads) {
    const int l = static_cast<int>(cw[c]);
    const int fl = (l == 0) ? 0 : merge[l - 1] + 1;
    cell_label[c] = fl;
    final_lab[c] = fl;
  }
  for (int i = tid; i < nseg; i += kCtaThreads) merge_out[i] = merge[i];
  // toImageLabels (plane_extractor.cpp:455-470) fused: a frame's pixels are painted as soon as its own region growing is
  // done, so the 4 B/pixel write stream of the faster frames overlaps the chains of the slower ones.  The slowest
  // sixteenth of the batch (by finishing order) leaves its pixels to the labeling kernel that follows: one CTA writing
  // 1.2 MB alone at the very end would only lengthen the tail.
  if (args.labels != nullptr) {
    if (tid == 0) misc[2] = atomicAdd(&args.tables.paint_state[0], 1);
    __syncthreads();
  }
  const bool paint_here = args.labels != nullptr && misc[2] < args.n_frames - labeling_deferred_frames(args.n_frames);
  if (args.labels != nullptr && !paint_here && tid == 0) {
    const int slot = atomicAdd(&args.tables.paint_state[1], 1);
    args.tables.paint_state[2 + slot] = frame;
  }
  if (paint_here) {
    const int p = g.patch, width = g.width;
    int32_t* out = args.labels + static_cast<long long>(frame) * g.n_points;
    const int groups = (width + 3) / 4;  // groups of 4 consecutive pixels of one image row
    const bool vec = args.labels_vec_ok != 0;
    for (int item = tid; item < nv * groups; item += kCtaThreads) {
      const int cr = item / groups, gi = item - cr * groups;
      const int col = gi * 4;
      int cq = col / p, rem = col - cq * p;
      int lab[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        lab[i] = (col + i < width) ? final_lab[cr * nh + cq] : 0;
        if (++rem == p) { rem = 0; ++cq; }
      }
      int32_t* dst = out + static_cast<long long>(cr) * p * width + col;
      if (vec) {
        const int4 v4 = make_int4(lab[0], lab[1], lab[2], lab[3]);
        for (int r = 0; r < p; ++r)
          asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(dst + static_cast<long long>(r) * width),
                       "r"(v4.x), "r"(v4.y), "r"(v4.z), "r"(v4.w)
                       : "memory");
      } else {
        for (int r = 0; r < p; ++r)
          for (int i = 0; i < 4; ++i)
            if (col + i < width) dst[static_cast<long long>(r) * width + i] = lab[i];
      }
    }
  }
  {
    const int n_smem = min(nseg, plan.rec_cap);
    for (int i = tid; i < n_smem * kSegFloats; i += kCtaThreads) segs_g[i] = recs[i];
  }
  if (prof && tid == 0) {
    long long* o = args.prof + static_cast<long long>(frame) * kRegionProfSlots;
    const long long t_end = clock64();
    o[0] = t_end - t_kernel0;          // whole frame
    o[1] = t_init;                     // setup
    o[2] = t_seed;                     // bin argmax + seed search (all seeds)
    o[3] = t_bfs;                      // BFS (all seeds)
    o[4] = t_wide;                     // part of [3] spent in wide BFS steps (mode 1)
    o[5] = t_fit_end - t_grow_end;     // plane fits + label painting
    o[6] = t_merge_end - t_fit_end;    // adjacency + merging
    o[7] = t_end - t_merge_end;        // final labels
#ifdef DPX_SEED_PROBE
    o[7] = t_pick_probe;               // probe builds: cycles from the top of the seed loop to the seed being known
    o[6] = n_refills_probe;            // probe builds: trips to the L2-resident sorted keys
#endif
    o[8] = n_seeds;
    o[9] = n_steps;
    o[10] = n_regions;
    o[11] = nseg;
  }
}

// Shared-memory layout of the CTA kernel; bytes == 0 means "does not fit, try the next mode / the generic kernel".
inline CtaPlan region_grow_cta_plan_w(const Geometry& g, const Thresholds& th, int mode, int win);
inline CtaPlan region_grow_cta_plan(const Geometry& g, const Thresholds& th, int mode) {
  if (mode == 0) return region_grow_cta_plan_w(g, th, 0, 0);
  // the widest key window per bin that fits: a window serves that many seeds of a bin per trip to the L2
  for (int win = 32; win >= 8; win /= 2) {
    const CtaPlan p = region_grow_cta_plan_w(g, th, mode, win);
    if (p.bytes > 0) return p;
  }
  return CtaPlan{};
}
inline CtaPlan region_grow_cta_plan_w(const Geometry& g, const Thresholds& th, int mode, int win) {
  CtaPlan p{};
  auto align16 = [](size_t v) { return (v + 15) & ~static_cast<size_t>(15); };
  const size_t B2 = static_cast<size_t>(th.histogram_bins_per_coord) * th.histogram_bins_per_coord;
  const size_t C = static_cast<size_t>(g.n_cells);
  // small frames keep several CTAs per SM; large ones may take (almost) a whole SM's shared memory
  const size_t limit = (mode == 0 ? 100u : 226u) * 1024;
  size_t off = 0;
  p.off_stage = static_cast<int>(off);   off = align16(off + static_cast<size_t>(kCtaWarps - 1) * 32 * 12 * 4);
  p.off_hkey = static_cast<int>(off);    off = align16(off + (B2 + 4) * 4);
  p.off_binslot = static_cast<int>(off); off = align16(off + B2 * 2);
  p.off_cursor = static_cast<int>(off);  off = align16(off + B2 * 4);
  p.off_binend = static_cast<int>(off);  off = align16(off + B2 * 4);
  p.off_cw = static_cast<int>(off);      off = align16(off + C * (mode == 3 ? 2 : 4));
  if (mode <= 1) {
    p.off_list = static_cast<int>(off);  off = align16(off + (C > B2 ? C : B2) * 4);
  } else {
    p.off_ring = static_cast<int>(off);  off = align16(off + kRing * 4);
  }
  if (mode == 0) {
    p.off_keys = static_cast<int>(off);  off = align16(off + C * 8);  // member runs: cell ids + their MSE
  } else {
    p.win = win;
    p.off_win = static_cast<int>(off);   off = align16(off + B2 * static_cast<size_t>(win) * 8);
    p.off_wpos = static_cast<int>(off);  off = align16(off + B2 * 4);
    p.off_wend = static_cast<int>(off);  off = align16(off + B2 * 4);
  }
  p.adj_bytes = static_cast<int>(C * 8);  // mode 0: the member runs; else the `pairs` scratch of this frame
  p.rec_cap = g.plane_cap < 128 ? g.plane_cap : 128;
  p.off_recs = static_cast<int>(off);    off = align16(off + static_cast<size_t>(p.rec_cap) * kRecFloats * 4);
  p.off_misc = static_cast<int>(off);    off = align16(off + (32 + 32) * 4);
  // merge labels last: in shared memory when there is room, else the kernel works on the global merge table
  const size_t merge_bytes = static_cast<size_t>(g.plane_cap) * 4;
  if (mode == 0 || off + merge_bytes <= limit) {
    p.merge_smem = 1;
    p.off_merge = static_cast<int>(off);
    off = align16(off + merge_bytes);
  }
  p.bytes = off <= limit ? off : 0;
  // 16-bit cell words hold the bin id / histogram slot in 10 bits
  if (mode == 3 && B2 > 1024) p.bytes = 0;
  return p;
}

}  // namespace
}  // namespace dpx
