#!/usr/bin/env python
"""bench.py -- frames/sec of the plane-extraction hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference CPU path (oracle port / reference build)

A "step" is one pass of the hot path over one batch of 256 synthetic 640x480 organized clouds
(BASELINE.json configs[2]).  `value` is device-resident throughput (inputs already in HBM); `e2e` is the
same metric through the host-pointer C-ABI call with pinned host buffers, H2D and D2H inside the timed
region.  One process per GPU; frames are independent, so ranks shard frames with no data-path collective
(weak scaling: every rank processes its own 256-frame batch per step).

The other BASELINE configs ride along in the same JSON line, each with the CPU path timed beside it:
  latency            configs[0] / configs[1]: the shipped TUM and ICL frames, one process() call at a time
  fhd_stress         configs[3]: 1920x1080 at patchSize 10, 8 and 5
  sharded_sequence   configs[4]: a 100 000-frame sequence in contiguous ranges over the ranks + the final gather (NCCL)
  e2e.pcie_ceiling   the e2e copy pattern with no kernels: what the box can move
A failure in any leg fails the run.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "frames/sec @640x480 (batched, 256 frames/batch)"
UNIT = "frames/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=256, help="frames per batch (per GPU)")
    ap.add_argument("--unique", type=int, default=64, help="distinct synthetic frames generated per rank, tiled to --frames")
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--patch", type=int, default=10)
    ap.add_argument("--layout", default="rowmajor", choices=["rowmajor", "colmajor"])
    ap.add_argument("--lanes", type=int, default=3,
                    help="extractors (each on its own stream) the device-resident leg feeds round-robin; 1 = strictly serial steps")
    ap.add_argument("--cpu-sample-frames", type=int, default=0, help="frames in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-latency", action="store_true", help="skip the single-frame latency legs (configs[0], configs[1])")
    ap.add_argument("--no-fhd", action="store_true", help="skip the 1920x1080 leg (configs[3])")
    ap.add_argument("--no-sequence", action="store_true", help="skip the sharded 100k-frame sequence leg (configs[4])")
    ap.add_argument("--seq-frames", type=int, default=100000, help="frames in the sharded sequence (configs[4])")
    ap.add_argument("--latency-calls", type=int, default=300)
    return ap.parse_args()


def workload_config(a, extra=None):
    cfg = {
        "workload": f"configs[2]: batched synthetic piecewise-planar {a.width}x{a.height} organized clouds, "
                    f"{a.frames} frames/batch, default Config (patchSize={a.patch}), noise + holes",
        "frames_per_batch": a.frames, "height": a.height, "width": a.width, "patch_size": a.patch,
        "layout": a.layout, "unique_frames": min(a.unique, a.frames),
        "l2": "inputs larger than L2 (batch input %.0f MB vs 126 MB L2)" % (a.frames * a.height * a.width * 12 / 1e6),
        "parallelism": "frame-sharded, one process per GPU, no data-path collective",
        "lanes": max(1, getattr(a, "lanes", 1)),
    }
    if extra:
        cfg.update(extra)
    return cfg


def make_host_batch(a, rank, with_depth=False):
    """`frames` synthetic organized clouds (the first `unique` generated, then tiled); optionally also the raw
    uint16 depth images they were back-projected from.  deplex_b200.synth is plain numpy (workload generation)."""
    from deplex_b200 import synth
    uniq = min(a.unique, a.frames)
    k = synth.intrinsics_for(a.height, a.width)
    depth = np.stack([synth.make_depth(a.height, a.width, rank * 100000 + i, k) for i in range(uniq)])
    base = np.stack([synth.depth_to_cloud(depth[i], k, a.layout) for i in range(uniq)])
    reps = (a.frames + uniq - 1) // uniq
    clouds = np.concatenate([base] * reps, axis=0)[: a.frames]
    if with_depth:
        return clouds, np.concatenate([depth] * reps, axis=0)[: a.frames], k
    return clouds


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the GPU runs the benchmark's steps
    (B200_PROFILING.md recipe).  Samples are time-stamped on arrival and selected by window."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.samples = []  # (arrival time, line)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def count(self, t0, t1):
        return sum(1 for t, _ in self.samples if t0 <= t <= t1)

    def stop(self, t0, t1, window):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, ln in self.samples:
            if not (t0 <= t <= t1):
                continue
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for n, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "window": window}


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


def stage_roofline(stage_ms, n_frames, n_px, n_cells, fused_frames, peak):
    """Per-stage algorithmic bytes over CUDA-event stage times."""
    alg = {
        "cell_stats": n_frames * n_px * 12,
        "region_grow": n_frames * n_cells * 82 + fused_frames * n_px * 4,
        "labeling": (n_frames - fused_frames) * n_px * 4,
    }
    out = {}
    for k, b in alg.items():
        ms = stage_ms.get(k, 0.0)
        gbs = b / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        out[k] = {"ms": ms, "algorithmic_bytes": b, "gbs": gbs, "frac": gbs / peak}
    return out


def fused_frames_of(ex, n_frames):
    """frames whose pixels the region-growing kernel paints itself (fused stage 3): all but max(2, F/16) per batch"""
    return (n_frames - min(n_frames, max(2, n_frames // 16))) if ex.info.fused_labeling else 0


def oracle_config_like(cfg):
    """OracleConfig with the same 16 fields as a deplex_b200.Config (checker side only)."""
    import oracle
    return oracle.OracleConfig(**cfg.as_dict())


def cpu_path(h, w, ocfg, batch, layout, threads, passes=1):
    """The reference CPU path on this box's host cores: oracle/_ref/libdeplex_ref.so (the reference's own sources) when
    it was built, else the oracle port (-O3 -DNDEBUG build).  Returns (frames/s, seconds, kind)."""
    import oracle
    if oracle.ref_available():
        run = lambda b: oracle.ref_process_batch(h, w, ocfg, b, layout, threads)  # noqa: E731
        kind = "reference"
    else:
        run = lambda b: oracle.process_batch(h, w, ocfg, b, layout, threads, timed_build=True)  # noqa: E731
        kind = "port"
    run(batch[: max(1, threads)])  # warm
    t0 = time.perf_counter()
    for _ in range(passes):
        run(batch)
    dt = time.perf_counter() - t0
    return passes * batch.shape[0] / dt, dt, kind


def latency_mode(dev, calls, cpu_base):
    """BASELINE.json configs[0] and configs[1]: the shipped TUM and ICL-NUIM frames (tests/golden fixtures), one
    process() call at a time through the host-pointer API (pinned buffers, H2D + kernels + D2H per call), the way
    examples/process_cloud.cpp:24-36 and process_sequence.cpp:30-53 time it: min / mean / max microseconds per frame;
    the same frame device-resident with per-stage CUDA-event times against the HBM roofline; and the CPU path (1 thread,
    the same frame, >= 200 calls) beside it."""
    import torch
    from deplex_b200 import Config, PlaneExtractor, synth, LAYOUT_ROWMAJOR
    golden = os.path.join(ROOT, "tests", "golden")
    peak, _ = hbm_peak()
    out = {}
    for name, cfgname in (("tum", "TUM_fr3_long_val"), ("icl", "ICL_living_room")):
        depth = np.load(os.path.join(golden, f"{name}_depth.npz"))["depth"]
        K = np.loadtxt(os.path.join(golden, cfgname + ".K"), dtype=np.float32)
        k = dict(fx=float(K[0, 0]), fy=float(K[1, 1]), cx=float(K[0, 2]), cy=float(K[1, 2]))
        cfg = Config(os.path.join(golden, cfgname + ".ini"))
        h, w = depth.shape
        cloud = synth.depth_to_cloud(depth, k, "rowmajor")
        xyz = torch.from_numpy(cloud).pin_memory()
        lab = torch.empty(h * w, dtype=torch.int32).pin_memory()
        ex = PlaneExtractor(h, w, cfg, device=dev.index)
        for _ in range(10):
            ex.process_batch_host_ptr(xyz.data_ptr(), 1, LAYOUT_ROWMAJOR, lab.data_ptr())
        ts = []
        for _ in range(calls):
            t0 = time.perf_counter()
            ex.process_batch_host_ptr(xyz.data_ptr(), 1, LAYOUT_ROWMAJOR, lab.data_ptr())
            ts.append(time.perf_counter() - t0)
        d_xyz = xyz.to(dev)
        d_lab = torch.empty(1, h * w, dtype=torch.int32, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(10):
            ex.process_batch_device(d_xyz, LAYOUT_ROWMAJOR, d_lab)
        e0.record()
        for _ in range(calls):
            ex.process_batch_device(d_xyz, LAYOUT_ROWMAJOR, d_lab)
        e1.record()
        torch.cuda.synchronize()
        dev_us = e0.elapsed_time(e1) * 1e3 / calls
        ex.set_profiling(True)
        acc, n_prof = {}, 30
        for _ in range(n_prof):
            ex.process_batch_device(d_xyz, LAYOUT_ROWMAJOR, d_lab)
            torch.cuda.synchronize()
            for kk, v in ex.stage_ms().items():
                acc[kk] = acc.get(kk, 0.0) + v / n_prof
        ex.set_profiling(False)
        us = np.array(ts) * 1e6
        n_px = h * w
        rec = {"frame": f"{cfgname} ({w}x{h}, patchSize={cfg.patch_size})", "calls": calls,
               "host_ptr_us": {"min": float(us.min()), "mean": float(us.mean()), "p99": float(np.percentile(us, 99)),
                               "max": float(us.max())},
               "host_ptr_fps": float(1e6 / us.mean()),
               "device_resident_us": float(dev_us), "device_resident_fps": float(1e6 / dev_us),
               "planes": int(lab.max().item()),
               "roofline": {"bound": "hbm", "peak": peak, "unit": "GB/s",
                            "pipeline": {"algorithmic_bytes": n_px * 16, "gbs": n_px * 16 / (dev_us * 1e-6) / 1e9,
                                         "frac": n_px * 16 / (dev_us * 1e-6) / 1e9 / peak},
                            "stages": stage_roofline(acc, 1, n_px, ex.info.n_cells, fused_frames_of(ex, 1), peak),
                            "note": "one frame cannot fill the GPU: region growing is one CTA on one SM (sequential by "
                                    "construction); the fractions say how far a single frame is from the bandwidth bound"}}
        if cpu_base:
            n_cpu = max(200, min(calls, 400))
            v, dt, kind = cpu_path(h, w, oracle_config_like(cfg), cloud[None], 1, 1, passes=n_cpu)
            rec["cpu_baseline"] = {"value": v, "unit": UNIT, "us_per_frame": 1e6 / v, "cores": 1, "kind": kind,
                                   "sample": f"{n_cpu} process() calls on this frame, 1 thread (methodology of "
                                             f"examples/process_cloud.cpp:24-36), {dt:.1f} s"}
            rec["speedup_vs_cpu"] = {"host_ptr": rec["host_ptr_fps"] / v, "device_resident": rec["device_resident_fps"] / v}
        ex.close()
        # the same frame with the shipped ini's RANSAC refinement switched on (plane_extractor.cpp:265-267, 472-509): stage 4
        rcfg = Config(os.path.join(golden, cfgname + ".ini"), ransac_refinement=1)
        rex = PlaneExtractor(h, w, rcfg, device=dev.index)
        for _ in range(3):
            rex.process_batch_device(d_xyz, LAYOUT_ROWMAJOR, d_lab)
        torch.cuda.synchronize()
        rex.set_profiling(True)
        racc, n_ref = {}, 10
        for _ in range(n_ref):
            rex.process_batch_device(d_xyz, LAYOUT_ROWMAJOR, d_lab)
            torch.cuda.synchronize()
            for kk, v in rex.stage_ms().items():
                racc[kk] = racc.get(kk, 0.0) + v / n_ref
        rex.set_profiling(False)
        labelled = int((d_lab != 0).sum().item())
        point_passes, rounds = rex.refine_work(0)
        rbytes = 12 * point_passes
        rrec = {"refine_ms": racc["refine"], "device_resident_us": sum(racc.values()) * 1e3,
                "labelled_pixels_after": labelled,
                "roofline": {"bound": "issue rate of the scoring loop / latency of the round pipeline (not HBM)",
                             "rounds": rounds, "point_passes": point_passes, "algorithmic_bytes": rbytes,
                             "gbs": rbytes / (racc["refine"] * 1e-3) / 1e9 if racc["refine"] > 0 else None,
                             "hypothesis_evaluations_per_s": 128 * point_passes / (racc["refine"] * 1e-3) if racc["refine"] > 0 else None,
                             "what": "12 B x (points of a label) x (rounds of 128 hypotheses scored on it), summed over the "
                                     "frame's labels (dpx_get_refine_work); the points come from L2 after the first round"},
                "note": "stage 4 is latency-bound (rounds of 128 hypotheses, prepared one round ahead by producer warps, two "
                        "cluster barriers each), not bandwidth-bound: each round re-reads the label's points (12 B each) from L2"}
        if cpu_base:
            v, dt, kind = cpu_path(h, w, oracle_config_like(rcfg), cloud[None], 1, 1, passes=30)
            rrec["cpu_baseline"] = {"value": v, "unit": UNIT, "us_per_frame": 1e6 / v, "cores": 1, "kind": kind,
                                    "sample": f"30 process() calls on this frame with ransacRefinement=1, 1 thread, {dt:.1f} s"}
            rrec["speedup_vs_cpu"] = (1e6 / rrec["device_resident_us"]) / v
        rex.close()
        rec["with_refinement"] = rrec
        out[name] = rec
    return out


def fhd_stress(dev, cpu_base, frames=8, steps=3):
    """BASELINE.json configs[3]: synthetic 1920x1080 clouds at the default patch and finer grids (patch 10 ->
    20 736 cells, patch 5 -> 82 944 cells): device-resident per-stage CUDA-event times against the HBM roofline, the
    host-pointer call (pinned buffers, copies in the timed region), and the CPU path on the same frames."""
    import torch
    from deplex_b200 import Config, PlaneExtractor, synth, LAYOUT_ROWMAJOR
    h, w = 1080, 1920
    n_px = h * w
    peak, _ = hbm_peak()
    base = synth.make_batch(h, w, 900000, 2, "rowmajor")
    host = np.concatenate([base] * (frames // 2), axis=0)
    pin_in = torch.from_numpy(host).pin_memory()
    pin_out = torch.empty((frames, n_px), dtype=torch.int32).pin_memory()
    d_xyz = pin_in.to(dev)
    d_lab = torch.empty((frames, n_px), dtype=torch.int32, device=dev)
    out = {}
    for patch in (10, 8, 5):
        cfg = Config(patch_size=patch)
        ex = PlaneExtractor(h, w, cfg, max_batch=frames, device=dev.index)
        for _ in range(2):
            ex.process_batch_device(d_xyz, LAYOUT_ROWMAJOR, d_lab)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            ex.process_batch_device(d_xyz, LAYOUT_ROWMAJOR, d_lab)
        e1.record()
        torch.cuda.synchronize()
        step_ms = e0.elapsed_time(e1) / steps
        ex.set_profiling(True)
        acc = {}
        for _ in range(steps):
            ex.process_batch_device(d_xyz, LAYOUT_ROWMAJOR, d_lab)
            torch.cuda.synchronize()
            for k, v in ex.stage_ms().items():
                acc[k] = acc.get(k, 0.0) + v / steps
        ex.set_profiling(False)
        ex.process_batch_host_ptr(pin_in.data_ptr(), frames, LAYOUT_ROWMAJOR, pin_out.data_ptr())
        t0 = time.perf_counter()
        for _ in range(steps):
            ex.process_batch_host_ptr(pin_in.data_ptr(), frames, LAYOUT_ROWMAJOR, pin_out.data_ptr())
        e2e_s = (time.perf_counter() - t0) / steps
        rec = {"cells": int(ex.info.n_cells), "frames": frames, "ms_per_step": step_ms,
               "frames_per_s": frames / (step_ms * 1e-3),
               "stage_ms": {k: round(v, 4) for k, v in acc.items()},
               "roofline": {"bound": "hbm", "peak": peak, "unit": "GB/s",
                            "pipeline": {"algorithmic_bytes": frames * n_px * 16,
                                         "frac": frames * n_px * 16 / (step_ms * 1e-3) / 1e9 / peak},
                            "stages": stage_roofline(acc, frames, n_px, ex.info.n_cells, fused_frames_of(ex, frames), peak)},
               "e2e": {"value": frames / e2e_s, "unit": UNIT, "h2d_bytes_per_step": frames * n_px * 12,
                       "d2h_bytes_per_step": frames * n_px * 4, "matches_device_path": bool(torch.equal(pin_out, d_lab.cpu()))}}
        if cpu_base:
            v, dt, kind = cpu_path(h, w, oracle_config_like(cfg), base, 1, 1)
            rec["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": kind,
                                   "sample": f"the 2 distinct frames of this batch, 1 thread, {dt:.1f} s"}
            rec["speedup_vs_cpu"] = {"device_resident": rec["frames_per_s"] / v, "e2e": rec["e2e"]["value"] / v}
        out[f"patch{patch}"] = rec
        ex.close()
    return out


def pcie_ceiling(dev, frames, h2d_bytes_per_frame, d2h_bytes_per_frame, h2d_chunk_frames, passes, barrier, all_reduce_max, world):
    """The copy pattern of the host-pointer path with NO kernels: chunked H2D of the inputs and D2H of the labels from /
    to pinned memory on two streams, all ranks at once.  What the box's PCIe links and host memory can move, so that
    `e2e` can be read as a fraction of it."""
    import torch
    h_in = torch.empty(frames * h2d_bytes_per_frame, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(frames * d2h_bytes_per_frame, dtype=torch.uint8).pin_memory()
    d_in = torch.empty_like(h_in, device=dev)
    d_out = torch.empty_like(h_out, device=dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def one_pass():
        for f0 in range(0, frames, h2d_chunk_frames):
            f1 = min(frames, f0 + h2d_chunk_frames)
            with torch.cuda.stream(s_in):
                d_in[f0 * h2d_bytes_per_frame:f1 * h2d_bytes_per_frame].copy_(h_in[f0 * h2d_bytes_per_frame:f1 * h2d_bytes_per_frame], non_blocking=True)
            with torch.cuda.stream(s_out):
                h_out[f0 * d2h_bytes_per_frame:f1 * d2h_bytes_per_frame].copy_(d_out[f0 * d2h_bytes_per_frame:f1 * d2h_bytes_per_frame], non_blocking=True)

    one_pass()
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(passes):
        one_pass()
    s_in.synchronize()
    s_out.synchronize()
    dt = all_reduce_max(time.perf_counter() - t0)
    return {"frames_per_s": world * passes * frames / dt,
            "h2d_gbs": world * passes * frames * h2d_bytes_per_frame / dt / 1e9,
            "d2h_gbs": world * passes * frames * d2h_bytes_per_frame / dt / 1e9,
            "what": f"copies only, {h2d_chunk_frames}-frame chunks, H2D and D2H concurrent, all {world} rank(s) at once"}


def sharded_sequence(a, dev, rank, world, lay, barrier, all_reduce_max):
    """BASELINE.json configs[4]: a frame-sharded synthetic 640x480 sequence of --seq-frames (100 000) frames.  Frame i of
    the sequence is synthetic frame i % U (U = --unique distinct frames, seeds 0..U-1, the same on every rank, resident in
    HBM -- 'frames generated up front on the device side'), so the result does not depend on the number of
    ranks.  Rank r processes the contiguous range sharding.frame_range(T, r, world) in 256-frame batches through the
    C-ABI pipeline (no data-path collective), then the labels are gathered to rank 0 point-to-point over NCCL, received
    in place into the (T, N) int32 result (sharding.gather_labels), verified there frame by frame and reduced to a
    checksum that must be the same for N = 1, 2, 4, 8."""
    import torch
    from deplex_b200 import Config, PipelinedExtractor, sharding, synth
    T, U, B = a.seq_frames, min(a.unique, a.frames), a.frames
    h, w = a.height, a.width
    n_px = h * w
    k = synth.intrinsics_for(h, w)
    uniq = np.stack([synth.depth_to_cloud(synth.make_depth(h, w, i, k), k, a.layout) for i in range(U)])
    reps = (B + U + U - 1) // U
    pool = torch.from_numpy(np.concatenate([uniq] * reps, axis=0)).to(dev)  # >= B + U frames, frame j = unique j % U
    begin, end = sharding.frame_range(T, rank, world)
    need = (T if rank == 0 else end - begin) * n_px * 4
    free, _ = torch.cuda.mem_get_info(dev)
    if need + (6 << 30) > free:
        raise RuntimeError(f"sharded_sequence: {need / 2**30:.0f} GiB of labels do not fit in {free / 2**30:.0f} GiB free HBM; "
                           f"lower --seq-frames")
    # rank 0 owns the whole result and works in place in its own slice of it; the others own their range only
    result = torch.empty((T, n_px), dtype=torch.int32, device=dev) if rank == 0 else None
    local = result[begin:end] if rank == 0 else torch.empty((end - begin, n_px), dtype=torch.int32, device=dev)
    cfg = Config(patch_size=a.patch)
    pipe = PipelinedExtractor(h, w, cfg, max_batch=B, device=dev.index, lanes=max(1, a.lanes))
    stream = torch.cuda.current_stream(dev)
    # the labels every unique frame must get, from one ordinary batch call
    base = pipe.lanes[0].process_batch_device(pool[:U].contiguous(), lay).clone()
    sharding.process_range_device(pipe, pool, U, begin, min(end, begin + 2 * B), lay, B, local, stream)  # warm-up
    pipe.join()
    barrier()
    launches0 = pipe.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sharding.process_range_device(pipe, pool, U, begin, end, lay, B, local, stream)
    pipe.join()
    e1.record(stream)
    barrier()
    ms = all_reduce_max(e0.elapsed_time(e1))
    launches = pipe.kernel_launches() - launches0
    if world > 1:
        # NCCL sets up its point-to-point channels on first use: one small exchange of the same pattern first
        warm = torch.zeros((world, 8), dtype=torch.int32, device=dev)
        sharding.gather_labels(warm[rank:rank + 1], world, dst=0, out=warm if rank == 0 else None)
        torch.cuda.synchronize()
        barrier()
    t0 = time.perf_counter()
    full = sharding.gather_labels(local, T, dst=0, out=result)
    torch.cuda.synchronize()
    barrier()
    gather_s = all_reduce_max(time.perf_counter() - t0)
    out = None
    if rank == 0:
        # verify: sequence frame i must carry the labels of unique frame i % U; checksum over the whole result
        checksum, mismatched = 0, 0
        rows = 16 * U
        for f0 in range(0, T, rows):
            blk = full[f0:min(T, f0 + rows)]
            idx = (torch.arange(f0, f0 + blk.shape[0], device=dev) % U)
            mismatched += int((blk != base[idx]).any(dim=1).sum().item())
            per_frame = blk.sum(dim=1, dtype=torch.int64)
            weight = (torch.arange(f0, f0 + blk.shape[0], device=dev, dtype=torch.int64) % 65521) + 1
            checksum = (checksum + int((per_frame * weight).sum().item())) % (1 << 61)
        out = {"workload": f"configs[4]: {T} synthetic {w}x{h} frames (frame i = unique frame i % {U}, resident in HBM), "
                           f"contiguous 1/{world} ranges per GPU, {B}-frame batches, {max(1, a.lanes)} lanes",
               "frames": T, "n_gpus": world, "ms": ms, "frames_per_s": T / (ms * 1e-3),
               "roofline_frac": T * n_px * 16 / (ms * 1e-3) / 1e9 / (world * hbm_peak()[0]),
               "gather": {"seconds": gather_s, "bytes": int((T - (end - begin)) * n_px * 4),
                          "gbs_into_rank0": (T - (end - begin)) * n_px * 4 / gather_s / 1e9 if world > 1 else None,
                          "how": "NCCL point-to-point, received in place into rank 0's (T, N) int32 result" if world > 1
                                 else "single rank: nothing to gather"},
               "frames_per_s_incl_gather": T / (ms * 1e-3 + gather_s),
               "frames_with_wrong_labels": mismatched, "checksum": checksum, "gpu_launches_rank0": launches}
        if mismatched:
            raise RuntimeError(f"sharded_sequence: {mismatched} frames differ from the single-batch labels")
    pipe.close()
    del result, local, full, pool
    torch.cuda.empty_cache()
    return out


def measured_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/traffic.json, written by
    tools/ncu_traffic.py from the same bench command); None when no capture is recorded."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return rec.get(kernel)
    except Exception:
        return None


def run_reference(a):
    """--impl reference: the reference's own CPU implementation of the path on all host threads: oracle/_ref's build of
    the reference sources when it exists (needs an Eigen tree, absent from this image), else the oracle port.  Nothing
    of deplex_b200's native code is loaded in this arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    threads = os.cpu_count() or 1
    host_batch = make_host_batch(a, 0)
    # a bounded sample per step, so that K steps + W warm-ups stay within ~20 s of wall clock at ~4 000 frames/s
    budget = max(threads, 80000 // max(1, a.steps + a.warmup))
    per_step = min(a.frames, max(threads, (budget // threads) * threads))
    per_step = min(per_step, host_batch.shape[0])
    cfg = oracle.OracleConfig(patch_size=a.patch)
    layout = 1 if a.layout == "rowmajor" else 0
    sample = host_batch[:per_step]
    if oracle.ref_available():
        kind, run = "reference", (lambda: oracle.ref_process_batch(a.height, a.width, cfg, sample, layout, threads))
    else:
        kind, run = "port", (lambda: oracle.process_batch(a.height, a.width, cfg, sample, layout, threads, timed_build=True))
    for _ in range(a.warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        run()
    dt = time.perf_counter() - t0
    value = a.steps * per_step / dt
    sample_desc = f"{per_step} frames/step of the same batch, {threads} threads, frame-parallel"
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(a),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample_desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def run_ours(a):
    import torch
    import torch.distributed as dist
    from deplex_b200 import Config, PipelinedExtractor, LAYOUT_COLMAJOR, LAYOUT_ROWMAJOR

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # host side of the e2e path: keep this rank's pinned buffers on the GPU's NUMA node
    from deplex_b200 import sharding
    numa_node = sharding.bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def all_reduce_max(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    lay = LAYOUT_ROWMAJOR if a.layout == "rowmajor" else LAYOUT_COLMAJOR
    host_np, depth_np, intr = make_host_batch(a, rank, with_depth=True)
    n_px = a.height * a.width
    pin_in = torch.empty(host_np.shape, dtype=torch.float32).pin_memory()
    pin_in.copy_(torch.from_numpy(host_np))
    pin_out = torch.empty((a.frames, n_px), dtype=torch.int32).pin_memory()
    d_xyz = pin_in.to(dev, non_blocking=False)
    d_lab = torch.empty((a.frames, n_px), dtype=torch.int32, device=dev)

    cfg = Config(patch_size=a.patch)
    # `lanes` extractors on their own streams, fed round-robin (the C-ABI's dpx_pipeline): the next batch's HBM-bound
    # cell-stats kernel fills the SMs the latency-bound region growing of the previous batch has already left
    pipe = PipelinedExtractor(a.height, a.width, cfg, max_batch=a.frames, device=local, lanes=max(1, a.lanes))
    ex = pipe.lanes[0]
    stream = torch.cuda.current_stream(dev)
    d_labs = [d_lab] + [torch.empty_like(d_lab) for _ in range(len(pipe.lanes) - 1)]

    def timed(run_step, n_steps):
        """K steps between two events on the current stream, barrier + synchronize on both sides"""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(n_steps):
            run_step(i)
        pipe.join()
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1)

    def lane_step(i):
        pipe.submit(d_xyz, lay, d_labs[i % len(d_labs)])

    def serial_step(i):
        ex.process_batch_device(d_xyz, lay, d_lab, stream)

    # ---- device-resident throughput ---------------------------------------------------------------
    for i in range(max(a.warmup, 3)):
        serial_step(i)
    torch.cuda.synchronize()
    # one extractor, one stream, steps strictly back to back: the number the per-stage roofline below decomposes
    single_ms = timed(serial_step, a.steps)
    for i in range(max(a.warmup, 3) * len(d_labs)):
        lane_step(i)
    pipe.synchronize()
    stage_acc = {}
    sampler = ClockSampler(local)
    sampler.start()
    for i in range(4):  # give nvidia-smi time to come up while the GPU is already under this load
        lane_step(i)
    pipe.synchronize()
    launches0 = pipe.kernel_launches()
    t_start = time.perf_counter()
    elapsed_ms = timed(lane_step, a.steps)
    t_end = time.perf_counter()
    launches = pipe.kernel_launches() - launches0
    window = "timed region"
    if sampler.count(t_start, t_end) < 3:
        # the timed region is shorter than a few sampler periods: keep the same step running for ~1.5 s
        # right after it (untimed) and read the clocks under that load
        t_start = time.perf_counter()
        while time.perf_counter() - t_start < 1.5:
            for i in range(6):
                lane_step(i)
            pipe.synchronize()
        t_end = time.perf_counter()
        window = "1.5 s of the same step run back to back right after the timed region (the timed region is shorter than the sampler period)"
    clocks = sampler.stop(t_start, t_end, window)
    # per-stage CUDA-event times (on the launching stream), averaged over a separate short run so the
    # event records do not sit inside the timed region above
    ex.set_profiling(True)
    n_prof = min(a.steps, 10)
    for _ in range(n_prof):
        ex.process_batch_device(d_xyz, lay, d_lab, stream)
        torch.cuda.synchronize()
        for k, v in ex.stage_ms().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v / n_prof
    ex.set_profiling(False)

    max_ms, single_max_ms = all_reduce_max(elapsed_ms), all_reduce_max(single_ms)
    value = world * a.steps * a.frames / (max_ms / 1e3)
    lanes_equal = all(bool(torch.equal(d_lab, other)) for other in d_labs[1:])
    if not lanes_equal:
        raise RuntimeError("lanes disagree: the labels depend on which extractor of the pipeline ran the batch")

    # ---- end to end: host pointers, pinned memory, H2D + D2H inside the timed region ------------------
    e2e = e2e_depth = None
    if not a.no_e2e:
        for _ in range(2):
            ex.process_batch_host_ptr(pin_in.data_ptr(), a.frames, lay, pin_out.data_ptr())
        k_e2e = max(3, min(a.steps, 10))
        barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            ex.process_batch_host_ptr(pin_in.data_ptr(), a.frames, lay, pin_out.data_ptr())
        torch.cuda.synchronize()
        dt = all_reduce_max(time.perf_counter() - t0)
        e2e = {"value": world * k_e2e * a.frames / dt, "unit": UNIT, "numa_node_rank0": numa_node,
               "h2d_bytes_per_step": int(a.frames * n_px * 12), "d2h_bytes_per_step": int(a.frames * n_px * 4),
               "steps": k_e2e, "api": "dpx_process_batch_host (pinned host buffers, tapered chunks, copy/compute overlap)",
               "labels_checksum": int(pin_out.view(-1)[:: 997].to(torch.int64).sum().item())}
        if not torch.equal(pin_out, d_lab.cpu()):
            raise RuntimeError("e2e: the host-pointer path and the device-resident path disagree")
        e2e["matches_device_path"] = True
        ceil = pcie_ceiling(dev, a.frames, n_px * 12, n_px * 4, 16, k_e2e, barrier, all_reduce_max, world)
        e2e["pcie_ceiling"] = ceil
        e2e["frac_of_pcie_ceiling"] = e2e["value"] / ceil["frames_per_s"]
        # the same workload entering as raw uint16 depth (toPointCloud fused on the device)
        pin_depth = torch.from_numpy(depth_np.view(np.int16)).pin_memory()
        pin_out2 = torch.empty_like(pin_out).pin_memory()
        for _ in range(2):
            ex.process_depth_batch_host_ptr(pin_depth.data_ptr(), a.frames, intr, pin_out2.data_ptr())
        barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            ex.process_depth_batch_host_ptr(pin_depth.data_ptr(), a.frames, intr, pin_out2.data_ptr())
        torch.cuda.synchronize()
        dt = all_reduce_max(time.perf_counter() - t0)
        if not torch.equal(pin_out2, pin_out):
            raise RuntimeError("e2e_depth16: the raw-depth path and the point path disagree")
        e2e_depth = {"value": world * k_e2e * a.frames / dt, "unit": UNIT,
                     "h2d_bytes_per_step": int(a.frames * n_px * 2), "d2h_bytes_per_step": int(a.frames * n_px * 4),
                     "api": "dpx_process_depth_batch_host (uint16 depth + intrinsics in, int32 labels out)",
                     "matches_point_path": True}
        ceil = pcie_ceiling(dev, a.frames, n_px * 2, n_px * 4, 34, k_e2e, barrier, all_reduce_max, world)
        e2e_depth["pcie_ceiling"] = ceil
        e2e_depth["frac_of_pcie_ceiling"] = e2e_depth["value"] / ceil["frames_per_s"]
        # ... and with the labels taken as uint16 (dpx_process_depth_batch_host_u16): 2 + 2 B/pixel over PCIe
        pin_out16 = torch.empty((a.frames, n_px), dtype=torch.int16).pin_memory()
        for _ in range(2):
            ex.process_depth_batch_host_ptr(pin_depth.data_ptr(), a.frames, intr, pin_out16.data_ptr(), labels_u16=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            ex.process_depth_batch_host_ptr(pin_depth.data_ptr(), a.frames, intr, pin_out16.data_ptr(), labels_u16=True)
        torch.cuda.synchronize()
        dt = all_reduce_max(time.perf_counter() - t0)
        if not torch.equal(pin_out16.to(torch.int32), pin_out):  # labels < 32768 here, so the int16 view is the value
            raise RuntimeError("e2e_depth16 (uint16 labels): differs from the int32 path")
        ceil16 = pcie_ceiling(dev, a.frames, n_px * 2, n_px * 2, 34, k_e2e, barrier, all_reduce_max, world)
        e2e_depth["labels_u16"] = {"value": world * k_e2e * a.frames / dt, "unit": UNIT,
                                   "h2d_bytes_per_step": int(a.frames * n_px * 2), "d2h_bytes_per_step": int(a.frames * n_px * 2),
                                   "api": "dpx_process_depth_batch_host_u16 (the caller takes uint16 labels: same values)",
                                   "matches_int32_path": True, "pcie_ceiling": ceil16,
                                   "frac_of_pcie_ceiling": world * k_e2e * a.frames / dt / ceil16["frames_per_s"]}

    # ---- configs[4]: the sharded 100k-frame sequence + final gather (all ranks) ------------------------------
    pipe.close()
    del d_labs
    seq = None
    if not a.no_sequence:
        seq = sharded_sequence(a, dev, rank, world, lay, barrier, all_reduce_max)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel -----------------------------------------------------------------
    peak, peak_src = hbm_peak()
    n_cells = ex.info.n_cells
    fused = fused_frames_of(ex, a.frames)
    stages = stage_roofline(stage_acc, a.frames, n_px, n_cells, fused, peak)
    dominant = max(stages, key=lambda k: stages[k]["ms"])
    step_ms = max_ms / a.steps
    pipeline_bytes = a.frames * n_px * 16  # 12 B/px read once + 4 B/px written once
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": stages[dominant]["gbs"], "peak": peak,
                "unit": "GB/s", "frac": stages[dominant]["frac"], "traffic": measured_traffic(dominant),
                "peak_source": peak_src, "stages": stages, "fused_label_frames": fused,
                "pipeline": {"algorithmic_bytes": pipeline_bytes, "gbs": pipeline_bytes / (step_ms * 1e-3) / 1e9,
                             "frac": pipeline_bytes / (step_ms * 1e-3) / 1e9 / peak,
                             "note": "whole step, 16 B/pixel/frame; the region-growing stage is latency-bound by construction; "
                                     "`stages` are timed on one lane running alone, `pipeline` on the overlapped lanes"}}
    for k in stages:
        stages[k]["traffic"] = measured_traffic(k)

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": max_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(a), "clocks": clocks, "gpu_launches": launches,
        "roofline": roofline, "latency_ms_per_frame_in_batch": max_ms / a.steps / a.frames,
        # the same K steps through ONE extractor on one stream (no overlap between consecutive batches): the step the
        # per-stage times in `roofline.stages` add up to, and the one ncu's serialised launch list reproduces
        "single_lane": {"value": world * a.steps * a.frames / (single_max_ms / 1e3), "unit": UNIT,
                        "ms_per_step": single_max_ms / a.steps},
        "lanes_agree": lanes_equal,
    }
    if e2e:
        out["e2e"] = e2e
        out["e2e_depth16"] = e2e_depth
    if seq:
        out["sharded_sequence"] = seq
    cpu_base = not a.no_cpu_baseline
    if world == 1 and not a.no_latency:
        out["latency"] = latency_mode(dev, a.latency_calls, cpu_base)
    if world == 1 and not a.no_fhd:
        out["fhd_stress"] = fhd_stress(dev, cpu_base)
    if world == 1 and cpu_base:
        n = a.cpu_sample_frames or a.frames
        reps = max(1, round(14 * (640 * 480 * 256) / (n_px * n)))  # ~10 s of single-thread work
        import oracle
        v, dt, kind = cpu_path(a.height, a.width, oracle.OracleConfig(patch_size=a.patch), host_np[:n],
                               1 if a.layout == "rowmajor" else 0, 1, passes=reps)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": kind,
                               "sample": f"{n * reps} frames ({n} frames of this batch x{reps} passes), 1 thread (the reference "
                                         f"is single-threaded by default), {dt:.1f} s; host has {os.cpu_count()} cores"}
    emit(out)
    if world > 1:
        dist.destroy_process_group()


_RESULT_FD = None


def emit(obj):
    """The one JSON line of this run, on the process's original stdout."""
    line = (json.dumps(obj) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, line)


def main():
    global _RESULT_FD
    a = parse_args()
    # libraries print to stdout too (NCCL's version banner under torchrun): keep fd 1 for the result line only
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
