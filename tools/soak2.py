"""Second one-off differential soak against the oracle (not part of the test suite): python tools/soak2.py
Covers what tools/soak.py does not: the column-major layout, the raw-depth entry points (fused back-projection)
and randomly drawn configs (histogram bins, candidate sizes, merge / planarity thresholds, refinement)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
from deplex_b200 import Config, PlaneExtractor, synth, LAYOUT_COLMAJOR

THREADS = os.cpu_count() or 1
SCALE = float(os.environ.get("SOAK_SCALE", "1"))  # fraction of the frame counts below (short re-runs)


def compare(tag, got, ref, first):
    bad = 0
    for f in range(len(ref)):
        if not np.array_equal(got[f], ref[f]):
            bad += 1
            print(f"  MISMATCH [{tag}] frame {first + f}: {int((got[f] != ref[f]).sum())} pixels", flush=True)
    return bad


def run_points(h, w, cfg, first, n, chunk):
    """column-major clouds through process_batch_host"""
    ex = PlaneExtractor(h, w, cfg, max_batch=chunk)
    ocfg = oracle.OracleConfig(**cfg.as_dict())
    bad = 0
    for f0 in range(0, n, chunk):
        nf = min(chunk, n - f0)
        batch = synth.make_batch(h, w, first + f0, nf, "colmajor")
        got = ex.process_batch_host(batch, LAYOUT_COLMAJOR)
        ref = oracle.process_batch(h, w, ocfg, batch, 0, THREADS)
        bad += compare("colmajor", got, ref, first + f0)
    return bad


def run_depth(h, w, cfg, first, n, chunk):
    """uint16 depth through process_depth_batch_host; the oracle sees toPointCloud's cloud"""
    ex = PlaneExtractor(h, w, cfg, max_batch=chunk)
    ocfg = oracle.OracleConfig(**cfg.as_dict())
    k = synth.intrinsics_for(h, w)
    bad = 0
    for f0 in range(0, n, chunk):
        nf = min(chunk, n - f0)
        depth = np.stack([synth.make_depth(h, w, first + f0 + f, k) for f in range(nf)])
        clouds = np.stack([synth.depth_to_cloud(depth[f], k, "rowmajor") for f in range(nf)])
        got = ex.process_depth_batch_host(depth, k)
        ref = oracle.process_batch(h, w, ocfg, clouds, 1, THREADS)
        bad += compare("depth16", got, ref, first + f0)
    return bad


def random_config(rng, refine):
    return Config(
        patch_size=int(rng.choice([5, 8, 10, 16, 20])),
        histogram_bins_per_coord=int(rng.choice([8, 12, 20, 32])),
        min_cos_angle_merge=float(rng.choice([0.9, 0.93, 0.97, 0.99])),
        max_merge_dist=float(rng.choice([100.0, 500.0, 2000.0])),
        min_region_growing_candidate_size=int(rng.choice([1, 3, 5, 9])),
        min_region_growing_cells_activated=int(rng.choice([2, 4, 8])),
        min_region_planarity_score=float(rng.choice([0.3, 0.55, 0.8])),
        depth_sigma_coeff=float(rng.choice([1.425e-6, 5e-7, 1e-5])),
        depth_sigma_margin=float(rng.choice([0.0, 10.0, 50.0])),
        min_pts_per_cell=int(rng.choice([1, 3, 10])),
        depth_discontinuity_threshold=float(rng.choice([10.0, 160.0, 1e4])),
        max_number_depth_discontinuity=int(rng.choice([0, 1, 3])),
        ransac_refinement=int(refine),
        ransac_max_iterations=int(rng.choice([16, 64, 200])),
        ransac_threshold=float(rng.choice([2.0, 6.0, 20.0])),
        ransac_inliers_ratio=float(rng.choice([0.3, 0.5, 0.8])),
    )


t0 = time.time()
total = 0
frames = 0
for name, fn, args in [
    ("VGA p10 colmajor x1000", run_points, (480, 640, Config(), 1100000, 1000, 100)),
    ("VGA p10 depth16 x1000", run_depth, (480, 640, Config(), 1200000, 1000, 100)),
    ("720p p10 depth16 x100", run_depth, (720, 1280, Config(), 1300000, 100, 50)),
    ("1080p p10 colmajor x24", run_points, (1080, 1920, Config(), 1400000, 24, 12)),
]:
    args = args[:4] + (max(args[5], int(args[4] * SCALE)) if SCALE < 1 else args[4], args[5])
    b = fn(*args)
    total += b
    frames += args[4]
    print(f"{name}: {b} mismatching frames  ({time.time() - t0:.0f} s)", flush=True)

rng = np.random.default_rng(20261018)
N_CONFIGS = max(4, int(40 * SCALE))
for i in range(N_CONFIGS):
    refine = i % 4 == 3
    cfg = random_config(rng, refine)
    n = 12 if refine else 40
    fn = run_depth if i % 2 else run_points
    b = fn(480, 640, cfg, 2000000 + 1000 * i, n, n)
    total += b
    frames += n
    if b:
        print("  config:", cfg.as_dict(), flush=True)
print(f"{N_CONFIGS} random configs: cumulative {total} mismatching frames  ({time.time() - t0:.0f} s)", flush=True)
print(f"TOTAL mismatching frames: {total} of {frames}")
sys.exit(1 if total else 0)
