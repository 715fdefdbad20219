"""Differential soak of the refinement stage against the oracle (not part of the test suite): random configurations
(patch size, thresholds, iteration caps that are not multiples of the round size, tiny minimum region sizes), batch sizes that
reach every launch variant of refine.cu, both libstdc++ sampling generations.  python tools/soak_refine.py [cases] [seed]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
from deplex_b200 import Config, PlaneExtractor, synth, LAYOUT_ROWMAJOR
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 20261018)
sizes = [(480, 640), (480, 640), (240, 320), (360, 480), (720, 1280)]
t0 = time.time()
bad = frames = 0
for case in range(n_cases):
    h, w = sizes[rng.integers(len(sizes))]
    patch = int(rng.choice([q for q in (4, 5, 6, 8, 10, 12, 16, 20) if h % q == 0 and w % q == 0]))
    tiny = rng.random() < 0.35
    cfg = Config(patch_size=patch, ransac_refinement=1,
                 ransac_threshold=float(rng.choice([0.5, 1.0, 2.0, 4.0, 10.0, 25.0])),
                 ransac_inliers_ratio=float(rng.choice([0.05, 0.15, 0.5, 0.8, 0.95, 1.0])),
                 ransac_max_iterations=int(rng.choice([0, 1, 31, 32, 100, 127, 128, 129, 300, 1000, 1500])),
                 min_region_growing_candidate_size=1 if tiny else 5,
                 min_region_growing_cells_activated=int(rng.choice([1, 2])) if tiny else 4)
    F = int(rng.choice([1, 3, 9, 12, 18, 24]))
    if h >= 720: F = min(F, 9)
    variant = int(rng.integers(2))
    first = int(rng.integers(1 << 20))
    batch = synth.make_batch(h, w, first, F, "rowmajor")
    ex = PlaneExtractor(h, w, cfg, max_batch=F)
    ex.set_rng_compat("libstdc++10" if variant else "libstdc++11")
    got = ex.process_batch_host(batch, LAYOUT_ROWMAJOR)
    ex.close()
    oracle.set_uniform_int_variant(variant)
    ref = oracle.process_batch(h, w, oracle.OracleConfig(**cfg.as_dict()), batch, 1, os.cpu_count() or 1)
    oracle.set_uniform_int_variant(0)
    nb = sum(not np.array_equal(got[f], ref[f]) for f in range(F))
    bad += nb
    frames += F
    print(f"case {case:3d}: {w}x{h} p{patch} F={F:2d} iters={cfg.ransac_max_iterations:4d} ratio={cfg.ransac_inliers_ratio:.2f} thr={cfg.ransac_threshold:4.1f} "
          f"tiny={int(tiny)} mapping={variant} labels<= {int(ref.max()):4d}  {'MISMATCH in %d frames' % nb if nb else 'ok'}", flush=True)
print(f"{frames} frames in {n_cases} cases, {bad} differing frames, {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
