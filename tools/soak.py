"""One-off large differential soak against the oracle (not part of the test suite): python tools/soak.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
from deplex_b200 import Config, PlaneExtractor, synth, LAYOUT_ROWMAJOR
def run(h, w, cfg, first, n, chunk):
    ex = PlaneExtractor(h, w, cfg, max_batch=chunk)
    ocfg = oracle.OracleConfig(**cfg.as_dict())
    bad = 0
    for f0 in range(0, n, chunk):
        nf = min(chunk, n - f0)
        batch = synth.make_batch(h, w, first + f0, nf, "rowmajor")
        got = ex.process_batch_host(batch, LAYOUT_ROWMAJOR)
        ref = oracle.process_batch(h, w, ocfg, batch, 1, os.cpu_count() or 1)
        for f in range(nf):
            if not np.array_equal(got[f], ref[f]):
                bad += 1
                print("  MISMATCH frame", first + f0 + f, int((got[f] != ref[f]).sum()), "pixels")
    return bad
t0 = time.time()
total = 0
for name, args in [
    ("VGA p10 default x1500", (480, 640, Config(), 100000, 1500, 100)),
    ("VGA p10 refine x150", (480, 640, Config(ransac_refinement=1, ransac_threshold=6.0, ransac_inliers_ratio=0.5, ransac_max_iterations=64), 200000, 150, 50)),
    ("VGA p8 minCos 0.97 x300", (480, 640, Config(patch_size=8, min_cos_angle_merge=0.97), 300000, 300, 100)),
    ("VGA p4 x60", (480, 640, Config(patch_size=4), 400000, 60, 30)),
    ("720p p10 x100", (720, 1280, Config(), 500000, 100, 50)),
    ("1080p p10 x24", (1080, 1920, Config(), 600000, 24, 12)),
]:
    b = run(*args)
    total += b
    print(f"{name}: {b} mismatching frames  ({time.time() - t0:.0f} s)", flush=True)
print("TOTAL mismatching frames:", total)
sys.exit(1 if total else 0)
