"""Differential soak on fine cell grids (many thousands of cells per frame: the large-frame storage modes of region growing,
many axis-aligned normals per frame) against the oracle.  python tools/soak_fine.py [scale]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
from deplex_b200 import Config, PlaneExtractor, synth, LAYOUT_ROWMAJOR
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
t0 = time.time()
total = frames = 0
for name, (h, w, cfg, first, n, chunk) in [
    ("720p p4", (720, 1280, Config(patch_size=4), 7_000_000, 180, 9)),
    ("720p p5", (720, 1280, Config(patch_size=5), 7_100_000, 120, 12)),
    ("720p p8 minCos 0.98", (720, 1280, Config(patch_size=8, min_cos_angle_merge=0.98), 7_200_000, 120, 12)),
    ("1080p p8", (1080, 1920, Config(patch_size=8), 7_300_000, 48, 8)),
    ("1080p p5", (1080, 1920, Config(patch_size=5), 7_400_000, 24, 4)),
    ("1080p p4 bins 16", (1080, 1920, Config(patch_size=4, histogram_bins_per_coord=16), 7_500_000, 16, 4)),
    ("VGA p4 bins 21", (480, 640, Config(patch_size=4, histogram_bins_per_coord=21), 7_600_000, 200, 50)),
    ("VGA p5 bins 17", (480, 640, Config(patch_size=5, histogram_bins_per_coord=17), 7_700_000, 200, 50)),
]:
    n = max(chunk, int(n * scale) // chunk * chunk)
    ex = PlaneExtractor(h, w, cfg, max_batch=chunk)
    ocfg = oracle.OracleConfig(**cfg.as_dict())
    bad = 0
    for f0 in range(0, n, chunk):
        batch = synth.make_batch(h, w, first + f0, chunk, "rowmajor")
        got = ex.process_batch_host(batch, LAYOUT_ROWMAJOR)
        ref = oracle.process_batch(h, w, ocfg, batch, 1, os.cpu_count() or 1)
        for f in range(chunk):
            if not np.array_equal(got[f], ref[f]):
                bad += 1
                print("  MISMATCH frame", first + f0 + f, int((got[f] != ref[f]).sum()), "pixels", flush=True)
    ex.close()
    total += bad
    frames += n
    print(f"{name} x{n}: {bad} mismatching frames  ({time.time() - t0:.0f} s)", flush=True)
print(f"TOTAL mismatching frames: {total} of {frames}")
sys.exit(1 if total else 0)
