"""Stage times with RANSAC refinement on: shipped TUM / ICL frames (shipped ini + ransacRefinement=1) and a synthetic batch."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deplex_b200 import Config, PlaneExtractor, synth, LAYOUT_ROWMAJOR
def stages(ex, d, n=5):
    for _ in range(2): ex.process_batch_device(d, LAYOUT_ROWMAJOR)
    torch.cuda.synchronize(); ex.set_profiling(True); acc = {}
    for _ in range(n):
        ex.process_batch_device(d, LAYOUT_ROWMAJOR); torch.cuda.synchronize()
        for k, v in ex.stage_ms().items(): acc[k] = acc.get(k, 0) + v / n
    return {k: round(v, 3) for k, v in acc.items()}
g = os.path.join(ROOT, "tests", "golden")
for name, cfgname in (("tum", "TUM_fr3_long_val"), ("icl", "ICL_living_room")):
    depth = np.load(os.path.join(g, f"{name}_depth.npz"))["depth"]
    K = np.loadtxt(os.path.join(g, cfgname + ".K"), dtype=np.float32)
    k = dict(fx=float(K[0, 0]), fy=float(K[1, 1]), cx=float(K[0, 2]), cy=float(K[1, 2]))
    xyz = torch.from_numpy(synth.depth_to_cloud(depth, k, "rowmajor")).cuda()
    ex = PlaneExtractor(480, 640, Config(os.path.join(g, cfgname + ".ini"), ransac_refinement=1))
    print(name, "shipped ini + refinement, 1 frame (ms):", stages(ex, xyz))
for (h, w, F, kw) in ((480, 640, 64, dict(ransac_threshold=1.0, ransac_inliers_ratio=0.15, ransac_max_iterations=1000)),
                      (480, 640, 64, dict(ransac_threshold=10.0, ransac_inliers_ratio=0.9, ransac_max_iterations=1000)),
                      (1080, 1920, 8, dict(ransac_threshold=10.0, ransac_inliers_ratio=0.9, ransac_max_iterations=1000))):
    batch = torch.from_numpy(synth.make_batch(h, w, 0, min(F, 8), "rowmajor")).cuda().repeat(F // min(F, 8), 1, 1)
    ex = PlaneExtractor(h, w, Config(ransac_refinement=1, **kw), max_batch=F)
    print(f"{w}x{h} x{F}", kw, "(ms):", stages(ex, batch, 3))
