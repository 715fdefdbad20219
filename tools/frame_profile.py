"""Stage times and region-growing phase counters of one shipped frame (tum | icl) or a synthetic size.
    python tools/frame_profile.py icl            python tools/frame_profile.py synth 1080 1920 10"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deplex_b200 import Config, PlaneExtractor, synth, LAYOUT_ROWMAJOR
name = sys.argv[1] if len(sys.argv) > 1 else "icl"
if name == "synth":
    h, w, p = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    xyz, cfg = synth.make_cloud(h, w, 3), Config(patch_size=p)
else:
    g = os.path.join(ROOT, "tests", "golden")
    cfgname = {"tum": "TUM_fr3_long_val", "icl": "ICL_living_room"}[name]
    depth = np.load(os.path.join(g, f"{name}_depth.npz"))["depth"]
    K = np.loadtxt(os.path.join(g, cfgname + ".K"), dtype=np.float32)
    k = dict(fx=float(K[0, 0]), fy=float(K[1, 1]), cx=float(K[0, 2]), cy=float(K[1, 2]))
    xyz, cfg = synth.depth_to_cloud(depth, k, "rowmajor"), Config(os.path.join(g, cfgname + ".ini"))
    h, w = depth.shape
ex = PlaneExtractor(h, w, cfg)
d = torch.from_numpy(xyz).cuda()
for _ in range(5):
    ex.process_batch_device(d, LAYOUT_ROWMAJOR)
torch.cuda.synchronize()
ex.set_profiling(True)
acc = {}
for _ in range(10):
    ex.process_batch_device(d, LAYOUT_ROWMAJOR); torch.cuda.synchronize()
    for k_, v in ex.stage_ms().items():
        acc[k_] = acc.get(k_, 0) + v / 10
print(name, h, w, "patch", cfg.patch_size, "cells", ex.info.n_cells, {k_: round(v * 1e3, 1) for k_, v in acc.items()}, "us")
print(ex.region_profile(0))
