"""Refinement stage time against the batch size: which launch variant (refine.cu launch_refine) each size gets and how well
co-resident clusters fill each other's barriers.  python tools/refine_batch_scaling.py [sizes...]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deplex_b200 import Config, PlaneExtractor, synth, LAYOUT_ROWMAJOR
sizes = [int(a) for a in sys.argv[1:]] or [1, 4, 9, 10, 18, 19, 36, 37, 64, 74, 128]
h, w = 480, 640
base = torch.from_numpy(synth.make_batch(h, w, 0, 8, "rowmajor")).cuda()
cfg = Config(ransac_refinement=1, ransac_threshold=10.0, ransac_inliers_ratio=0.9, ransac_max_iterations=1000)
for F in sizes:
    batch = base.repeat((F + 7) // 8, 1, 1)[:F].contiguous()
    ex = PlaneExtractor(h, w, cfg, max_batch=F)
    for _ in range(2): ex.process_batch_device(batch, LAYOUT_ROWMAJOR)
    torch.cuda.synchronize(); ex.set_profiling(True); acc = 0.0
    for _ in range(3):
        ex.process_batch_device(batch, LAYOUT_ROWMAJOR); torch.cuda.synchronize()
        acc += ex.stage_ms()["refine"] / 3
    print(f"{F:4d} frames: refine {acc:7.3f} ms  = {acc / F * 1e3:7.1f} us/frame")
    ex.close()
