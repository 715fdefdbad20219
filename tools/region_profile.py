"""SM-cycle counters of the region-growing phases for the frames of the bench batch (mean, max, slowest frames):
    python tools/region_profile.py   (from the repo root, on the GPU box)"""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import bench
from deplex_b200 import Config, PlaneExtractor, LAYOUT_ROWMAJOR
class A: pass
a = A(); a.unique=64; a.frames=256; a.height=480; a.width=640; a.layout='rowmajor'
host = bench.make_host_batch(a, 0)
d = torch.from_numpy(host).cuda()
ex = PlaneExtractor(480, 640, Config(), max_batch=256)
for _ in range(3): ex.process_batch_device(d, LAYOUT_ROWMAJOR)
torch.cuda.synchronize()
ex.set_profiling(True)
ex.process_batch_device(d, LAYOUT_ROWMAJOR); torch.cuda.synchronize()
print(ex.stage_ms())
P = [ex.region_profile(f) for f in range(256)]
keys = list(P[0].keys())
M = np.array([[p[k] for k in keys] for p in P], dtype=np.float64)
print(keys)
print('mean', np.round(M.mean(0)).astype(int).tolist())
print('max ', np.round(M.max(0)).astype(int).tolist())
order = np.argsort(-M[:,0])[:6]
for f in order: print('frame', f, M[f].astype(int).tolist())
