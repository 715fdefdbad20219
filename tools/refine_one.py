"""The TUM frame with the shipped ini + ransacRefinement=1, three device-resident calls: the short program behind the ncu
capture of the refinement kernel (profiles/r02f_refine_ncu.txt)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import frame_cloud
from deplex_b200 import Config, PlaneExtractor, LAYOUT_ROWMAJOR
xyz, ini = frame_cloud(sys.argv[1] if len(sys.argv) > 1 else "tum")
ex = PlaneExtractor(480, 640, Config(ini, ransac_refinement=1))
d = torch.from_numpy(xyz[None]).cuda()
for _ in range(3):
    lab = ex.process_batch_device(d, LAYOUT_ROWMAJOR)
torch.cuda.synchronize()
print("labelled pixels", int((lab != 0).sum()), "work", ex.refine_work(0))
