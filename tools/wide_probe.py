"""Probe build only (make -C deplex_b200/csrc NVFLAGS_EXTRA=-DDPX_BFS_PROBE): where bfs_wide_step's cycles go on the ICL frame."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import frame_cloud
from deplex_b200 import Config, PlaneExtractor, _capi, LAYOUT_ROWMAJOR
lib = _capi.load()
xyz, ini = frame_cloud("icl")
ex = PlaneExtractor(480, 640, Config(ini))
d = torch.from_numpy(xyz).cuda()
for _ in range(3):
    ex.process_batch_device(d, LAYOUT_ROWMAJOR)
buf = (C.c_longlong * 8)()
lib.dpx_debug_wide_probe(buf, 1)
ex.set_profiling(True)
ex.process_batch_device(d, LAYOUT_ROWMAJOR)
lib.dpx_debug_wide_probe(buf, 0)
print("wide-step phases (cycles): read+probe %d, claims %d, prefix ballots %d, stores+atomics %d, tail syncwarp %d" % tuple(buf[:5]))
print(ex.region_profile(0))
