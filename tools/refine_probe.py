"""Probe build only (make -C deplex_b200/csrc NVFLAGS_EXTRA=-DDPX_REFINE_PROBE): where a refinement round's cycles go
(leader CTA, thread 0) on the TUM frame."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import frame_cloud
from deplex_b200 import Config, PlaneExtractor, _capi, LAYOUT_ROWMAJOR
lib = _capi.load()
for name in ("tum", "icl"):
    xyz, ini = frame_cloud(name)
    ex = PlaneExtractor(480, 640, Config(ini, ransac_refinement=1))
    d = torch.from_numpy(xyz).cuda()
    for _ in range(2):
        ex.process_batch_device(d, LAYOUT_ROWMAJOR)
    buf = (C.c_longlong * 24)()
    lib.dpx_debug_refine_probe(buf, 1)
    ex.process_batch_device(d, LAYOUT_ROWMAJOR)
    lib.dpx_debug_refine_probe(buf, 0)
    names = ["sampling", "pixels+models", "wait for the CTA", "flush + barrier 1", "evaluate", "barrier 2", "label setup",
             "first-round barrier", "scoring", "generator blocks", "group sampling + producer barriers", "round entry"]
    for who, off in (("producer thread 0", 0), ("first scoring thread", 12)):
        tot = sum(buf[off:off + 12])
        print(name, who, "cycles:", {n: int(v) for n, v in zip(names, buf[off:off + 12]) if v}, "total", tot, "= %.2f ms" % (tot / 1.965e6))
