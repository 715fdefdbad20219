"""Where a frame's GPU result first departs from the oracle's: per-cell tables, bins, region labels, planes, merge labels.
python tools/diag_frame.py <height> <width> <patch> <synthetic frame index> [field=value ...]"""
import os, sys, math
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
from deplex_b200 import Config, PlaneExtractor, synth, LAYOUT_ROWMAJOR
h, w, patch, index = (int(a) for a in sys.argv[1:5])
extra = {k: (float(v) if "." in v else int(v)) for k, v in (a.split("=") for a in sys.argv[5:])}
cfg = Config(patch_size=patch, **extra)
xyz = synth.make_batch(h, w, index, 1, "rowmajor")[0]
ex = PlaneExtractor(h, w, cfg)
got = ex.process_batch_device(torch.from_numpy(xyz[None]).cuda(), LAYOUT_ROWMAJOR).cpu().numpy()[0]
ref, dbg = oracle.process(h, w, oracle.OracleConfig(**cfg.as_dict()), xyz, debug=True)
cells, planes = ex.cells(0), ex.planes(0)
nh = w // patch
bits = lambda a: np.ascontiguousarray(a).view(np.uint32)
valid = dbg["cell_valid"].astype(bool)
print("pixels differing:", int((got != ref).sum()), " cells:", len(cells), " planes gpu/oracle:", len(planes), dbg["n_planes"])
for name, r in (("sum", dbg["cell_sum"]), ("mean", dbg["cell_mean"]), ("normal", dbg["cell_normal"]), ("d", dbg["cell_d"]),
                ("mse", dbg["cell_mse"]), ("score", dbg["cell_score"]), ("merge_tolerance", dbg["cell_tol"])):
    d = bits(cells[name]) != bits(r)
    d = d.reshape(d.shape[0], -1).any(axis=1) & valid
    print(f"  {name}: {int(d.sum())} cells differ in bits", np.nonzero(d)[0][:8])
print("  valid:", int((cells["valid"].astype(bool) != valid).sum()), " planar:", int((cells["planar"].astype(bool) != dbg["cell_planar"].astype(bool)).sum()),
      " bin:", int((cells["bin"] != dbg["cell_bin"]).sum()))
for c in np.nonzero(cells["bin"] != dbg["cell_bin"])[0][:5]:
    n = cells["normal"][c]
    print("    bin differs at cell", c, (c // nh, c % nh), "normal", n, "gpu", cells["bin"][c], "oracle", dbg["cell_bin"][c])
sd = np.nonzero(cells["seg_label"] != dbg["cell_seglabel"])[0]
print("  region labels: ", len(sd), "cells differ", [(int(c // nh), int(c % nh)) for c in sd[:8]], "gpu", cells["seg_label"][sd[:8]], "oracle", dbg["cell_seglabel"][sd[:8]])
if len(planes) == dbg["n_planes"]:
    for name, r in (("normal", dbg["plane_normal"]), ("d", dbg["plane_d"]), ("mean", dbg["plane_mean"]), ("mse", dbg["plane_mse"]), ("score", dbg["plane_score"])):
        d = bits(planes[name]) != bits(r)
        d = d.reshape(d.shape[0], -1).any(axis=1)
        print(f"  plane {name}: {int(d.sum())} differ in bits", np.nonzero(d)[0][:8])
    print("  plane n_points differ:", int((planes["n_points"] != dbg["plane_npts"]).sum()))
    md = np.nonzero(planes["merge_label"] != dbg["merge_labels"])[0]
    print("  merge labels differ for planes", md[:10], "gpu", planes["merge_label"][md[:10]], "oracle", dbg["merge_labels"][md[:10]])
    for i in md[:3]:
        for j in {int(planes["merge_label"][i]), int(dbg["merge_labels"][i])}:
            pn, qn = dbg["plane_normal"][i].astype(np.float64), dbg["plane_normal"][j].astype(np.float64)
            print(f"    plane {i} vs {j}: cos {float(np.float32(pn[0])*np.float32(qn[0]) + (np.float32(pn[1])*np.float32(qn[1]) + np.float32(pn[2])*np.float32(qn[2]))):.9f}",
                  "gpu normals", planes["normal"][i], planes["normal"][j], "oracle", dbg["plane_normal"][i], dbg["plane_normal"][j])
