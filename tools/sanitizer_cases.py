"""Small set of hot-path invocations for compute-sanitizer (tools/gpu_sanitize.sh): the shipped TUM and ICL frames,
a 4-frame VGA batch (fused painting + labeling kernel), raw-depth input, 1080p with a fine grid (region growing with
its state in global memory) and RANSAC refinement (cluster kernel).  Every case is checked against the oracle so a
sanitizer run is also a parity run.  Usage: python tools/sanitizer_cases.py [case ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402  (checker only)
from conftest import frame_cloud, load_frame  # noqa: E402
from deplex_b200 import Config, PlaneExtractor, synth, LAYOUT_ROWMAJOR  # noqa: E402


def ocfg(cfg):
    return oracle.OracleConfig(**cfg.as_dict())


def check(name, got, ref):
    bad = int((np.asarray(got) != np.asarray(ref)).sum())
    print(f"{name}: {'ok' if bad == 0 else 'MISMATCH'} ({bad} labels differ)", flush=True)
    return bad == 0


def case_tum():
    xyz, ini = frame_cloud("tum")
    cfg = Config(ini)
    return check("tum", PlaneExtractor(480, 640, cfg).process(xyz), oracle.process(480, 640, ocfg(cfg), xyz))


def case_icl():
    xyz, ini = frame_cloud("icl")
    cfg = Config(ini)
    return check("icl", PlaneExtractor(480, 640, cfg).process(xyz), oracle.process(480, 640, ocfg(cfg), xyz))


def case_batch():
    cfg = Config()
    xyz = synth.make_batch(480, 640, 4242, 4, "rowmajor")
    got = PlaneExtractor(480, 640, cfg, max_batch=4).process_batch_host(xyz, LAYOUT_ROWMAJOR)
    ref = np.stack([oracle.process(480, 640, ocfg(cfg), xyz[i]) for i in range(4)])
    return check("vga batch of 4", got, ref)


def case_depth():
    depth, k, ini = load_frame("tum")
    cfg = Config(ini)
    got = PlaneExtractor(480, 640, cfg).process_depth_batch_host(depth[None], k)[0]
    xyz = synth.depth_to_cloud(depth, k, "rowmajor")
    return check("tum raw depth", got, oracle.process(480, 640, ocfg(cfg), xyz))


def case_fhd():
    cfg = Config(patch_size=8)
    xyz = synth.make_batch(1080, 1920, 900000, 1, "rowmajor")
    got = PlaneExtractor(1080, 1920, cfg).process(xyz[0])
    return check("1080p patch 8", got, oracle.process(1080, 1920, ocfg(cfg), xyz[0]))


def case_refine():
    xyz, ini = frame_cloud("tum")
    cfg = Config(ini, ransac_refinement=1)
    return check("tum refinement", PlaneExtractor(480, 640, cfg).process(xyz), oracle.process(480, 640, ocfg(cfg), xyz))


CASES = {"tum": case_tum, "icl": case_icl, "batch": case_batch, "depth": case_depth, "fhd": case_fhd,
         "refine": case_refine}

if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    ok = all([CASES[n]() for n in names])
    sys.exit(0 if ok else 1)
