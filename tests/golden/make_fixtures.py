"""Regenerates the fixtures in tests/golden/ (run in the build container, where /root/reference is mounted).

  tum_depth.npz / icl_depth.npz   raw uint16 depth of the reference's two shipped frames
                                  (data/tum/1341848230.910894.png, data/icl_nuim/0.png), decoded with cv2
  *.ini / *.K                      the reference's shipped configs and intrinsics (data/configs/)
  oracle_tum.npz / oracle_icl.npz  outputs of the CPU oracle (oracle/) on those frames: labels, per-cell
                                  table and plane list.  The reference itself cannot be built here (Eigen 3.4
                                  is fetched from the network by its build), so these are ORACLE outputs, pinned
                                  by the reference's own golden value max(labels)==34 on the TUM frame.
"""
import os
import shutil
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/data"

import oracle  # noqa: E402


def main():
    frames = {
        "tum": ("tum/1341848230.910894.png", "TUM_fr3_long_val"),
        "icl": ("icl_nuim/0.png", "ICL_living_room"),
    }
    for name, (png, cfgname) in frames.items():
        depth = cv2.imread(os.path.join(REF, png), cv2.IMREAD_UNCHANGED)
        assert depth.dtype == np.uint16 and depth.shape == (480, 640)
        np.savez_compressed(os.path.join(HERE, f"{name}_depth.npz"), depth=depth)
        for ext in (".ini", ".K"):
            shutil.copyfile(os.path.join(REF, "configs", cfgname + ext), os.path.join(HERE, cfgname + ext))
            os.chmod(os.path.join(HERE, cfgname + ext), 0o644)
        K = np.loadtxt(os.path.join(HERE, cfgname + ".K"), dtype=np.float32)
        cfg = oracle.load_ini(os.path.join(HERE, cfgname + ".ini"))
        xyz = oracle.depth_to_cloud(depth, K[0, 0], K[1, 1], K[0, 2], K[1, 2])
        labels, dbg = oracle.process(480, 640, cfg, xyz, debug=True)
        keep = {k: v for k, v in dbg.items() if isinstance(v, np.ndarray)}
        keep["labels"] = labels
        keep["n_seeds"] = np.int32(dbg["n_seeds"])
        keep["n_ql_fallback"] = np.int32(dbg["n_ql_fallback"])
        np.savez_compressed(os.path.join(HERE, f"oracle_{name}.npz"), **keep)
        print(name, "max label", labels.max(), "planes", dbg["n_planes"], "seeds", dbg["n_seeds"])
    # default-config run on the TUM frame: the reference's own golden value (test_plane_extractor.cpp:27-33)
    depth = np.load(os.path.join(HERE, "tum_depth.npz"))["depth"]
    K = np.loadtxt(os.path.join(HERE, "TUM_fr3_long_val.K"), dtype=np.float32)
    xyz = oracle.depth_to_cloud(depth, K[0, 0], K[1, 1], K[0, 2], K[1, 2])
    assert oracle.process(480, 640, oracle.OracleConfig(), xyz).max() == 34


if __name__ == "__main__":
    main()
