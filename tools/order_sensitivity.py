"""How much do the labels depend on the three Eigen reduction orders the oracle had to restate?
CPU only:   python tools/order_sensitivity.py [n_synthetic_frames]

The oracle is run three times on the same inputs -- with the restated Eigen 3.4 orders (the parity reference), with plain
left-to-right fp32 sums, and with fp64 accumulation rounded once -- and the per-pixel labels are compared.  If Eigen's real
orders differed from the restatement, the effect on the labels would be of the size reported here."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
from conftest import load_frame
from deplex_b200 import synth

VARIANTS = {0: "restated Eigen 3.4 orders", 1: "plain left-to-right fp32", 2: "fp64 accumulation, rounded once"}
THREADS = os.cpu_count() or 1


def labels_for(variant, h, w, cfg, clouds):
    oracle.set_sum_variant(variant)
    try:
        return oracle.process_batch(h, w, cfg, clouds, 1, THREADS)
    finally:
        oracle.set_sum_variant(0)


def partition_mismatch(a, b):
    """pixels that still differ after every label of b is renamed to the label of a it overlaps most: what is left is a
    real difference between the two segmentations, not a renumbering (the north star compares 'up to label permutation')"""
    lut = np.zeros(int(b.max()) + 1, dtype=a.dtype)
    for lb in np.unique(b):
        vals, cnt = np.unique(a[b == lb], return_counts=True)
        lut[lb] = vals[cnt.argmax()]
    return int((lut[b] != a).sum())


def report(name, h, w, cfg, clouds):
    base = labels_for(0, h, w, cfg, clouds)
    print(f"{name}: {len(clouds)} frame(s), {h}x{w}, max label of frame 0 = {int(base[0].max())}")
    for v in (1, 2):
        other = labels_for(v, h, w, cfg, clouds)
        raw = other != base
        frames = int(raw.reshape(len(clouds), -1).any(axis=1).sum())
        real = [partition_mismatch(base[f], other[f]) for f in range(len(clouds))]
        print(f"   vs {VARIANTS[v]:34s}: {frames} frame(s) not bit-identical; after renaming labels "
              f"{sum(1 for r in real if r)} frame(s) differ, {sum(real)} of {raw.size} pixels = {100.0 * sum(real) / raw.size:.4f} % "
              f"(worst frame {100.0 * max(real) / base[0].size:.3f} %); max label of frame 0 = {int(other[0].max())}")


def main():
    n_synth = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    oracle.build()
    for name in ("tum", "icl"):
        depth, k, ini = load_frame(name)
        cloud = synth.depth_to_cloud(depth, k, "rowmajor")[None]
        report(f"shipped {name.upper()} frame, shipped ini", 480, 640, oracle.load_ini(ini), cloud)
        if name == "tum":
            report("shipped TUM frame, default config (golden 34)", 480, 640, oracle.OracleConfig(), cloud)
    for (h, w, n) in ((480, 640, n_synth), (720, 1280, max(1, n_synth // 8)), (1080, 1920, max(1, n_synth // 16))):
        report("synthetic", h, w, oracle.OracleConfig(), synth.make_batch(h, w, 3000000, n, "rowmajor"))


if __name__ == "__main__":
    main()
