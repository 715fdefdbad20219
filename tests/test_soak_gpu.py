"""Differential soak: many synthetic frames per geometry against the oracle (run frame-parallel on the host cores).
Exercises every storage mode of the region-growing kernel (shared memory; member runs in global memory + wide BFS
steps; everything in global memory), the fused label painting with many frames finishing in different orders, and both
input layouts.  Bit-exact labels are required."""
import numpy as np
import pytest

from conftest import to_oracle_cfg

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("hw,patch,n_frames,layout", [
    ((480, 640), 10, 160, "rowmajor"),     # mode 0, the bench configuration
    ((480, 640), 5, 48, "colmajor"),       # mode 1 (12 288 cells), scalar cell walk
    ((480, 640), 4, 24, "rowmajor"),       # mode 1 + wide steps (19 200 cells)
    ((720, 1280), 8, 12, "rowmajor"),      # mode 1 (14 400 cells)
    ((1080, 1920), 8, 4, "colmajor"),      # mode 2 (32 400 cells)
])
def test_many_frames_match_oracle(oracle_mod, hw, patch, n_frames, layout):
    import os
    from deplex_b200 import Config, PlaneExtractor, synth, LAYOUT_COLMAJOR, LAYOUT_ROWMAJOR
    h, w = hw
    cfg = Config(patch_size=patch)
    batch = synth.make_batch(h, w, 31000 + 17 * patch, n_frames, layout)
    lay = LAYOUT_ROWMAJOR if layout == "rowmajor" else LAYOUT_COLMAJOR
    ex = PlaneExtractor(h, w, cfg, max_batch=n_frames)
    got = ex.process_batch_host(batch, lay)
    ref = oracle_mod.process_batch(h, w, to_oracle_cfg(oracle_mod, cfg), batch, 1 if layout == "rowmajor" else 0,
                                   os.cpu_count() or 1)
    bad = [(f, int((got[f] != ref[f]).sum())) for f in range(n_frames) if not np.array_equal(got[f], ref[f])]
    assert not bad, f"frames with differing labels (frame, pixels): {bad[:8]}"
    assert int(got.max()) > 0
