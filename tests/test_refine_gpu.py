"""RANSAC refinement (ransacRefinement=1) on the GPU against the oracle: plane_extractor.cpp:472-509,
libs/rtl RANSAC.hpp / Plane.hpp, std::mt19937 + libstdc++ uniform_int_distribution.  The bar is bit-exact labels:
the random stream, the sample order, the fp32 model arithmetic and the early-exit rule all have to agree."""
import time

import numpy as np
import pytest

from conftest import frame_cloud, to_oracle_cfg

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["tum", "icl"])
@pytest.mark.parametrize("layout", ["rowmajor", "colmajor"])
def test_shipped_frames_refined_labels_identical(oracle_mod, name, layout):
    from deplex_b200 import Config, PlaneExtractor
    xyz, ini = frame_cloud(name)
    cfg = Config(ini, ransac_refinement=1)   # shipped ini: threshold 1, ratio 0.15, 1000 iterations
    host = xyz if layout == "rowmajor" else np.asfortranarray(xyz)
    labels = PlaneExtractor(480, 640, cfg).process(host)
    ref = oracle_mod.process(480, 640, to_oracle_cfg(oracle_mod, cfg), xyz)
    coarse = PlaneExtractor(480, 640, Config(ini)).process(host)
    assert ((labels == 0) | (labels == coarse)).all()      # refinement only removes labels
    assert (labels != coarse).any()                          # ... and it does remove some
    assert np.array_equal(labels, ref), f"{(labels != ref).sum()} pixels differ"


@pytest.mark.parametrize("threshold,ratio,iters", [(1.0, 0.15, 1000), (25.0, 0.9, 40), (3.0, 0.5, 7), (1.0, 0.9, 0)])
def test_refinement_parameters_and_batches(oracle_mod, threshold, ratio, iters):
    """Different early-exit regimes (immediate hit, iteration cap in the middle of a 32-hypothesis round, zero
    iterations), several frames per batch: every frame restarts the generator, like one process() call each."""
    from deplex_b200 import Config, PlaneExtractor, synth, LAYOUT_ROWMAJOR
    h, w, F = 480, 640, 4
    cfg = Config(ransac_refinement=1, ransac_threshold=threshold, ransac_inliers_ratio=ratio, ransac_max_iterations=iters)
    batch = synth.make_batch(h, w, 4242, F, "rowmajor")
    labels = PlaneExtractor(h, w, cfg, max_batch=F).process_batch_host(batch, LAYOUT_ROWMAJOR)
    ocfg = to_oracle_cfg(oracle_mod, cfg)
    for f in range(F):
        ref = oracle_mod.process(h, w, ocfg, batch[f])
        assert np.array_equal(labels[f], ref), f"frame {f}: {(labels[f] != ref).sum()} pixels differ"


@pytest.mark.parametrize("frames", [12, 20])
def test_refinement_batch_sizes_reach_every_kernel_variant(oracle_mod, frames):
    """The launcher picks the cluster size and register budget by batch size (refine.cu launch_refine): up to 9 frames
    get 16-CTA clusters (the tests above), up to 18 get 8-CTA clusters with the full register budget, larger batches
    8-CTA clusters at two CTAs per SM.  Same labels from all of them."""
    from deplex_b200 import Config, PlaneExtractor, synth, LAYOUT_ROWMAJOR
    h, w = 480, 640
    cfg = Config(ransac_refinement=1, ransac_threshold=3.0, ransac_inliers_ratio=0.5, ransac_max_iterations=200)
    batch = synth.make_batch(h, w, 9090 + frames, frames, "rowmajor")
    labels = PlaneExtractor(h, w, cfg, max_batch=frames).process_batch_host(batch, LAYOUT_ROWMAJOR)
    ocfg = to_oracle_cfg(oracle_mod, cfg)
    for f in range(frames):
        ref = oracle_mod.process(h, w, ocfg, batch[f])
        assert np.array_equal(labels[f], ref), f"frame {f} of {frames}: {(labels[f] != ref).sum()} pixels differ"


def test_refine_work_counter(oracle_mod):
    """dpx_get_refine_work: rounds of 128 hypotheses and point passes of one frame.  ransacMaxIterations = 300 with a ratio
    no hypothesis reaches means every label runs ceil(300 / 128) = 3 rounds over all of its points."""
    from deplex_b200 import Config, PlaneExtractor, UnsupportedError
    xyz, ini = frame_cloud("tum")
    cfg = Config(ini, ransac_refinement=1, ransac_inliers_ratio=2.0, ransac_max_iterations=300)
    ex = PlaneExtractor(480, 640, cfg)
    labels_before = PlaneExtractor(480, 640, Config(ini)).process(xyz)
    ex.process(xyz)
    point_passes, rounds = ex.refine_work(0)
    n_labels = len(np.unique(labels_before[labels_before > 0]))
    assert rounds == 3 * n_labels
    assert point_passes == 3 * int((labels_before > 0).sum())
    with pytest.raises(UnsupportedError):
        PlaneExtractor(480, 640, Config(ini)).refine_work(0)


def test_refinement_mse_not_worse():
    """cpp/tests/test_refinement.cpp:43-75 on the GPU path: the MSE of the points labelled 1 does not grow."""
    from deplex_b200 import Config, PlaneExtractor
    for name in ("tum", "icl"):
        xyz, ini = frame_cloud(name)

        def plane_mse(cfg):
            labels = PlaneExtractor(480, 640, cfg).process(xyz)
            pts = xyz[labels == 1].astype(np.float64)
            return np.linalg.eigvalsh(np.cov(pts.T, bias=True))[0]

        assert plane_mse(Config(ini, ransac_refinement=1)) <= plane_mse(Config(ini))


def test_refinement_fine_grid_and_fhd(oracle_mod):
    from deplex_b200 import Config, PlaneExtractor, synth
    for (h, w, patch) in ((480, 640, 4), (1080, 1920, 10)):
        cfg = Config(patch_size=patch, ransac_refinement=1, ransac_threshold=4.0, ransac_inliers_ratio=0.6, ransac_max_iterations=50)
        xyz = synth.make_cloud(h, w, 77)
        t0 = time.time()
        labels = PlaneExtractor(h, w, cfg).process(xyz)
        ref = oracle_mod.process(h, w, to_oracle_cfg(oracle_mod, cfg), xyz)
        assert np.array_equal(labels, ref), f"{h}x{w}/p{patch}: {(labels != ref).sum()} pixels differ ({time.time() - t0:.1f}s)"


def test_refinement_many_tiny_labels(oracle_mod):
    """Hundreds of labels of one to a few 4x4 cells (n = 16, 32, ... points): sampling three distinct ranks out of 16
    repeats all the time, so the groups of 32 hypotheses take extra draws, settle over several passes or fall back to
    the draw-by-draw path, and the generator is rewound in every label; more labels than the shared-memory label sort
    holds.  Both uniform_int mappings."""
    from deplex_b200 import Config, PlaneExtractor, synth
    h, w = 480, 640
    for seed, iters in ((5, 300), (77, 130)):
        cfg = Config(patch_size=4, min_region_growing_candidate_size=1, min_region_growing_cells_activated=1,
                     ransac_refinement=1, ransac_threshold=2.0, ransac_inliers_ratio=0.95, ransac_max_iterations=iters)
        xyz = synth.make_cloud(h, w, seed)
        ocfg = to_oracle_cfg(oracle_mod, cfg)
        ex = PlaneExtractor(h, w, cfg)
        for variant in ("libstdc++11", "libstdc++10"):
            ex.set_rng_compat(variant)
            oracle_mod.set_uniform_int_variant(0 if variant == "libstdc++11" else 1)
            try:
                ref = oracle_mod.process(h, w, ocfg, xyz)
            finally:
                oracle_mod.set_uniform_int_variant(0)
            labels = ex.process(xyz)
            assert labels.max() > 256
            assert np.array_equal(labels, ref), f"seed {seed}, mapping {variant}: {(labels != ref).sum()} pixels differ"


def test_refinement_with_pre_gcc11_uniform_int_mapping(oracle_mod):
    """std::uniform_int_distribution is implementation-defined.  With dpx_set_rng_compat(libstdc++10) the
    kernel reproduces the oracle's explicit restatement of the GCC <= 10 mapping (scaling + rejection), for the tape fast
    path and the sequential fallback alike; the default stays the GCC >= 11 mapping."""
    from deplex_b200 import Config, PlaneExtractor
    for name in ("tum", "icl"):
        xyz, ini = frame_cloud(name)
        cfg = Config(ini, ransac_refinement=1)
        ocfg = to_oracle_cfg(oracle_mod, cfg)
        ex = PlaneExtractor(480, 640, cfg)
        default = ex.process(xyz)
        assert np.array_equal(default, oracle_mod.process(480, 640, ocfg, xyz))
        ex.set_rng_compat("libstdc++10")
        legacy = ex.process(xyz)
        try:
            oracle_mod.set_uniform_int_variant(1)
            want = oracle_mod.process(480, 640, ocfg, xyz)
        finally:
            oracle_mod.set_uniform_int_variant(0)
        assert np.array_equal(legacy, want)
        assert (legacy != default).any()
        ex.set_rng_compat("libstdc++11")
        assert np.array_equal(ex.process(xyz), default)
