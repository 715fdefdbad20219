"""The CPU oracle against the reference's own golden values / known-answer tests and
against the reference's 3x3 eigensolver sources compiled into oracle/_ref/.  Runs without a GPU."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import GOLDEN, frame_cloud, load_frame


def test_tum_default_config_max_label_34(oracle_mod):
    # cpp/tests/test_plane_extractor.cpp:27-33, python/tests/test_plane_extraction.py:44-47
    xyz, _ = frame_cloud("tum")
    labels = oracle_mod.process(480, 640, oracle_mod.OracleConfig(), xyz)
    assert labels.max() == 34
    assert labels.size == 640 * 480


def test_depth_cloud_z_range(oracle_mod):
    # cpp/tests/test_depth_image.cpp:42-48
    depth, k, _ = load_frame("tum")
    xyz = oracle_mod.depth_to_cloud(depth, k["fx"], k["fy"], k["cx"], k["cy"])
    assert xyz[:, 2].max() == 46655 and xyz[:, 2].min() == 0


def test_synth_cloud_matches_oracle_backprojection(oracle_mod):
    from deplex_b200 import synth
    depth, k, _ = load_frame("icl")
    a = synth.depth_to_cloud(depth, k)
    b = oracle_mod.depth_to_cloud(depth, k["fx"], k["fy"], k["cx"], k["cy"])
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_zero_leading_config(oracle_mod):
    # test_plane_extractor.cpp:35-45
    xyz, ini = frame_cloud("tum")
    cfg = oracle_mod.load_ini(ini)
    cfg.min_region_planarity_score = 5000
    labels = oracle_mod.process(480, 640, cfg, xyz)
    assert not labels.any() and labels.size == xyz.shape[0]


def test_zero_patch_size_throws(oracle_mod):
    # test_plane_extractor.cpp:47-53
    xyz, ini = frame_cloud("tum")
    cfg = oracle_mod.load_ini(ini)
    cfg.patch_size = 0
    with pytest.raises(oracle_mod.OracleError) as e:
        oracle_mod.process(480, 640, cfg, xyz)
    assert e.value.code == 1
    assert str(e.value) == "Error! Invalid config parameter: patchSize(0). patchSize has to be positive."


def test_enormous_patch_size(oracle_mod):
    # test_plane_extractor.cpp:55-65
    xyz, ini = frame_cloud("tum")
    cfg = oracle_mod.load_ini(ini)
    cfg.patch_size = 1000000
    labels = oracle_mod.process(480, 640, cfg, xyz)
    assert not labels.any() and labels.size == xyz.shape[0]


def test_zero_value_points(oracle_mod):
    # test_plane_extractor.cpp:67-74
    xyz = np.zeros((640 * 480, 3), dtype=np.float32)
    labels = oracle_mod.process(480, 640, oracle_mod.OracleConfig(), xyz)
    assert not labels.any() and labels.size == xyz.shape[0]


def test_empty_and_wrong_shape_throw(oracle_mod):
    # test_plane_extractor.cpp:76-88
    with pytest.raises(oracle_mod.OracleError) as e:
        oracle_mod.process(480, 640, oracle_mod.OracleConfig(), np.zeros((0, 3), dtype=np.float32))
    assert str(e.value) == "Error! Number of points doesn't match image shape: 0 != 480 x 640"
    with pytest.raises(oracle_mod.OracleError) as e:
        oracle_mod.process(240, 320, oracle_mod.OracleConfig(), np.zeros((640 * 480, 3), dtype=np.float32))
    assert str(e.value) == "Error! Number of points doesn't match image shape: 307200 != 240 x 320"


def test_config_ini(oracle_mod, tmp_path):
    # cpp/tests/test_config.cpp:24-29 + data/invalid/MissingParameters.ini semantics
    with pytest.raises(oracle_mod.OracleError) as e:
        oracle_mod.load_ini("/no/such/file.ini")
    assert str(e.value) == "Couldn't open ini file: /no/such/file.ini"
    p = tmp_path / "partial.ini"
    p.write_text("[Parameters]\npatchSize=12\n;minCosAngleForMerge=0.5\n#maxMergeDist=1\nbogusKey=3\n=7\nransacRefinement=1")
    cfg = oracle_mod.load_ini(str(p))
    assert cfg.patch_size == 12 and cfg.ransac_refinement == 1
    assert cfg.min_cos_angle_merge == np.float32(0.90) and cfg.max_merge_dist == 500
    icl = oracle_mod.load_ini(os.path.join(GOLDEN, "ICL_living_room.ini"))
    assert icl.patch_size == 4 and icl.min_cos_angle_merge == np.float32(0.93)
    assert icl.min_region_planarity_score == 0.5 and icl.ransac_inliers_ratio == np.float32(0.15)


@pytest.mark.parametrize("name", ["tum", "icl"])
def test_golden_fixture_regression(oracle_mod, name):
    """The committed oracle outputs (tests/golden/oracle_*.npz) are reproduced bit for bit."""
    xyz, ini = frame_cloud(name)
    cfg = oracle_mod.load_ini(ini)
    labels, dbg = oracle_mod.process(480, 640, cfg, xyz, debug=True)
    gold = np.load(os.path.join(GOLDEN, f"oracle_{name}.npz"))
    assert np.array_equal(labels, gold["labels"])
    for key in ("cell_planar", "cell_bin", "cell_seglabel", "merge_labels"):
        assert np.array_equal(dbg[key], gold[key]), key
    for key in ("cell_normal", "cell_d", "cell_mse", "cell_sum", "cell_var", "plane_normal", "plane_d"):
        assert np.array_equal(dbg[key].view(np.uint32), gold[key].view(np.uint32)), key
    assert dbg["n_seeds"] == int(gold["n_seeds"])
    # survey probes: TUM 88 seeds / 34 planes / 1 merge; ICL 53 seeds, 571 QL fallbacks
    if name == "tum":
        assert (dbg["n_seeds"], dbg["n_planes"], int(dbg["cell_planar"].sum())) == (88, 34, 1773)
        assert int((dbg["merge_labels"] != np.arange(34)).sum()) == 1
    else:
        assert (dbg["n_seeds"], dbg["n_ql_fallback"], int(dbg["cell_planar"].sum())) == (53, 571, 18527)


def test_summation_order_switch_is_off_by_default_and_matters(oracle_mod):
    """tools/order_sensitivity.py's switch: other summation orders change per-cell moments (the labels depend on the order
    the reference's Eigen build uses), and setting it back restores the committed fixture bit for bit."""
    xyz, ini = frame_cloud("tum")
    cfg = oracle_mod.load_ini(ini)
    gold = np.load(os.path.join(GOLDEN, "oracle_tum.npz"))
    try:
        for variant in (1, 2):
            oracle_mod.set_sum_variant(variant)
            labels, dbg = oracle_mod.process(480, 640, cfg, xyz, debug=True)
            assert not np.array_equal(dbg["cell_sum"].view(np.uint32), gold["cell_sum"].view(np.uint32))
            assert int(labels.max()) == 34          # the reference's golden value does not discriminate between orders
            assert 0 < int((labels != gold["labels"]).sum())
    finally:
        oracle_mod.set_sum_variant(0)
    labels, dbg = oracle_mod.process(480, 640, cfg, xyz, debug=True)
    assert np.array_equal(labels, gold["labels"])
    assert np.array_equal(dbg["cell_sum"].view(np.uint32), gold["cell_sum"].view(np.uint32))


def test_layouts_agree(oracle_mod):
    xyz, ini = frame_cloud("tum")
    cfg = oracle_mod.load_ini(ini)
    a = oracle_mod.process(480, 640, cfg, xyz)
    b = oracle_mod.process(480, 640, cfg, np.asfortranarray(xyz))
    assert np.array_equal(a, b)


@pytest.mark.parametrize("name", ["tum", "icl"])
def test_refinement_mse_not_worse(oracle_mod, name):
    # cpp/tests/test_refinement.cpp:43-75: MSE of the points labelled 1 does not grow with RANSAC refinement
    xyz, ini = frame_cloud(name)
    cfg = oracle_mod.load_ini(ini)

    def plane_mse(c):
        labels = oracle_mod.process(480, 640, c, xyz)
        pts = xyz[labels == 1].astype(np.float64)
        cov = np.cov(pts.T, bias=True)
        return np.linalg.eigvalsh(cov)[0], labels

    coarse, l0 = plane_mse(cfg)
    cfg.ransac_refinement = 1
    refined, l1 = plane_mse(cfg)
    assert refined <= coarse
    assert ((l1 == 0) | (l1 == l0)).all()  # refinement only removes labels


def _ref_dsyev(oracle_mod):
    path = oracle_mod.ref_dsyev_path()
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libdsyev_ref.so not built (reference tree not mounted)")
    lib = C.CDLL(path)
    lib.dsyevh3.argtypes = [C.c_void_p] * 3
    return lib


def test_eig3_restatement_matches_reference_dsyev(oracle_mod):
    """Bit-exact agreement of the oracle's 3x3 solver with libs/dsyev compiled from the reference's sources."""
    ref = _ref_dsyev(oracle_mod)
    rng = np.random.default_rng(1234)
    mats = []
    for i in range(6000):
        n = int(rng.integers(3, 60))
        X = rng.normal(size=(n, 3)) * rng.uniform(1e-3, 1e4, size=3)
        if i % 3 == 0:  # nearly planar: tiny smallest eigenvalue
            X[:, 2] = 0.3 * X[:, 0] - 0.2 * X[:, 1] + rng.normal(size=n) * 1e-7
        if i % 5 == 0:  # exactly rank deficient / repeated eigenvalues -> QL fallback
            X = np.round(X)
            X[:, 1] = X[:, 0]
        mats.append(X.T @ X)
    mats += [np.zeros((3, 3)), np.eye(3), np.diag([1.0, 1.0, 2.0]), np.diag([3.0, 1.0, 1.0]), np.ones((3, 3)),
             np.array([[2.0, 1, 0], [1, 2, 0], [0, 0, 3]]), np.diag([1e-300, 1.0, 1e300])]
    # covariance matrices of real cells exercise the production regime
    gold = np.load(os.path.join(GOLDEN, "oracle_icl.npz"))
    S, V, valid = gold["cell_sum"], gold["cell_var"], gold["cell_valid"].astype(bool)
    for c in np.flatnonzero(valid)[::7]:
        cov = (V[c].reshape(3, 3) - np.outer(S[c], S[c]) / np.float32(16)).astype(np.float32).astype(np.float64)
        mats.append(cov)
    n_ql = 0
    for A in mats:
        A = np.ascontiguousarray(A, dtype=np.float64)
        Q, w = np.zeros((3, 3)), np.zeros(3)
        ref.dsyevh3(A.ctypes.data, Q.ctypes.data, w.ctypes.data)
        q2, w2, ql = oracle_mod.eig3(A)
        n_ql += ql
        assert np.array_equal(w.view(np.uint64), w2.view(np.uint64)), A
        assert np.array_equal(Q.view(np.uint64), q2.view(np.uint64)), A
    assert n_ql > 100  # the QL branch is really exercised


def test_synthetic_determinism():
    from deplex_b200 import synth
    a = synth.make_depth(480, 640, 5)
    b = synth.make_depth(480, 640, 5)
    assert np.array_equal(a, b) and a.dtype == np.uint16
    assert not np.array_equal(a, synth.make_depth(480, 640, 6))
    cm = synth.make_cloud(120, 160, 1, "colmajor")
    rm = synth.make_cloud(120, 160, 1, "rowmajor")
    assert cm.shape == (3, 120 * 160) and np.array_equal(cm.T, rm)


def test_oracle_matches_reference_build(oracle_mod):
    """The pin VERDICT r01 asked for: when oracle/_ref/libdeplex_ref.so exists (the UNMODIFIED reference sources built by
    `make -C oracle ref_full EIGEN3_INCLUDE_DIR=...`, oracle/ref_full_shim.cpp), the oracle must reproduce its labels and
    the raw coord_sum_ / variance_ bits of every cell -- the three Eigen reductions (cell_segment_stat.cpp:29-35,56,74)
    that could only be restated, never run, in an image without Eigen.  Skipped, not passed, until such a build exists."""
    if not oracle_mod.ref_available():
        pytest.skip("oracle/_ref/libdeplex_ref.so not built: no Eigen >= 3.4 tree in this image "
                    "(external/eigen3/CMakeLists.txt:1-16 fetches it from the network)")
    from deplex_b200 import synth
    cases = []
    for name in ("tum", "icl"):
        xyz, ini = frame_cloud(name)
        cases.append((name, 480, 640, oracle_mod.load_ini(ini), xyz))
    for i in range(32):
        h, w, patch = [(480, 640, 10), (480, 640, 4), (720, 1280, 8), (480, 640, 6)][i % 4]
        cases.append((f"synth{i}", h, w, oracle_mod.OracleConfig(patch_size=patch), synth.make_cloud(h, w, 7000 + i)))
    for name, h, w, cfg, xyz in cases:
        labels, dbg = oracle_mod.process(h, w, cfg, xyz, debug=True)
        ref = oracle_mod.ref_process(h, w, cfg, xyz)
        assert np.array_equal(labels, ref), f"{name}: {(labels != ref).sum()} labels differ from the reference build"
        st = oracle_mod.ref_cell_stats(h, w, cfg, xyz)
        valid = dbg["cell_valid"].astype(bool)
        assert np.array_equal(st["cell_planar"].astype(bool), dbg["cell_planar"].astype(bool)), name
        for key in ("cell_sum", "cell_var", "cell_mse", "cell_score", "cell_d", "cell_normal"):
            a = np.ascontiguousarray(st[key][valid]).view(np.uint32)
            b = np.ascontiguousarray(dbg[key][valid]).view(np.uint32)
            assert np.array_equal(a, b), f"{name}: {key} differs in {(a != b).any(axis=-1).sum() if a.ndim > 1 else (a != b).sum()} cells"


def test_release_build_of_the_oracle_gives_identical_labels(oracle_mod):
    """bench.py times the -O3 -DNDEBUG build (CMake Release); parity is judged on the -O2 one.
    Both keep contraction off and use no -march, so they must agree bit for bit."""
    from deplex_b200 import synth
    cfg = oracle_mod.OracleConfig()
    batch = synth.make_batch(480, 640, 31000, 4, "rowmajor")
    a = oracle_mod.process_batch(480, 640, cfg, batch, 1, 2)
    b = oracle_mod.process_batch(480, 640, cfg, batch, 1, 2, timed_build=True)
    assert np.array_equal(a, b)
    xyz, ini = frame_cloud("icl")
    cfg = oracle_mod.load_ini(ini)
    a = oracle_mod.process_batch(480, 640, cfg, xyz[None], 1, 1)
    b = oracle_mod.process_batch(480, 640, cfg, xyz[None], 1, 1, timed_build=True)
    assert np.array_equal(a, b)


def test_uniform_int_mapping_is_explicit_and_pinned(oracle_mod):
    """std::uniform_int_distribution is implementation-defined; the oracle restates both libstdc++ generations
    explicitly.  Variant 0 (GCC >= 11) is pinned against this host's <random>; variant 1 (GCC <= 10)
    is a different stream, and refinement labels depend on the choice."""
    for n in (3, 7, 100, 4097, 307200, (1 << 31) - 1):
        assert oracle_mod.uniform_selftest(n, 20000) == 0, n
    xyz, ini = frame_cloud("tum")
    cfg = oracle_mod.load_ini(ini)
    cfg.ransac_refinement = 1
    try:
        a = oracle_mod.process(480, 640, cfg, xyz)
        oracle_mod.set_uniform_int_variant(1)
        b = oracle_mod.process(480, 640, cfg, xyz)
    finally:
        oracle_mod.set_uniform_int_variant(0)
    assert np.array_equal(a, oracle_mod.process(480, 640, cfg, xyz))
    assert a.max() == b.max() and (a != b).any()  # same planes, different sampled triples
