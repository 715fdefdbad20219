"""Host-side logic of the product and the shape of the C-ABI, without a GPU: the library loads, exports every
symbol include/deplex_b200.h declares, parses configs like the reference, and refuses to run without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


def test_library_exports_every_declared_symbol(lib_built):
    from deplex_b200 import _capi
    header = open(os.path.join(ROOT, "include", "deplex_b200.h")).read()
    declared = set(re.findall(r"DPX_API[^;(]*?\b(dpx_\w+)\s*\(", header))
    assert declared == set(_capi.EXPORTS)
    lib = C.CDLL(lib_built)
    for name in declared:
        assert hasattr(lib, name), name
    assert _capi.load().dpx_version() == 200


def test_struct_layouts_match_header(lib_built):
    from deplex_b200 import _capi
    assert C.sizeof(_capi.dpx_config) == 64
    assert C.sizeof(_capi.dpx_cell) == 4 * (3 + 6 + 3 + 3 + 4 + 5)
    assert C.sizeof(_capi.dpx_plane) == 4 * 11
    assert C.sizeof(_capi.dpx_info) == 44
    assert C.sizeof(_capi.dpx_intrinsics) == 16


def test_config_defaults_and_ini(lib_built, tmp_path):
    from deplex_b200 import Config
    d = Config()
    # config.h:51-81
    assert (d.patch_size, d.histogram_bins_per_coord, d.min_region_growing_candidate_size) == (10, 20, 5)
    assert (d.min_region_growing_cells_activated, d.min_pts_per_cell, d.max_number_depth_discontinuity) == (4, 3, 1)
    assert d.min_cos_angle_merge == np.float32(0.90) and d.max_merge_dist == 500
    assert d.min_region_planarity_score == np.float32(0.55) and d.depth_sigma_coeff == np.float32(1.425e-6)
    assert d.depth_sigma_margin == 10 and d.depth_discontinuity_threshold == 160
    assert (d.ransac_refinement, d.ransac_max_iterations) == (0, 1000)
    assert d.ransac_threshold == 1 and d.ransac_inliers_ratio == np.float32(0.9)

    icl = Config(os.path.join(GOLDEN, "ICL_living_room.ini"))
    assert icl.patch_size == 4 and icl.min_cos_angle_merge == np.float32(0.93)
    assert icl.min_region_planarity_score == 0.5 and icl.ransac_inliers_ratio == np.float32(0.15)

    # cpp/tests/test_config.cpp:24-29
    with pytest.raises(RuntimeError, match="Couldn't open ini file: /no/such.ini"):
        Config("/no/such.ini")
    # config.cpp:33-79: header / comments / unknown keys / leading '=' are skipped, CRLF tolerated, no final newline
    p = tmp_path / "partial.ini"
    p.write_bytes(b"[Parameters]\r\npatchSize=12\r\n;minCosAngleForMerge=0.5\n#maxMergeDist=1\nbogusKey=3\n=7\n\nransacRefinement=5")
    c = Config(str(p))
    assert c.patch_size == 12 and c.ransac_refinement == 1 and c.max_merge_dist == 500
    assert c.min_cos_angle_merge == np.float32(0.90)


def test_config_matches_oracle_parser(lib_built, oracle_mod, tmp_path):
    from deplex_b200 import Config
    for name in ("TUM_fr3_long_val.ini", "ICL_living_room.ini"):
        a = Config(os.path.join(GOLDEN, name)).as_dict()
        b = oracle_mod.load_ini(os.path.join(GOLDEN, name))
        for k, v in a.items():
            assert v == getattr(b, k), (name, k)


def test_constructor_errors_match_reference(lib_built):
    from deplex_b200 import Config, PlaneExtractor
    # plane_extractor.cpp:161-164 (thrown before any device work)
    with pytest.raises(RuntimeError) as e:
        PlaneExtractor(480, 640, Config(patch_size=0))
    assert str(e.value) == "Error! Invalid config parameter: patchSize(0). patchSize has to be positive."


def test_no_cpu_fallback(lib_built):
    """Without a CUDA device the product refuses to run instead of falling back to the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from deplex_b200 import CudaError, PipelinedExtractor, PlaneExtractor
    with pytest.raises(CudaError):
        PlaneExtractor(480, 640)
    with pytest.raises(CudaError):
        PipelinedExtractor(480, 640, lanes=2)


def test_product_does_not_touch_the_oracle():
    """Nothing under deplex_b200/ may import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "deplex_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "deplex_oracle" not in text and "dpxo_" not in text, f
