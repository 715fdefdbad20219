"""GPU checks of the drop-in boundary: pybind module == ctypes mirror == oracle, and the C++ class through its
test executable (cpp/tests/test_plane_extractor.cpp:27-88 restated in deplex_b200/cpp/tests/test_api.cpp)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, frame_cloud, load_frame, to_oracle_cfg

pytestmark = pytest.mark.gpu
PYPKG = os.path.join(ROOT, "deplex_b200", "python")


@pytest.fixture(scope="module")
def deplex_mod(lib_built):
    if PYPKG not in sys.path:
        sys.path.insert(0, PYPKG)
    import deplex
    return deplex


def test_pybind_tum_default_config_34(deplex_mod, oracle_mod):
    """python/tests/test_plane_extraction.py:44-52: float64 C-order input, default config, max(labels) == 34."""
    import deplex_b200
    xyz, _ = frame_cloud("tum")
    algorithm = deplex_mod.PlaneExtractor(image_height=480, image_width=640)
    labels = algorithm.process(xyz.astype(np.float64))
    assert labels.dtype == np.int32 and labels.shape == (480 * 640,)
    assert max(labels) == 34
    ref = oracle_mod.process(480, 640, to_oracle_cfg(oracle_mod, deplex_b200.Config()), xyz)
    assert np.array_equal(labels, ref)
    # Fortran-ordered input (what the reference's transform_to_pcd returns) takes the column-major path
    assert np.array_equal(algorithm.process(np.asfortranarray(xyz)), ref)
    assert len(algorithm.planes()) == 34


def test_pybind_ini_config_and_errors(deplex_mod, oracle_mod):
    import deplex_b200
    xyz, ini = frame_cloud("icl")
    labels = deplex_mod.PlaneExtractor(480, 640, config=deplex_mod.Config(ini)).process(xyz)
    ref = oracle_mod.process(480, 640, to_oracle_cfg(oracle_mod, deplex_b200.Config(ini)), xyz)
    assert np.array_equal(labels, ref)
    algorithm = deplex_mod.PlaneExtractor(480, 640)
    with pytest.raises(RuntimeError, match="Number of points doesn't match image shape: 3 != 480 x 640"):
        algorithm.process(np.empty((3, 3)))
    c = deplex_mod.Config()
    c.patch_size = 0
    with pytest.raises(RuntimeError, match=r"patchSize\(0\)"):
        deplex_mod.PlaneExtractor(480, 640, c)


def test_cpp_class_reference_cases(tmp_path):
    cv2 = pytest.importorskip("cv2")
    depth, _, _ = load_frame("tum")
    png = str(tmp_path / "tum.png")
    cv2.imwrite(png, depth)
    exe = os.path.join(ROOT, "deplex_b200", "cpp", "build", "test_api")
    r = subprocess.run([exe, png, os.path.join(GOLDEN, "TUM_fr3_long_val.K"), os.path.join(GOLDEN, "MissingParameters.ini"), "34"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all cases passed" in r.stdout
    exe = os.path.join(ROOT, "deplex_b200", "cpp", "build", "process_cloud")
    r = subprocess.run([exe, png, os.path.join(GOLDEN, "TUM_fr3_long_val.K"), os.path.join(GOLDEN, "TUM_fr3_long_val.ini"), "5"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "Number of found planes: 34" in r.stdout, r.stdout + r.stderr


def test_cpp_process_sequence_example(tmp_path):
    """examples/process_sequence.cpp workflow: a directory of PNGs, latency mode and raw-depth batch mode agree."""
    cv2 = pytest.importorskip("cv2")
    from deplex_b200 import synth
    seq = tmp_path / "seq"
    seq.mkdir()
    depth, _, _ = load_frame("tum")
    cv2.imwrite(str(seq / "000.png"), depth)
    for i in range(1, 5):
        cv2.imwrite(str(seq / f"{i:03d}.png"), synth.make_depth(480, 640, 60 + i))
    exe = os.path.join(ROOT, "deplex_b200", "cpp", "build", "process_sequence")
    r = subprocess.run([exe, str(seq), os.path.join(GOLDEN, "TUM_fr3_long_val.K"), os.path.join(GOLDEN, "TUM_fr3_long_val.ini"), "3"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "modes agree: yes" in r.stdout and "5 frames" in r.stdout
