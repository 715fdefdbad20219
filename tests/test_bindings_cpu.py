"""Host-side checks of the C++ drop-in layer and the pybind module `deplex.pybind` (no GPU needed).

Reference surface: cpp/pybind/deplex_pybind.cpp:20-24, plane_extraction/plane_extraction.cpp:28-37,
utils/utils.cpp:29-36, python/deplex/__init__.py:1-2; tests: cpp/tests/test_config.cpp:24-29,
cpp/tests/test_depth_image.cpp:24-48."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_frame

PYPKG = os.path.join(ROOT, "deplex_b200", "python")


@pytest.fixture(scope="module")
def deplex_mod(lib_built):
    if PYPKG not in sys.path:
        sys.path.insert(0, PYPKG)
    try:
        import deplex
    except ImportError:
        import __graft_entry__
        __graft_entry__.build()
        import deplex
    return deplex


def test_module_layout_matches_reference(deplex_mod):
    import deplex.pybind as pb
    assert hasattr(pb, "plane_extraction") and hasattr(pb, "utils")
    assert deplex_mod.PlaneExtractor is pb.plane_extraction.PlaneExtractor
    assert deplex_mod.Config is pb.plane_extraction.Config
    assert deplex_mod.utils.DepthImage is pb.utils.DepthImage
    # the extension links the in-tree CUDA library, not a copy of it
    maps = open("/proc/self/maps").read()
    assert os.path.join(ROOT, "deplex_b200", "libdeplex_b200.so") in maps


def test_config_from_ini(deplex_mod, capfd):
    c = deplex_mod.Config(path=os.path.join(GOLDEN, "MissingParameters.ini"))
    d = deplex_mod.Config(os.path.join(GOLDEN, "TUM_fr3_long_val.ini"))
    assert c.patch_size == 12 and d.patch_size == 10
    # ';'-commented and unknown keys leave the defaults and are reported on stderr (config.cpp:77)
    assert c.depth_sigma_coeff == pytest.approx(1.425e-6) and c.max_merge_dist == 500 and c.min_pts_per_cell == 3
    assert "Unknown parameter name: ;maxMergeDist" in capfd.readouterr().err
    with pytest.raises(RuntimeError, match="Couldn't open ini file: __INVALID_PATH"):
        deplex_mod.Config("__INVALID_PATH")


def test_depth_image_png16_and_backprojection(deplex_mod, tmp_path):
    cv2 = pytest.importorskip("cv2")
    from deplex_b200 import synth
    for name in ("tum", "icl"):
        depth, k, _ = load_frame(name)
        path = str(tmp_path / f"{name}.png")
        assert cv2.imwrite(path, depth)
        img = deplex_mod.utils.DepthImage(path)
        assert (img.height, img.width) == (480, 640)
        K = np.array([[k["fx"], 0, k["cx"]], [0, k["fy"], k["cy"]], [0, 0, 1]])
        pts = img.transform_to_pcd(K)
        assert pts.shape == (480 * 640, 3) and pts.dtype == np.float32
        # Eigen::MatrixX3f comes back Fortran-ordered from the reference's module (cpp/pybind/utils/utils.cpp:34)
        assert pts.flags.f_contiguous and not pts.flags.c_contiguous
        assert np.array_equal(pts[:, 2], depth.reshape(-1).astype(np.float32))
        want = synth.depth_to_cloud(depth, k, "rowmajor")
        assert np.array_equal(pts.view(np.uint32), want.view(np.uint32))
    # cpp/tests/test_depth_image.cpp:42-48: z range of the TUM frame
    depth, k, _ = load_frame("tum")
    assert depth.max() == 46655 and depth.min() == 0


def test_depth_image_other_png_flavours(deplex_mod, tmp_path):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    g8 = rng.integers(0, 256, (37, 53), dtype=np.uint8)
    cv2.imwrite(str(tmp_path / "g8.png"), g8)
    img = deplex_mod.utils.DepthImage(str(tmp_path / "g8.png"))
    z = img.transform_to_pcd(np.eye(3))[:, 2].reshape(37, 53)
    assert np.array_equal(z, g8.astype(np.float32) * 257)  # 8-bit samples widen as v * 257 (stb_image)
    rgb16 = rng.integers(0, 65536, (21, 34, 3), dtype=np.uint16)
    cv2.imwrite(str(tmp_path / "rgb16.png"), rgb16[:, :, ::-1])  # cv2 writes BGR
    img.reset(str(tmp_path / "rgb16.png"))
    z = img.transform_to_pcd(np.eye(3))[:, 2].reshape(21, 34)
    r, g, b = (rgb16[:, :, i].astype(np.uint32) for i in range(3))
    assert np.array_equal(z, ((r * 77 + g * 150 + b * 29) >> 8).astype(np.float32))
    # 8-bit colour: stb reduces to luma on the 8-bit samples first, then widens by 257 (r=1,g=0,b=0 -> 0, not 77)
    rgb8 = rng.integers(0, 256, (19, 31, 3), dtype=np.uint8)
    rgb8[0, 0] = (1, 0, 0)
    cv2.imwrite(str(tmp_path / "rgb8.png"), rgb8[:, :, ::-1])
    img.reset(str(tmp_path / "rgb8.png"))
    z = img.transform_to_pcd(np.eye(3))[:, 2].reshape(19, 31)
    r, g, b = (rgb8[:, :, i].astype(np.uint32) for i in range(3))
    assert np.array_equal(z, (((r * 77 + g * 150 + b * 29) >> 8) * 257).astype(np.float32)) and z[0, 0] == 0
    # invalid / empty files throw (cpp/tests/test_depth_image.cpp:30-40)
    (tmp_path / "empty.png").write_bytes(b"")
    for bad in ("empty.png", "missing.png"):
        with pytest.raises(RuntimeError, match="Error: Couldn't read image"):
            deplex_mod.utils.DepthImage(str(tmp_path / bad))


def test_png_reader_matches_reference_stb_image(deplex_mod, tmp_path):
    """Differential test against the reference's own decoder: oracle/_ref/libstb_ref.so is the vendored stb_image.h
    compiled where it lies (oracle/Makefile), called exactly like depth_image.cpp:32 (stbi_load_16, STBI_grey)."""
    import ctypes as C
    Image = pytest.importorskip("PIL.Image")
    so = os.path.join(ROOT, "oracle", "_ref", "libstb_ref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libstb_ref.so not built (reference tree not mounted)")
    stb = C.CDLL(so)
    stb.stb_ref_load16_grey.argtypes = [C.c_char_p, C.c_void_p, C.c_long, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    rng = np.random.default_rng(11)
    h, w = 23, 37
    files = {}
    files["grey8"] = Image.fromarray(rng.integers(0, 256, (h, w), dtype=np.uint8), "L")
    files["grey16"] = Image.fromarray(rng.integers(0, 65536, (h, w), dtype=np.uint16))
    files["rgb8"] = Image.fromarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8), "RGB")
    files["rgba8"] = Image.fromarray(rng.integers(0, 256, (h, w, 4), dtype=np.uint8), "RGBA")
    files["la8"] = Image.fromarray(rng.integers(0, 256, (h, w, 2), dtype=np.uint8), "LA")
    files["bilevel"] = Image.fromarray(rng.integers(0, 2, (h, w), dtype=np.uint8) * 255, "L").convert("1")
    pal = Image.fromarray(rng.integers(0, 256, (h, w), dtype=np.uint8), "P")
    pal.putpalette([int(v) for v in rng.integers(0, 256, 768)])
    files["palette8"] = pal
    pal4 = Image.fromarray(rng.integers(0, 16, (h, w), dtype=np.uint8), "P")
    pal4.putpalette([int(v) for v in rng.integers(0, 256, 48)])
    files["palette4"] = pal4
    checked = 0
    for name, im in files.items():
        path = str(tmp_path / f"{name}.png")
        im.save(path, bits=4) if name == "palette4" else im.save(path)
        ref = np.zeros(h * w, dtype=np.uint16)
        rw, rh = C.c_int(0), C.c_int(0)
        assert stb.stb_ref_load16_grey(path.encode(), ref.ctypes.data, ref.size, C.byref(rw), C.byref(rh)) == 1, name
        img = deplex_mod.utils.DepthImage(path)
        assert (img.height, img.width) == (rh.value, rw.value) == (h, w), name
        z = np.asarray(img.transform_to_pcd(np.eye(3)))[:, 2]
        assert np.array_equal(z, ref.astype(np.float32)), name
        checked += 1
    # the shipped frames as 16-bit grey: the format deplex is actually used with
    cv2 = pytest.importorskip("cv2")
    for name in ("tum", "icl"):
        depth, _, _ = load_frame(name)
        path = str(tmp_path / f"{name}.png")
        cv2.imwrite(path, depth)
        ref = np.zeros(depth.size, dtype=np.uint16)
        rw, rh = C.c_int(0), C.c_int(0)
        assert stb.stb_ref_load16_grey(path.encode(), ref.ctypes.data, ref.size, C.byref(rw), C.byref(rh)) == 1
        assert np.array_equal(ref.reshape(depth.shape), depth)
    assert checked == len(files)


def test_no_cpu_fallback_through_the_bindings(deplex_mod):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        deplex_mod.PlaneExtractor(image_height=480, image_width=640)


def test_cpp_binaries_built(lib_built):
    for exe in ("process_cloud", "process_sequence", "test_api"):
        assert os.access(os.path.join(ROOT, "deplex_b200", "cpp", "build", exe), os.X_OK)


def test_wheel_is_self_contained(deplex_mod, tmp_path):
    """python/setup.py gathers the module and both native libraries into one wheel (reference: python/setup.py);
    installed on its own, `import deplex` resolves all three from the wheel, not from this tree."""
    import glob
    import subprocess
    import zipfile
    dist, site = tmp_path / "dist", tmp_path / "site"
    r = subprocess.run([sys.executable, "setup.py", "-q", "bdist_wheel", "--dist-dir", str(dist), "--bdist-dir",
                        str(tmp_path / "bdist"), "build", "--build-base", str(tmp_path / "build"), "egg_info",
                        "--egg-base", str(tmp_path)], cwd=PYPKG, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    (wheel,) = glob.glob(str(dist / "deplex-*.whl"))
    names = zipfile.ZipFile(wheel).namelist()
    assert "deplex/libdeplex.so" in names and "deplex/libdeplex_b200.so" in names
    assert any(n.startswith("deplex/pybind.") and n.endswith(".so") for n in names)
    zipfile.ZipFile(wheel).extractall(site)
    probe = ("import deplex; c = deplex.Config(); assert c.patch_size == 10\n"
             "print([l.split()[-1] for l in open('/proc/self/maps') if 'deplex' in l])")
    env = {k: v for k, v in os.environ.items() if k not in ("PYTHONPATH", "LD_LIBRARY_PATH")}
    env["PYTHONPATH"] = str(site)
    r = subprocess.run([sys.executable, "-c", probe], cwd=str(tmp_path), env=env, capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    loaded = set(eval(r.stdout.strip().splitlines()[-1]))
    assert loaded and all(p.startswith(str(site)) for p in loaded), loaded
    assert len({os.path.basename(p) for p in loaded}) == 3


def test_cpp_text_io_roundtrip(lib_built, tmp_path):
    """deplex::utils::{save,read}PointCloudCSV and readIntrinsics (reference: utils/eigen_io.cpp:22-60) through libdeplex.so:
    float-exact round trip, the reference's error texts."""
    import subprocess
    src = tmp_path / "io_check.cpp"
    src.write_text(r'''
#include <cstdio>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>
#include "deplex/utils/eigen_io.h"
int main(int argc, char** argv) {
  using namespace deplex::utils;
  const std::string dir = argv[1];
  std::vector<float> pts = {0.1f, -2.5e-7f, 46655.f, 1.f / 3.f, 3.4028235e38f, 0.f, 7.f, 8.f, 9.f};
  savePointCloudCSV(pts, dir + "/cloud.csv");
  if (readPointCloudCSV(dir + "/cloud.csv", ',') != pts) return 1;
  { std::ofstream f(dir + "/bad.csv"); f << "1,2,3\n4,5\n"; }
  try { readPointCloudCSV(dir + "/bad.csv", ','); return 2; }
  catch (std::runtime_error const& e) { if (std::string(e.what()) != "Error reading file: Invalid points shape") return 3; }
  { std::ofstream f(dir + "/k.K"); f << "525.0 0 319.5\n0 525.0 239.5\n0 0 1\n"; }
  const auto k = readIntrinsics(dir + "/k.K");
  if (k[0] != 525.0f || k[2] != 319.5f || k[4] != 525.0f || k[5] != 239.5f || k[8] != 1.0f) return 4;
  try { readIntrinsics(dir + "/missing.K"); return 5; }
  catch (std::runtime_error const& e) {
    if (std::string(e.what()) != "Error: Couldn't open intrinsics file " + dir + "/missing.K") return 6;
  }
  std::puts("io ok");
  return 0;
}
''')
    pkg = os.path.join(ROOT, "deplex_b200")
    exe = tmp_path / "io_check"
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-DDEPLEX_NO_EIGEN", f"-I{pkg}/cpp/include", f"-I{ROOT}/include", str(src), "-o",
                        str(exe), f"-L{pkg}", "-ldeplex", "-ldeplex_b200", f"-Wl,-rpath,{pkg}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([str(exe), str(tmp_path)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "io ok" in r.stdout, (r.returncode, r.stderr[-500:])
