"""ctypes binding of oracle/libdeplex_oracle.so (see deplex_oracle.h).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libdeplex_oracle.so")
_lib = None

LAYOUT_COLMAJOR = 0
LAYOUT_ROWMAJOR = 1


class OracleError(RuntimeError):
    """The reference would have thrown std::runtime_error (code 1) or the input is outside the restated domain (2)."""

    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class OracleConfig(C.Structure):
    _fields_ = [
        ("patch_size", C.c_int32), ("histogram_bins_per_coord", C.c_int32),
        ("min_cos_angle_merge", C.c_float), ("max_merge_dist", C.c_float),
        ("min_region_growing_candidate_size", C.c_int32), ("min_region_growing_cells_activated", C.c_int32),
        ("min_region_planarity_score", C.c_float), ("depth_sigma_coeff", C.c_float),
        ("depth_sigma_margin", C.c_float), ("min_pts_per_cell", C.c_int32),
        ("depth_discontinuity_threshold", C.c_float), ("max_number_depth_discontinuity", C.c_int32),
        ("ransac_refinement", C.c_int32), ("ransac_max_iterations", C.c_int32),
        ("ransac_threshold", C.c_float), ("ransac_inliers_ratio", C.c_float),
    ]

    def __init__(self, **kw):
        super().__init__()
        _load().dpxo_config_default(C.byref(self))
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)


class _Debug(C.Structure):
    _fields_ = [
        ("cell_valid", C.c_void_p), ("cell_planar", C.c_void_p), ("cell_sum", C.c_void_p),
        ("cell_var", C.c_void_p), ("cell_mean", C.c_void_p), ("cell_normal", C.c_void_p),
        ("cell_d", C.c_void_p), ("cell_mse", C.c_void_p), ("cell_score", C.c_void_p),
        ("cell_tol", C.c_void_p), ("cell_eval", C.c_void_p), ("cell_bin", C.c_void_p),
        ("cell_seglabel", C.c_void_p),
        ("plane_capacity", C.c_int32), ("n_planes", C.c_void_p), ("plane_normal", C.c_void_p),
        ("plane_d", C.c_void_p), ("plane_mean", C.c_void_p), ("plane_mse", C.c_void_p),
        ("plane_score", C.c_void_p), ("plane_npts", C.c_void_p), ("merge_labels", C.c_void_p),
        ("n_seeds", C.c_void_p), ("n_ql_fallback", C.c_void_p),
    ]


def build(force=False):
    """Compile the oracle (and oracle/_ref when the reference tree is mounted)."""
    src = [os.path.join(_HERE, f) for f in ("deplex_oracle.cpp", "deplex_oracle.h", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "all"], check=True, capture_output=True)
    elif os.path.isdir("/root/reference/libs/dsyev/src") and not os.path.exists(ref_dsyev_path()):
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)


def ref_dsyev_path():
    return os.path.join(_HERE, "_ref", "libdsyev_ref.so")


def ref_full_path():
    """oracle/_ref/libdeplex_ref.so: the UNMODIFIED reference hot path behind oracle/ref_full_shim.cpp.  Exists only
    after `make -C oracle ref_full EIGEN3_INCLUDE_DIR=...` found an Eigen >= 3.4 tree (absent from this image)."""
    return os.path.join(_HERE, "_ref", "libdeplex_ref.so")


_ref_full = None


def _load_ref_full():
    global _ref_full
    if _ref_full is None:
        lib = C.CDLL(ref_full_path())
        lib.ref_process.argtypes = [C.c_int32, C.c_int32, C.POINTER(OracleConfig), C.c_void_p, C.c_int64, C.c_void_p,
                                    C.c_char_p, C.c_int]
        lib.ref_cell_stats.argtypes = [C.c_int32, C.c_int32, C.POINTER(OracleConfig)] + [C.c_void_p] * 8 + [C.c_char_p, C.c_int]
        lib.ref_eigen_version.restype = C.c_char_p
        _ref_full = lib
    return _ref_full


def ref_available():
    return os.path.exists(ref_full_path())


def ref_process(height, width, cfg, xyz):
    """deplex::PlaneExtractor(height, width, cfg).process(xyz) on the reference's own build.  xyz: (N,3) float32 in
    either order (converted to Eigen's column-major)."""
    lib = _load_ref_full()
    a = np.asfortranarray(np.asarray(xyz, dtype=np.float32))
    labels = np.empty(max(a.shape[0], 1), dtype=np.int32)
    err = C.create_string_buffer(512)
    rc = lib.ref_process(height, width, C.byref(cfg), a.ctypes.data if a.size else None, a.shape[0], labels.ctypes.data, err, 512)
    if rc:
        raise OracleError(rc, err.value.decode())
    return labels[: a.shape[0]]


def ref_cell_stats(height, width, cfg, xyz):
    """Per-cell coord_sum_, variance_, normal, mse, score, d, planar of the reference's own CellGrid for one frame."""
    lib = _load_ref_full()
    a = np.asfortranarray(np.asarray(xyz, dtype=np.float32))
    p = max(cfg.patch_size, 1)
    nc = max((width // p) * (height // p), 1)
    out = {"cell_sum": np.zeros((nc, 3), np.float32), "cell_var": np.zeros((nc, 9), np.float32),
           "cell_normal": np.zeros((nc, 3), np.float32), "cell_mse": np.zeros(nc, np.float32),
           "cell_score": np.zeros(nc, np.float32), "cell_d": np.zeros(nc, np.float32), "cell_planar": np.zeros(nc, np.uint8)}
    err = C.create_string_buffer(512)
    rc = lib.ref_cell_stats(height, width, C.byref(cfg), a.ctypes.data, *[out[k].ctypes.data for k in
                            ("cell_sum", "cell_var", "cell_normal", "cell_mse", "cell_score", "cell_d", "cell_planar")], err, 512)
    if rc:
        raise OracleError(rc, err.value.decode())
    return out


def ref_process_batch(height, width, cfg, xyz_batch, layout, n_threads=1):
    """Frame-parallel run of the reference build (one PlaneExtractor per call; ctypes drops the GIL): the CPU baseline
    of bench.py when oracle/_ref/libdeplex_ref.so exists (cpu_baseline.kind = "reference")."""
    from concurrent.futures import ThreadPoolExecutor
    f = xyz_batch.shape[0]
    n = height * width
    frames = [xyz_batch[i].reshape(n, 3) if layout == LAYOUT_ROWMAJOR else xyz_batch[i].reshape(3, n).T for i in range(f)]
    with ThreadPoolExecutor(max(1, n_threads)) as pool:
        return np.stack(list(pool.map(lambda x: ref_process(height, width, cfg, x), frames)))


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        lib = C.CDLL(_LIB_PATH)
        lib.dpxo_config_default.argtypes = [C.POINTER(OracleConfig)]
        lib.dpxo_config_load_ini.argtypes = [C.c_char_p, C.POINTER(OracleConfig), C.c_char_p, C.c_int]
        lib.dpxo_process.argtypes = [C.c_int32, C.c_int32, C.POINTER(OracleConfig), C.c_void_p, C.c_int64,
                                     C.c_int, C.c_void_p, C.POINTER(_Debug), C.c_char_p, C.c_int]
        lib.dpxo_process_batch.argtypes = [C.c_int32, C.c_int32, C.POINTER(OracleConfig), C.c_void_p, C.c_int32,
                                           C.c_int, C.c_void_p, C.c_int32, C.c_char_p, C.c_int]
        lib.dpxo_eig3.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.dpxo_set_sum_variant.argtypes = [C.c_int]
        lib.dpxo_set_sum_variant.restype = None
        lib.dpxo_set_uniform_int_variant.argtypes = [C.c_int]
        lib.dpxo_set_uniform_int_variant.restype = None
        lib.dpxo_uniform_selftest.argtypes = [C.c_uint32, C.c_int]
        lib.dpxo_depth_to_cloud.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float,
                                            C.c_float, C.c_void_p]
        _lib = lib
    return _lib


def set_sum_variant(variant):
    """0 = the restated Eigen 3.4 reduction orders (default, what parity is judged against); 1 = plain left-to-right
    fp32; 2 = fp64 accumulation rounded once.  Process-global; only tools/order_sensitivity.py changes it."""
    _load().dpxo_set_sum_variant(int(variant))


def set_uniform_int_variant(variant):
    """Which libstdc++ generation's std::uniform_int_distribution<int> the RANSAC refinement draws with: 0 = GCC >= 11
    (Lemire multiply-shift, default), 1 = GCC <= 10 (scaling + rejection).  Process-global."""
    _load().dpxo_set_uniform_int_variant(int(variant))


def uniform_selftest(n, draws):
    """Mismatches between the explicit variant-0 mapping and this host's std::uniform_int_distribution<int>(0, n-1)."""
    return int(_load().dpxo_uniform_selftest(int(n), int(draws)))


def load_ini(path):
    cfg = OracleConfig()
    err = C.create_string_buffer(512)
    rc = _load().dpxo_config_load_ini(os.fsencode(path), C.byref(cfg), err, 512)
    if rc:
        raise OracleError(rc, err.value.decode())
    return cfg


def _layout_of(xyz):
    """(N,3) C-order -> row-major; (N,3) F-order or (3,N) C-order -> column-major (Eigen MatrixX3f)."""
    if xyz.ndim != 2:
        raise ValueError("expected a 2-D array")
    if xyz.shape[1] == 3 and xyz.flags.c_contiguous:
        return LAYOUT_ROWMAJOR, xyz.shape[0]
    if xyz.shape[1] == 3 and xyz.flags.f_contiguous:
        return LAYOUT_COLMAJOR, xyz.shape[0]
    raise ValueError("expected a contiguous (N,3) float32 array")


_CELL_FIELDS = {
    "cell_valid": (np.uint8, 1), "cell_planar": (np.uint8, 1), "cell_sum": (np.float32, 3),
    "cell_var": (np.float32, 9), "cell_mean": (np.float32, 3), "cell_normal": (np.float32, 3),
    "cell_d": (np.float32, 1), "cell_mse": (np.float32, 1), "cell_score": (np.float32, 1),
    "cell_tol": (np.float32, 1), "cell_eval": (np.float64, 3), "cell_bin": (np.int32, 1),
    "cell_seglabel": (np.int32, 1),
}
_PLANE_FIELDS = {
    "plane_normal": (np.float32, 3), "plane_d": (np.float32, 1), "plane_mean": (np.float32, 3),
    "plane_mse": (np.float32, 1), "plane_score": (np.float32, 1), "plane_npts": (np.int32, 1),
    "merge_labels": (np.int32, 1),
}


def process(height, width, cfg, xyz, debug=False):
    """PlaneExtractor(height, width, cfg).process(xyz) on the CPU oracle.

    Returns labels (int32, N) or (labels, debug_dict) when debug=True."""
    lib = _load()
    xyz = np.asarray(xyz)
    if xyz.dtype != np.float32:
        xyz = xyz.astype(np.float32)
    if xyz.size == 0:
        layout, n = LAYOUT_ROWMAJOR, 0
    else:
        layout, n = _layout_of(xyz)
    labels = np.empty(max(n, 1), dtype=np.int32)
    err = C.create_string_buffer(512)
    dbg = None
    out = {}
    if debug:
        p = max(cfg.patch_size, 1)
        nc = (width // p) * (height // p)
        dbg = _Debug()
        for name, (dt, k) in _CELL_FIELDS.items():
            arr = np.zeros((max(nc, 1), k) if k > 1 else (max(nc, 1),), dtype=dt)
            out[name] = arr
            setattr(dbg, name, arr.ctypes.data)
        cap = max(nc, 1)
        dbg.plane_capacity = cap
        for name, (dt, k) in _PLANE_FIELDS.items():
            arr = np.zeros((cap, k) if k > 1 else (cap,), dtype=dt)
            out[name] = arr
            setattr(dbg, name, arr.ctypes.data)
        for name in ("n_planes", "n_seeds", "n_ql_fallback"):
            arr = np.zeros(1, dtype=np.int32)
            out[name] = arr
            setattr(dbg, name, arr.ctypes.data)
    rc = lib.dpxo_process(height, width, C.byref(cfg), xyz.ctypes.data if xyz.size else None, n, layout,
                          labels.ctypes.data, C.byref(dbg) if dbg is not None else None, err, 512)
    if rc:
        raise OracleError(rc, err.value.decode())
    labels = labels[:n]
    if debug:
        P = int(out["n_planes"][0])
        nc = (width // max(cfg.patch_size, 1)) * (height // max(cfg.patch_size, 1))
        for name in _CELL_FIELDS:
            out[name] = out[name][:nc]
        for name in _PLANE_FIELDS:
            out[name] = out[name][:P]
        for name in ("n_planes", "n_seeds", "n_ql_fallback"):
            out[name] = int(out[name][0])
        return labels, out
    return labels


_lib_o3 = None


def _load_o3():
    """The -O3 -DNDEBUG build of the same source (CMake Release flags): what bench.py times."""
    global _lib_o3
    if _lib_o3 is None:
        path = os.path.join(_HERE, "libdeplex_oracle_o3.so")
        if not os.path.exists(path):
            build(force=True)
        lib = C.CDLL(path)
        lib.dpxo_process_batch.argtypes = _load().dpxo_process_batch.argtypes
        _lib_o3 = lib
    return _lib_o3


def process_batch(height, width, cfg, xyz_batch, layout, n_threads=1, timed_build=False):
    """Frame-parallel CPU baseline: xyz_batch is (F, N, 3) row-major or (F, 3, N) column-major float32.
    timed_build=True runs the -O3 -DNDEBUG build (identical labels, tests/test_oracle_cpu.py)."""
    lib = _load_o3() if timed_build else _load()
    xyz_batch = np.ascontiguousarray(xyz_batch, dtype=np.float32)
    f = xyz_batch.shape[0]
    n = height * width
    assert xyz_batch.size == f * n * 3
    labels = np.empty((f, n), dtype=np.int32)
    err = C.create_string_buffer(512)
    rc = lib.dpxo_process_batch(height, width, C.byref(cfg), xyz_batch.ctypes.data, f, layout, labels.ctypes.data,
                                n_threads, err, 512)
    if rc:
        raise OracleError(rc, err.value.decode())
    return labels


def eig3(a):
    """The oracle's 3x3 symmetric eigensolver restatement: returns (Q, w, took_ql_branch)."""
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(3, 3)
    q = np.zeros((3, 3), dtype=np.float64)
    w = np.zeros(3, dtype=np.float64)
    ql = _load().dpxo_eig3(a.ctypes.data, q.ctypes.data, w.ctypes.data)
    return q, w, bool(ql)


def depth_to_cloud(depth_u16, fx, fy, cx, cy):
    """DepthImage::toPointCloud restated: (H,W) uint16 -> (N,3) float32 row-major."""
    depth_u16 = np.ascontiguousarray(depth_u16, dtype=np.uint16)
    h, w = depth_u16.shape
    out = np.empty((h * w, 3), dtype=np.float32)
    _load().dpxo_depth_to_cloud(depth_u16.ctypes.data, h, w, fx, fy, cx, cy, out.ctypes.data)
    return out
