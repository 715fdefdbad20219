"""ctypes declarations for libdeplex_b200.so (include/deplex_b200.h).  No arithmetic happens in Python."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DPX_LIB_PATH: an explicit alternative build of the same library (the debug build with the kernels' own bounds checks,
# `make -C deplex_b200/csrc debug`); there is still no fallback of any kind
LIB_PATH = os.environ.get("DPX_LIB_PATH") or os.path.join(_HERE, "libdeplex_b200.so")

DPX_OK, DPX_ERR_RUNTIME, DPX_ERR_UNSUPPORTED, DPX_ERR_CUDA, DPX_ERR_ARGUMENT = range(5)
LAYOUT_COLMAJOR, LAYOUT_ROWMAJOR = 0, 1
N_STAGES = 4
STAGE_NAMES = ("cell_stats", "region_grow", "labeling", "refine")

# every symbol include/deplex_b200.h declares (checked by tests/test_capi_cpu.py)
EXPORTS = (
    "dpx_config_default", "dpx_config_load_ini", "dpx_create", "dpx_destroy", "dpx_last_error", "dpx_get_info",
    "dpx_process_host", "dpx_process_batch_host", "dpx_process_batch_device", "dpx_process_depth_batch_host",
    "dpx_process_depth_batch_device", "dpx_get_cells", "dpx_get_planes", "dpx_get_seed_order", "dpx_get_refine_work",
    "dpx_set_profiling", "dpx_get_stage_ms", "dpx_get_region_profile", "dpx_kernel_launches", "dpx_host_alloc", "dpx_host_free", "dpx_version",
    "dpx_set_label_transport", "dpx_set_rng_compat", "dpx_process_batch_host_u16", "dpx_process_depth_batch_host_u16",
    "dpx_pipeline_create", "dpx_pipeline_destroy", "dpx_pipeline_last_error", "dpx_pipeline_lanes", "dpx_pipeline_lane",
    "dpx_pipeline_submit_device", "dpx_pipeline_submit_depth_device", "dpx_pipeline_join", "dpx_pipeline_synchronize",
    "dpx_pipeline_kernel_launches",
    "dpx_sequence_create", "dpx_sequence_destroy", "dpx_sequence_last_error", "dpx_sequence_devices", "dpx_sequence_range",
    "dpx_sequence_process_host", "dpx_sequence_process_depth_host", "dpx_sequence_process_device",
    "dpx_device_count", "dpx_device_alloc", "dpx_device_free", "dpx_memcpy_to_device", "dpx_memcpy_to_host",
)
LABELS_AUTO, LABELS_I32, LABELS_U16 = 0, 1, 2


class dpx_config(C.Structure):
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


class dpx_intrinsics(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float)]


class dpx_info(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("height", "width", "patch_size", "cells_x", "cells_y", "n_cells",
                                          "plane_capacity", "max_batch", "device", "sm_count", "fused_labeling")]


class dpx_cell(C.Structure):
    _fields_ = [
        ("sum", C.c_float * 3), ("var", C.c_float * 6), ("mean", C.c_float * 3), ("normal", C.c_float * 3),
        ("d", C.c_float), ("mse", C.c_float), ("score", C.c_float), ("merge_tolerance", C.c_float),
        ("bin", C.c_int32), ("valid", C.c_int32), ("planar", C.c_int32), ("seg_label", C.c_int32),
        ("final_label", C.c_int32),
    ]


class dpx_plane(C.Structure):
    _fields_ = [
        ("normal", C.c_float * 3), ("d", C.c_float), ("mean", C.c_float * 3), ("mse", C.c_float),
        ("score", C.c_float), ("n_points", C.c_int32), ("merge_label", C.c_int32),
    ]


_lib = None


def load():
    """Load the CUDA library.  There is no fallback: a missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C deplex_b200/csrc`).  deplex_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    sig = {
        "dpx_config_default": (None, [C.POINTER(dpx_config)]),
        "dpx_config_load_ini": (C.c_int, [C.c_char_p, C.POINTER(dpx_config)]),
        "dpx_create": (C.c_int, [i32, i32, C.POINTER(dpx_config), i32, i32, C.POINTER(vp)]),
        "dpx_destroy": (None, [vp]),
        "dpx_last_error": (C.c_char_p, [vp]),
        "dpx_get_info": (C.c_int, [vp, C.POINTER(dpx_info)]),
        "dpx_process_host": (C.c_int, [vp, vp, i64, C.c_int, vp]),
        "dpx_process_batch_host": (C.c_int, [vp, vp, i32, C.c_int, vp]),
        "dpx_process_batch_device": (C.c_int, [vp, vp, i32, C.c_int, vp, vp]),
        "dpx_process_depth_batch_host": (C.c_int, [vp, vp, i32, C.POINTER(dpx_intrinsics), vp]),
        "dpx_process_depth_batch_device": (C.c_int, [vp, vp, i32, C.POINTER(dpx_intrinsics), vp, vp]),
        "dpx_get_cells": (C.c_int, [vp, i32, C.POINTER(dpx_cell), i32]),
        "dpx_get_planes": (C.c_int, [vp, i32, C.POINTER(dpx_plane), i32, C.POINTER(i32)]),
        "dpx_get_seed_order": (C.c_int, [vp, i32, vp, i32]),
        "dpx_get_refine_work": (C.c_int, [vp, i32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
        "dpx_set_profiling": (C.c_int, [vp, i32]),
        "dpx_get_stage_ms": (C.c_int, [vp, C.POINTER(C.c_float * N_STAGES)]),
        "dpx_get_region_profile": (C.c_int, [vp, i32, C.POINTER(C.c_int64 * 12)]),
        "dpx_kernel_launches": (i64, [vp]),
        "dpx_host_alloc": (C.c_int, [C.POINTER(vp), C.c_size_t]),
        "dpx_host_free": (None, [vp]),
        "dpx_version": (i32, []),
        "dpx_set_label_transport": (C.c_int, [vp, i32]),
        "dpx_set_rng_compat": (C.c_int, [vp, i32]),
        "dpx_process_batch_host_u16": (C.c_int, [vp, vp, i32, C.c_int, vp]),
        "dpx_process_depth_batch_host_u16": (C.c_int, [vp, vp, i32, C.POINTER(dpx_intrinsics), vp]),
        "dpx_pipeline_create": (C.c_int, [i32, i32, C.POINTER(dpx_config), i32, i32, i32, C.POINTER(vp)]),
        "dpx_pipeline_destroy": (None, [vp]),
        "dpx_pipeline_last_error": (C.c_char_p, [vp]),
        "dpx_pipeline_lanes": (i32, [vp]),
        "dpx_pipeline_lane": (vp, [vp, i32]),
        "dpx_pipeline_submit_device": (C.c_int, [vp, vp, i32, C.c_int, vp, vp]),
        "dpx_pipeline_submit_depth_device": (C.c_int, [vp, vp, i32, C.POINTER(dpx_intrinsics), vp, vp]),
        "dpx_pipeline_join": (C.c_int, [vp, vp]),
        "dpx_pipeline_synchronize": (C.c_int, [vp]),
        "dpx_pipeline_kernel_launches": (i64, [vp]),
        "dpx_sequence_create": (C.c_int, [i32, i32, C.POINTER(dpx_config), C.POINTER(i32), i32, i32, C.POINTER(vp)]),
        "dpx_sequence_destroy": (None, [vp]),
        "dpx_sequence_last_error": (C.c_char_p, [vp]),
        "dpx_sequence_devices": (i32, [vp]),
        "dpx_sequence_range": (None, [vp, i64, i32, C.POINTER(i64), C.POINTER(i64)]),
        "dpx_sequence_process_host": (C.c_int, [vp, vp, i64, C.c_int, vp]),
        "dpx_sequence_process_depth_host": (C.c_int, [vp, vp, i64, C.POINTER(dpx_intrinsics), vp]),
        "dpx_sequence_process_device": (C.c_int, [vp, C.POINTER(vp), i64, C.c_int, C.POINTER(vp), i32, C.POINTER(C.c_float)]),
        "dpx_device_count": (i32, []),
        "dpx_device_alloc": (C.c_int, [i32, C.POINTER(vp), C.c_size_t]),
        "dpx_device_free": (None, [i32, vp]),
        "dpx_memcpy_to_device": (C.c_int, [i32, vp, vp, C.c_size_t]),
        "dpx_memcpy_to_host": (C.c_int, [i32, vp, vp, C.c_size_t]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
