"""Host-side mirror of the reference's Python interface, over the C-ABI.

Reference surface (cpp/pybind/plane_extraction/plane_extraction.cpp:28-37, python/deplex/__init__.py:1-2):
    deplex.Config(path)
    deplex.PlaneExtractor(image_height, image_width, config=Config()).process(pcd_array) -> int32 (N,)
std::runtime_error surfaces as RuntimeError, with the reference's texts.  Additions beyond the reference:
keyword construction of Config, batched / device-resident processing, and the per-cell / per-plane tables.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import LAYOUT_COLMAJOR, LAYOUT_ROWMAJOR


class UnsupportedError(RuntimeError):
    """Input that is undefined behaviour in the reference or outside the parity domain (DPX_ERR_UNSUPPORTED)."""


class CudaError(RuntimeError):
    """CUDA failure, including 'no device': there is no CPU path."""


def _raise(status, msg):
    msg = msg.decode() if isinstance(msg, bytes) else msg
    if status == _capi.DPX_ERR_RUNTIME:
        raise RuntimeError(msg)
    if status == _capi.DPX_ERR_UNSUPPORTED:
        raise UnsupportedError(msg)
    if status == _capi.DPX_ERR_CUDA:
        raise CudaError(msg)
    raise ValueError(msg or f"dpx status {status}")


class Config:
    """deplex.Config: parameters of the plane extraction algorithm (config.h:29-82).

    Config(path) reads an .ini exactly like the reference; Config() holds the defaults; keyword arguments
    (e.g. Config(patch_size=4)) override single fields."""

    def __init__(self, path=None, **fields):
        lib = _capi.load()
        self._c = _capi.dpx_config()
        if path is None:
            lib.dpx_config_default(C.byref(self._c))
        else:
            st = lib.dpx_config_load_ini(str(path).encode(), C.byref(self._c))
            if st != _capi.DPX_OK:
                _raise(st, lib.dpx_last_error(None))
        for k, v in fields.items():
            setattr(self, k, v)

    def copy(self):
        c = Config()
        C.memmove(C.byref(c._c), C.byref(self._c), C.sizeof(self._c))
        return c

    def as_dict(self):
        return {n: getattr(self._c, n) for n, _ in self._c._fields_}

    def __repr__(self):
        return "Config(" + ", ".join(f"{k}={v!r}" for k, v in self.as_dict().items()) + ")"


def _cfg_property(name):
    def get(self):
        return getattr(self._c, name)

    def set_(self, v):
        setattr(self._c, name, v)

    return property(get, set_)


for _n, _ in _capi.dpx_config._fields_:
    setattr(Config, _n, _cfg_property(_n))


def _host_layout(a):
    if a.ndim == 2 and a.shape[1] == 3 and a.flags.c_contiguous:
        return LAYOUT_ROWMAJOR
    if a.ndim == 2 and a.shape[1] == 3 and a.flags.f_contiguous:
        return LAYOUT_COLMAJOR
    return None


class PlaneExtractor:
    """deplex.PlaneExtractor (plane_extractor.h:28-56) running on one B200.

    PlaneExtractor(image_height, image_width, config=Config()).process(pcd_array) is the reference call;
    `max_batch` / `device` size the device scratch for the batched entry points."""

    def __init__(self, image_height, image_width, config=None, *, max_batch=1, device=-1):
        lib = _capi.load()
        self._lib = lib
        self._h = C.c_void_p()
        cfg = config if config is not None else Config()
        st = lib.dpx_create(int(image_height), int(image_width), C.byref(cfg._c), int(device), int(max_batch),
                            C.byref(self._h))
        if st != _capi.DPX_OK:
            self._h = C.c_void_p()
            _raise(st, lib.dpx_last_error(None))
        info = _capi.dpx_info()
        lib.dpx_get_info(self._h, C.byref(info))
        self.info = info
        self.height, self.width = int(image_height), int(image_width)
        self.n_points = self.height * self.width

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.dpx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st != _capi.DPX_OK:
            _raise(st, self._lib.dpx_last_error(self._h))

    # ---- the reference call -------------------------------------------------------------------------
    def process(self, pcd_array):
        """One organized cloud, any numeric (N,3) array (converted to float32 like the pybind Eigen caster).
        Returns int32 labels of shape (N,): 0 = non-planar, k > 0 = plane id (not compacted)."""
        a = np.asarray(pcd_array)
        # pybind11's Eigen::MatrixX3f caster (cpp/pybind/plane_extraction/plane_extraction.cpp:36) takes an (N, 3)
        # array and nothing else: a (3, N), flat or (H, W, 3) array is a TypeError there, not a silent reshape
        if a.ndim != 2 or a.shape[1] != 3:
            raise TypeError(f"process(): pcd_array must have shape (N, 3), got {a.shape}")
        if a.dtype != np.float32:
            a = a.astype(np.float32)
        n = a.shape[0]
        layout = _host_layout(a) if a.size else LAYOUT_ROWMAJOR
        if layout is None:
            a = np.ascontiguousarray(a)
            layout = LAYOUT_ROWMAJOR
        labels = np.empty(max(n, 0), dtype=np.int32)
        self._check(self._lib.dpx_process_host(self._h, a.ctypes.data if a.size else None, n, layout,
                                               labels.ctypes.data if labels.size else None))
        return labels

    # ---- batched entry points (additions) -----------------------------------------------------------
    def process_batch_host(self, xyz, layout, labels=None):
        """xyz: host float32 array holding F frames back to back ((F,N,3) row-major or (F,3,N) column-major).
        Copies are chunked and overlapped with the kernels.  Returns (F,N) int32."""
        a = np.ascontiguousarray(xyz, dtype=np.float32)
        f = a.size // (3 * self.n_points) if self.n_points else 0
        assert a.size == f * 3 * self.n_points, "batch does not hold a whole number of frames"
        if labels is None:
            labels = np.empty((f, self.n_points), dtype=np.int32)
        self._check(self._lib.dpx_process_batch_host(self._h, a.ctypes.data, f, layout, labels.ctypes.data))
        return labels

    def process_batch_host_ptr(self, xyz_ptr, n_frames, layout, labels_ptr):
        """Raw host pointers (e.g. pinned torch tensors' data_ptr())."""
        self._check(self._lib.dpx_process_batch_host(self._h, xyz_ptr, n_frames, layout, labels_ptr))

    def process_batch_device(self, xyz, layout, labels=None, stream=None):
        """xyz: CUDA torch.float32 tensor with F frames ((F,N,3) or (F,3,N)); asynchronous on `stream`
        (default: torch's current stream).  Returns a CUDA int32 tensor (F,N)."""
        import torch
        assert xyz.is_cuda and xyz.dtype == torch.float32 and xyz.is_contiguous()
        f = xyz.numel() // (3 * self.n_points)
        assert xyz.numel() == f * 3 * self.n_points
        if labels is None:
            labels = torch.empty((f, self.n_points), dtype=torch.int32, device=xyz.device)
        if stream is None:
            stream = torch.cuda.current_stream(xyz.device)
        self._check(self._lib.dpx_process_batch_device(self._h, xyz.data_ptr(), f, layout, labels.data_ptr(),
                                                       stream.cuda_stream))
        return labels

    # ---- raw depth in (DepthImage::toPointCloud evaluated on the device, depth_image.cpp:55-78) ------------
    @staticmethod
    def _intrinsics(k):
        """dict(fx, fy, cx, cy) or a 3x3 matrix -> dpx_intrinsics (values rounded to float32 like Eigen::Matrix3f)."""
        if isinstance(k, dict):
            return _capi.dpx_intrinsics(k["fx"], k["fy"], k["cx"], k["cy"])
        m = np.asarray(k, dtype=np.float32)
        return _capi.dpx_intrinsics(m[0, 0], m[1, 1], m[0, 2], m[1, 2])

    def process_depth_batch_host(self, depth, intrinsics, labels=None):
        """depth: host uint16 array of F frames (F,H,W).  Returns (F,N) int32 labels, identical to process() on the
        clouds DepthImage.transform_to_pcd would produce; only 2 bytes per pixel cross PCIe."""
        d = np.ascontiguousarray(depth, dtype=np.uint16)
        f = d.size // self.n_points if self.n_points else 0
        assert d.size == f * self.n_points, "batch does not hold a whole number of frames"
        if labels is None:
            labels = np.empty((f, self.n_points), dtype=np.int32)
        k = self._intrinsics(intrinsics)
        self._check(self._lib.dpx_process_depth_batch_host(self._h, d.ctypes.data, f, C.byref(k), labels.ctypes.data))
        return labels

    def process_depth_batch_host_ptr(self, depth_ptr, n_frames, intrinsics, labels_ptr, labels_u16=False):
        """Raw host pointers; labels_u16=True: labels_ptr receives uint16 labels (dpx_process_depth_batch_host_u16)."""
        k = self._intrinsics(intrinsics)
        fn = self._lib.dpx_process_depth_batch_host_u16 if labels_u16 else self._lib.dpx_process_depth_batch_host
        self._check(fn(self._h, depth_ptr, n_frames, C.byref(k), labels_ptr))

    def process_batch_host_u16(self, xyz, layout):
        """As process_batch_host, returning uint16 labels (same values; 2 B/pixel over PCIe and in host memory)."""
        a = np.ascontiguousarray(xyz, dtype=np.float32)
        f = a.size // (3 * self.n_points) if self.n_points else 0
        assert a.size == f * 3 * self.n_points, "batch does not hold a whole number of frames"
        labels = np.empty((f, self.n_points), dtype=np.uint16)
        self._check(self._lib.dpx_process_batch_host_u16(self._h, a.ctypes.data, f, layout, labels.ctypes.data))
        return labels

    def process_depth_batch_host_u16(self, depth, intrinsics):
        d = np.ascontiguousarray(depth, dtype=np.uint16)
        f = d.size // self.n_points if self.n_points else 0
        assert d.size == f * self.n_points, "batch does not hold a whole number of frames"
        labels = np.empty((f, self.n_points), dtype=np.uint16)
        k = self._intrinsics(intrinsics)
        self._check(self._lib.dpx_process_depth_batch_host_u16(self._h, d.ctypes.data, f, C.byref(k), labels.ctypes.data))
        return labels

    def process_depth_batch_device(self, depth, intrinsics, labels=None, stream=None):
        """depth: CUDA int16/uint16 tensor of F frames; asynchronous on `stream`."""
        import torch
        assert depth.is_cuda and depth.element_size() == 2 and depth.is_contiguous()
        f = depth.numel() // self.n_points
        if labels is None:
            labels = torch.empty((f, self.n_points), dtype=torch.int32, device=depth.device)
        if stream is None:
            stream = torch.cuda.current_stream(depth.device)
        k = self._intrinsics(intrinsics)
        self._check(self._lib.dpx_process_depth_batch_device(self._h, depth.data_ptr(), f, C.byref(k), labels.data_ptr(),
                                                             stream.cuda_stream))
        return labels

    # ---- tables the reference computes and discards --------------------------------------------------
    def cells(self, frame=0):
        n = self.info.n_cells
        buf = (_capi.dpx_cell * max(n, 1))()
        self._check(self._lib.dpx_get_cells(self._h, frame, buf, n))
        rec = np.frombuffer(buf, dtype=np.dtype([
            ("sum", "f4", 3), ("var", "f4", 6), ("mean", "f4", 3), ("normal", "f4", 3), ("d", "f4"), ("mse", "f4"),
            ("score", "f4"), ("merge_tolerance", "f4"), ("bin", "i4"), ("valid", "i4"), ("planar", "i4"),
            ("seg_label", "i4"), ("final_label", "i4")]))[:n]
        return rec.copy()

    def planes(self, frame=0):
        cap = self.info.plane_capacity
        buf = (_capi.dpx_plane * max(cap, 1))()
        n = C.c_int32(0)
        self._check(self._lib.dpx_get_planes(self._h, frame, buf, cap, C.byref(n)))
        rec = np.frombuffer(buf, dtype=np.dtype([
            ("normal", "f4", 3), ("d", "f4"), ("mean", "f4", 3), ("mse", "f4"), ("score", "f4"), ("n_points", "i4"),
            ("merge_label", "i4")]))[:n.value]
        return rec.copy()

    def seed_order(self, frame=0):
        """(n_cells,) uint64 keys [bin:15][MSE order:32][cell:17] of the frame's cells in seed order (dpx_get_seed_order)."""
        n = self.info.n_cells
        out = np.empty(max(n, 1), dtype=np.uint64)
        self._check(self._lib.dpx_get_seed_order(self._h, frame, out.ctypes.data, n))
        return out[:n]

    def refine_work(self, frame=0):
        """(point_passes, rounds) of the refinement stage on one frame of the last batch (dpx_get_refine_work): every point
        pass reads 12 bytes and scores 128 hypotheses."""
        pp, rr = C.c_uint64(0), C.c_uint64(0)
        self._check(self._lib.dpx_get_refine_work(self._h, frame, C.byref(pp), C.byref(rr)))
        return int(pp.value), int(rr.value)

    # ---- measurement ----------------------------------------------------------------------------------
    def set_profiling(self, enabled):
        self._check(self._lib.dpx_set_profiling(self._h, 1 if enabled else 0))

    def stage_ms(self):
        ms = (C.c_float * _capi.N_STAGES)()
        self._check(self._lib.dpx_get_stage_ms(self._h, C.byref(ms)))
        return dict(zip(_capi.STAGE_NAMES, [float(x) for x in ms]))

    def region_profile(self, frame=0):
        """Cycle counters of the region-growing stage for one frame of the last profiled batch."""
        buf = (C.c_int64 * 12)()
        self._check(self._lib.dpx_get_region_profile(self._h, frame, C.byref(buf)))
        names = ("total", "init", "seed", "bfs", "wide", "fit", "merge", "final", "n_seeds", "n_steps",
                 "n_regions", "n_segments")
        return dict(zip(names, [int(x) for x in buf]))

    def kernel_launches(self):
        return int(self._lib.dpx_kernel_launches(self._h))

    def set_label_transport(self, mode):
        """'auto' | 'i32' | 'u16': how the batched host entry points bring the labels back over PCIe (the result in
        the caller's buffer is int32 either way; dpx_set_label_transport)."""
        m = {"auto": _capi.LABELS_AUTO, "i32": _capi.LABELS_I32, "u16": _capi.LABELS_U16}[mode]
        self._check(self._lib.dpx_set_label_transport(self._h, m))

    def set_rng_compat(self, which):
        """'libstdc++11' (default) | 'libstdc++10': which std::uniform_int_distribution mapping the RANSAC refinement's
        sampling reproduces (dpx_set_rng_compat)."""
        self._check(self._lib.dpx_set_rng_compat(self._h, {"libstdc++11": 0, "libstdc++10": 1}[which]))

    @classmethod
    def _borrow(cls, handle, height, width):
        """A view of an extractor owned by someone else (a pipeline lane): never destroyed from here."""
        self = cls.__new__(cls)
        self._lib = _capi.load()
        self._h = C.c_void_p()  # close() / __del__ must not destroy the borrowed handle
        self._borrowed = C.c_void_p(handle)
        info = _capi.dpx_info()
        self._lib.dpx_get_info(self._borrowed, C.byref(info))
        self.info = info
        self.height, self.width = int(height), int(width)
        self.n_points = self.height * self.width
        self._h = self._borrowed
        self.close = lambda: None
        return self


class PipelinedExtractor:
    """Several batches in flight on one GPU: a thin owner of a C-ABI `dpx_pipeline` (include/deplex_b200.h), which
    holds `lanes` extractors, each with its own device tables and CUDA stream, and deals the submitted batches to them
    round-robin (an addition; the reference's caller loops over process(), examples/process_sequence.cpp:30-43).
    Within one batch the HBM-bound cell-stats kernel and the latency-bound region growing run back to back, and the
    region-growing kernel's tail -- a few long frames on a few SMs -- leaves most of the GPU idle; with several
    batches in flight the next batch's cell-stats CTAs start on every SM region growing has already left.

    submit*() is asynchronous: the lane's stream first waits for the work already queued on torch's current stream (the
    producer of the input), and nothing after the call on the current stream is ordered behind the batch until join().
    Results do not depend on the lane count."""

    def __init__(self, image_height, image_width, config=None, *, max_batch=1, device=-1, lanes=3):
        lib = _capi.load()
        self._lib = lib
        self._p = C.c_void_p()
        cfg = config if config is not None else Config()
        st = lib.dpx_pipeline_create(int(image_height), int(image_width), C.byref(cfg._c), int(device), int(max_batch),
                                     int(lanes), C.byref(self._p))
        if st != _capi.DPX_OK:
            self._p = C.c_void_p()
            _raise(st, lib.dpx_pipeline_last_error(None))
        self.lanes = [PlaneExtractor._borrow(lib.dpx_pipeline_lane(self._p, i), image_height, image_width)
                      for i in range(lib.dpx_pipeline_lanes(self._p))]
        self.info = self.lanes[0].info
        self.n_points = self.lanes[0].n_points

    def _check(self, st):
        if st != _capi.DPX_OK:
            _raise(st, self._lib.dpx_pipeline_last_error(self._p))

    def _current_stream(self, tensor):
        import torch
        return torch.cuda.current_stream(tensor.device).cuda_stream

    def submit(self, xyz, layout, labels=None):
        """Queue one batch of F <= max_batch device-resident frames; returns the (F,N) int32 CUDA tensor the labels
        will be in after join().  Keep `xyz` and the result alive until then."""
        import torch
        assert xyz.is_cuda and xyz.dtype == torch.float32 and xyz.is_contiguous()
        f = xyz.numel() // (3 * self.n_points)
        assert xyz.numel() == f * 3 * self.n_points
        if labels is None:
            labels = torch.empty((f, self.n_points), dtype=torch.int32, device=xyz.device)
        self._check(self._lib.dpx_pipeline_submit_device(self._p, xyz.data_ptr(), f, layout, labels.data_ptr(),
                                                         self._current_stream(xyz)))
        return labels

    def submit_ptr(self, xyz_ptr, n_frames, layout, labels_ptr, stream_ptr):
        """Raw device pointers (a slice of a larger resident buffer) and a raw cudaStream_t."""
        self._check(self._lib.dpx_pipeline_submit_device(self._p, xyz_ptr, n_frames, layout, labels_ptr, stream_ptr))

    def submit_depth(self, depth, intrinsics, labels=None):
        """As submit(), for raw uint16 depth frames (dpx_pipeline_submit_depth_device)."""
        import torch
        assert depth.is_cuda and depth.element_size() == 2 and depth.is_contiguous()
        f = depth.numel() // self.n_points
        if labels is None:
            labels = torch.empty((f, self.n_points), dtype=torch.int32, device=depth.device)
        k = PlaneExtractor._intrinsics(intrinsics)
        self._check(self._lib.dpx_pipeline_submit_depth_device(self._p, depth.data_ptr(), f, C.byref(k), labels.data_ptr(),
                                                               self._current_stream(depth)))
        return labels

    def join(self, stream=None):
        """Order `stream` (default: torch's current stream) behind every batch submitted so far (device-side wait)."""
        import torch
        if stream is None:
            stream = torch.cuda.current_stream(torch.device("cuda", self.info.device))
        self._check(self._lib.dpx_pipeline_join(self._p, stream.cuda_stream))

    def synchronize(self):
        self._check(self._lib.dpx_pipeline_synchronize(self._p))

    def kernel_launches(self):
        return int(self._lib.dpx_pipeline_kernel_launches(self._p))

    def close(self):
        if getattr(self, "_p", None) and self._p.value:
            self._lib.dpx_pipeline_destroy(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SequenceExtractor:
    """A frame sequence sharded over the GPUs of one box from ONE process: a thin owner of a C-ABI `dpx_sequence`
    (contiguous frame ranges, one worker thread per device, no inter-GPU traffic).  Replaces the loop of
    examples/process_sequence.cpp:30-43; labels are identical to process() on every frame in order."""

    def __init__(self, image_height, image_width, config=None, *, devices=None, max_batch=64):
        lib = _capi.load()
        self._lib = lib
        self._s = C.c_void_p()
        cfg = config if config is not None else Config()
        n = len(devices) if devices else 0
        arr = (C.c_int32 * max(n, 1))(*(devices or [0]))
        st = lib.dpx_sequence_create(int(image_height), int(image_width), C.byref(cfg._c), arr if n else None, n,
                                     int(max_batch), C.byref(self._s))
        if st != _capi.DPX_OK:
            self._s = C.c_void_p()
            _raise(st, lib.dpx_sequence_last_error(None))
        self.n_devices = int(lib.dpx_sequence_devices(self._s))
        self.n_points = int(image_height) * int(image_width)

    def _check(self, st):
        if st != _capi.DPX_OK:
            _raise(st, self._lib.dpx_sequence_last_error(self._s))

    def frame_range(self, n_frames, slot):
        b, e = C.c_int64(0), C.c_int64(0)
        self._lib.dpx_sequence_range(self._s, int(n_frames), int(slot), C.byref(b), C.byref(e))
        return b.value, e.value

    def process_host(self, xyz, layout, labels=None):
        a = np.ascontiguousarray(xyz, dtype=np.float32)
        f = a.size // (3 * self.n_points) if self.n_points else 0
        assert a.size == f * 3 * self.n_points, "sequence does not hold a whole number of frames"
        if labels is None:
            labels = np.empty((f, self.n_points), dtype=np.int32)
        self._check(self._lib.dpx_sequence_process_host(self._s, a.ctypes.data, f, layout, labels.ctypes.data))
        return labels

    def process_depth_host(self, depth, intrinsics, labels=None):
        d = np.ascontiguousarray(depth, dtype=np.uint16)
        f = d.size // self.n_points if self.n_points else 0
        assert d.size == f * self.n_points, "sequence does not hold a whole number of frames"
        if labels is None:
            labels = np.empty((f, self.n_points), dtype=np.int32)
        k = PlaneExtractor._intrinsics(intrinsics)
        self._check(self._lib.dpx_sequence_process_depth_host(self._s, d.ctypes.data, f, C.byref(k), labels.ctypes.data))
        return labels

    def process_host_ptr(self, xyz_ptr, n_frames, layout, labels_ptr):
        self._check(self._lib.dpx_sequence_process_host(self._s, xyz_ptr, int(n_frames), layout, labels_ptr))

    def process_depth_host_ptr(self, depth_ptr, n_frames, intrinsics, labels_ptr):
        k = PlaneExtractor._intrinsics(intrinsics)
        self._check(self._lib.dpx_sequence_process_depth_host(self._s, depth_ptr, int(n_frames), C.byref(k), labels_ptr))

    def close(self):
        if getattr(self, "_s", None) and self._s.value:
            self._lib.dpx_sequence_destroy(self._s)
            self._s = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
