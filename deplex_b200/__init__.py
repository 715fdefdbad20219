"""deplex_b200 -- B200-native (sm_100a) plane extraction behind prime-slam/deplex's API.

    from deplex_b200 import Config, PlaneExtractor
    labels = PlaneExtractor(480, 640, Config("TUM_fr3_long_val.ini")).process(points)

Everything numerical runs in hand-written CUDA kernels inside libdeplex_b200.so, reached through the
C-ABI in include/deplex_b200.h.  There is no CPU fallback: importing works without a GPU (so that
configuration and host logic can be tested), creating an extractor does not.
"""
from ._capi import LAYOUT_COLMAJOR, LAYOUT_ROWMAJOR, LIB_PATH, load  # noqa: F401
from .extractor import Config, CudaError, PipelinedExtractor, PlaneExtractor, SequenceExtractor, UnsupportedError  # noqa: F401

__all__ = ["Config", "PlaneExtractor", "PipelinedExtractor", "SequenceExtractor", "UnsupportedError", "CudaError", "LAYOUT_COLMAJOR", "LAYOUT_ROWMAJOR"]
