"""deplex -- drop-in for the reference's Python package (python/deplex/__init__.py:1-2), running on a B200.

    import deplex
    labels = deplex.PlaneExtractor(image_height=480, image_width=640, config=deplex.Config(path)).process(points)

The compiled module `deplex.pybind` (deplex_b200/pybind/deplex_pybind.cpp) goes through the same C-ABI layer
as the C++ class.  Put deplex_b200/python on sys.path (or install it) to use this package name."""
from deplex.plane_extraction import *  # noqa: F401,F403
import deplex.plane_extraction  # noqa: F401
import deplex.utils  # noqa: F401
