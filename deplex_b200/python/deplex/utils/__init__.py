"""deplex.utils: the names the reference package exposes here (python/deplex/utils/__init__.py), bound to the compiled
module deplex.pybind (deplex_b200/pybind/deplex_pybind.cpp)."""
from deplex.pybind.utils import DepthImage

__all__ = ["DepthImage"]
