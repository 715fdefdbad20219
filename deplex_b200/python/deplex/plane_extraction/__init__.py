"""deplex.plane_extraction: the names the reference package exposes here (python/deplex/plane_extraction/__init__.py),
bound to the compiled module deplex.pybind (deplex_b200/pybind/deplex_pybind.cpp)."""
from deplex.pybind.plane_extraction import Config, PlaneExtractor

__all__ = ["Config", "PlaneExtractor"]
