from deplex.pybind.plane_extraction import *  # noqa: F401,F403
