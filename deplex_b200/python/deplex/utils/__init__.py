from deplex.pybind.utils import *  # noqa: F401,F403
