"""Wheel of the drop-in `deplex` package (the reference packages its module the same way: python/setup.py:1-60).

    make -C ../csrc && make -C ../cpp && make -C ../pybind      # or: python ../../__graft_entry__.py build
    python -m pip wheel . --no-deps --no-build-isolation -w dist/

The wheel carries deplex/pybind.*.so together with libdeplex.so and libdeplex_b200.so (placed next to it; the
module's run path lists $ORIGIN for that), so an installed `import deplex` needs nothing from this source tree.
The libraries are built by the Makefiles, not by setuptools: this file only gathers them."""
import glob
import os
import shutil

from setuptools import Distribution, find_packages, setup
from setuptools.command.build_py import build_py as _build_py

HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.dirname(HERE)  # deplex_b200/
NATIVE = ["libdeplex.so", "libdeplex_b200.so"]


class build_py(_build_py):
    """Copies the prebuilt native libraries beside the extension module in the build tree."""

    def run(self):
        super().run()
        target = os.path.join(self.build_lib, "deplex")
        os.makedirs(target, exist_ok=True)
        module = glob.glob(os.path.join(HERE, "deplex", "pybind.*.so"))
        if not module:
            raise RuntimeError("deplex/pybind.*.so is missing: build deplex_b200/pybind first")
        for path in module:
            shutil.copy2(path, target)
        for name in NATIVE:
            path = os.path.join(PKG_ROOT, name)
            if not os.path.exists(path):
                raise RuntimeError(f"{name} is missing: build deplex_b200/csrc and deplex_b200/cpp first")
            shutil.copy2(path, target)


class BinaryDistribution(Distribution):
    def has_ext_modules(self):  # platform wheel, not a pure one
        return True


setup(
    name="deplex",
    version="1.0.0+b200",
    description="deplex plane extraction, B200 (sm_100a) implementation behind the reference Python API",
    packages=find_packages(where=HERE),
    package_dir={"": "."},
    install_requires=["numpy"],
    zip_safe=False,
    cmdclass={"build_py": build_py},
    distclass=BinaryDistribution,
)
