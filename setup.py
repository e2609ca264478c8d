"""Packaging for triton_dist (B200 / sm_100a).  ``python setup.py build_ext --inplace`` (or ``pip install -e . --no-build-isolation``)
compiles the two native libraries in-tree with nvcc/g++ through triton_dist/_build.py -- the same path ``__graft_entry__.build()`` uses --
so the shared objects always live next to the package (``triton_dist/lib``) and no JIT cache is involved."""
import os
import sys

from setuptools import Command, find_packages, setup
from setuptools.command.build_ext import build_ext
from setuptools.command.build_py import build_py

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def _build_native():
    from triton_dist import _build
    host, cuda = _build.build_all(verbose=True)
    print("built:", host, cuda)


class BuildNative(build_ext):
    def run(self):
        _build_native()


class BuildPyWithNative(build_py):
    def run(self):
        _build_native()
        super().run()


setup(
    name="triton_dist_b200",
    version="0.1.0",
    description="Blackwell-native compute-communication overlap framework (tcgen05 / TMEM / TMA kernels over an NVLink symmetric heap)",
    packages=find_packages(include=["triton_dist", "triton_dist.*"]),
    package_data={"triton_dist": ["lib/*.so"]},
    python_requires=">=3.10",
    install_requires=[],          # torch is expected from the environment (no index access here)
    cmdclass={"build_ext": BuildNative, "build_py": BuildPyWithNative},
)
