"""Builds the `jax_ext` pybind module against libqmha.so (pure pybind11 — unlike the reference's
extensions/jax/setup.py:2 it does not need torch to build).
    cd extensions/jax && python setup.py build_ext --inplace
"""
import os
import subprocess

import pybind11
from setuptools import Extension, setup

here = os.path.dirname(os.path.abspath(__file__))
root = os.path.abspath(os.path.join(here, "..", ".."))
libdir = os.path.join(root, "quantizedmha_b200", "lib")
if not os.path.exists(os.path.join(libdir, "libqmha.so")):
    subprocess.run(["make", "-C", root, "lib"], check=True)

setup(
    name="jax_ext",
    ext_modules=[Extension(
        "jax_ext", [os.path.join(here, "jax_ext.cpp")],
        include_dirs=[pybind11.get_include(), os.path.join(root, "include")],
        library_dirs=[libdir], libraries=["qmha"], runtime_library_dirs=[libdir],
        extra_compile_args=["-O2", "-std=c++17"], language="c++")],
)
