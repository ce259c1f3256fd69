"""Builds the `torch_ext` module (same name as the reference's extensions/torch/setup.py:62-71)
against the C-ABI library.  KERNEL=<name> (reference convention, setup.py:10) only selects the
default variant of solve(); flash_solve(kernel=...) dispatches at run time.

    cd extensions/torch && python setup.py build_ext --inplace
"""
import os
import subprocess

from setuptools import setup
from torch.utils.cpp_extension import BuildExtension, CUDAExtension

here = os.path.dirname(os.path.abspath(__file__))
root = os.path.abspath(os.path.join(here, "..", ".."))
libdir = os.path.join(root, "quantizedmha_b200", "lib")
if not os.path.exists(os.path.join(libdir, "libqmha.so")):
    subprocess.run(["make", "-C", root, "lib"], check=True)

setup(
    name="torch_ext",
    ext_modules=[CUDAExtension(
        name="torch_ext",
        sources=[os.path.join(here, "torch_ext.cpp")],
        include_dirs=[os.path.join(root, "include")],
        depends=[os.path.join(root, "include", "qmha.h")],   # qmha_args carries its size: a stale build fails at run time
        library_dirs=[libdir],
        libraries=["qmha"],
        runtime_library_dirs=[libdir],
    )],
    cmdclass={"build_ext": BuildExtension},
)
