"""gid_b200 -- B200-native JPEG pixel reconstruction behind GID's Load_Image_Contents.

The product is the C-ABI library ``gid_b200/lib/libgidcuda.so`` (CUDA kernels +
shim, see ``include/gidcuda.h``) and the C++ host mirror ``libgidhost.so``.
This Python package is a thin ctypes binding used by tests and bench.py.
"""
from .capi import (GidCuda, GidCudaError, FrameDesc, load_library, library_path)  # noqa: F401

__all__ = ["GidCuda", "GidCudaError", "FrameDesc", "load_library", "library_path"]
