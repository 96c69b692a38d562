"""ctypes binding of the C++ host mirror (gid_b200/host, libgidhost.so).

The mirror offers GID's own interface for the JPEG path -- Load_Image_Header /
Load_Image_Contents with Set_X_Y / Put_Pixel / Feedback -- above the C ABI of
libgidcuda.so.  The flat entry points bound here wrap it for Python callers and
tests: header probing, full decode into a top-down RGB buffer (the client of
test/benchmark.adb:57-108) and entropy-decode-only into coefficient planes.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass

import numpy as np

from . import capi

_HERE = os.path.dirname(os.path.abspath(__file__))

# status codes of gidhost_* = the exceptions of gid.ads:73-77
ERRORS = {
    -1: "unknown_image_format", -2: "known_but_unsupported_image_format",
    -3: "unsupported_image_subformat", -4: "error_in_image_data",
    -5: "invalid_primary_color_range", -7: "Constraint_Error", -8: "bad_argument",
    -9: "cuda_device_error",
}
EXPORTS = ["gidhost_probe", "gidhost_decode_rgb", "gidhost_decode_planes", "gidhost_free"]


class GidHostError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code
        self.name = ERRORS.get(code, str(code))
        self.msg = msg


def library_path() -> str:
    return os.path.join(_HERE, "lib", "libgidhost.so")


_lib = None


def load_library() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    capi.load_library()  # libgidcuda.so first (same directory; also found through the rpath)
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `python -m gid_b200.build`")
    L = ctypes.CDLL(path)
    i32p, vp = ctypes.POINTER(ctypes.c_int32), ctypes.c_void_p
    L.gidhost_probe.argtypes = [vp, ctypes.c_size_t, i32p, i32p, i32p, i32p, i32p, ctypes.c_char_p, ctypes.c_size_t]
    L.gidhost_decode_rgb.argtypes = [vp, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint32, vp, ctypes.c_size_t,
                                     i32p, ctypes.c_int32, i32p, ctypes.c_char_p, ctypes.c_size_t]
    L.gidhost_decode_planes.argtypes = [vp, ctypes.c_size_t, ctypes.POINTER(capi.FrameDesc),
                                        ctypes.POINTER(vp), i32p, i32p, ctypes.c_char_p, ctypes.c_size_t]
    L.gidhost_free.argtypes = [vp]
    L.gidhost_set_option.argtypes = [ctypes.c_char_p, ctypes.c_int]
    L.gidhost_set_option.restype = ctypes.c_int
    L.gidhost_free.restype = None
    for n in EXPORTS[:3]:
        getattr(L, n).restype = ctypes.c_int
    _lib = L
    return L


@dataclass
class Header:
    width: int
    height: int
    ncomp: int
    progressive: bool
    orientation: int  # 0 Unchanged, 1 Rotation_90, 2 Rotation_180, 3 Rotation_270


def _buf(data: bytes):
    return (ctypes.c_uint8 * len(data)).from_buffer_copy(data)


def probe(data: bytes) -> Header:
    """GID.Load_Image_Header (no GPU needed)."""
    L = load_library()
    v = [ctypes.c_int32() for _ in range(5)]
    err = ctypes.create_string_buffer(512)
    rc = L.gidhost_probe(_buf(data), len(data), *[ctypes.byref(x) for x in v], err, 512)
    if rc != 0:
        raise GidHostError(rc, err.value.decode(errors="replace"))
    return Header(v[0].value, v[1].value, v[2].value, bool(v[3].value), v[4].value)


def set_option(name: str, value: int) -> None:
    """"sparse_input": 1 (default) = coefficients reach the device as records, 0 = as dense planes."""
    rc = load_library().gidhost_set_option(name.encode(), value)
    if rc != 0:
        raise GidHostError(rc, f"unknown option {name}")


def decode_rgb(data: bytes, device: int = 0, modulus: int = 256):
    """Load_Image_Header + Load_Image_Contents on the GPU path.  Returns (rgb (H, W, 3) uint8,
    list of Feedback values)."""
    L = load_library()
    h = probe(data)
    rgb = np.zeros((h.height, h.width, 3), dtype=np.uint8)
    cap = 4 * h.height + 4096
    log = (ctypes.c_int32 * cap)()
    n = ctypes.c_int32()
    err = ctypes.create_string_buffer(512)
    rc = L.gidhost_decode_rgb(_buf(data), len(data), device, modulus, rgb.ctypes.data, rgb.nbytes, log, cap,
                              ctypes.byref(n), err, 512)
    if rc != 0:
        raise GidHostError(rc, err.value.decode(errors="replace"))
    return rgb, list(log[:min(n.value, cap)])


def decode_planes(data: bytes):
    """Entropy decode only (host cores, no GPU).  Returns (FrameDesc, [plane (by, bx, 64) int16 ...])."""
    L = load_library()
    desc = capi.FrameDesc()
    planes = (ctypes.c_void_p * 3)()
    bx, by = (ctypes.c_int32 * 3)(), (ctypes.c_int32 * 3)()
    err = ctypes.create_string_buffer(512)
    rc = L.gidhost_decode_planes(_buf(data), len(data), ctypes.byref(desc), planes, bx, by, err, 512)
    if rc != 0:
        raise GidHostError(rc, err.value.decode(errors="replace"))
    out = []
    for c in range(desc.ncomp):
        n = bx[c] * by[c] * 64
        a = np.frombuffer((ctypes.c_int16 * n).from_address(planes[c]), dtype=np.int16).reshape(by[c], bx[c], 64).copy()
        out.append(a)
        L.gidhost_free(planes[c])
    return desc, out
