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
EXPORTS = ["gidhost_probe", "gidhost_decode_rgb", "gidhost_decode_planes", "gidhost_free", "gidhost_set_option",
           "gidhost_counter", "gidhost_decode_batch", "gidhost_batch_create",
           "gidhost_batch_decode", "gidhost_batch_destroy"]


class HostBatchStats(ctypes.Structure):
    """struct gidhost_batch_stats (gid_b200/host/gid_batch.cpp)"""
    _fields_ = [("seconds", ctypes.c_double), ("megapixels", ctypes.c_double), ("jpeg_bytes", ctypes.c_int64),
                ("images_failed", ctypes.c_int32), ("threads", ctypes.c_int32), ("devices", ctypes.c_int32),
                ("entropy_threads", ctypes.c_int32)]


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
    L.gidhost_decode_batch.argtypes = [ctypes.c_int32, ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_size_t), i32p,
                                       ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(vp),
                                       ctypes.POINTER(ctypes.c_size_t), i32p, ctypes.POINTER(HostBatchStats),
                                       ctypes.c_char_p, ctypes.c_size_t]
    L.gidhost_decode_batch.restype = ctypes.c_int
    L.gidhost_batch_create.argtypes = [i32p, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(vp)]
    L.gidhost_batch_decode.argtypes = [vp, ctypes.c_int32, ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_size_t),
                                       ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_size_t), i32p,
                                       ctypes.POINTER(HostBatchStats), ctypes.c_char_p, ctypes.c_size_t]
    L.gidhost_batch_destroy.argtypes = [vp]
    for n_ in ("gidhost_batch_create", "gidhost_batch_decode", "gidhost_batch_destroy"):
        getattr(L, n_).restype = ctypes.c_int
    L.gidhost_counter.argtypes = [ctypes.c_char_p]
    L.gidhost_counter.restype = ctypes.c_long
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


def counter(name: str) -> int:
    """"parallel_scans" / "parallel_fallbacks": restart-interval-parallel baseline scans so far."""
    return int(load_library().gidhost_counter(name.encode()))


def set_option(name: str, value: int) -> None:
    """"sparse_input": 1 (default) = coefficients reach the device as records, 0 = as dense planes;
    "entropy_threads": host threads for the entropy decode of one image (1 = serial, 0 = all)."""
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


class HostBatch:
    """Many .jpg files -> top-down RGB8 arrays (gid_b200/host/gid_batch.cpp): `threads` host threads
    (0 = all) entropy-decode, the B200(s) reconstruct, three frames in flight per thread.  The
    workers' device contexts (pinned + device buffers) persist from one decode() to the next."""

    def __init__(self, devices=(0,), threads: int = 0):
        self.lib = load_library()
        devs = (ctypes.c_int32 * len(devices))(*devices)
        self.handle = ctypes.c_void_p()
        rc = self.lib.gidhost_batch_create(devs, len(devices), threads, ctypes.byref(self.handle))
        if rc != 0:
            raise GidHostError(rc, "gidhost_batch_create")

    def decode(self, files, out=None):
        """Returns (list of (H, W, 3) uint8 arrays -- None where the file failed --, list of status
        codes, HostBatchStats).  `out`: optional preallocated arrays to decode into; arrays in pinned
        memory (capi.pinned_empty) receive their pixels straight from the device."""
        n = len(files)
        if out is None:
            out = []
            for f in files:
                try:
                    h = probe(f)
                    out.append(np.empty((h.height, h.width, 3), dtype=np.uint8))
                except GidHostError:
                    out.append(np.empty((0, 0, 3), dtype=np.uint8))  # the batch reports the error class
        bufs = [(ctypes.c_uint8 * max(1, len(f))).from_buffer_copy(f.ljust(1, b"\0")) for f in files]
        data = (ctypes.c_void_p * n)(*[ctypes.addressof(b) for b in bufs])
        lens = (ctypes.c_size_t * n)(*[len(f) for f in files])
        outs = (ctypes.c_void_p * n)(*[a.ctypes.data for a in out])
        caps = (ctypes.c_size_t * n)(*[a.nbytes for a in out])
        status = (ctypes.c_int32 * n)()
        stats = HostBatchStats()
        err = ctypes.create_string_buffer(512)
        rc = self.lib.gidhost_batch_decode(self.handle, n, data, lens, outs, caps, status, ctypes.byref(stats), err, 512)
        if rc != 0:
            raise GidHostError(rc, err.value.decode(errors="replace"))
        st = list(status)
        return [a if st[i] == 0 else None for i, a in enumerate(out)], st, stats

    def close(self) -> None:
        if self.handle:
            self.lib.gidhost_batch_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def decode_batch(files, devices=(0,), threads: int = 0):
    """One-shot HostBatch.decode."""
    with HostBatch(devices, threads) as b:
        return b.decode(files)
