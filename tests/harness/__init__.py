"""Test harness: ctypes bindings to the synthetic JPEG encoder (synth_jpeg.c)
and to the CPU oracle (oracle/gidref.c).

Test infrastructure only.  Nothing under ``gid_b200/`` imports this package.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_ORACLE_DIR = os.path.join(_ROOT, "oracle")


def _build(target_dir: str, lib: str, sources: list[str], flags: list[str]) -> str:
    """Compile ``lib`` in ``target_dir`` if it is missing or older than its sources."""
    out = os.path.join(target_dir, lib)
    srcs = [os.path.join(target_dir, s) for s in sources]
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in srcs):
        return out
    cmd = ["gcc", "-O2", "-fPIC", "-std=c11", "-shared", "-o", out, srcs[0], "-lm"] + flags
    subprocess.run(cmd, check=True, cwd=target_dir)
    return out


def _build_emu() -> str:
    """Host build of the kernel body (test-only; see emu_recon.cpp)."""
    out = os.path.join(_HERE, "libemurecon.so")
    csrc = os.path.join(_ROOT, "gid_b200", "csrc")
    deps = [os.path.join(_HERE, "emu_recon.cpp")] + [
        os.path.join(csrc, f) for f in ("recon_math.cuh", "recon_body.cuh", "frame_setup.hpp")
    ] + [os.path.join(_ROOT, "include", "gidcuda.h")]
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                    "-o", out, deps[0]], check=True, cwd=_HERE)
    return out


def build_all() -> None:
    _build(_ORACLE_DIR, "libgidref.so", ["gidref.c", "gidref_load.inc", "gidref.h"], ["-fwrapv"])
    _build(_HERE, "libsynthjpeg.so", ["synth_jpeg.c"], [])
    _build_emu()


# --------------------------------------------------------------------------- oracle


class _RefFrame(ctypes.Structure):
    _fields_ = [
        ("width", ctypes.c_int32), ("height", ctypes.c_int32), ("ncomp", ctypes.c_int32),
        ("samp_h", ctypes.c_int32 * 3), ("samp_v", ctypes.c_int32 * 3),
        ("blocks_x", ctypes.c_int32 * 3), ("blocks_y", ctypes.c_int32 * 3),
        ("qt", (ctypes.c_uint16 * 64) * 3),
        ("progressive", ctypes.c_int32), ("restart_interval", ctypes.c_int32),
        ("orientation", ctypes.c_int32), ("scan_count", ctypes.c_int32),
        ("coef_overflow", ctypes.c_int32),
    ]


class _RefClient(ctypes.Structure):
    _fields_ = [
        ("user", ctypes.c_void_p),
        ("set_x_y", ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_int, ctypes.c_int)),
        ("put_pixel", ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_uint, ctypes.c_uint,
                                       ctypes.c_uint, ctypes.c_uint)),
        ("feedback", ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_int)),
        ("modulus", ctypes.c_uint),
    ]


@dataclass
class Frame:
    """Geometry + quant tables of one JPEG frame (what the patched front end emits)."""
    width: int
    height: int
    ncomp: int
    samp_h: list[int]
    samp_v: list[int]
    blocks_x: list[int]
    blocks_y: list[int]
    qt: np.ndarray  # (ncomp, 64) uint16, natural order
    progressive: bool = False
    restart_interval: int = 0
    orientation: int = 0
    scan_count: int = 0
    coef_overflow: bool = False
    planes: list[np.ndarray] = field(default_factory=list)  # per component (by, bx, 64) int16

    def algorithmic_bytes(self) -> int:
        """128 * sum_c blocks_c + 3 * W * H."""
        return 128 * sum(self.blocks_x[c] * self.blocks_y[c] for c in range(self.ncomp)) \
            + 3 * self.width * self.height


class OracleError(Exception):
    def __init__(self, code: int, msg: str):
        super().__init__(f"gidref error {code}: {msg}")
        self.code = code
        self.msg = msg


class Oracle:
    """CPU restatement of GID's JPEG path (oracle/gidref.c)."""

    ERRORS = {
        -1: "unknown_image_format", -2: "known_but_unsupported_image_format",
        -3: "unsupported_image_subformat", -4: "error_in_image_data",
        -5: "invalid_primary_color_range", -6: "End_Error", -7: "Constraint_Error",
        -8: "bad_argument",
    }

    def __init__(self) -> None:
        build_all()
        self.lib = ctypes.CDLL(os.path.join(_ORACLE_DIR, "libgidref.so"))
        L = self.lib
        L.gidref_probe.restype = ctypes.c_int
        L.gidref_decode_rgb.restype = ctypes.c_int
        L.gidref_decode_client.restype = ctypes.c_int
        L.gidref_reconstruct.restype = ctypes.c_int
        L.gidref_reconstruct.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_size_t]
        L.gidref_free.argtypes = [ctypes.c_void_p]
        L.gidref_zig_zag.restype = ctypes.POINTER(ctypes.c_int32)
        L.gidref_undo_zig_zag.restype = ctypes.POINTER(ctypes.c_int32)
        L.gidref_version.restype = ctypes.c_char_p

    @staticmethod
    def _frame_from_c(f: _RefFrame) -> Frame:
        n = f.ncomp if f.ncomp in (1, 3) else 0
        return Frame(
            width=f.width, height=f.height, ncomp=f.ncomp,
            samp_h=[f.samp_h[c] for c in range(n)], samp_v=[f.samp_v[c] for c in range(n)],
            blocks_x=[f.blocks_x[c] for c in range(n)], blocks_y=[f.blocks_y[c] for c in range(n)],
            qt=np.array([[f.qt[c][i] for i in range(64)] for c in range(n)], dtype=np.uint16).reshape(n, 64),
            progressive=bool(f.progressive), restart_interval=f.restart_interval,
            orientation=f.orientation, scan_count=f.scan_count,
            coef_overflow=bool(f.coef_overflow))

    @staticmethod
    def _frame_to_c(fr: Frame) -> _RefFrame:
        f = _RefFrame()
        f.width, f.height, f.ncomp = fr.width, fr.height, fr.ncomp
        for c in range(fr.ncomp):
            f.samp_h[c], f.samp_v[c] = fr.samp_h[c], fr.samp_v[c]
            f.blocks_x[c], f.blocks_y[c] = fr.blocks_x[c], fr.blocks_y[c]
            for i in range(64):
                f.qt[c][i] = int(fr.qt[c][i])
        return f

    def probe(self, data: bytes) -> Frame:
        f = _RefFrame()
        err = ctypes.create_string_buffer(256)
        rc = self.lib.gidref_probe(data, ctypes.c_size_t(len(data)), ctypes.byref(f), err, ctypes.c_size_t(256))
        if rc != 0:
            raise OracleError(rc, err.value.decode())
        return self._frame_from_c(f)

    def decode(self, data: bytes, want_planes: bool = True):
        """Full GID decode.  Returns (frame, rgb[H, W, 3] uint8, feedback list)."""
        fr = self.probe(data)
        rgb = np.zeros((fr.height, fr.width, 3), dtype=np.uint8)
        f = _RefFrame()
        err = ctypes.create_string_buffer(256)
        cap = 1 << 16
        fb = (ctypes.c_int * cap)()
        nfb = ctypes.c_int(0)
        planes = (ctypes.POINTER(ctypes.c_int16) * 3)()
        rc = self.lib.gidref_decode_rgb(
            data, ctypes.c_size_t(len(data)), rgb.ctypes.data_as(ctypes.c_void_p), ctypes.byref(f),
            planes if want_planes else None, fb, ctypes.c_int(cap), ctypes.byref(nfb), err,
            ctypes.c_size_t(256))
        if rc != 0:
            raise OracleError(rc, err.value.decode())
        out = self._frame_from_c(f)
        if want_planes:
            for c in range(out.ncomp):
                n = out.blocks_x[c] * out.blocks_y[c] * 64
                arr = np.ctypeslib.as_array(planes[c], shape=(n,)).copy()
                out.planes.append(arr.reshape(out.blocks_y[c], out.blocks_x[c], 64))
                self.lib.gidref_free(planes[c])
        feedback = [fb[i] for i in range(min(nfb.value, cap))]
        return out, rgb, feedback

    def decode_client(self, data: bytes, modulus: int = 256):
        """Full GID decode through function-pointer callbacks; returns the call log
        as (frame, image[H, W, 4] uint32 filled through Set_X_Y/Put_Pixel, feedback)."""
        fr = self.probe(data)
        img = np.zeros((fr.height, fr.width, 4), dtype=np.uint32)
        state = {"x": 0, "y": 0}
        feedback: list[int] = []

        def set_x_y(_u, x, y):
            state["x"], state["y"] = x, y

        def put_pixel(_u, r, g, b, a):
            img[fr.height - 1 - state["y"], state["x"]] = (r, g, b, a)
            state["x"] += 1

        def fb(_u, p):
            feedback.append(p)

        cl = _RefClient()
        cl.user = None
        cl.set_x_y = type(cl.set_x_y)(set_x_y)
        cl.put_pixel = type(cl.put_pixel)(put_pixel)
        cl.feedback = type(cl.feedback)(fb)
        cl.modulus = modulus
        f = _RefFrame()
        err = ctypes.create_string_buffer(256)
        rc = self.lib.gidref_decode_client(data, ctypes.c_size_t(len(data)), ctypes.byref(cl),
                                           ctypes.byref(f), err, ctypes.c_size_t(256))
        if rc != 0:
            raise OracleError(rc, err.value.decode())
        return self._frame_from_c(f), img, feedback

    def reconstruct(self, fr: Frame, planes: list[np.ndarray] | None = None,
                    stride: int | None = None) -> np.ndarray:
        """Reconstruction only (dequant + IDCT + upsample + colour) from int16 planes."""
        planes = planes if planes is not None else fr.planes
        stride = stride or 3 * fr.width
        f = self._frame_to_c(fr)
        keep = [np.ascontiguousarray(p, dtype=np.int16) for p in planes]
        ptrs = (ctypes.c_void_p * 3)()
        for c in range(fr.ncomp):
            assert keep[c].size == fr.blocks_x[c] * fr.blocks_y[c] * 64
            ptrs[c] = keep[c].ctypes.data
        out = np.zeros((fr.height, stride), dtype=np.uint8)
        rc = self.lib.gidref_reconstruct(ctypes.byref(f), ptrs, out.ctypes.data, stride)
        if rc != 0:
            raise OracleError(rc, "gidref_reconstruct")
        return out[:, :3 * fr.width].reshape(fr.height, fr.width, 3)

    def idct(self, block) -> np.ndarray:
        b = (ctypes.c_int32 * 64)(*[int(v) for v in block])
        self.lib.gidref_idct_8x8(b)
        return np.array(b[:], dtype=np.int32)

    def idct_variant(self, block, no_row=False, no_col=False, shifts=False) -> np.ndarray:
        b = (ctypes.c_int32 * 64)(*[int(v) for v in block])
        self.lib.gidref_idct_8x8_variant(b, int(no_row), int(no_col), int(shifts))
        return np.array(b[:], dtype=np.int32)

    def ycbcr_to_rgb(self, y: int, cb: int, cr: int) -> tuple[int, int, int]:
        out = (ctypes.c_uint8 * 3)()
        self.lib.gidref_ycbcr_to_rgb(y, cb, cr, out)
        return out[0], out[1], out[2]

    def zig_zag(self) -> list[int]:
        p = self.lib.gidref_zig_zag()
        return [p[i] for i in range(64)]

    def undo_zig_zag(self) -> list[int]:
        p = self.lib.gidref_undo_zig_zag()
        return [p[i] for i in range(64)]


# --------------------------------------------------------------------------- encoder


class _Scan(ctypes.Structure):
    _fields_ = [("ncomp", ctypes.c_int32), ("comp", ctypes.c_int32 * 3), ("ss", ctypes.c_int32),
                ("se", ctypes.c_int32), ("ah", ctypes.c_int32), ("al", ctypes.c_int32)]


class _Params(ctypes.Structure):
    _fields_ = [
        ("width", ctypes.c_int32), ("height", ctypes.c_int32), ("ncomp", ctypes.c_int32),
        ("samp_h", ctypes.c_int32 * 3), ("samp_v", ctypes.c_int32 * 3),
        ("quality", ctypes.c_int32), ("progressive", ctypes.c_int32),
        ("restart_interval", ctypes.c_int32), ("optimize_huffman", ctypes.c_int32),
        ("dqt16", ctypes.c_int32), ("nscans", ctypes.c_int32), ("scans", _Scan * 32),
        ("qt_override", (ctypes.c_uint16 * 64) * 2), ("use_qt_override", ctypes.c_int32),
    ]


class _Result(ctypes.Structure):
    _fields_ = [
        ("jpg", ctypes.POINTER(ctypes.c_uint8)), ("jpg_len", ctypes.c_size_t),
        ("planes", ctypes.POINTER(ctypes.c_int16) * 3),
        ("blocks_x", ctypes.c_int32 * 3), ("blocks_y", ctypes.c_int32 * 3),
        ("qt", (ctypes.c_uint16 * 64) * 3),
    ]


SAMPLING = {
    "444": ([1, 1, 1], [1, 1, 1]),
    "422": ([2, 1, 1], [1, 1, 1]),   # h2v1
    "420": ([2, 1, 1], [2, 1, 1]),
    "440": ([1, 1, 1], [2, 1, 1]),   # h1v2
    "411": ([4, 1, 1], [1, 1, 1]),
    "410": ([4, 1, 1], [2, 1, 1]),   # h4v2: no fused kernel, the generic two-pass path
    "grey": ([1], [1]),
}


class Synth:
    """Synthetic image source + JPEG encoder of the test harness."""

    def __init__(self) -> None:
        build_all()
        self.lib = ctypes.CDLL(os.path.join(_HERE, "libsynthjpeg.so"))
        self.lib.synth_image.argtypes = [ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        for fn in ("synth_encode_pixels", "synth_encode_planes", "synth_planes_only"):
            getattr(self.lib, fn).restype = ctypes.c_int

    def image(self, index: int, width: int, height: int) -> np.ndarray:
        rgb = np.empty((height, width, 3), dtype=np.uint8)
        self.lib.synth_image(index, width, height, rgb.ctypes.data)
        return rgb

    @staticmethod
    def _params(width, height, sampling, quality, progressive, restart_interval, optimize,
                dqt16, scans, qt_override) -> _Params:
        sh, sv = SAMPLING[sampling] if isinstance(sampling, str) else sampling
        p = _Params()
        p.width, p.height, p.ncomp = width, height, len(sh)
        for c in range(len(sh)):
            p.samp_h[c], p.samp_v[c] = sh[c], sv[c]
        p.quality, p.progressive = quality, int(progressive)
        p.restart_interval, p.optimize_huffman, p.dqt16 = restart_interval, int(optimize), int(dqt16)
        if scans:
            p.nscans = len(scans)
            for i, (comps, ss, se, ah, al) in enumerate(scans):
                p.scans[i].ncomp = len(comps)
                for j, c in enumerate(comps):
                    p.scans[i].comp[j] = c
                p.scans[i].ss, p.scans[i].se, p.scans[i].ah, p.scans[i].al = ss, se, ah, al
        if qt_override is not None:
            q = np.asarray(qt_override, dtype=np.uint16).reshape(2, 64)
            for t in range(2):
                for i in range(64):
                    p.qt_override[t][i] = int(q[t, i])
            p.use_qt_override = 1
        return p

    def _collect(self, p: _Params, res: _Result, with_jpg: bool):
        n = p.ncomp
        fr = Frame(
            width=p.width, height=p.height, ncomp=n,
            samp_h=[p.samp_h[c] for c in range(n)], samp_v=[p.samp_v[c] for c in range(n)],
            blocks_x=[res.blocks_x[c] for c in range(n)], blocks_y=[res.blocks_y[c] for c in range(n)],
            qt=np.array([[res.qt[c][i] for i in range(64)] for c in range(n)], dtype=np.uint16).reshape(n, 64),
            progressive=bool(p.progressive), restart_interval=p.restart_interval)
        for c in range(n):
            cnt = fr.blocks_x[c] * fr.blocks_y[c] * 64
            arr = np.ctypeslib.as_array(res.planes[c], shape=(cnt,)).copy()
            fr.planes.append(arr.reshape(fr.blocks_y[c], fr.blocks_x[c], 64))
        jpg = bytes(np.ctypeslib.as_array(res.jpg, shape=(res.jpg_len,))) if with_jpg else b""
        self.lib.synth_result_free(ctypes.byref(res))
        return jpg, fr

    def encode(self, pixels: np.ndarray, sampling="420", quality=75, progressive=False,
               restart_interval=0, optimize=False, dqt16=False, scans=None, qt_override=None):
        """pixels: (H, W, 3) RGB or (H, W) grey uint8.  Returns (jpg bytes, Frame with planes)."""
        pixels = np.ascontiguousarray(pixels, dtype=np.uint8)
        h, w = pixels.shape[:2]
        p = self._params(w, h, sampling, quality, progressive, restart_interval, optimize, dqt16,
                         scans, qt_override)
        assert (p.ncomp == 1) == (pixels.ndim == 2)
        res = _Result()
        rc = self.lib.synth_encode_pixels(ctypes.byref(p), pixels.ctypes.data_as(ctypes.c_void_p),
                                          ctypes.byref(res))
        if rc != 0:
            raise RuntimeError("synth_encode_pixels failed")
        return self._collect(p, res, True)

    def encode_planes(self, width, height, planes, sampling="420", progressive=False,
                      restart_interval=0, optimize=True, dqt16=False, scans=None,
                      qt_override=None, quality=75):
        """Entropy-code given quantised planes (list of (by, bx, 64) int16)."""
        p = self._params(width, height, sampling, quality, progressive, restart_interval, optimize,
                         dqt16, scans, qt_override)
        keep = [np.ascontiguousarray(a, dtype=np.int16) for a in planes]
        ptrs = (ctypes.c_void_p * 3)()
        for c in range(p.ncomp):
            ptrs[c] = keep[c].ctypes.data
        res = _Result()
        rc = self.lib.synth_encode_planes(ctypes.byref(p), ptrs, ctypes.byref(res))
        if rc != 0:
            raise RuntimeError("synth_encode_planes failed")
        return self._collect(p, res, True)

    def planes_only(self, pixels: np.ndarray, sampling="420", quality=75) -> Frame:
        pixels = np.ascontiguousarray(pixels, dtype=np.uint8)
        h, w = pixels.shape[:2]
        p = self._params(w, h, sampling, quality, False, 0, False, False, None, None)
        res = _Result()
        rc = self.lib.synth_planes_only(ctypes.byref(p), pixels.ctypes.data_as(ctypes.c_void_p),
                                        ctypes.byref(res))
        if rc != 0:
            raise RuntimeError("synth_planes_only failed")
        return self._collect(p, res, False)[1]

    def grid(self, width: int, height: int, sampling) -> tuple[list[int], list[int]]:
        sh, sv = SAMPLING[sampling] if isinstance(sampling, str) else sampling
        hmax, vmax = max(sh), max(sv)
        mx = -(-width // (8 * hmax))
        my = -(-height // (8 * vmax))
        return [mx * h for h in sh], [my * v for v in sv]


# --------------------------------------------------------------------------- kernel body on the host


class Emu:
    """The kernel's per-thread body compiled for the host (tests only)."""

    def __init__(self) -> None:
        build_all()
        self.lib = ctypes.CDLL(os.path.join(_HERE, "libemurecon.so"))
        self.lib.emu_reconstruct.restype = ctypes.c_int

    def set_turn_tile_rows(self, rows: int) -> None:
        """Force the rows per tile of frames turned by a quarter (32, 16, 8, 4; 0 = make_geometry chooses)."""
        self.lib.emu_set_turn_tile_rows(ctypes.c_int(rows))

    def layout(self, desc) -> int:
        """The kernel family a frame runs (recon_body.cuh, Layout; 6 = the generic two-pass kernels)."""
        return int(self.lib.emu_layout(ctypes.byref(desc)))

    def reconstruct(self, desc, planes, stride: int, height: int) -> np.ndarray:
        """desc: gid_b200.FrameDesc; returns (H, stride) uint8 rows."""
        keep = [np.ascontiguousarray(p, dtype=np.int16) for p in planes]
        ptrs = (ctypes.c_void_p * 3)()
        for c in range(desc.ncomp):
            ptrs[c] = keep[c].ctypes.data
        out = np.zeros((height, stride), dtype=np.uint8)
        err = ctypes.create_string_buffer(512)
        rc = self.lib.emu_reconstruct(ctypes.byref(desc), ptrs, out.ctypes.data_as(ctypes.c_void_p),
                                      err, ctypes.c_size_t(512))
        if rc != 0:
            raise RuntimeError(f"emu_reconstruct {rc}: {err.value.decode()}")
        return out

    def idct(self, block) -> np.ndarray:
        b = (ctypes.c_int32 * 64)(*[int(v) for v in block])
        self.lib.emu_idct_block(b)
        return np.array(b[:], dtype=np.int32)

    def chroma_exhaustive(self) -> int:
        return self.lib.emu_chroma_exhaustive()


def frame_desc(fr: Frame, flags: int = 0, out_stride: int = 0):
    """gid_b200.FrameDesc for a harness Frame."""
    from gid_b200 import FrameDesc, capi
    d = FrameDesc.make(fr.width, fr.height, fr.samp_h, fr.samp_v, fr.qt, flags, out_stride)
    # the producer of the planes states their range (GIDCUDA_FLAG_WIDE_COEF)
    if getattr(fr, "planes", None) is not None and len(fr.planes) and capi.coefficients_are_wide(d, fr.planes):
        d.flags |= capi.FLAG_WIDE_COEF
    return d
