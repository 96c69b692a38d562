"""ctypes binding of include/gidcuda.h.

The library must exist (``python -m gid_b200.build``); importing this module
never falls back to a CPU implementation -- if the CUDA extension is missing,
``load_library`` raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

GIDCUDA_OK = 0
ERR_NAMES = {
    -1: "BAD_ARGUMENT", -2: "UNSUPPORTED", -3: "NO_DEVICE", -4: "CUDA", -5: "OUT_OF_MEMORY",
    -6: "STATE",
    -7: "DATA",
}
OUT_RGB8 = 0
OUT_BGR8_BOTTOM_UP = 1
OUT_RGBA8 = 2
FLAG_PACKED_ROWS = 0x100
FLAG_ROTATE_90, FLAG_ROTATE_180, FLAG_ROTATE_270 = 0x200, 0x400, 0x600  # GID.Display_Orientation applied to the pixels
FLAG_WIDE_COEF = 0x800  # some coefficient exceeds gidcuda_fast_coef_limit: the exact kernel instances run

# every symbol include/gidcuda.h declares
EXPORTS = [
    "gidcuda_abi_version", "gidcuda_device_count", "gidcuda_last_error", "gidcuda_set_option", "gidcuda_counter", "gidcuda_create",
    "gidcuda_destroy", "gidcuda_plane_geometry", "gidcuda_output_geometry", "gidcuda_job_begin",
    "gidcuda_job_plane", "gidcuda_job_hint_rows", "gidcuda_fast_coef_limit", "gidcuda_job_wide_coefficients", "gidcuda_job_records", "gidcuda_job_records_commit", "gidcuda_job_records_commit_runs",
    "gidcuda_job_pscan_reserve", "gidcuda_job_pscan", "gidcuda_job_pscan_masks", "gidcuda_job_pscan_refine",
    "gidcuda_job_pscan_refine_commit", "gidcuda_job_output", "gidcuda_job_scan", "gidcuda_host_is_pinned", "gidcuda_job_submit", "gidcuda_job_wait", "gidcuda_job_release",
    "gidcuda_plan_create", "gidcuda_plan_launch", "gidcuda_plan_launch_count", "gidcuda_plan_uses_tma",
    "gidcuda_plan_destroy", "gidcuda_synchronize", "gidcuda_batch_create", "gidcuda_batch_run", "gidcuda_batch_run_records",
    "gidcuda_batch_destroy", "gidcuda_host_alloc", "gidcuda_host_free",
]


class FrameDesc(ctypes.Structure):
    """struct gidcuda_frame_desc"""
    _fields_ = [
        ("width", ctypes.c_int32), ("height", ctypes.c_int32), ("ncomp", ctypes.c_int32),
        ("samp_h", ctypes.c_int32 * 3), ("samp_v", ctypes.c_int32 * 3),
        ("qt", (ctypes.c_uint16 * 64) * 3), ("flags", ctypes.c_uint32),
        ("out_stride", ctypes.c_int32),
    ]

    @classmethod
    def make(cls, width: int, height: int, samp_h: Sequence[int], samp_v: Sequence[int],
             qt: np.ndarray, flags: int = 0, out_stride: int = 0) -> "FrameDesc":
        d = cls()
        d.width, d.height, d.ncomp = int(width), int(height), len(samp_h)
        q = np.asarray(qt, dtype=np.uint16).reshape(len(samp_h), 64)
        for c in range(len(samp_h)):
            d.samp_h[c], d.samp_v[c] = int(samp_h[c]), int(samp_v[c])
            ctypes.memmove(d.qt[c], q[c].ctypes.data, 128)
        d.flags, d.out_stride = flags, out_stride
        return d


class HuffmanTable(ctypes.Structure):
    """struct gidcuda_huffman_table"""
    _fields_ = [("counts", ctypes.c_uint8 * 16), ("symbols", ctypes.c_uint8 * 256)]


class ScanDesc(ctypes.Structure):
    """struct gidcuda_scan_desc"""
    _fields_ = [("restart_interval", ctypes.c_int32), ("nintervals", ctypes.c_int32),
                ("dc", HuffmanTable * 3), ("ac", HuffmanTable * 3)]


class BatchStats(ctypes.Structure):
    _fields_ = [
        ("seconds", ctypes.c_double), ("megapixels", ctypes.c_double),
        ("h2d_bytes", ctypes.c_int64), ("d2h_bytes", ctypes.c_int64),
        ("kernel_launches", ctypes.c_int32), ("images_failed", ctypes.c_int32),
        ("images_per_device", ctypes.c_int32 * 16),
    ]


class GidCudaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"gidcuda {ERR_NAMES.get(code, code)}: {msg}")
        self.code = code
        self.msg = msg


def library_path() -> str:
    """In-tree library; GIDCUDA_LIB selects another in-tree build (kernel tuning experiments)."""
    return os.environ.get("GIDCUDA_LIB") or os.path.join(_HERE, "lib", "libgidcuda.so")


_lib = None


def load_library() -> ctypes.CDLL:
    """Load libgidcuda.so.  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build the CUDA extension with `python -m gid_b200.build` "
            "(there is no CPU fallback for the reconstruction path)")
    L = ctypes.CDLL(path)
    vp, i32p = ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32)
    L.gidcuda_abi_version.restype = ctypes.c_int
    L.gidcuda_counter.argtypes = [ctypes.c_char_p]
    L.gidcuda_counter.restype = ctypes.c_long
    L.gidcuda_device_count.restype = ctypes.c_int
    L.gidcuda_last_error.restype = ctypes.c_char_p
    L.gidcuda_last_error.argtypes = [vp]
    L.gidcuda_create.argtypes = [ctypes.c_int, ctypes.POINTER(vp)]
    L.gidcuda_set_option.argtypes = [ctypes.c_char_p, ctypes.c_int]
    L.gidcuda_destroy.argtypes = [vp]
    L.gidcuda_plane_geometry.argtypes = [ctypes.POINTER(FrameDesc), ctypes.c_int, i32p, i32p]
    L.gidcuda_output_geometry.argtypes = [ctypes.POINTER(FrameDesc), ctypes.POINTER(ctypes.c_size_t),
                                          ctypes.POINTER(ctypes.c_size_t)]
    L.gidcuda_job_begin.argtypes = [vp, ctypes.POINTER(FrameDesc), ctypes.POINTER(vp)]
    L.gidcuda_job_plane.restype = vp
    L.gidcuda_job_plane.argtypes = [vp, ctypes.c_int, i32p, i32p]
    L.gidcuda_job_hint_rows.argtypes = [vp, ctypes.c_int, ctypes.c_int]
    L.gidcuda_fast_coef_limit.argtypes = [ctypes.POINTER(FrameDesc), ctypes.c_int]
    L.gidcuda_fast_coef_limit.restype = ctypes.c_int32
    L.gidcuda_job_wide_coefficients.argtypes = [vp]
    L.gidcuda_job_records.argtypes = [vp, ctypes.c_int, ctypes.POINTER(vp), ctypes.POINTER(vp),
                                      ctypes.POINTER(ctypes.c_size_t), i32p]
    L.gidcuda_job_records_commit.argtypes = [vp, ctypes.c_int, ctypes.c_size_t]
    L.gidcuda_job_records_commit_runs.argtypes = [vp, ctypes.c_int, ctypes.c_int32, ctypes.POINTER(ctypes.c_size_t),
                                                  ctypes.POINTER(ctypes.c_size_t)]
    L.gidcuda_job_output.argtypes = [vp, vp, ctypes.c_size_t]
    L.gidcuda_job_scan.argtypes = [vp, ctypes.POINTER(ScanDesc), vp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint32),
                                   ctypes.POINTER(ctypes.c_uint32)]
    L.gidcuda_host_is_pinned.argtypes = [vp]
    L.gidcuda_job_submit.argtypes = [vp]
    L.gidcuda_job_wait.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_size_t)]
    L.gidcuda_job_release.argtypes = [vp]
    L.gidcuda_plan_create.argtypes = [vp, ctypes.c_int32, ctypes.POINTER(FrameDesc),
                                      ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp)]
    L.gidcuda_plan_launch.argtypes = [vp, vp]
    L.gidcuda_plan_launch_count.argtypes = [vp]
    L.gidcuda_plan_uses_tma.argtypes = [vp]
    L.gidcuda_plan_destroy.argtypes = [vp]
    L.gidcuda_synchronize.argtypes = [vp]
    L.gidcuda_batch_create.argtypes = [i32p, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(vp)]
    L.gidcuda_batch_run.argtypes = [vp, ctypes.c_int32, ctypes.POINTER(FrameDesc), ctypes.POINTER(vp),
                                    ctypes.POINTER(vp), i32p, ctypes.POINTER(BatchStats)]
    L.gidcuda_batch_run_records.argtypes = [vp, ctypes.c_int32, ctypes.POINTER(FrameDesc), ctypes.POINTER(vp),
                                            ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(vp),
                                            i32p, ctypes.POINTER(BatchStats)]
    L.gidcuda_batch_destroy.argtypes = [vp]
    L.gidcuda_host_alloc.argtypes = [ctypes.c_size_t, ctypes.POINTER(vp)]
    L.gidcuda_host_free.argtypes = [vp]
    for name in EXPORTS:
        if name not in ("gidcuda_last_error", "gidcuda_job_plane"):
            getattr(L, name).restype = ctypes.c_int
    _lib = L
    return L


def set_option(name: str, value: int) -> None:
    L = load_library()
    rc = L.gidcuda_set_option(name.encode(), value)
    if rc != GIDCUDA_OK:
        raise GidCudaError(rc, (L.gidcuda_last_error(None) or b"").decode())


def coefficients_are_wide(desc: "FrameDesc", planes: Sequence[np.ndarray]) -> bool:
    """What an entropy decoder tracks while it stores coefficients: does any magnitude of component c
    exceed gidcuda_fast_coef_limit(desc, c)?  (include/gidcuda.h, GIDCUDA_FLAG_WIDE_COEF)"""
    L = load_library()
    for c in range(desc.ncomp):
        p = np.asarray(planes[c])
        if p.size and int(np.abs(p.astype(np.int32)).max()) > L.gidcuda_fast_coef_limit(ctypes.byref(desc), c):
            return True
    return False


def rows_used(plane: np.ndarray) -> int:
    """1 + the last block row (natural order) holding a non-zero coefficient anywhere in the plane."""
    nz = np.flatnonzero(np.asarray(plane).reshape(-1, 8, 8).any(axis=(0, 2)))
    return int(nz[-1]) + 1 if nz.size else 1


def records_from_plane(plane: np.ndarray, samp_h: int, samp_v: int) -> tuple[np.ndarray, np.ndarray]:
    """Dense plane (by, bx, 64) int16 -> (first, records) of the sparse input form
    (include/gidcuda.h): blocks in the order an interleaved scan visits them (MCUs in
    raster order, inside an MCU the component's blocks row by row); one uint32 record per
    non-zero coefficient = (uint16)value << 16 | natural index; first[seq] = index of block
    seq's first record, first[nblocks] = number of records.  What an entropy decoder
    emits directly; here derived from planes for tests and the bench."""
    p = np.asarray(plane, dtype=np.int16)
    by, bx = p.shape[0], p.shape[1]
    seq = (p.reshape(by // samp_v, samp_v, bx // samp_h, samp_h, 64)
            .transpose(0, 2, 1, 3, 4).reshape(-1, 64))
    mask = seq != 0
    first = np.zeros(seq.shape[0] + 1, dtype=np.uint32)
    np.cumsum(mask.sum(axis=1), out=first[1:])
    blk, k = np.nonzero(mask)
    records = (seq[blk, k].astype(np.uint16).astype(np.uint32) << 16) | k.astype(np.uint32)
    return first, records


class GidCuda:
    """One context (= one host thread's view of one GPU)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self.ctx = ctypes.c_void_p()
        self._check(self.lib.gidcuda_create(device, ctypes.byref(self.ctx)), None)
        self.device = device

    def _check(self, rc: int, ctx) -> None:
        if rc != GIDCUDA_OK:
            msg = self.lib.gidcuda_last_error(ctx)
            raise GidCudaError(rc, msg.decode() if msg else "")

    def close(self) -> None:
        if self.ctx:
            self._check(self.lib.gidcuda_destroy(self.ctx), self.ctx)
            self.ctx = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- geometry ---------------------------------------------------------------
    def plane_geometry(self, desc: FrameDesc, comp: int) -> tuple[int, int]:
        bx, by = ctypes.c_int32(), ctypes.c_int32()
        self._check(self.lib.gidcuda_plane_geometry(ctypes.byref(desc), comp, ctypes.byref(bx),
                                                    ctypes.byref(by)), None)
        return bx.value, by.value

    def output_geometry(self, desc: FrameDesc) -> tuple[int, int]:
        st, nb = ctypes.c_size_t(), ctypes.c_size_t()
        self._check(self.lib.gidcuda_output_geometry(ctypes.byref(desc), ctypes.byref(st),
                                                     ctypes.byref(nb)), None)
        return st.value, nb.value

    # -- one frame through the job API (host buffers, H2D + kernel + D2H) -----------
    def job_fill_records(self, job, desc: FrameDesc, comp: int, plane: np.ndarray) -> int:
        """Supply component `comp` of a begun job in the sparse form; returns the bytes that will
        cross the host link for it.  Raises OverflowError when the plane is too dense for the
        record buffer (the caller then supplies the dense plane)."""
        first, records = records_from_plane(plane, desc.samp_h[comp], desc.samp_v[comp])
        pf, pr, cap, nb = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_size_t(), ctypes.c_int32()
        self._check(self.lib.gidcuda_job_records(job, comp, ctypes.byref(pf), ctypes.byref(pr),
                                                 ctypes.byref(cap), ctypes.byref(nb)), self.ctx)
        if nb.value + 1 != first.size:
            raise ValueError(f"plane {comp}: {first.size - 1} blocks, expected {nb.value}")
        if records.size > cap.value:
            raise OverflowError(f"plane {comp}: {records.size} records exceed the capacity {cap.value}")
        ctypes.memmove(pf, first.ctypes.data, first.nbytes)
        if records.size:
            ctypes.memmove(pr, records.ctypes.data, records.nbytes)
        self._check(self.lib.gidcuda_job_records_commit(job, comp, records.size), self.ctx)
        return first.nbytes + records.nbytes

    def reconstruct(self, desc: FrameDesc, planes: Sequence[np.ndarray], hint_rows: bool = False,
                    sparse: Sequence[bool] | bool | str = False) -> np.ndarray:
        """planes[c]: (by, bx, 64) int16.  Returns the pixel rows as (H, stride) uint8.
        hint_rows: tell the library which block rows are empty (gidcuda_job_hint_rows).
        sparse: supply the components (all, or per component) as coefficient records; "auto" = records
        unless a component is too dense for the record buffer (what a decoder does on overflow)."""
        job = ctypes.c_void_p()
        self._check(self.lib.gidcuda_job_begin(self.ctx, ctypes.byref(desc), ctypes.byref(job)), self.ctx)
        auto = isinstance(sparse, str) and sparse == "auto"
        sp = [bool(sparse)] * desc.ncomp if isinstance(sparse, (bool, int, str)) else list(sparse)
        try:
            for c in range(desc.ncomp):
                if sp[c]:
                    try:
                        self.job_fill_records(job, desc, c, np.asarray(planes[c]))
                        continue
                    except OverflowError:
                        if not auto:
                            raise
                bx, by = ctypes.c_int32(), ctypes.c_int32()
                p = self.lib.gidcuda_job_plane(job, c, ctypes.byref(bx), ctypes.byref(by))
                if not p:
                    self._check(-1, self.ctx)
                src = np.ascontiguousarray(planes[c], dtype=np.int16)
                if src.size != bx.value * by.value * 64:
                    raise ValueError(f"plane {c}: {src.size} coefficients, expected "
                                     f"{bx.value * by.value * 64}")
                ctypes.memmove(p, src.ctypes.data, src.nbytes)
                if hint_rows:
                    self._check(self.lib.gidcuda_job_hint_rows(job, c, rows_used(src)), self.ctx)
            if not desc.flags & FLAG_WIDE_COEF and coefficients_are_wide(desc, planes):
                self._check(self.lib.gidcuda_job_wide_coefficients(job), self.ctx)  # as a decoder would, after decoding
            self._check(self.lib.gidcuda_job_submit(job), self.ctx)
            pix, stride = ctypes.c_void_p(), ctypes.c_size_t()
            self._check(self.lib.gidcuda_job_wait(job, ctypes.byref(pix), ctypes.byref(stride)), self.ctx)
            rows = desc.width if desc.flags & 0x200 else desc.height  # FLAG_ROTATE_90 / _270 swap the sides
            buf = (ctypes.c_uint8 * (stride.value * rows)).from_address(pix.value)
            out = np.frombuffer(buf, dtype=np.uint8).reshape(rows, stride.value).copy()
        finally:
            self._check(self.lib.gidcuda_job_release(job), self.ctx)
        return out

    def reconstruct_rgb(self, desc: FrameDesc, planes: Sequence[np.ndarray], hint_rows: bool = False,
                        sparse: Sequence[bool] | bool | str = False) -> np.ndarray:
        """Same, cropped to (H, W, 3) -- (H, W, 4) for the RGBA format."""
        rows = self.reconstruct(desc, planes, hint_rows, sparse)
        bpp = 4 if (desc.flags & 0xFF) == OUT_RGBA8 else 3
        w, h = (desc.height, desc.width) if desc.flags & 0x200 else (desc.width, desc.height)
        return rows[:, :bpp * w].reshape(h, w, bpp)

    # -- device-resident plans ----------------------------------------------------
    def plan_create(self, descs: Sequence[FrameDesc], d_planes: Sequence[int],
                    d_pixels: Sequence[int]) -> ctypes.c_void_p:
        n = len(descs)
        arr = (FrameDesc * n)(*descs)
        pl = (ctypes.c_void_p * (3 * n))(*[ctypes.c_void_p(int(p) if p else None) for p in d_planes])
        px = (ctypes.c_void_p * n)(*[ctypes.c_void_p(int(p)) for p in d_pixels])
        plan = ctypes.c_void_p()
        self._check(self.lib.gidcuda_plan_create(self.ctx, n, arr, pl, px, ctypes.byref(plan)), self.ctx)
        return plan

    def plan_launch(self, plan, stream: int | None = None) -> None:
        self._check(self.lib.gidcuda_plan_launch(plan, ctypes.c_void_p(stream) if stream else None), self.ctx)

    def plan_launch_count(self, plan) -> int:
        return self.lib.gidcuda_plan_launch_count(plan)

    def plan_uses_tma(self, plan) -> int:
        return self.lib.gidcuda_plan_uses_tma(plan)

    def plan_destroy(self, plan) -> None:
        self._check(self.lib.gidcuda_plan_destroy(plan), self.ctx)

    def synchronize(self) -> None:
        self._check(self.lib.gidcuda_synchronize(self.ctx), self.ctx)


class Batch:
    """Batch driver over one or several GPUs of this box (no NCCL)."""

    def __init__(self, devices: Sequence[int], chunk_images: int = 8):
        self.lib = load_library()
        self.h = ctypes.c_void_p()
        dev = (ctypes.c_int32 * len(devices))(*devices)
        rc = self.lib.gidcuda_batch_create(dev, len(devices), chunk_images, ctypes.byref(self.h))
        if rc != GIDCUDA_OK:
            raise GidCudaError(rc, (self.lib.gidcuda_last_error(None) or b"").decode())

    def run(self, descs: Sequence[FrameDesc], planes: Sequence[int], pixels: Sequence[int]):
        """planes: 3*n host addresses (0 for absent), pixels: n host addresses."""
        n = len(descs)
        arr = (FrameDesc * n)(*descs)
        pl = (ctypes.c_void_p * (3 * n))(*[ctypes.c_void_p(int(p) if p else None) for p in planes])
        px = (ctypes.c_void_p * n)(*[ctypes.c_void_p(int(p) if p else None) for p in pixels])
        status = (ctypes.c_int32 * n)()
        stats = BatchStats()
        rc = self.lib.gidcuda_batch_run(self.h, n, arr, pl, px, status, ctypes.byref(stats))
        if rc != GIDCUDA_OK:
            raise GidCudaError(rc, (self.lib.gidcuda_last_error(None) or b"").decode())
        return list(status), stats

    def run_records(self, descs: Sequence[FrameDesc], first: Sequence[int], records: Sequence[int],
                    counts: Sequence[int], pixels: Sequence[int]):
        """first / records: 3*n host addresses (0 for absent), counts: 3*n record counts, pixels: n host addresses."""
        n = len(descs)
        arr = (FrameDesc * n)(*descs)
        fa = (ctypes.c_void_p * (3 * n))(*[ctypes.c_void_p(int(p) if p else None) for p in first])
        ra = (ctypes.c_void_p * (3 * n))(*[ctypes.c_void_p(int(p) if p else None) for p in records])
        ca = (ctypes.c_size_t * (3 * n))(*[int(c) for c in counts])
        px = (ctypes.c_void_p * n)(*[ctypes.c_void_p(int(p) if p else None) for p in pixels])
        status = (ctypes.c_int32 * n)()
        stats = BatchStats()
        rc = self.lib.gidcuda_batch_run_records(self.h, n, arr, fa, ra, ca, px, status, ctypes.byref(stats))
        if rc != GIDCUDA_OK:
            raise GidCudaError(rc, (self.lib.gidcuda_last_error(None) or b"").decode())
        return list(status), stats

    def close(self) -> None:
        if self.h:
            self.lib.gidcuda_batch_destroy(self.h)
            self.h = ctypes.c_void_p()


def counter(name: str) -> int:
    """"chain_repairs" / "chain_repair_subsequences" (gidcuda_counter)."""
    return int(load_library().gidcuda_counter(name.encode()))


def pinned_empty(nbytes: int) -> tuple[np.ndarray, int]:
    """A pinned host byte buffer (numpy view, address).  Free with pinned_free."""
    L = load_library()
    p = ctypes.c_void_p()
    rc = L.gidcuda_host_alloc(nbytes, ctypes.byref(p))
    if rc != GIDCUDA_OK:
        raise GidCudaError(rc, (L.gidcuda_last_error(None) or b"").decode())
    buf = (ctypes.c_uint8 * nbytes).from_address(p.value)
    return np.frombuffer(buf, dtype=np.uint8), p.value


def pinned_free(addr: int) -> None:
    load_library().gidcuda_host_free(ctypes.c_void_p(addr))
