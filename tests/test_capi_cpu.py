"""The C-ABI library without a GPU: it loads, exports every symbol
include/gidcuda.h declares, answers geometry questions, and refuses to compute
without a device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gidcuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gidcuda_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(gidcuda_lib):
    from gid_b200 import capi
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(gidcuda_lib, s), s
    assert sorted(capi.EXPORTS) == syms
    assert gidcuda_lib.gidcuda_abi_version() == 5


def test_geometry_matches_gid(gidcuda_lib, synth):
    """Block grids = what GID's Read_SOS computes, for the five configs + odd sizes."""
    from gid_b200 import FrameDesc
    from tests.harness import SAMPLING
    q = np.ones((3, 64), dtype=np.uint16)
    table = [(512, 512, "420", [(64, 64), (32, 32), (32, 32)], 1536),
             (4096, 4096, "444", [(512, 512)] * 3, 12288),
             (1920, 1080, "420", [(240, 136), (120, 68), (120, 68)], 5760),
             (8192, 8192, "grey", [(1024, 1024)], 24576),
             (2048, 2048, "422", [(256, 256), (128, 256), (128, 256)], 6144),
             (17, 9, "420", [(4, 2), (2, 1), (2, 1)], 64),
             (1, 1, "444", [(1, 1)] * 3, 16)]
    for w, h, samp, grids, stride in table:
        sh, sv = SAMPLING[samp]
        d = FrameDesc.make(w, h, sh, sv, q[:len(sh)])
        for c, (ex, ey) in enumerate(grids):
            bx, by = ctypes.c_int32(), ctypes.c_int32()
            assert gidcuda_lib.gidcuda_plane_geometry(ctypes.byref(d), c, ctypes.byref(bx), ctypes.byref(by)) == 0
            assert (bx.value, by.value) == (ex, ey)
            assert [bx.value, by.value] == [synth.grid(w, h, samp)[0][c], synth.grid(w, h, samp)[1][c]]
        st, nb = ctypes.c_size_t(), ctypes.c_size_t()
        assert gidcuda_lib.gidcuda_output_geometry(ctypes.byref(d), ctypes.byref(st), ctypes.byref(nb)) == 0
        assert st.value == stride and nb.value == stride * h


def test_bad_descriptors_are_rejected(gidcuda_lib):
    from gid_b200 import FrameDesc
    q = np.ones((3, 64), dtype=np.uint16)
    st = ctypes.c_size_t()

    def rc(d):
        return gidcuda_lib.gidcuda_output_geometry(ctypes.byref(d), ctypes.byref(st), None)

    assert rc(FrameDesc.make(0, 10, [1], [1], q[:1])) == -1
    assert rc(FrameDesc.make(70000, 10, [1], [1], q[:1])) == -1
    assert rc(FrameDesc.make(10, 10, [1, 1], [1, 1], q[:2])) == -2          # 2 components
    assert rc(FrameDesc.make(10, 10, [5, 1, 1], [1, 1, 1], q)) == -1        # sampling factor 5
    assert rc(FrameDesc.make(10, 10, [3, 2, 1], [1, 1, 1], q)) == -2        # 3/2 not integer
    assert rc(FrameDesc.make(10, 10, [2], [2], q[:1])) == -2                # grey with 2x2
    assert rc(FrameDesc.make(10, 10, [1], [1], q[:1], out_stride=29)) == -1 # stride < 3*W
    assert b"stride" in gidcuda_lib.gidcuda_last_error(None)


def test_no_cpu_fallback(gidcuda_lib):
    """On a machine without a CUDA device every compute entry point fails loudly."""
    if gidcuda_lib.gidcuda_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from gid_b200 import GidCuda, GidCudaError, capi
    with pytest.raises(GidCudaError) as e:
        GidCuda(0)
    assert e.value.code == -3 and "no CPU fallback" in e.value.msg
    with pytest.raises(GidCudaError):
        capi.Batch([0])
    with pytest.raises(GidCudaError):
        capi.pinned_empty(16)


def test_product_does_not_touch_the_oracle():
    """Nothing under gid_b200/ or include/ may reference oracle/ or the test harness."""
    bad = []
    for base in ("gid_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".so", ".o", ".pyc")):
                    continue
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"gidref|oracle/|tests[./]harness|libemurecon", text):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_records_from_plane_scan_order():
    """records_from_plane: blocks in interleaved-scan order, one record per non-zero coefficient."""
    import numpy as np
    from gid_b200 import capi
    rng = np.random.default_rng(4)
    by, bx, sh, sv = 4, 6, 2, 2
    p = np.zeros((by, bx, 64), dtype=np.int16)
    m = rng.random(p.shape) < 0.1
    p[m] = rng.integers(-300, 300, size=int(m.sum()))
    first, rec = capi.records_from_plane(p, sh, sv)
    assert first[0] == 0 and first[-1] == rec.size == np.count_nonzero(p)
    back = np.zeros_like(p)
    seq = 0
    for my in range(by // sv):
        for mx in range(bx // sh):
            for sy in range(sv):
                for sx in range(sh):
                    for w in rec[first[seq]:first[seq + 1]]:
                        back[my * sv + sy, mx * sh + sx, int(w) & 63] = np.int16(np.uint16(int(w) >> 16))
                    seq += 1
    assert np.array_equal(back, p)


def test_output_geometry_of_turned_frames(gidcuda_lib):
    """GIDCUDA_FLAG_ROTATE_90 / _270 swap the sides of the pixel buffer; _180 keeps them."""
    from gid_b200 import FrameDesc, capi
    q = np.ones((3, 64), dtype=np.uint16)
    for flag, (ow, oh) in ((0, (131, 67)), (capi.FLAG_ROTATE_90, (67, 131)), (capi.FLAG_ROTATE_180, (131, 67)),
                           (capi.FLAG_ROTATE_270, (67, 131))):
        for fmt, bpp, align in ((capi.OUT_RGB8 | capi.FLAG_PACKED_ROWS, 3, 1), (capi.OUT_BGR8_BOTTOM_UP, 3, 4), (capi.OUT_RGBA8, 4, 16)):
            d = FrameDesc.make(131, 67, [2, 1, 1], [2, 1, 1], q, flags=fmt | flag)
            st, nb = ctypes.c_size_t(), ctypes.c_size_t()
            assert gidcuda_lib.gidcuda_output_geometry(ctypes.byref(d), ctypes.byref(st), ctypes.byref(nb)) == 0
            want = -(-(bpp * ow) // align) * align
            assert (st.value, nb.value) == (want, want * oh), (flag, fmt)


def test_job_entry_points_reject_null_jobs(gidcuda_lib):
    """Without a device there are no jobs: every job entry point must answer a null handle with an
    error code, never crash (the Ada binding turns the code into an exception)."""
    L = gidcuda_lib
    null = ctypes.c_void_p()
    assert L.gidcuda_job_submit(null) == -1
    assert L.gidcuda_job_wait(null, None, None) == -1
    assert L.gidcuda_job_hint_rows(null, 0, 4) == -1
    assert L.gidcuda_job_records_commit(null, 0, 0) == -1
    assert L.gidcuda_job_records_commit_runs(null, 0, 1, None, None) == -1
    assert L.gidcuda_job_output(null, None, 0) == -1
    assert L.gidcuda_job_scan(null, None, None, 0, None, None) == -1
    assert L.gidcuda_host_is_pinned(None) == 0
    assert L.gidcuda_counter(b"chain_repairs") == 0 and L.gidcuda_counter(b"no such counter") == -1
    assert L.gidcuda_set_option(b"chain_repair", 1) == 0 and L.gidcuda_set_option(b"no such option", 1) == -1


def test_ada_binding_constants_match_the_header():
    """gid_b200/ada/gid-cuda.ads (the binding a GID maintainer would add) names the same flag and
    format values as include/gidcuda.h."""
    import re
    from gid_b200 import capi
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ads = open(os.path.join(root, "gid_b200", "ada", "gid-cuda.ads")).read()
    hdr = open(os.path.join(root, "include", "gidcuda.h")).read()
    ada = {m.group(1).upper(): int(m.group(2).replace("16#", "").replace("#", ""), 16 if "16#" in m.group(2) else 10)
           for m in re.finditer(r"^\s*((?:Out|Flag)_\w+)\s*:\s*constant\s*:=\s*([0-9#A-Fa-f]+);", ads, re.M)}
    c = {m.group(1): int(m.group(2).rstrip("u"), 0)
         for m in re.finditer(r"#define GIDCUDA_((?:OUT|FLAG)_\w+) (0x[0-9A-Fa-f]+u?|[0-9]+u?)\b", hdr)}
    assert len(ada) >= 8
    for name, value in ada.items():
        assert c[name] == value, name
    assert capi.FLAG_ROTATE_270 == c["FLAG_ROTATE_270"] and capi.OUT_RGBA8 == c["OUT_RGBA8"]
