"""The kernel's per-thread body, compiled for the host (tests/harness/emu_recon.cpp),
against the oracle.  Same source as the device code (gid_b200/csrc/recon_*.cuh),
same frame descriptors as the shim builds; bit-exact bar.  No GPU needed.
"""
import ctypes
import zlib

import numpy as np
import pytest

from tests.cases import GEOMETRIES, edge_frames, pixels_for, random_planes
from tests.harness import Frame, SAMPLING, frame_desc


def _stride(lib, d):
    st, nb = ctypes.c_size_t(), ctypes.c_size_t()
    assert lib.gidcuda_output_geometry(ctypes.byref(d), ctypes.byref(st), ctypes.byref(nb)) == 0
    return st.value


def _check(emu, oracle, lib, fr, flags=0, out_stride=0):
    from gid_b200 import capi
    ref = oracle.reconstruct(fr)
    d = frame_desc(fr, flags, out_stride)
    rows = emu.reconstruct(d, fr.planes, _stride(lib, d), fr.height)
    if (flags & 0xFF) == capi.OUT_RGBA8:
        got = rows[:, :4 * fr.width].reshape(fr.height, fr.width, 4)
        assert (got[:, :, 3] == 255).all()
        got = got[:, :, :3]
    else:
        got = rows[:, :3 * fr.width].reshape(fr.height, fr.width, 3)
    if (flags & 0xFF) == capi.OUT_BGR8_BOTTOM_UP:
        got = got[::-1, :, ::-1]
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("w,h,samp", GEOMETRIES)
def test_geometries(emu, oracle, synth, gidcuda_lib, w, h, samp):
    fr = synth.planes_only(pixels_for(synth, 11, w, h, samp), samp)
    _check(emu, oracle, gidcuda_lib, fr)


@pytest.mark.parametrize("kind", ["dense", "sparse", "dc_only", "rows", "extreme"])
@pytest.mark.parametrize("samp", ["grey", "444", "422", "420", "440"])
def test_random_coefficients(emu, oracle, synth, gidcuda_lib, kind, samp):
    rng = np.random.default_rng(zlib.crc32(f"{kind}-{samp}".encode()))
    w, h = 200, 136
    bx, by = synth.grid(w, h, samp)
    sh, sv = SAMPLING[samp]
    qt = rng.integers(1, 256, size=(len(sh), 64)).astype(np.uint16)
    fr = Frame(w, h, len(sh), list(sh), list(sv), bx, by, qt)
    fr.planes = random_planes(rng, bx, by, kind)
    _check(emu, oracle, gidcuda_lib, fr)


def test_edge_of_the_domain(emu, oracle, synth, gidcuda_lib):
    """Tables with entries of 0, products that are multiples of 2**21, DC terms that wrap, full-range garbage,
    the rim of the fast domain: identical bytes everywhere (the fast instances inside their domain, the
    exact ones outside it)."""
    from gid_b200 import capi
    seen_fast = seen_exact = 0
    for name, fr in edge_frames(synth):
        d = frame_desc(fr)
        fast = not (d.flags & capi.FLAG_WIDE_COEF) and int(fr.qt.min()) >= 1 and int(fr.qt.max()) <= 255
        seen_fast += fast
        seen_exact += not fast
        if name.startswith("rim_of_the_fast_domain"):
            assert fast, name
        if name.startswith(("q8_wide", "dc_terms", "full_range", "all_ac", "scattered", "q16")):
            assert not fast, name
        try:
            _check(emu, oracle, gidcuda_lib, fr)
        except AssertionError as e:
            raise AssertionError(f"{name}: kernel body != oracle") from e
    assert seen_fast >= 3 and seen_exact >= 10


def test_16bit_quant_tables(emu, oracle, synth, gidcuda_lib):
    rng = np.random.default_rng(7)
    for samp in ("grey", "420", "444"):
        w, h = 120, 72
        bx, by = synth.grid(w, h, samp)
        sh, sv = SAMPLING[samp]
        qt = rng.integers(1, 2000, size=(len(sh), 64)).astype(np.uint16)
        fr = Frame(w, h, len(sh), list(sh), list(sv), bx, by, qt)
        fr.planes = [np.clip(p, -400, 400) for p in random_planes(rng, bx, by, "sparse")]
        _check(emu, oracle, gidcuda_lib, fr)


@pytest.mark.parametrize("samp", ["grey", "444", "422", "420", "411"])
def test_output_formats_and_strides(emu, oracle, synth, gidcuda_lib, samp):
    from gid_b200 import capi
    fr = synth.planes_only(pixels_for(synth, 3, 211, 77, samp), samp)
    _check(emu, oracle, gidcuda_lib, fr, flags=capi.FLAG_PACKED_ROWS)
    _check(emu, oracle, gidcuda_lib, fr, flags=capi.OUT_BGR8_BOTTOM_UP)
    _check(emu, oracle, gidcuda_lib, fr, out_stride=3 * 211 + 29)
    _check(emu, oracle, gidcuda_lib, fr, flags=capi.OUT_RGBA8)
    _check(emu, oracle, gidcuda_lib, fr, flags=capi.OUT_RGBA8, out_stride=4 * 211 + 4)


@pytest.mark.parametrize("samp", ["grey", "444", "422", "420", "440", "411"])
@pytest.mark.parametrize("w,h", [(64, 48), (211, 77), (40, 9), (1056, 16)])
def test_half_turn_in_the_fused_kernels(emu, oracle, synth, gidcuda_lib, samp, w, h):
    """GIDCUDA_FLAG_ROTATE_180 (Display_Orientation = Rotation_180 applied as test/to_bmp.adb:104-122 does in its
    Set_X_Y): pixel (x, r) lands at column W - 1 - x, row H - 1 - r -- done by the fused kernels themselves
    (mirrored runs, rows bottom to top), for whole and cut blocks, aligned and unaligned widths, every format."""
    from gid_b200 import capi
    fr = synth.planes_only(pixels_for(synth, 17, w, h, samp), samp)
    _check_turned(emu, gidcuda_lib, fr, capi.FLAG_ROTATE_180, oracle.reconstruct(fr)[::-1, ::-1], samp)


def _check_turned(emu, gidcuda_lib, fr, flag, ref, samp):
    from gid_b200 import capi
    oh, ow = ref.shape[:2]
    for flags, stride in ((0, 0), (capi.FLAG_PACKED_ROWS, 0), (capi.OUT_BGR8_BOTTOM_UP, 0), (capi.OUT_RGBA8, 0),
                          (capi.OUT_RGBA8, 4 * ow + 4), (0, 3 * ow + 5)):
        d = frame_desc(fr, flags | flag, stride)
        assert emu.layout(d) != 6, "the frame must run a fused kernel body"
        rows = emu.reconstruct(d, fr.planes, _stride(gidcuda_lib, d), oh)
        bpp = 4 if (flags & 0xFF) == capi.OUT_RGBA8 else 3
        got = rows[:, :bpp * ow].reshape(oh, ow, bpp)[:, :, :3]
        if (flags & 0xFF) == capi.OUT_BGR8_BOTTOM_UP:
            got = got[::-1, :, ::-1]
        assert np.array_equal(got, ref), (samp, fr.width, fr.height, flags, stride)
        if bpp == 4:
            assert (rows[:, :4 * ow].reshape(oh, ow, 4)[:, :, 3] == 255).all()
        assert not rows[:, bpp * ow:].any(), "bytes beyond a row's pixels were written"


@pytest.mark.parametrize("samp", ["grey", "444", "422", "420", "440", "411"])
@pytest.mark.parametrize("w,h", [(64, 48), (211, 77), (40, 9), (9, 40), (1056, 16), (24, 1064)])
def test_quarter_turns_in_the_fused_kernels(emu, oracle, synth, gidcuda_lib, samp, w, h):
    """GIDCUDA_FLAG_ROTATE_90 / _270 (test/to_bmp.adb:104-122, Set_X_Y): pixel (x, r) lands at column r, row
    W - 1 - x / at column H - 1 - r, row x of an image H wide and W high.  The fused kernels store the samples of
    a block transposed and emit its columns as the rows of the turned image: whole and cut blocks, heights that
    are and are not multiples of 8 (the mirrored runs of _270), every format."""
    from gid_b200 import capi
    fr = synth.planes_only(pixels_for(synth, 23, w, h, samp), samp)
    ref = oracle.reconstruct(fr)
    _check_turned(emu, gidcuda_lib, fr, capi.FLAG_ROTATE_90, np.rot90(ref, 1), samp)
    _check_turned(emu, gidcuda_lib, fr, capi.FLAG_ROTATE_270, np.rot90(ref, -1), samp)


@pytest.mark.parametrize("rows", [32, 16, 8, 4])
@pytest.mark.parametrize("samp", ["grey", "444", "422", "420", "440", "411"])
def test_quarter_turn_tiles_of_every_shape(emu, oracle, synth, gidcuda_lib, samp, rows):
    """Frames turned by a quarter are worked through in tiles of `rows` MCUs down x 32 / rows across (make_geometry
    picks the shape that fills best; forced here): every shape, frames with cut tiles in both directions."""
    from gid_b200 import capi
    emu.set_turn_tile_rows(rows)
    try:
        for w, h in ((211, 77), (72, 1088), (600, 40)):
            fr = synth.planes_only(pixels_for(synth, 29, w, h, samp), samp)
            ref = oracle.reconstruct(fr)
            for flag, want in ((capi.FLAG_ROTATE_90, np.rot90(ref, 1)), (capi.FLAG_ROTATE_270, np.rot90(ref, -1))):
                d = frame_desc(fr, flag | capi.FLAG_PACKED_ROWS)
                assert emu.layout(d) != 6
                got = emu.reconstruct(d, fr.planes, 3 * h, w)[:, :3 * h].reshape(w, h, 3)
                assert np.array_equal(got, want), (samp, rows, w, h, flag)
    finally:
        emu.set_turn_tile_rows(0)


def test_idct_blocks_match_oracle(emu, oracle):
    """Block-level: the kernel IDCT (shortcut folded into the rounding constant,
    column shortcut dropped, re-associated rotations) == GID's, on blocks built
    to hit every special case."""
    rng = np.random.default_rng(99)
    for k in range(3000):
        blk = np.zeros(64, dtype=np.int64)
        mode = k % 6
        if mode == 0:
            blk[:] = rng.integers(-2000, 2001, size=64)
        elif mode == 1:
            blk[0] = rng.integers(-500000, 500001)           # DC only, large
        elif mode == 2:
            blk[::8] = rng.integers(-3000, 3001, size=8)     # first column only: all rows shortcut
        elif mode == 3:
            blk[:8] = rng.integers(-3000, 3001, size=8)      # first row only: column shortcut territory
        elif mode == 4:
            idx = rng.choice(64, size=3, replace=False)
            blk[idx] = rng.integers(-30000, 30001, size=3)
        else:
            blk[:] = rng.integers(-3, 4, size=64)            # tiny values around the truncation points
        assert np.array_equal(emu.idct(blk), oracle.idct(blk)), (k, blk.tolist())


def test_colour_identity_exhaustive(emu):
    """Clip(Y + (T >> 8)) == Clip((Y*256 + T) / 256) for all 2**24 (Y, Cb, Cr)."""
    assert emu.chroma_exhaustive() == 0
