"""GPU parity tests: the CUDA path (through the C ABI) vs the oracle, bit-exact.

Bar: identical RGB bytes (integer path, no tolerance).
"""
import ctypes

import numpy as np
import pytest

from tests.cases import CONFIG_SMALL, GEOMETRIES, edge_frames, pixels_for, random_planes
from tests.harness import frame_desc

pytestmark = pytest.mark.gpu


def _check(gpu, oracle, fr, flags=0, out_stride=0):
    from gid_b200 import capi
    ref = oracle.reconstruct(fr)
    d = frame_desc(fr, flags, out_stride)
    got = gpu.reconstruct_rgb(d, fr.planes)
    if (flags & 0xFF) == capi.OUT_BGR8_BOTTOM_UP:
        got = got[::-1, :, ::-1]
    if (flags & 0xFF) == capi.OUT_RGBA8:
        assert (got[:, :, 3] == 255).all()
        got = got[:, :, :3]
    assert got.shape == ref.shape
    if not np.array_equal(got, ref):
        bad = np.argwhere(got != ref)
        raise AssertionError(f"{len(bad)} differing samples, first at (y, x, c) = {bad[0].tolist()}: "
                             f"gpu {got[tuple(bad[0])]} oracle {ref[tuple(bad[0])]}")


@pytest.fixture(params=["tma", "ldg"])
def path(request):
    """Both coefficient paths: TMA-staged ring (default) and plain 128-bit loads."""
    from gid_b200 import capi
    capi.set_option("tma", 1 if request.param == "tma" else 0)
    yield request.param
    capi.set_option("tma", 1)


@pytest.mark.parametrize("w,h,samp", GEOMETRIES)
def test_geometries_match_oracle(gpu, oracle, synth, path, w, h, samp):
    fr = synth.planes_only(pixels_for(synth, 11, w, h, samp), samp)
    _check(gpu, oracle, fr)


@pytest.mark.parametrize("cfg", CONFIG_SMALL, ids=[c["name"] for c in CONFIG_SMALL])
def test_configs_through_jpeg_files(gpu, oracle, synth, cfg):
    """synthetic .jpg -> oracle front end (GID's entropy decode) -> planes -> GPU;
    compared with the oracle's full MCU-by-MCU decode of the same file."""
    jpg, enc = synth.encode(pixels_for(synth, 5, cfg["w"], cfg["h"], cfg["samp"]), cfg["samp"],
                            progressive=cfg["prog"], restart_interval=cfg["ri"])
    fr, rgb, _ = oracle.decode(jpg)
    for c in range(fr.ncomp):
        assert np.array_equal(fr.planes[c], enc.planes[c])
    got = gpu.reconstruct_rgb(frame_desc(fr), fr.planes)
    assert np.array_equal(got, rgb)


@pytest.mark.parametrize("kind", ["dense", "sparse", "dc_only", "rows", "extreme"])
@pytest.mark.parametrize("samp", ["grey", "444", "422", "420", "440"])
def test_random_coefficients(gpu, oracle, synth, path, kind, samp):
    from tests.harness import Frame, SAMPLING
    import zlib
    rng = np.random.default_rng(zlib.crc32(f"{kind}-{samp}".encode()))
    w, h = 200, 136
    bx, by = synth.grid(w, h, samp)
    sh, sv = SAMPLING[samp]
    qt = rng.integers(1, 256, size=(len(sh), 64)).astype(np.uint16)
    fr = Frame(w, h, len(sh), list(sh), list(sv), bx, by, qt)
    fr.planes = random_planes(rng, bx, by, kind)
    # |coef * q| <= 2047 * 255 < 2**20: inside the documented domain
    _check(gpu, oracle, fr)


def test_edge_of_the_domain(gpu, oracle, synth, path):
    """Tables with entries of 0 (all AC entries, scattered), 16-bit tables whose products are multiples of
    2**21, wide coefficients, DC terms that wrap, full-range garbage, the rim of the fast domain: identical
    bytes (the exact kernel instances outside the fast domain, the fast ones on its rim)."""
    for name, fr in edge_frames(synth):
        try:
            _check(gpu, oracle, fr)
            if fr.ncomp == 3:
                from gid_b200 import capi
                _check(gpu, oracle, fr, flags=capi.OUT_RGBA8)
        except AssertionError as e:
            raise AssertionError(f"{name}: {e}") from e


def test_wide_coefficients_declared_after_begin(gpu, oracle, synth):
    """gidcuda_job_wide_coefficients (what a decoder calls once it has seen the coefficients) == the flag in
    the descriptor; without either, the same planes run the fast instances and the bytes differ (which is
    why the producer has to say so)."""
    from gid_b200 import capi
    name, fr = next((n, f) for n, f in edge_frames(synth) if n == "dc_terms_that_wrap_444")
    ref = oracle.reconstruct(fr)
    d = frame_desc(fr)
    assert d.flags & capi.FLAG_WIDE_COEF
    assert np.array_equal(gpu.reconstruct_rgb(d, fr.planes), ref)
    d.flags &= ~capi.FLAG_WIDE_COEF
    assert np.array_equal(gpu.reconstruct_rgb(d, fr.planes), ref)      # the helper notices and calls the setter
    L = gpu.lib
    for c in range(3):
        assert 257 <= L.gidcuda_fast_coef_limit(ctypes.byref(d), c) <= 32767


def test_recycled_job_with_row_hints_of_another_layout(gpu, oracle, synth):
    """A hinted plane (rows 4..7 not moved) relies on device rows that were zeroed for THIS frame: a job
    recycled from a frame of another layout and size must zero them again."""
    rng = np.random.default_rng(5)
    frames = []
    for w, h, samp in ((4096, 2304, "420"), (4096, 3072, "grey"), (4096, 2304, "420")):
        fr = synth.planes_only(pixels_for(synth, 21 + len(frames), w, h, samp), samp)
        for c in range(fr.ncomp):          # rows 4..7 empty: the hint applies; the first frame's chroma is dense
            fr.planes[c][..., 32:] = 0
        if len(frames) == 0:
            fr.planes[1][..., 32:] = rng.integers(-3, 4, size=fr.planes[1][..., 32:].shape)
            fr.planes[2][..., 32:] = rng.integers(-3, 4, size=fr.planes[2][..., 32:].shape)
        frames.append(fr)
    for fr in frames:
        got = gpu.reconstruct_rgb(frame_desc(fr), fr.planes, hint_rows=True)
        assert np.array_equal(got, oracle.reconstruct(fr)), (fr.width, fr.height, fr.ncomp)


def test_16bit_quant_tables(gpu, oracle, synth):
    from tests.harness import Frame, SAMPLING
    rng = np.random.default_rng(7)
    for samp in ("grey", "420", "444"):
        w, h = 120, 72
        bx, by = synth.grid(w, h, samp)
        sh, sv = SAMPLING[samp]
        qt = rng.integers(1, 2000, size=(len(sh), 64)).astype(np.uint16)
        fr = Frame(w, h, len(sh), list(sh), list(sv), bx, by, qt)
        fr.planes = [np.clip(p, -400, 400) for p in random_planes(rng, bx, by, "sparse")]
        _check(gpu, oracle, fr)


@pytest.mark.parametrize("samp", ["grey", "444", "422", "420", "411"])
def test_output_formats_and_strides(gpu, oracle, synth, samp):
    from gid_b200 import capi
    fr = synth.planes_only(pixels_for(synth, 3, 211, 77, samp), samp)
    _check(gpu, oracle, fr, flags=capi.FLAG_PACKED_ROWS)         # unaligned rows
    _check(gpu, oracle, fr, flags=capi.OUT_BGR8_BOTTOM_UP)       # to_bmp layout
    _check(gpu, oracle, fr, out_stride=3 * 211 + 29)             # odd caller-chosen stride
    _check(gpu, oracle, fr, out_stride=1024)
    _check(gpu, oracle, fr, flags=capi.OUT_RGBA8)                # R, G, B, 255
    _check(gpu, oracle, fr, flags=capi.OUT_RGBA8 | capi.FLAG_PACKED_ROWS)
    _check(gpu, oracle, fr, flags=capi.OUT_RGBA8, out_stride=4 * 211 + 4)


@pytest.mark.parametrize("w,h,samp", [(65535, 9, "grey"), (9, 65535, "420"), (65535, 17, "422"), (33, 65535, "444")])
def test_maximum_dimensions(gpu, oracle, synth, w, h, samp):
    """The largest width / height a JPEG frame header can carry (16 bits, :206-215), with partial
    MCUs on the far edge; dense and record input."""
    rng = np.random.default_rng(w ^ h)
    gx, gy = synth.grid(w, h, samp)
    from tests.harness import Frame
    ref_frame = synth.planes_only(pixels_for(synth, 1, 64, 64, samp), samp)   # for the tables
    planes = random_planes(rng, gx, gy, "sparse")
    n = ref_frame.ncomp
    fr = Frame(width=w, height=h, ncomp=n, samp_h=ref_frame.samp_h, samp_v=ref_frame.samp_v,
               blocks_x=list(gx[:n]), blocks_y=list(gy[:n]), qt=ref_frame.qt, planes=planes[:n])
    ref = oracle.reconstruct(fr)
    assert np.array_equal(gpu.reconstruct_rgb(frame_desc(fr), fr.planes), ref)
    assert np.array_equal(gpu.reconstruct_rgb(frame_desc(fr), fr.planes, sparse="auto"), ref)


def _to_bmp_buffer(ref, orientation):
    """The byte buffer test/to_bmp.adb builds from GID's callbacks (:94-146): Set_X_Y (x, y) with y
    counted from the bottom picks idx per orientation, Put_Pixel stores (blue, green, red)."""
    h, w = ref.shape[:2]
    pls_x, pls_y = 4 * -(-(w * 3) // 4), 4 * -(-(h * 3) // 4)
    out_h = w if orientation in (1, 3) else h
    buf = np.zeros((pls_y if orientation in (1, 3) else pls_x) * out_h, dtype=np.uint8)
    x = np.arange(w)[None, :].repeat(h, 0)
    y = (h - 1 - np.arange(h))[:, None].repeat(w, 1)     # GID's y: 0 = bottom row
    rev_x, rev_y = w - (x + 1), h - (y + 1)
    idx = [3 * x + pls_x * y, 3 * rev_y + pls_y * x, 3 * rev_x + pls_x * rev_y, 3 * y + pls_y * rev_x][orientation]
    for k in range(3):
        buf[(idx + k).ravel()] = ref[:, :, 2 - k].ravel()
    return buf


@pytest.mark.parametrize("samp", ["grey", "444", "422", "420", "440", "411"])
@pytest.mark.parametrize("w,h", [(512, 40), (1064, 56), (2080, 24)])
def test_turned_frames_of_several_tile_columns(gpu, oracle, synth, samp, w, h):
    """Quarter-turned frames list their tiles column by column (launch_set.hpp, append_tiles): frames wide enough
    for several tiles per MCU row, with and without a cut last tile, MCU counts that are and are not multiples
    of 32 (grey / 4:4:4: the linear runs then wrap and keep their order)."""
    from gid_b200 import capi
    fr = synth.planes_only(pixels_for(synth, 31, w, h, samp), samp)
    ref = oracle.reconstruct(fr)
    for k, flag in ((1, capi.FLAG_ROTATE_90), (-1, capi.FLAG_ROTATE_270)):
        got = gpu.reconstruct_rgb(frame_desc(fr, flag), fr.planes)
        assert np.array_equal(got, np.rot90(ref, k)), (samp, w, h, k)
    got = gpu.reconstruct_rgb(frame_desc(fr, capi.FLAG_ROTATE_180 | capi.OUT_RGBA8), fr.planes)
    assert np.array_equal(got[:, :, :3], ref[::-1, ::-1])


@pytest.mark.parametrize("samp", ["grey", "444", "422", "420", "440"])
def test_display_orientation_applied_to_the_pixels(gpu, oracle, synth, samp):
    """GIDCUDA_FLAG_ROTATE_*: GID.Display_Orientation (EXIF) applied while the pixels are written, as
    the client of test/to_bmp.adb does per pixel in Set_X_Y / Put_Pixel.  The bottom-up B, G, R
    buffer must be byte for byte the one to_bmp's formulas fill; the top-down formats its mirror."""
    from gid_b200 import capi
    fr = synth.planes_only(pixels_for(synth, 6, 131, 67, samp), samp)
    ref = oracle.reconstruct(fr)
    turned = {1: np.rot90(ref, 1), 2: ref[::-1, ::-1], 3: np.rot90(ref, -1)}
    for orientation, flag in ((1, capi.FLAG_ROTATE_90), (2, capi.FLAG_ROTATE_180), (3, capi.FLAG_ROTATE_270)):
        want = turned[orientation]
        got = gpu.reconstruct_rgb(frame_desc(fr, flag | capi.FLAG_PACKED_ROWS), fr.planes)
        assert got.shape == want.shape and np.array_equal(got, want), (samp, orientation)
        got = gpu.reconstruct_rgb(frame_desc(fr, flag | capi.OUT_RGBA8), fr.planes)
        assert np.array_equal(got[:, :, :3], want) and (got[:, :, 3] == 255).all()
        rows = gpu.reconstruct(frame_desc(fr, flag | capi.OUT_BGR8_BOTTOM_UP), fr.planes)
        used = 3 * want.shape[1]  # (the bytes that pad a row to a multiple of 4 are nobody's)
        assert np.array_equal(rows[:, :used], _to_bmp_buffer(ref, orientation).reshape(rows.shape)[:, :used]), (samp, orientation)
    rows = gpu.reconstruct(frame_desc(fr, capi.OUT_BGR8_BOTTOM_UP), fr.planes)
    assert np.array_equal(rows[:, :3 * 131], _to_bmp_buffer(ref, 0).reshape(rows.shape)[:, :3 * 131])


def test_job_recycling_and_state_errors(gpu, oracle, synth):
    from gid_b200 import GidCudaError
    a = synth.planes_only(pixels_for(synth, 1, 320, 240, "420"), "420")
    b = synth.planes_only(pixels_for(synth, 2, 64, 48, "444"), "444")
    for fr in (a, b, a, b):
        _check(gpu, oracle, fr)
    L = gpu.lib
    job = ctypes.c_void_p()
    d = frame_desc(a)
    assert L.gidcuda_job_begin(gpu.ctx, ctypes.byref(d), ctypes.byref(job)) == 0
    pix, st = ctypes.c_void_p(), ctypes.c_size_t()
    assert L.gidcuda_job_wait(job, ctypes.byref(pix), ctypes.byref(st)) == -6  # STATE: not submitted
    assert L.gidcuda_job_release(job) == 0
    bad = frame_desc(a)
    bad.samp_h[0] = 3  # 3 / 1 chroma is fine, but 3x.. with chroma 1 -> ok; make it invalid instead:
    bad.samp_h[1] = 2  # hmax 3 not divisible by 2
    with pytest.raises(GidCudaError):
        gpu.reconstruct_rgb(bad, a.planes)


def test_plan_device_resident_batch(gpu, oracle, synth):
    """Device-resident path: planes and pixels in HBM (torch tensors), one launch
    per layout over a mixed batch."""
    import torch
    frames = [synth.planes_only(pixels_for(synth, 20 + i, w, h, s), s)
              for i, (w, h, s) in enumerate([(320, 200, "420"), (97, 61, "420"), (128, 64, "444"),
                                             (200, 120, "grey"), (64, 64, "422"), (90, 50, "440"),
                                             (640, 360, "420"), (130, 40, "411"), (96, 64, "410")])]
    dev = torch.device("cuda:0")
    descs, d_planes, d_pixels, keep = [], [], [], []
    for fr in frames:
        d = frame_desc(fr)
        stride, nbytes = gpu.output_geometry(d)
        descs.append(d)
        for c in range(3):
            if c < fr.ncomp:
                t = torch.from_numpy(fr.planes[c].reshape(-1)).to(dev)
                keep.append(t)
                d_planes.append(t.data_ptr())
            else:
                d_planes.append(0)
        out = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        keep.append(out)
        d_pixels.append(out.data_ptr())
    torch.cuda.synchronize()
    plan = gpu.plan_create(descs, d_planes, d_pixels)
    assert gpu.plan_launch_count(plan) == 6 + 2  # grey, 444, 422, 420, 440, 411 fused + generic 4:1:0 (2 passes)
    assert gpu.plan_uses_tma(plan) == 6          # torch allocations are 512-byte aligned
    refs = [oracle.reconstruct(fr) for fr in frames]
    # the same plan launched several times back to back: the kernels hand out tiles through
    # launch-wide counters that the last warp resets, so every launch must be complete
    for rep in range(4):
        for t in keep:
            if t.dtype == torch.uint8:
                t.zero_()
        gpu.plan_launch(plan, torch.cuda.current_stream().cuda_stream)
        if rep == 2:  # and once without any host synchronisation in between
            gpu.plan_launch(plan, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        k = 0
        for fr, d, ref in zip(frames, descs, refs):
            stride, _ = gpu.output_geometry(d)
            while keep[k].dtype != torch.uint8:
                k += 1
            rows = keep[k].cpu().numpy().reshape(fr.height, stride)
            k += 1
            got = rows[:, :3 * fr.width].reshape(fr.height, fr.width, 3)
            assert np.array_equal(got, ref), rep
    gpu.plan_destroy(plan)


def test_plan_with_turned_frames(gpu, oracle, synth):
    """A plan over frames of every orientation: one launch per (layout, orientation) group, every frame where
    to_bmp's Set_X_Y would put it."""
    import torch
    from gid_b200 import capi
    flags = [0, capi.FLAG_ROTATE_90, capi.FLAG_ROTATE_180, capi.FLAG_ROTATE_270]
    spec = [(320, 200, "420"), (1064, 72, "420"), (256, 64, "444"), (200, 120, "grey")]
    frames, descs, d_planes, d_pixels, outs, keep = [], [], [], [], [], []
    dev = torch.device("cuda:0")
    for i, (w, h, s) in enumerate(spec):
        for flag in flags:
            fr = synth.planes_only(pixels_for(synth, 40 + i, w, h, s), s)
            d = frame_desc(fr, flag)
            _, nbytes = gpu.output_geometry(d)
            frames.append((fr, flag))
            descs.append(d)
            for c in range(3):
                if c < fr.ncomp:
                    t = torch.from_numpy(fr.planes[c].reshape(-1)).to(dev)
                    keep.append(t)
                    d_planes.append(t.data_ptr())
                else:
                    d_planes.append(0)
            out = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
            outs.append(out)
            d_pixels.append(out.data_ptr())
    torch.cuda.synchronize()
    plan = gpu.plan_create(descs, d_planes, d_pixels)
    assert gpu.plan_launch_count(plan) == 3 * 4  # three layouts x four orientations
    for rep in range(2):
        gpu.plan_launch(plan, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    for (fr, flag), d, out in zip(frames, descs, outs):
        ref = oracle.reconstruct(fr)
        want = {0: ref, capi.FLAG_ROTATE_90: np.rot90(ref, 1), capi.FLAG_ROTATE_180: ref[::-1, ::-1],
                capi.FLAG_ROTATE_270: np.rot90(ref, -1)}[flag]
        stride, _ = gpu.output_geometry(d)
        oh, ow = want.shape[:2]
        got = out.cpu().numpy().reshape(oh, stride)[:, :3 * ow].reshape(oh, ow, 3)
        assert np.array_equal(got, want), (fr.width, fr.height, flag)
    gpu.plan_destroy(plan)


@pytest.mark.parametrize("samp", ["444", "420", "grey", "440"])
def test_row_hints(gpu, oracle, synth, samp):
    """gidcuda_job_hint_rows: planes whose lower block rows are empty travel as pitched copies;
    the result must not change, also when a recycled job switches between hinted and full
    planes (the device copy of the skipped rows has to be zero, not stale)."""
    from gid_b200 import capi
    # planes of at least 8 MiB: smaller ones are always copied whole
    W, H = (2304, 1296) if samp != "grey" else (3072, 3072)
    if samp in ("420", "440"):
        W, H = 4096, 2304
    fr = synth.planes_only(pixels_for(synth, 70, W, H, samp), samp)
    dense = synth.planes_only(pixels_for(synth, 71, W, H, samp), samp)
    rng = np.random.default_rng(3)
    dense.planes = [rng.integers(-40, 41, size=p.shape).astype(np.int16) for p in dense.planes]
    sparse = synth.planes_only(pixels_for(synth, 72, W, H, samp), samp)
    for p in sparse.planes:
        p[:, :, 16:] = 0          # rows 2 .. 7 empty everywhere
    assert all(capi.rows_used(p) <= 2 for p in sparse.planes)
    for f in (fr, dense, sparse, fr, sparse, dense, sparse):
        got = gpu.reconstruct_rgb(frame_desc(f), f.planes, hint_rows=True)
        assert np.array_equal(got, oracle.reconstruct(f))
    # a wrong component / row count is rejected
    L = gpu.lib
    job = ctypes.c_void_p()
    d = frame_desc(fr)
    assert L.gidcuda_job_begin(gpu.ctx, ctypes.byref(d), ctypes.byref(job)) == 0
    assert L.gidcuda_job_hint_rows(job, 0, 0) == -1 and L.gidcuda_job_hint_rows(job, 3, 4) == -1
    assert L.gidcuda_job_hint_rows(job, 0, 4) == 0
    assert L.gidcuda_job_release(job) == 0


def test_job_resubmission(gpu, oracle, synth):
    """A job submitted again after wait (what bench.py's e2e leg does) reconstructs again, also
    after its planes were rewritten in place."""
    a = synth.planes_only(pixels_for(synth, 61, 1000, 700, "420"), "420")
    b = synth.planes_only(pixels_for(synth, 62, 1000, 700, "420"), "420")
    L = gpu.lib
    job = ctypes.c_void_p()
    d = frame_desc(a)
    assert L.gidcuda_job_begin(gpu.ctx, ctypes.byref(d), ctypes.byref(job)) == 0
    pix, st = ctypes.c_void_p(), ctypes.c_size_t()
    for rep, fr in enumerate((a, a, b, a)):
        for c in range(fr.ncomp):
            p = L.gidcuda_job_plane(job, c, None, None)
            ctypes.memmove(p, fr.planes[c].ctypes.data, fr.planes[c].nbytes)
        assert L.gidcuda_job_submit(job) == 0
        assert L.gidcuda_job_wait(job, ctypes.byref(pix), ctypes.byref(st)) == 0
        rows = np.frombuffer((ctypes.c_uint8 * (st.value * fr.height)).from_address(pix.value), dtype=np.uint8)
        got = rows.reshape(fr.height, st.value)[:, :3 * fr.width].reshape(fr.height, fr.width, 3)
        assert np.array_equal(got, oracle.reconstruct(fr)), rep
    assert L.gidcuda_job_release(job) == 0


def test_batch_driver(oracle, synth, gidcuda_lib):
    """Batch driver: host buffers in, host buffers out, dynamic sharding; a bad
    image is reported and skipped."""
    from gid_b200 import capi
    n_dev = gidcuda_lib.gidcuda_device_count()
    assert n_dev >= 1
    geos = [(320, 200, "420"), (97, 61, "420"), (128, 64, "444"), (200, 120, "grey"), (64, 64, "422"),
            (90, 50, "440")]
    frames = [synth.planes_only(pixels_for(synth, 40 + i, *geos[i % len(geos)][:2], geos[i % len(geos)][2]),
                                geos[i % len(geos)][2]) for i in range(23)]
    descs = [frame_desc(fr) for fr in frames]
    planes, pixels, outs = [], [], []
    ctx = capi.GidCuda(0)
    for fr, d in zip(frames, descs):
        for c in range(3):
            planes.append(fr.planes[c].ctypes.data if c < fr.ncomp else 0)
        stride, nbytes = ctx.output_geometry(d)
        o = np.zeros(nbytes, dtype=np.uint8)
        outs.append((o, stride))
        pixels.append(o.ctypes.data)
    ctx.close()
    pixels[5] = 0  # broken request: no output buffer
    batch = capi.Batch(list(range(n_dev)), chunk_images=4)
    status, stats = batch.run(descs, planes, pixels)
    batch.close()
    assert status[5] != 0 and stats.images_failed == 1
    assert sum(stats.images_per_device[:n_dev]) == 22
    for i, (fr, (o, stride)) in enumerate(zip(frames, outs)):
        if i == 5:
            continue
        assert status[i] == 0
        got = o.reshape(fr.height, stride)[:, :3 * fr.width].reshape(fr.height, fr.width, 3)
        assert np.array_equal(got, oracle.reconstruct(fr)), i


def test_batch_driver_every_device_pinned_and_records(oracle, synth, gidcuda_lib):
    """The in-process batch driver on EVERY visible device: per-GPU worker, work queue and three chunks in
    flight; images_per_device is non-zero on each device.  Buffers in pinned memory are read and written by
    the copy engines directly (no staging copy); the records variant moves what an entropy decoder emits."""
    from gid_b200 import capi
    n_dev = gidcuda_lib.gidcuda_device_count()
    geos = [(640, 360, "420"), (512, 256, "444"), (400, 300, "grey"), (320, 200, "422")]
    distinct = [synth.planes_only(pixels_for(synth, 60 + i, *g[:2], g[2]), g[2]) for i, g in enumerate(geos)]
    refs = [oracle.reconstruct(fr) for fr in distinct]
    n = 24 * n_dev
    frames = [distinct[i % len(distinct)] for i in range(n)]
    descs = [frame_desc(fr) for fr in frames]
    ctx = capi.GidCuda(0)
    geo = [ctx.output_geometry(d) for d in descs]
    ctx.close()
    allocs = []

    def pinned_copy(a):
        buf, addr = capi.pinned_empty(max(a.nbytes, 16))
        allocs.append(addr)
        buf[:a.nbytes] = np.frombuffer(a.tobytes(), dtype=np.uint8)
        return addr

    def check(outs, status, stats):
        assert all(st == 0 for st in status) and stats.images_failed == 0
        assert sum(stats.images_per_device[:n_dev]) == n
        assert all(stats.images_per_device[k] > 0 for k in range(n_dev)), list(stats.images_per_device[:n_dev])
        for i, fr in enumerate(frames):
            stride = geo[i][0]
            got = outs[i][:fr.height * stride].reshape(fr.height, stride)[:, :3 * fr.width].reshape(fr.height, fr.width, 3)
            assert np.array_equal(got, refs[i % len(distinct)]), i

    batch = capi.Batch(list(range(n_dev)), chunk_images=2)
    try:
        # dense planes, pinned in and out
        src = [[pinned_copy(fr.planes[c]) for c in range(fr.ncomp)] for fr in distinct]
        planes, pixels, outs = [], [], []
        for i, fr in enumerate(frames):
            planes += [src[i % len(distinct)][c] if c < fr.ncomp else 0 for c in range(3)]
            o, addr = capi.pinned_empty(geo[i][1])
            allocs.append(addr)
            o[:] = 0
            outs.append(o)
            pixels.append(addr)
        status, stats = batch.run(descs, planes, pixels)
        check(outs, status, stats)
        assert stats.d2h_bytes == sum(g[1] for g in geo)
        # records, pageable in, pinned out
        recs = []
        for fr in distinct:
            per = []
            for c in range(fr.ncomp):
                first, rec = capi.records_from_plane(fr.planes[c], fr.samp_h[c], fr.samp_v[c])
                per.append((np.ascontiguousarray(first, dtype=np.uint32), np.ascontiguousarray(rec, dtype=np.uint32)))
            recs.append(per)
        fa, ra, ca = [], [], []
        for i, fr in enumerate(frames):
            per = recs[i % len(distinct)]
            for c in range(3):
                fa.append(per[c][0].ctypes.data if c < fr.ncomp else 0)
                ra.append(per[c][1].ctypes.data if c < fr.ncomp else 0)
                ca.append(len(per[c][1]) if c < fr.ncomp else 0)
        for o in outs:
            o[:] = 0
        status, stats2 = batch.run_records(descs, fa, ra, ca, pixels)
        check(outs, status, stats2)
        assert stats2.h2d_bytes < stats.h2d_bytes / 2     # the link carries the records, not the planes
    finally:
        batch.close()
        for addr in allocs:
            capi.pinned_free(addr)


FULL_SIZE = [
    ("cfg1", 512, 512, "420", False, 0),
    ("cfg2", 4096, 4096, "444", False, 0),
    ("cfg3", 1920, 1080, "420", False, 0),
    ("cfg4", 8192, 8192, "grey", False, 64),
    ("cfg5", 2048, 2048, "422", True, 0),
]


@pytest.mark.parametrize("name,w,h,samp,prog,ri", FULL_SIZE, ids=[c[0] for c in FULL_SIZE])
def test_baseline_configs_full_size(gpu, oracle, synth, name, w, h, samp, prog, ri):
    """The five BASELINE.json configs at their full frame size: synthetic .jpg ->
    GID front end (oracle) -> planes -> GPU, against the oracle's own decode."""
    jpg, enc = synth.encode(pixels_for(synth, 77, w, h, samp), samp, progressive=prog, restart_interval=ri)
    fr, rgb, fb = oracle.decode(jpg)
    for c in range(fr.ncomp):
        assert np.array_equal(fr.planes[c], enc.planes[c])
    got = gpu.reconstruct_rgb(frame_desc(fr), fr.planes)
    assert np.array_equal(got, rgb)
    assert fb == sorted(fb) and fb[-1] == 100


def test_batch_properties_full_size(oracle, synth, gidcuda_lib):
    """Batch configs (3 and 5) at scale through the batch driver: the batch is built
    from a small pool of distinct frames in a shuffled order, so every output must
    equal the (oracle-checked) output of its pool member -- a size-independent
    check of sharding and indexing."""
    from gid_b200 import capi
    n_dev = gidcuda_lib.gidcuda_device_count()
    for (w, h, samp, count) in ((1920, 1080, "420", 96), (2048, 2048, "422", 24)):
        pool = [synth.planes_only(pixels_for(synth, 300 + i, w, h, samp), samp) for i in range(3)]
        refs = [oracle.reconstruct(fr) for fr in pool]
        rng = np.random.default_rng(5)
        order = rng.integers(0, len(pool), size=count)
        ctx = capi.GidCuda(0)
        descs = [frame_desc(pool[k]) for k in order]
        stride, nbytes = ctx.output_geometry(descs[0])
        ctx.close()
        outs = [np.zeros(nbytes, dtype=np.uint8) for _ in order]
        planes = []
        for k in order:
            planes += [pool[k].planes[c].ctypes.data for c in range(3)]
        batch = capi.Batch(list(range(n_dev)), chunk_images=8)
        status, stats = batch.run(descs, planes, [o.ctypes.data for o in outs])
        batch.close()
        assert all(s == 0 for s in status) and stats.images_failed == 0
        assert sum(stats.images_per_device[:n_dev]) == count
        for k, o in zip(order, outs):
            got = o.reshape(h, stride)[:, :3 * w].reshape(h, w, 3)
            assert np.array_equal(got, refs[k])


# ---- sparse coefficient input (gidcuda_job_records) -----------------------------------


@pytest.mark.parametrize("w,h,samp", GEOMETRIES)
def test_sparse_records_match_oracle(gpu, oracle, synth, w, h, samp):
    """The same frames supplied as coefficient records instead of dense planes."""
    fr = synth.planes_only(pixels_for(synth, 13, w, h, samp), samp)
    ref = oracle.reconstruct(fr)
    # "auto": a component denser than 32 records per block (tiny noisy frames) goes as a plane
    got = gpu.reconstruct_rgb(frame_desc(fr), fr.planes, sparse="auto")
    assert np.array_equal(got, ref)


def test_sparse_records_mixed_and_resubmitted(gpu, oracle, synth):
    """Components may be mixed (records / dense plane / hinted dense plane); a job can be
    filled again after a wait with different data in the other form."""
    fr = synth.planes_only(pixels_for(synth, 17, 333, 217, "420"), "420")
    ref = oracle.reconstruct(fr)
    d = frame_desc(fr)
    for sp in ([True, False, False], [False, True, True], [True, True, False]):
        got = gpu.reconstruct_rgb(d, fr.planes, sparse=sp, hint_rows=True)
        assert np.array_equal(got, ref), sp
    # same job object recycled: dense after sparse and sparse after dense give the same pixels
    assert np.array_equal(gpu.reconstruct_rgb(d, fr.planes), ref)
    assert np.array_equal(gpu.reconstruct_rgb(d, fr.planes, sparse=True), ref)


def test_sparse_records_random_and_limits(gpu, oracle, synth):
    from tests.harness import Frame, SAMPLING
    rng = np.random.default_rng(21)
    for samp in ("grey", "444", "422", "420", "411"):
        w, h = 200, 136
        bx, by = synth.grid(w, h, samp)
        sh, sv = SAMPLING[samp]
        qt = rng.integers(1, 256, size=(len(sh), 64)).astype(np.uint16)
        fr = Frame(w, h, len(sh), list(sh), list(sv), bx, by, qt)
        fr.planes = random_planes(rng, bx, by, "sparse")
        ref = oracle.reconstruct(fr)
        assert np.array_equal(gpu.reconstruct_rgb(frame_desc(fr), fr.planes, sparse=True), ref), samp
    # exactly at capacity (32 records per block on average) is accepted ...
    bx, by = synth.grid(64, 64, "grey")
    fr = Frame(64, 64, 1, [1], [1], bx, by, np.full((1, 64), 3, dtype=np.uint16))
    p = np.zeros((by[0], bx[0], 64), dtype=np.int16)
    p[:, :, :32] = rng.integers(1, 50, size=(by[0], bx[0], 32))
    fr.planes = [p]
    assert np.array_equal(gpu.reconstruct_rgb(frame_desc(fr), fr.planes, sparse=True), oracle.reconstruct(fr))
    # ... one more is refused by the binding (the decoder then uses the dense plane)
    p[0, 0, 40] = 7
    with pytest.raises(OverflowError):
        gpu.reconstruct_rgb(frame_desc(fr), fr.planes, sparse=True)
    # a commit whose first[] does not frame the records is rejected
    L = gpu.lib
    job = ctypes.c_void_p()
    d = frame_desc(fr)
    assert L.gidcuda_job_begin(gpu.ctx, ctypes.byref(d), ctypes.byref(job)) == 0
    pf, pr, cap, nb = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_size_t(), ctypes.c_int32()
    assert L.gidcuda_job_records(job, 0, ctypes.byref(pf), ctypes.byref(pr), ctypes.byref(cap), ctypes.byref(nb)) == 0
    assert nb.value == bx[0] * by[0] and cap.value == 32 * nb.value
    ctypes.memset(pf, 0, 4 * (nb.value + 1))
    assert L.gidcuda_job_records_commit(job, 0, 5) == -1
    assert L.gidcuda_job_submit(job) == -6      # neither planes nor records
    assert L.gidcuda_job_release(job) == 0
    with pytest.raises(ValueError):             # wrong number of blocks for the frame
        gpu.reconstruct_rgb(frame_desc(fr), [p[:1]], sparse=True)


def test_job_scan_argument_checks(gpu, synth):
    """gidcuda_job_scan validates its geometry, offsets and tables (no decode launched)."""
    import ctypes
    from gid_b200 import capi
    L = gpu.lib
    fr = synth.planes_only(synth.image(1, 64, 64)[:, :, 1].copy(), "grey")
    desc = frame_desc(fr)
    job = ctypes.c_void_p()
    assert L.gidcuda_job_begin(gpu.ctx, ctypes.byref(desc), ctypes.byref(job)) == 0
    scan = capi.ScanDesc()
    scan.restart_interval, scan.nintervals = 8, 8   # 64 MCUs
    data = (ctypes.c_uint8 * 16)()
    b = (ctypes.c_uint32 * 8)(*range(8))
    e = (ctypes.c_uint32 * 8)(*range(1, 9))
    scan.nintervals = 7
    assert L.gidcuda_job_scan(job, ctypes.byref(scan), data, 16, b, e) == -1      # intervals do not cover the frame
    scan.nintervals = 8
    e[3] = 99
    assert L.gidcuda_job_scan(job, ctypes.byref(scan), data, 16, b, e) == -1      # offset outside the data
    e[3] = 4
    scan.dc[0].counts[0] = 3                                                      # three 1-bit codes: not a prefix code
    assert L.gidcuda_job_scan(job, ctypes.byref(scan), data, 16, b, e) == -1
    scan.dc[0].counts[0] = 0
    assert L.gidcuda_job_scan(job, ctypes.byref(scan), data, 16, b, e) == 0       # empty tables are a valid (useless) code
    assert L.gidcuda_job_submit(job) == 0
    pix, st = ctypes.c_void_p(), ctypes.c_size_t()
    assert L.gidcuda_job_wait(job, ctypes.byref(pix), ctypes.byref(st)) == -7     # no code matches: left to the host
    assert L.gidcuda_job_release(job) == 0
