"""The C++ host mirror of GID's API (gid_b200/host): Load_Image_Header /
Load_Image_Contents over the C ABI, with its own entropy decoder.

CPU part: header fields, error classes (= GID's exceptions), and the entropy
decoder's coefficient planes against the oracle's front end.  GPU part: the whole
path -- .jpg in, Put_Pixel out -- against the oracle's decode, bit-exact.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from tests.cases import CONFIG_SMALL, pixels_for

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def host(gidcuda_lib):
    from gid_b200 import build, host_api
    if not os.path.exists(host_api.library_path()):
        build.build_host()
    host_api.load_library()
    return host_api


def _golden_files():
    index = json.load(open(os.path.join(GOLDEN, "index.json")))
    return [(e["jpg"], e) for e in index["files"]]


def test_library_exports(host):
    L = host.load_library()
    for name in host.EXPORTS:
        assert hasattr(L, name), name


@pytest.mark.parametrize("name,ent", _golden_files(), ids=[n for n, _ in _golden_files()])
def test_header_and_planes_match_oracle_on_golden(host, oracle, name, ent):
    data = open(os.path.join(GOLDEN, name), "rb").read()
    h = host.probe(data)
    assert (h.width, h.height) == (ent["width"], ent["height"])
    fr, _, _ = oracle.decode(data)
    assert h.ncomp == fr.ncomp and h.progressive == fr.progressive
    desc, planes = host.decode_planes(data)
    assert desc.ncomp == fr.ncomp
    for c in range(fr.ncomp):
        assert (desc.samp_h[c], desc.samp_v[c]) == (fr.samp_h[c], fr.samp_v[c])
        assert list(desc.qt[c]) == fr.qt[c].tolist()
        assert np.array_equal(planes[c], fr.planes[c]), (name, c)
    hsh = hashlib.sha256()
    for p in planes:
        hsh.update(p.tobytes())
    assert hsh.hexdigest() == ent["planes_sha256"]


@pytest.mark.parametrize("samp,prog,ri,w,h", [
    ("444", False, 0, 97, 61), ("420", False, 0, 130, 70), ("422", False, 7, 200, 40), ("grey", False, 5, 67, 45),
    ("440", False, 0, 64, 80), ("411", False, 3, 133, 31), ("420", True, 0, 161, 99), ("422", True, 0, 96, 96),
    ("444", True, 0, 40, 24), ("grey", True, 0, 75, 50),
])
def test_entropy_decoder_matches_oracle_front_end(host, oracle, synth, samp, prog, ri, w, h):
    """synthetic .jpg (baseline / progressive with DC+AC refinement / restart intervals):
    the product's entropy decoder and GID's (restated in the oracle) emit the same planes,
    which are also the planes the encoder quantised."""
    jpg, enc = synth.encode(pixels_for(synth, 9, w, h, samp), samp, progressive=prog, restart_interval=ri)
    fr, _, _ = oracle.decode(jpg)
    desc, planes = host.decode_planes(jpg)
    assert (desc.width, desc.height, desc.ncomp) == (w, h, fr.ncomp)
    for c in range(fr.ncomp):
        assert np.array_equal(planes[c], fr.planes[c])
        assert np.array_equal(planes[c], enc.planes[c])
        assert list(desc.qt[c]) == fr.qt[c].tolist()


def test_error_classes_follow_gid(host, oracle, synth):
    jpg, _ = synth.encode(pixels_for(synth, 1, 64, 48, "420"), "420")
    # not an image at all / another format GID knows
    with pytest.raises(host.GidHostError) as e:
        host.probe(b"\x00\x01\x02\x03 nothing")
    assert e.value.name == "unknown_image_format"
    with pytest.raises(host.GidHostError) as e:
        host.probe(b"GIF89a" + bytes(64))
    assert e.value.name == "known_but_unsupported_image_format"
    # SOF1 (extended sequential) -> unsupported_image_subformat (gid-decoding_jpg.adb:197)
    bad = bytearray(jpg)
    i = bad.index(b"\xff\xc0")
    bad[i + 1] = 0xC1
    with pytest.raises(host.GidHostError) as e:
        host.probe(bytes(bad))
    assert e.value.name == "unsupported_image_subformat"
    with pytest.raises(Exception):
        oracle.probe(bytes(bad))
    # 12-bit precision -> unsupported (:201); zero width -> error_in_image_data (:209)
    bad = bytearray(jpg)
    bad[i + 4] = 12
    with pytest.raises(host.GidHostError) as e:
        host.probe(bytes(bad))
    assert e.value.name == "unsupported_image_subformat"
    bad = bytearray(jpg)
    bad[i + 7] = bad[i + 8] = 0
    with pytest.raises(host.GidHostError) as e:
        host.probe(bytes(bad))
    assert e.value.name == "error_in_image_data"
    # truncated entropy data still decodes (End_Error is swallowed and padded, :665-670);
    # garbage where a marker should be -> error_in_image_data (:138)
    bad = bytearray(jpg)
    j = bad.index(b"\xff\xda")
    bad[j] = 0x12
    with pytest.raises(host.GidHostError) as e:
        host.decode_planes(bytes(bad))
    assert e.value.name == "error_in_image_data"
    # a restart marker with the wrong index -> error_in_image_data (:1155)
    jr, _ = synth.encode(pixels_for(synth, 1, 64, 48, "420"), "420", restart_interval=2)
    bad = bytearray(jr)
    k = bad.index(b"\xff\xd0")
    bad[k + 1] = 0xD3
    with pytest.raises(host.GidHostError) as e:
        host.decode_planes(bytes(bad))
    assert e.value.name == "error_in_image_data"


# --------------------------------------------------------------------------- restart-interval-parallel decode


@pytest.fixture
def entropy_threads(host):
    def set_threads(n):
        host.set_option("entropy_threads", n)
    yield set_threads
    host.set_option("entropy_threads", 1)


def _outcome(host, data):
    try:
        _, planes = host.decode_planes(data)
        return ("ok", hashlib.sha256(b"".join(p.tobytes() for p in planes)).hexdigest())
    except host.GidHostError as e:
        return ("err", e.name)


@pytest.mark.parametrize("samp,w,h,ri,threads", [
    ("grey", 1024, 512, 4, 4), ("420", 1024, 800, 3, 3), ("444", 640, 520, 16, 8), ("422", 1000, 600, 64, 5),
    ("grey", 520, 8 * 300 + 3, 1, 7), ("420", 16 * 90 + 5, 16 * 30 + 9, 2700, 8),
])
def test_restart_parallel_scan_matches_serial_and_oracle(host, oracle, synth, entropy_threads, samp, w, h, ri, threads):
    """the restart intervals of a baseline scan (Check_Restart,
    gid-decoding_jpg.adb:1133-1177) decoded by several host threads give the planes of the
    serial loop, = GID's front end (oracle), = what the encoder quantised."""
    jpg, enc = synth.encode(pixels_for(synth, 4, w, h, samp), samp, restart_interval=ri)
    fr, _, _ = oracle.decode(jpg)
    entropy_threads(1)
    before = host.counter("parallel_scans")
    _, serial = host.decode_planes(jpg)
    assert host.counter("parallel_scans") == before
    entropy_threads(threads)
    _, par = host.decode_planes(jpg)
    # (the last case has a single restart interval: too few for the parallel path, decoded serially)
    assert host.counter("parallel_scans") == before + (1 if ri < 2700 else 0)
    for c in range(fr.ncomp):
        assert np.array_equal(par[c], serial[c])
        assert np.array_equal(par[c], fr.planes[c])
        assert np.array_equal(par[c], enc.planes[c])


def test_restart_parallel_scan_keeps_the_serial_error_behaviour(host, oracle, synth, entropy_threads):
    """Damaged streams: whatever the serial decoder does with them (same planes GID's front end
    produces, or the same exception class), the threaded decoder does too -- it hands anything
    unusual back to the serial loop."""
    rng = np.random.default_rng(11)
    checked = 0
    for samp, w, h, ri in (("grey", 1024, 512, 4), ("420", 1024, 800, 3), ("444", 640, 520, 16)):
        jpg, _ = synth.encode(pixels_for(synth, 3, w, h, samp), samp, restart_interval=ri)
        sos = jpg.index(b"\xff\xda")
        for t in range(24):
            bad = bytearray(jpg)
            pos = int(rng.integers(sos + 14, len(bad) - 2))
            kind = t % 4
            if kind == 0:
                bad[pos] = int(rng.integers(0, 256))
            elif kind == 1:
                del bad[pos]
            elif kind == 2:
                bad.insert(pos, int(rng.integers(0, 256)))
            else:  # renumber (or, below, drop) the next restart marker
                k = bad.find(b"\xff", pos)
                while k > 0 and not 0xD0 <= bad[k + 1] <= 0xD7:
                    k = bad.find(b"\xff", k + 2)
                if k > 0 and t % 8 == 3:
                    bad[k + 1] = 0xD0 + int(rng.integers(0, 8))
                elif k > 0:
                    del bad[k:k + 2]
            bad = bytes(bad)
            entropy_threads(1)
            serial = _outcome(host, bad)
            entropy_threads(4)
            assert _outcome(host, bad) == serial, (samp, t)
            try:
                fr, _, _ = oracle.decode(bad)
                ref = ("ok", hashlib.sha256(b"".join(np.asarray(p).tobytes() for p in fr.planes[:fr.ncomp])).hexdigest())
            except Exception:
                ref = ("err", None)
            assert serial[0] == ref[0] and (serial[0] == "err" or serial[1] == ref[1]), (samp, t)
            checked += 1
    assert checked == 72


def test_exif_orientation(host, synth):
    jpg, _ = synth.encode(pixels_for(synth, 1, 32, 32, "444"), "444")
    for value, want in ((1, 0), (8, 1), (3, 2), (6, 3), (2, 0)):
        tiff = b"II*\x00\x08\x00\x00\x00" + b"\x01\x00" + b"\x12\x01\x03\x00\x01\x00\x00\x00" + bytes([value, 0, 0, 0]) + bytes(4)
        body = b"Exif\x00\x00" + tiff
        app1 = b"\xff\xe1" + (len(body) + 2).to_bytes(2, "big") + body
        assert host.probe(jpg[:2] + app1 + jpg[2:]).orientation == want


# --------------------------------------------------------------------------- GPU


@pytest.fixture(params=["records", "planes"])
def coef_form(request, host):
    """Both forms in which the entropy decoder hands coefficients to the device."""
    host.set_option("sparse_input", 1 if request.param == "records" else 0)
    yield request.param
    host.set_option("sparse_input", 1)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", CONFIG_SMALL, ids=[c["name"] for c in CONFIG_SMALL])
def test_load_image_contents_matches_gid(host, oracle, synth, gpu, coef_form, cfg):
    """.jpg -> gid::Load_Image_Header + gid::Load_Image_Contents (entropy decode on the host,
    reconstruction on the B200, pixels replayed through Set_X_Y / Put_Pixel) == GID's decode."""
    jpg, _ = synth.encode(pixels_for(synth, 5, cfg["w"], cfg["h"], cfg["samp"]), cfg["samp"],
                          progressive=cfg["prog"], restart_interval=cfg["ri"])
    _, ref, ref_fb = oracle.decode(jpg)
    rgb, fb = host.decode_rgb(jpg)
    assert np.array_equal(rgb, ref)
    assert fb == sorted(fb) and fb[-1] == 100 and min(fb) >= 0


@pytest.mark.gpu
@pytest.mark.parametrize("prog", [False, True], ids=["baseline", "progressive"])
def test_dense_frames_outgrow_the_record_buffer(host, oracle, synth, gpu, prog):
    """Quality 100 on noise: more than 32 non-zero coefficients per block on average, so the
    decoder runs out of record capacity part-way and carries on in the dense plane
    (baseline: replaying the records it had; progressive: at compaction time)."""
    rng = np.random.default_rng(8)
    pix = rng.integers(0, 256, size=(96, 160, 3), dtype=np.uint8)
    jpg, enc = synth.encode(pix, "420", quality=100, progressive=prog)
    assert sum(int(np.count_nonzero(p)) for p in enc.planes) > 32 * sum(p.shape[0] * p.shape[1] for p in enc.planes)
    _, ref, _ = oracle.decode(jpg)
    rgb, _ = host.decode_rgb(jpg)
    assert np.array_equal(rgb, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("name,ent", _golden_files(), ids=[n for n, _ in _golden_files()])
def test_golden_files_through_host_mirror(host, gpu, coef_form, name, ent):
    data = open(os.path.join(GOLDEN, name), "rb").read()
    rgb, _ = host.decode_rgb(data)
    assert hashlib.sha256(rgb.tobytes()).hexdigest() == ent["rgb_sha256"]


@pytest.mark.gpu
def test_primary_color_ranges(host, oracle, synth, gpu):
    """16-bit Primary_Color_Range: values are the 8-bit ones times 257 (:909-938); any other
    modulus raises invalid_primary_color_range."""
    jpg, _ = synth.encode(pixels_for(synth, 3, 90, 50, "420"), "420")
    _, ref, _ = oracle.decode(jpg)
    rgb16, _ = host.decode_rgb(jpg, modulus=65536)
    assert np.array_equal(rgb16, ref)
    with pytest.raises(host.GidHostError) as e:
        host.decode_rgb(jpg, modulus=1024)
    assert e.value.name == "invalid_primary_color_range"


@pytest.mark.gpu
@pytest.mark.parametrize("samp,w,h,ri,threads", [("grey", 1024, 520, 4, 6), ("420", 1030, 810, 3, 5), ("444", 640, 520, 16, 8)])
def test_restart_parallel_scan_through_the_device(host, oracle, synth, gpu, coef_form, entropy_threads, samp, w, h, ri, threads):
    """The threaded scan hands its records over as one run per worker
    (gidcuda_job_records_commit_runs) or fills the dense planes; pixels == GID's decode."""
    jpg, _ = synth.encode(pixels_for(synth, 6, w, h, samp), samp, restart_interval=ri)
    _, ref, _ = oracle.decode(jpg)
    entropy_threads(threads)
    host.set_option("device_entropy", 0)  # this test is about the HOST threads
    before = host.counter("parallel_scans")
    try:
        rgb, fb = host.decode_rgb(jpg)
    finally:
        host.set_option("device_entropy", 1)
    assert host.counter("parallel_scans") == before + 1
    assert np.array_equal(rgb, ref)
    assert fb == sorted(fb) and fb[-1] == 100


@pytest.mark.gpu
def test_file_batch_matches_gid(host, oracle, synth, gpu):
    """gidhost_batch_decode: a mixed set of files on several host threads, three frames in flight per
    thread; every image == GID's decode, a damaged file is reported with GID's exception class and
    does not disturb the others; the contexts are recycled by a second call."""
    specs = [("420", 640, 360, False, 0), ("444", 97, 61, False, 0), ("grey", 300, 200, False, 5), ("422", 320, 240, True, 0),
             ("420", 161, 99, True, 0), ("440", 64, 80, False, 0), ("grey", 1024, 520, False, 4), ("420", 1919, 1079, False, 0)]
    files = []
    for k, (samp, w, h, prog, ri) in enumerate(specs * 2):
        files.append(synth.encode(pixels_for(synth, 20 + k, w, h, samp), samp, progressive=prog, restart_interval=ri)[0])
    bad = bytearray(files[0])
    i = bad.index(b"\xff\xda")
    bad[i] = 0x12                       # garbage where a marker should be -> error_in_image_data
    files.insert(3, bytes(bad))
    files.insert(9, b"GIF89a" + bytes(64))  # another format GID knows
    refs = []
    for f in files:
        try:
            refs.append(oracle.decode(f)[1])
        except Exception:
            refs.append(None)
    with host.HostBatch(devices=(0,), threads=4) as hb:
        for _ in range(2):
            imgs, st, stats = hb.decode(files)
            assert stats.images_failed == 2 and stats.threads == 4
            assert host.ERRORS[st[3]] == "error_in_image_data" and host.ERRORS[st[9]] == "known_but_unsupported_image_format"
            for k, (img, ref) in enumerate(zip(imgs, refs)):
                if ref is None:
                    assert img is None and st[k] != 0
                else:
                    assert st[k] == 0 and np.array_equal(img, ref), k
            assert abs(stats.megapixels - sum(r.shape[0] * r.shape[1] for r in refs if r is not None) / 1e6) < 1e-6
    # fewer files than threads: the spare threads go into the restart-interval-parallel scan
    big = synth.encode(pixels_for(synth, 31, 2048, 1024, "grey"), "grey", restart_interval=64)[0]
    host.set_option("device_entropy", 0)
    try:
        before = host.counter("parallel_scans")
        imgs, st, stats = host.decode_batch([big], threads=8)
    finally:
        host.set_option("device_entropy", 1)
    assert st == [0] and stats.entropy_threads == 8 and host.counter("parallel_scans") == before + 1
    assert np.array_equal(imgs[0], oracle.decode(big)[1])
    # ... or, by default, the scan is entropy-decoded on the device
    before = host.counter("device_scans")
    imgs, st, stats = host.decode_batch([big, big], threads=2)
    assert st == [0, 0] and host.counter("device_scans") == before + 2
    assert np.array_equal(imgs[0], oracle.decode(big)[1]) and np.array_equal(imgs[1], imgs[0])


@pytest.mark.gpu
def test_pixels_land_in_the_callers_pinned_buffer(host, oracle, synth, gpu):
    """gidcuda_job_output: the client's own (pinned) image buffer receives the pixels straight from
    the device; ordinary memory is refused there, and the batch driver copies instead."""
    import ctypes
    from gid_b200 import capi
    from tests.harness import frame_desc
    L = gpu.lib
    jpg, enc = synth.encode(pixels_for(synth, 12, 333, 211, "420"), "420")
    _, ref, _ = oracle.decode(jpg)
    desc = frame_desc(enc, flags=capi.FLAG_PACKED_ROWS)
    buf, addr = capi.pinned_empty(ref.nbytes)
    try:
        assert L.gidcuda_host_is_pinned(ctypes.c_void_p(addr)) == 1
        plain = np.zeros(ref.nbytes, dtype=np.uint8)
        assert L.gidcuda_host_is_pinned(ctypes.c_void_p(plain.ctypes.data)) == 0
        job = ctypes.c_void_p()
        assert L.gidcuda_job_begin(gpu.ctx, ctypes.byref(desc), ctypes.byref(job)) == 0
        assert L.gidcuda_job_output(job, ctypes.c_void_p(plain.ctypes.data), plain.nbytes) == -1   # not pinned
        assert L.gidcuda_job_output(job, ctypes.c_void_p(addr), ref.nbytes - 1) == -1                # too small
        assert L.gidcuda_job_output(job, ctypes.c_void_p(addr), ref.nbytes) == 0
        for c in range(enc.ncomp):
            p = L.gidcuda_job_plane(job, c, None, None)
            ctypes.memmove(p, enc.planes[c].ctypes.data, enc.planes[c].nbytes)
        assert L.gidcuda_job_submit(job) == 0
        assert L.gidcuda_job_output(job, None, 0) == -6                                                # state
        pix, st = ctypes.c_void_p(), ctypes.c_size_t()
        assert L.gidcuda_job_wait(job, ctypes.byref(pix), ctypes.byref(st)) == 0
        assert pix.value == addr and st.value == 3 * 333
        assert np.array_equal(buf.reshape(ref.shape), ref)
        assert L.gidcuda_job_release(job) == 0
        # the batch driver: pinned and ordinary client buffers side by side
        buf[:] = 0
        imgs, st, _ = host.HostBatch(devices=(0,), threads=2).decode([jpg, jpg], [buf.reshape(ref.shape), plain.reshape(ref.shape)])
        assert st == [0, 0] and np.array_equal(imgs[0], ref) and np.array_equal(imgs[1], ref)
    finally:
        capi.pinned_free(addr)


# --------------------------------------------------------------------------- entropy decode on the device


@pytest.mark.gpu
@pytest.mark.parametrize("samp,w,h,ri", [
    ("grey", 1024, 520, 4), ("420", 1030, 810, 3), ("444", 640, 520, 16), ("422", 1000, 600, 64), ("440", 333, 700, 2),
    ("411", 1024, 256, 1), ("grey", 520, 8 * 300 + 3, 1), ("420", 16 * 90 + 5, 16 * 30 + 9, 11),
])
def test_device_entropy_decode_matches_gid(host, oracle, synth, gpu, samp, w, h, ri):
    """SURVEY.md 8f rank 1, second half: the restart intervals of a baseline scan Huffman-decoded
    by one device thread each (gidcuda_job_scan) -- only the compressed bytes cross the link --
    then reconstructed: pixels == GID's decode of the file."""
    jpg, _ = synth.encode(pixels_for(synth, 8, w, h, samp), samp, restart_interval=ri)
    _, ref, _ = oracle.decode(jpg)
    scans, back = host.counter("device_scans"), host.counter("device_scan_fallbacks")
    rgb, fb = host.decode_rgb(jpg)
    assert host.counter("device_scans") == scans + 1 and host.counter("device_scan_fallbacks") == back
    assert np.array_equal(rgb, ref)
    assert fb == sorted(fb) and fb[-1] == 100


@pytest.mark.gpu
def test_device_entropy_decode_optimised_tables_and_dense_data(host, oracle, synth, gpu):
    """Per-image optimised Huffman tables (codes up to 16 bits: the canonical search behind the
    10-bit table) and quality-100 noise (long runs of large values)."""
    rng = np.random.default_rng(5)
    pix = rng.integers(0, 256, size=(256, 512, 3), dtype=np.uint8)
    for quality, optimize in ((100, True), (100, False), (30, True)):
        jpg, _ = synth.encode(pix, "420", quality=quality, restart_interval=1, optimize=optimize)
        scans = host.counter("device_scans")
        rgb, _ = host.decode_rgb(jpg)
        assert host.counter("device_scans") == scans + 1
        assert np.array_equal(rgb, oracle.decode(jpg)[1])


@pytest.mark.gpu
def test_device_entropy_decode_leaves_damaged_streams_to_the_host(host, oracle, synth, gpu):
    """Whatever the host decoder does with a damaged stream (GID's behaviour, checked against the
    oracle on the CPU side), the device path does too: it declines and the host decodes."""
    rng = np.random.default_rng(21)
    declined = 0
    for samp, w, h, ri in (("grey", 1024, 512, 4), ("420", 1024, 800, 3)):
        jpg, _ = synth.encode(pixels_for(synth, 3, w, h, samp), samp, restart_interval=ri)
        sos = jpg.index(b"\xff\xda")
        for t in range(12):
            bad = bytearray(jpg)
            pos = int(rng.integers(sos + 14, len(bad) - 2))
            if t % 3 == 0:
                bad[pos] = int(rng.integers(0, 256))
            elif t % 3 == 1:
                del bad[pos]
            else:
                bad.insert(pos, int(rng.integers(1, 255)))
            bad = bytes(bad)

            def outcome():
                try:
                    return ("ok", hashlib.sha256(host.decode_rgb(bad)[0].tobytes()).hexdigest())
                except host.GidHostError as e:
                    return ("err", e.name)
            host.set_option("device_entropy", 0)
            try:
                want = outcome()
            finally:
                host.set_option("device_entropy", 1)
            back = host.counter("device_scan_fallbacks")
            assert outcome() == want, (samp, t)
            declined += host.counter("device_scan_fallbacks") - back
            try:
                ref = ("ok", hashlib.sha256(oracle.decode(bad)[1].tobytes()).hexdigest())
            except Exception:
                ref = ("err", None)
            assert want[0] == ref[0] and (want[0] == "err" or want[1] == ref[1]), (samp, t)
    assert declined > 0


@pytest.mark.gpu
@pytest.mark.parametrize("samp,w,h,quality,optimize", [
    ("420", 1920, 1080, 75, False), ("444", 1024, 768, 75, False), ("grey", 2048, 1024, 75, False),
    ("422", 1280, 720, 90, True), ("440", 640, 800, 50, False), ("411", 1024, 512, 75, True),
    ("420", 2999, 2001, 95, False), ("444", 333, 211, 100, True),
])
def test_device_entropy_decode_without_restart_markers(host, oracle, synth, gpu, samp, w, h, quality, optimize):
    """A baseline scan without RSTn is one bit stream: the device cuts it into 1024-bit
    subsequences and lets the Huffman code re-synchronise (entropy_kernel.cu, second half).
    The verified chain is the serial decode: pixels == GID's decode, nothing left to the host."""
    jpg, _ = synth.encode(pixels_for(synth, 14, w, h, samp), samp, quality=quality, optimize=optimize)
    _, ref, _ = oracle.decode(jpg)
    scans, back = host.counter("device_scans"), host.counter("device_scan_fallbacks")
    rgb, fb = host.decode_rgb(jpg)
    assert np.array_equal(rgb, ref)
    assert host.counter("device_scans") == scans + 1 and host.counter("device_scan_fallbacks") == back
    assert fb == sorted(fb) and fb[-1] == 100


@pytest.mark.gpu
def test_device_entropy_decode_noise_and_damage_without_restart_markers(host, oracle, synth, gpu):
    """Quality-100 noise (16-bit codes, 63-coefficient blocks), then damaged streams: the outcome is
    the host decoder's (pixels or exception class), which the CPU suite pins to the oracle."""
    rng = np.random.default_rng(9)
    pix = rng.integers(0, 256, size=(512, 768, 3), dtype=np.uint8)
    for samp in ("420", "444"):
        jpg, _ = synth.encode(pix, samp, quality=100)
        scans = host.counter("device_scans")
        rgb, _ = host.decode_rgb(jpg)
        assert host.counter("device_scans") == scans + 1
        assert np.array_equal(rgb, oracle.decode(jpg)[1])
    jpg, _ = synth.encode(pixels_for(synth, 3, 1024, 800, "420"), "420")
    sos = jpg.index(b"\xff\xda")
    for t in range(12):
        bad = bytearray(jpg)
        pos = int(rng.integers(sos + 14, len(bad) - 2))
        if t % 3 == 0:
            bad[pos] = int(rng.integers(0, 255))
        elif t % 3 == 1:
            del bad[pos]
        else:
            del bad[pos:pos + 1000]
        bad = bytes(bad)

        def outcome():
            try:
                return ("ok", hashlib.sha256(host.decode_rgb(bad)[0].tobytes()).hexdigest())
            except host.GidHostError as e:
                return ("err", e.name)
        host.set_option("device_entropy", 0)
        try:
            want = outcome()
        finally:
            host.set_option("device_entropy", 1)
        assert outcome() == want, t


@pytest.mark.gpu
def test_configs_full_size_from_files(host, oracle, synth, gpu):
    """The five BASELINE.json configs at their full frame size as .jpg files through the file batch
    driver: device Huffman decoding (self-synchronising for configs 1-3, restart-interval-parallel for
    config 4, first scans + DC refinement of the progressive config 5); every image == GID's decode."""
    from tests.test_gpu_parity import FULL_SIZE
    files, refs = [], []
    for name, w, h, samp, prog, ri in FULL_SIZE:
        jpg, _ = synth.encode(pixels_for(synth, 78, w, h, samp), samp, progressive=prog, restart_interval=ri)
        files.append(jpg)
        refs.append(oracle.decode(jpg)[1])
    scans, back = host.counter("device_scans"), host.counter("device_scan_fallbacks")
    imgs, st, stats = host.decode_batch(files, devices=(0,), threads=3)
    assert st == [0] * 5 and stats.images_failed == 0
    assert host.counter("device_scans") == scans + 5 and host.counter("device_scan_fallbacks") == back
    for k, (img, ref) in enumerate(zip(imgs, refs)):
        assert np.array_equal(img, ref), FULL_SIZE[k][0]


def test_progressive_scans_through_the_wide_bit_buffer(host, oracle, synth):
    """Progressive scans are read through the 64-bit buffer, which afterwards puts the stream where
    the reference's byte-wise buffer would have stopped: on valid files the planes are GID's
    (oracle); on damaged files the outcome -- planes or exception class -- is that of the byte-wise
    reader, whatever the segment loop meets next."""
    rng = np.random.default_rng(2)
    same = 0
    for samp, w, h in (("420", 161, 99), ("422", 256, 192), ("444", 120, 88), ("grey", 200, 150)):
        jpg, enc = synth.encode(pixels_for(synth, 3, w, h, samp), samp, progressive=True)
        _, planes = host.decode_planes(jpg)
        fr, _, _ = oracle.decode(jpg)
        for c in range(fr.ncomp):
            assert np.array_equal(planes[c], fr.planes[c]) and np.array_equal(planes[c], enc.planes[c])
        sos = jpg.index(b"\xff\xda")
        for t in range(60):
            bad = bytearray(jpg)
            pos = int(rng.integers(sos + 10, len(bad) - 2))
            kind = t % 5
            if kind == 0:
                bad[pos] = int(rng.integers(0, 256))
            elif kind == 1:
                del bad[pos]
            elif kind == 2:
                bad.insert(pos, int(rng.integers(0, 256)))
            elif kind == 3:
                del bad[pos:pos + int(rng.integers(2, 40))]
            else:
                for _ in range(3):
                    bad.insert(pos, int(rng.integers(0, 255)))
            bad = bytes(bad)
            host.set_option("wide_bit_buffer", 0)
            try:
                want = _outcome(host, bad)
            finally:
                host.set_option("wide_bit_buffer", 1)
            assert _outcome(host, bad) == want, (samp, t)
            same += 1
    assert same == 240


def _scan_ends(jpg):
    """Positions of the markers that end the entropy-coded segments of a file."""
    out, j = [], jpg.index(b"\xff\xda") + 2
    while j < len(jpg) - 1:
        if jpg[j] == 0xFF and jpg[j + 1] != 0 and not 0xD0 <= jpg[j + 1] <= 0xD7:
            out.append(j)
        j += 1
    return out


def _damage_near(jpg, end, rng, t):
    bad = bytearray(jpg)
    k = int(rng.integers(1, 4))
    bad[end - k] = int(rng.integers(0, 255))
    if t % 3 == 0 and k > 1:
        bad[end - k + 1] = int(rng.integers(0, 255))
    return bytes(bad)


@pytest.mark.parametrize("wide", [1, 0], ids=["wide_buffer", "bytewise_buffer"])
def test_damage_at_the_end_of_a_scan_is_decoded_as_gid_does(host, oracle, synth, wide):
    """The last bytes of a scan damaged: the decoder then reads INTO the marker that ends the scan, meets
    refinement symbols of sizes other than 1, runs that leave their band or the image, end-of-band runs longer
    than the scan.  GID decodes much of that silently, in its own way (Show_Bits :609-677 puts the marker's two
    bytes into the bit buffer and carries on behind them; Progressive_AC_Scan :1515-1689 takes any size, walks
    runs past the band, adds correction bits to the magnitude; Append :1402-1415 holds 2000 points): the host
    mirror must give the oracle's planes, or raise where it raises.  (A frame whose coefficients leave 16 bits is
    the documented exception: the oracle flags it, the product path reports it as unsupported.)"""
    rng = np.random.default_rng(11)
    host.set_option("wide_bit_buffer", wide)
    ok = err = 0
    try:
        for prog in (False, True):
            for samp, w, h in (("420", 96, 80), ("444", 64, 48), ("grey", 72, 40)):
                jpg, _ = synth.encode(pixels_for(synth, 5, w, h, samp), samp, progressive=prog)
                for end in _scan_ends(jpg)[-6:]:
                    for t in range(40):
                        bad = _damage_near(jpg, end, rng, t)
                        try:
                            fr, _, _ = oracle.decode(bad)
                        except Exception:
                            fr = None
                        try:
                            _, planes = host.decode_planes(bad)
                        except host.GidHostError:
                            planes = None
                        if fr is not None and fr.coef_overflow:
                            continue
                        assert (fr is None) == (planes is None), (prog, samp, end, t)
                        if fr is None:
                            err += 1
                            continue
                        ok += 1
                        for c in range(fr.ncomp):
                            assert np.array_equal(planes[c], fr.planes[c]), (prog, samp, end, t, c)
    finally:
        host.set_option("wide_bit_buffer", 1)
    assert ok > 300 and err > 300


def test_dc_scans_of_single_components_are_decoded_as_t81_says(host, oracle, synth):
    """A KNOWN DIFFERENCE from GID: in a DC scan that carries some but not all components of
    a colour frame, Progressive_DC_Scan numbers the layers of its image array by the scan's components
    (gid-decoding_jpg.adb:1286, :1352), so the DC terms of a Cb-only scan land in the Y layer.  The mirror decodes
    such scans as T.81 says: its chroma planes are the encoder's.  (Should GID ever change, the last assertion
    says so.)"""
    Y, CB, CR = [0], [1], [2]
    sc = [(Y, 0, 0, 0, 1), (CB, 0, 0, 0, 1), (CR, 0, 0, 0, 1), (Y, 1, 63, 0, 0), (CB, 1, 63, 0, 0), (CR, 1, 63, 0, 0),
          (Y, 0, 0, 1, 0), (CB, 0, 0, 1, 0), (CR, 0, 0, 1, 0)]
    jpg, enc = synth.encode(pixels_for(synth, 14, 200, 120, "420"), "420", quality=85, progressive=True, scans=sc)
    _, planes = host.decode_planes(jpg)
    for c in (1, 2):
        assert np.array_equal(planes[c], enc.planes[c])
    fr, _, _ = oracle.decode(jpg)
    assert not np.array_equal(fr.planes[1], enc.planes[1])


def test_damage_around_restart_markers_is_decoded_as_gid_does(host, oracle, synth):
    """Baseline files with restart intervals, damaged in and around their RSTn markers (a byte changed, dropped,
    inserted, a stretch cut out, an FF planted): Check_Restart (:1133-1177) consumes the 16 bits it compares and
    leaves the rest of the bit buffer alone -- 16 bits that merely LOOK like the marker (a stuffed FF followed by
    Dn) are followed by data the buffer has already loaded.  Outcome and planes must be the oracle's."""
    rng = np.random.default_rng(5)
    ok = err = 0
    for samp, w, h, ri in (("420", 161, 99, 2), ("422", 256, 192, 5), ("444", 120, 88, 1), ("grey", 200, 150, 7)):
        jpg, _ = synth.encode(pixels_for(synth, 3, w, h, samp), samp, restart_interval=ri)
        sos = jpg.index(b"\xff\xda")
        marks = [j for j in range(sos, len(jpg) - 1) if jpg[j] == 0xFF and jpg[j + 1] != 0]
        for t in range(120):
            bad = bytearray(jpg)
            if t % 2:
                pos = int(rng.integers(sos + 10, len(bad) - 2))
            else:
                e = marks[int(rng.integers(0, len(marks)))]
                pos = max(sos + 10, min(len(bad) - 3, e + int(rng.integers(-3, 3))))
            kind = t % 5
            if kind == 0:
                bad[pos] = int(rng.integers(0, 256))
            elif kind == 1:
                del bad[pos]
            elif kind == 2:
                bad.insert(pos, int(rng.integers(0, 256)))
            elif kind == 3:
                del bad[pos:pos + int(rng.integers(2, 40))]
            else:
                bad[pos] = 0xFF
            bad = bytes(bad)
            try:
                fr, _, _ = oracle.decode(bad)
            except Exception:
                fr = None
            try:
                _, planes = host.decode_planes(bad)
            except host.GidHostError:
                planes = None
            if fr is not None and fr.coef_overflow:
                continue
            assert (fr is None) == (planes is None), (samp, t, kind, pos)
            if fr is None:
                err += 1
                continue
            ok += 1
            for c in range(fr.ncomp):
                assert np.array_equal(planes[c], fr.planes[c]), (samp, t, kind, pos, c)
    assert ok > 100 and err > 200


@pytest.mark.gpu
def test_flat_and_periodic_images_stay_exact(host, oracle, synth, gpu):
    """Flat regions make the bit stream periodic, and a periodic stream need not re-synchronise a
    decoder that started in the wrong phase: the device's chain of states is then left unsettled, the
    host walks those stretches serially (gidcuda_job_wait, "chain_repair") and the output pass verifies
    the repaired chain.  Whichever way an image goes, the pixels are GID's."""
    h, w = 540, 960
    base = synth.image(3, w, h)
    flat = np.full((h, w, 3), 128, np.uint8)
    bars = base.copy()
    bars[:135] = 0
    bars[-135:] = 0
    stripes = np.zeros((h, w, 3), np.uint8)
    stripes[:, ::16] = 255
    sky = base.copy()
    sky[:200] = 255
    cases = [("flat", flat, "420"), ("flat", flat[:, :, 0].copy(), "grey"), ("flat", flat, "444"),
             ("black", np.zeros((h, w, 3), np.uint8), "420"), ("bars", bars, "420"), ("bars", bars, "444"),
             ("stripes", stripes, "420"), ("sky", sky, "422")]
    from gid_b200 import capi
    scans, back = host.counter("device_scans"), host.counter("device_scan_fallbacks")
    repairs = capi.counter("chain_repairs")
    for name, pix, samp in cases:
        jpg, _ = synth.encode(pix, samp)
        rgb, _ = host.decode_rgb(jpg)
        assert np.array_equal(rgb, oracle.decode(jpg)[1]), (name, samp)
    assert host.counter("device_scans") == scans + len(cases)     # every scan went to the device ...
    assert host.counter("device_scan_fallbacks") == back          # ... and none came back to the host decoder:
    assert capi.counter("chain_repairs") > repairs                # the host repaired the unsettled chains
    assert capi.counter("chain_repair_subsequences") > 0
    # without the repair the same scans are handed back whole (and still come out exact)
    capi.set_option("chain_repair", 0)
    try:
        for name, pix, samp in cases[:4]:
            jpg, _ = synth.encode(pix, samp)
            assert np.array_equal(host.decode_rgb(jpg)[0], oracle.decode(jpg)[1]), (name, samp)
    finally:
        capi.set_option("chain_repair", 1)
    assert host.counter("device_scan_fallbacks") > back


# --------------------------------------------------------------------------- files from another encoder (libjpeg via Pillow)


def _pillow_files(synth):
    """JPEG files written by libjpeg(-turbo) through Pillow: other Huffman tables (standard and
    per-image optimised), JFIF / APPn segments, its progressive scan script, restart markers."""
    Image = pytest.importorskip("PIL.Image")
    import io
    from PIL import ImageFile
    ImageFile.MAXBLOCK = max(ImageFile.MAXBLOCK, 4 << 20)  # (libjpeg "suspension" with optimised / progressive output)
    photo = synth.image(11, 640, 424)
    rng = np.random.default_rng(3)
    noisy = np.clip(photo.astype(np.int16) + rng.integers(-40, 41, size=photo.shape), 0, 255).astype(np.uint8)
    out = []
    for name, pix, kw in (
            ("q85_420", photo, dict(quality=85)),
            ("q85_420_optimised", photo, dict(quality=85, optimize=True)),
            ("q92_444", noisy, dict(quality=92, subsampling=0)),
            ("q75_422", photo, dict(quality=75, subsampling=1)),
            ("q60_grey", photo[:, :, 1], dict(quality=60)),
            ("q80_420_restart", noisy, dict(quality=80, restart_marker_blocks=3)),
            ("q80_444_restart_rows", photo, dict(quality=80, subsampling=0, restart_marker_rows=1)),
            ("q100_420", noisy, dict(quality=100)),
            ("q30_420_progressive", photo, dict(quality=30, progressive=True)),
            ("q90_444_progressive_optimised", noisy, dict(quality=90, subsampling=0, progressive=True, optimize=True))):
        b = io.BytesIO()
        Image.fromarray(pix).save(b, "JPEG", **kw)
        out.append((name, b.getvalue()))
    return out


def test_libjpeg_files_entropy_decoder_matches_oracle(host, oracle, synth):
    for name, jpg in _pillow_files(synth):
        fr, _, _ = oracle.decode(jpg)
        desc, planes = host.decode_planes(jpg)
        assert (desc.width, desc.height, desc.ncomp) == (fr.width, fr.height, fr.ncomp), name
        for c in range(fr.ncomp):
            assert np.array_equal(planes[c], fr.planes[c]), (name, c)
            assert list(desc.qt[c]) == fr.qt[c].tolist(), (name, c)


@pytest.mark.gpu
def test_libjpeg_files_through_the_device(host, oracle, synth, gpu):
    """The same files end to end: baseline ones are Huffman-decoded on the device (with restart
    markers only when there are at least 64 intervals, else by the self-synchronising decoder's
    sibling on the host), progressive ones on the host; pixels == GID's decode of the file."""
    scans, back = host.counter("device_scans"), host.counter("device_scan_fallbacks")
    files = _pillow_files(synth)
    for name, jpg in files:
        rgb, _ = host.decode_rgb(jpg)
        assert np.array_equal(rgb, oracle.decode(jpg)[1]), name
    assert host.counter("device_scans") - scans >= 6 and host.counter("device_scan_fallbacks") == back
    imgs, st, _ = host.decode_batch([f for _, f in files], threads=3)
    assert st == [0] * len(files)
    for (name, jpg), img in zip(files, imgs):
        assert np.array_equal(img, oracle.decode(jpg)[1]), name


# --------------------------------------------------------------------------- the edge of the coefficient range


def _wide_planes(synth, rng, w, h, samp):
    from tests.harness import SAMPLING
    bx, by = synth.grid(w, h, samp)
    planes = []
    for c in range(len(SAMPLING[samp][0])):
        p = np.zeros((by[c], bx[c], 64), dtype=np.int64)
        m = rng.random(p.shape) < 0.05
        p[m] = rng.integers(-40, 41, size=int(m.sum()))
        big = rng.random(p.shape) < 0.002
        p[big] = rng.choice(np.array([-20000, -5000, 4500, 9000, 16384, -16384]), size=int(big.sum()))
        p[..., 0] = rng.integers(-3000, 3001, size=p.shape[:2])
        planes.append(p.astype(np.int16))
    return planes


@pytest.mark.gpu
@pytest.mark.parametrize("samp,ri,prog", [("420", 0, False), ("444", 3, False), ("grey", 0, False), ("422", 0, True)])
def test_streams_with_wide_coefficients(host, oracle, synth, gpu, coef_form, samp, ri, prog):
    """Coefficients far beyond any JPEG of 8-bit samples (categories 13 .. 15, which Decode_8x8_Block accepts,
    gid-decoding_jpg.adb:727-733): the device Huffman decoders hand the scan to the host decoder, which marks
    the frame (gidcuda_job_wide_coefficients), and the exact kernel instances reproduce the reference's
    32-bit wrap-around: pixels == GID's decode of the file."""
    rng = np.random.default_rng(31)
    w, h = 264, 136
    jpg, _ = synth.encode_planes(w, h, _wide_planes(synth, rng, w, h, samp), samp, progressive=prog,
                                 restart_interval=ri, optimize=True)
    _, ref, _ = oracle.decode(jpg)
    rgb, _ = host.decode_rgb(jpg)
    assert np.array_equal(rgb, ref)


def _two_block_stream_with_runaway_predictor() -> bytes:
    """16 x 8 grey baseline file, two blocks, each a DC difference of +32767 (category 15) and EOB: the
    predictor reaches 65 534.  DC table: the one code '0' -> category 15; AC table: '0' -> EOB."""
    def seg(marker, body):
        return bytes([0xFF, marker]) + (len(body) + 2).to_bytes(2, "big") + body
    dqt = seg(0xDB, bytes([0]) + bytes([1] * 64))
    sof = seg(0xC0, bytes([8]) + (8).to_bytes(2, "big") + (16).to_bytes(2, "big") + bytes([1, 1, 0x11, 0]))
    dht_dc = seg(0xC4, bytes([0x00]) + bytes([1] + [0] * 15) + bytes([15]))
    dht_ac = seg(0xC4, bytes([0x10]) + bytes([1] + [0] * 15) + bytes([0x00]))
    sos = seg(0xDA, bytes([1, 1, 0x00, 0, 63, 0]))
    bits = ("0" + "1" * 15 + "0") * 2
    bits += "1" * (-len(bits) % 8)
    data = bytearray()
    for i in range(0, len(bits), 8):
        b = int(bits[i:i + 8], 2)
        data.append(b)
        if b == 0xFF:
            data.append(0)
    return b"\xff\xd8" + dqt + sof + dht_dc + dht_ac + sos + bytes(data) + b"\xff\xd9"


@pytest.mark.gpu
def test_dc_predictor_beyond_16_bits_is_reported(host, oracle, gpu):
    """The reference keeps a 32-bit predictor (dc_predictor * qt_local (0), :766-771) and paints this file
    white.  The int16 planes of the boundary cannot carry 65 534: the host mirror says so
    (unsupported_image_subformat) instead of painting a wrapped value; the device decoders leave the scan to
    it.  Documented limit of the plane format."""
    jpg = _two_block_stream_with_runaway_predictor()
    _, ref, _ = oracle.decode(jpg)
    assert ref.shape == (8, 16, 3) and (ref[:, :8] == 255).all() and (ref[:, 8:] == 255).all()
    with pytest.raises(host.GidHostError) as e:
        host.decode_rgb(jpg)
    assert e.value.name == "unsupported_image_subformat" and "16 bits" in str(e.value)


# --------------------------------------------------------------------------- progressive scans on the device


@pytest.fixture
def device_progressive(host):
    """Every progressive file takes the device path, however small."""
    host.set_option("device_progressive_min_bytes", 0)
    yield
    host.set_option("device_progressive_min_bytes", 96 * 1024)


def _rgb_outcome(host, data):
    try:
        rgb, _ = host.decode_rgb(data)
        return ("ok", hashlib.sha256(rgb.tobytes()).hexdigest())
    except host.GidHostError as e:
        return ("err", e.name)


@pytest.mark.gpu
@pytest.mark.parametrize("name,ent", [(n, e) for n, e in _golden_files() if e["progressive"] or "libjpeg_script" in n],
                         ids=[n for n, e in _golden_files() if e["progressive"] or "libjpeg_script" in n])
def test_golden_progressive_files_on_the_device(host, oracle, gpu, device_progressive, name, ent):
    """the first scans (DC first, AC first) and the DC refinement scans of a progressive file are
    decoded by the device's self-synchronising decoder into planes that accumulate on the device; AC
    refinement scans are decoded on the host from the non-zero history the device hands back and applied there.
    No coefficient crosses the link; pixels == the committed hashes == GID's decode of the file."""
    jpg = open(os.path.join(GOLDEN, ent["jpg"]), "rb").read()
    scans, back = host.counter("device_scans"), host.counter("device_scan_fallbacks")
    rgb, fb = host.decode_rgb(jpg)
    assert host.counter("device_scans") == scans + 1 and host.counter("device_scan_fallbacks") == back, name
    assert hashlib.sha256(rgb.tobytes()).hexdigest() == ent["rgb_sha256"], name
    assert np.array_equal(rgb, oracle.decode(jpg)[1])
    assert fb == sorted(fb) and fb[-1] == 100


@pytest.mark.gpu
@pytest.mark.parametrize("samp,w,h,quality", [
    ("420", 640, 480, 75), ("422", 512, 512, 75), ("444", 333, 217, 90), ("grey", 520, 300, 60), ("440", 200, 310, 75),
    ("411", 1024, 256, 85), ("420", 1919, 1079, 75), ("422", 2048, 2048, 75), ("444", 17, 9, 75), ("420", 5, 5, 75),
])
def test_progressive_files_on_the_device_match_gid(host, oracle, synth, gpu, device_progressive, samp, w, h, quality):
    """The default scan script (DC with Al = 1, AC 1-5, DC refinement, AC 6-63 with Al = 1, AC refinement) on
    every sampling layout, partial MCUs (a single-component scan walks the component's own block grid,
    :1952-1971), the full cfg5 frame."""
    jpg, _ = synth.encode(pixels_for(synth, 13, w, h, samp), samp, quality=quality, progressive=True)
    scans, back = host.counter("device_scans"), host.counter("device_scan_fallbacks")
    rgb, _ = host.decode_rgb(jpg)
    assert host.counter("device_scans") == scans + 1 and host.counter("device_scan_fallbacks") == back
    assert np.array_equal(rgb, oracle.decode(jpg)[1])


@pytest.mark.gpu
def test_progressive_scan_scripts_on_the_device(host, oracle, synth, gpu, device_progressive):
    """Other scripts: no successive approximation at all; narrow bands; two
    refinement passes of the same band (libjpeg's default: Al = 2, then 1, then 0); a first scan of another
    band AFTER a refinement of the same component (the host fetches the non-zero history again)."""
    pix = pixels_for(synth, 14, 400, 304, "420")
    Y, CB, CR = [0], [1], [2]
    scripts = {
        "no_approximation": [([0, 1, 2], 0, 0, 0, 0)] + [(c, 1, 63, 0, 0) for c in (Y, CB, CR)],
        "libjpeg_default": [([0, 1, 2], 0, 0, 0, 1), (Y, 1, 5, 0, 2), (CR, 1, 63, 0, 1), (CB, 1, 63, 0, 1), (Y, 6, 63, 0, 2),
                            (Y, 1, 63, 2, 1), ([0, 1, 2], 0, 0, 1, 0), (CR, 1, 63, 1, 0), (CB, 1, 63, 1, 0), (Y, 1, 63, 1, 0)],
        "first_after_refinement": [([0, 1, 2], 0, 0, 0, 0), (Y, 1, 9, 0, 1), (Y, 1, 9, 1, 0), (Y, 10, 63, 0, 1), (Y, 10, 63, 1, 0),
                                   (CB, 1, 63, 0, 0), (CR, 1, 63, 0, 0)],
        "narrow_bands": [([0, 1, 2], 0, 0, 0, 0)] + [(c, a, b, 0, 0) for c in (Y, CB, CR) for a, b in ((1, 1), (2, 2), (3, 20), (21, 63))],
    }
    for name, sc in scripts.items():
        jpg, _ = synth.encode(pix, "420", quality=85, progressive=True, scans=sc)
        scans, back = host.counter("device_scans"), host.counter("device_scan_fallbacks")
        rgb, _ = host.decode_rgb(jpg)
        assert host.counter("device_scans") == scans + 1 and host.counter("device_scan_fallbacks") == back, name
        assert np.array_equal(rgb, oracle.decode(jpg)[1]), name


@pytest.mark.gpu
def test_damaged_progressive_files_are_handed_back(host, oracle, synth, gpu, device_progressive):
    """Anything the device decoders do not decode silently (a code in no table, a coefficient outside its
    band, a scan that does not fill its blocks, a chain of states that does not settle, a truncated file)
    hands the whole frame to the host decoder: the outcome -- pixels or exception class -- is the one the
    host decoder alone gives."""
    rng = np.random.default_rng(8)
    jpg, _ = synth.encode(pixels_for(synth, 15, 320, 240, "420"), "420", progressive=True)
    sos = jpg.index(b"\xff\xda")
    same = handed_back = 0
    for t in range(40):
        bad = bytearray(jpg)
        pos = int(rng.integers(sos + 10, len(bad) - 2))
        kind = t % 4
        if kind == 0:
            bad[pos] = int(rng.integers(0, 255))
        elif kind == 1:
            del bad[pos]
        elif kind == 2:
            del bad[pos:pos + int(rng.integers(2, 40))]
        else:
            bad = bad[:pos]
        bad = bytes(bad)
        host.set_option("device_progressive_min_bytes", -1)
        want = _rgb_outcome(host, bad)
        host.set_option("device_progressive_min_bytes", 0)
        back = host.counter("device_scan_fallbacks")
        got = _rgb_outcome(host, bad)
        handed_back += host.counter("device_scan_fallbacks") - back
        assert got == want, (t, kind, got, want)
        same += 1
    assert same == 40 and handed_back >= 10


@pytest.mark.gpu
def test_damage_at_the_end_of_a_scan_on_the_device_path(host, oracle, synth, gpu, device_progressive):
    """... and with the scans of the progressive frame on the device: what is unusual is handed back, the outcome is
    the host decoder's (which the CPU suite holds against the oracle)."""
    rng = np.random.default_rng(12)
    jpg, _ = synth.encode(pixels_for(synth, 15, 320, 240, "420"), "420", progressive=True)
    same = 0
    for end in _scan_ends(jpg)[-6:]:
        for t in range(20):
            bad = _damage_near(jpg, end, rng, t)
            host.set_option("device_progressive_min_bytes", -1)
            want = _rgb_outcome(host, bad)
            host.set_option("device_progressive_min_bytes", 0)
            got = _rgb_outcome(host, bad)
            assert got == want, (end, t, got, want)
            same += 1
    assert same == 120


@pytest.mark.gpu
def test_progressive_files_in_the_file_batch_driver(host, oracle, synth, gpu):
    """cfg5 through gidhost_batch_decode with the default threshold: the 2048 x 2048 files take the device
    path, the small one the host decoder; pixels == GID's."""
    big, _ = synth.encode(pixels_for(synth, 16, 2048, 2048, "422"), "422", progressive=True)
    small, _ = synth.encode(pixels_for(synth, 17, 96, 64, "422"), "422", progressive=True)
    files = [big, small, big, big]
    scans = host.counter("device_scans")
    imgs, st, _ = host.decode_batch(files, threads=3)
    assert st == [0] * 4 and host.counter("device_scans") == scans + 3
    ref_big, ref_small = oracle.decode(big)[1], oracle.decode(small)[1]
    for f, img in zip(files, imgs):
        assert np.array_equal(img, ref_big if f is big else ref_small)


def test_refinement_from_the_nonzero_history_selftest(host, oracle, synth):
    """The host half of the device path for progressive files, checked WITHOUT a device: with the option
    "selftest_refine" the host decoder decodes every AC refinement scan a first time from the non-zero history
    alone (which positions read a correction bit of 1, which became +-(1 << Al): what it would send to
    pscan_apply_kernel), applies that to a copy of the plane as the kernel does, and compares the copy -- and the
    stream position the scan leaves behind -- with what the decoder proper makes of the scan.  Both forms of the
    history decoder (bit arithmetic with BMI2, portable loops), valid and damaged files."""
    files = []
    for name, ent in _golden_files():
        if ent["progressive"] or "libjpeg_script" in name:
            files.append((name, open(os.path.join(GOLDEN, ent["jpg"]), "rb").read()))
    for samp, w, h in (("420", 333, 217), ("422", 256, 192), ("444", 120, 88), ("grey", 200, 150), ("411", 160, 40)):
        files.append((f"synth_{samp}", synth.encode(pixels_for(synth, 5, w, h, samp), samp, progressive=True, quality=85)[0]))
    Y, CB, CR = [0], [1], [2]
    libjpeg = [([0, 1, 2], 0, 0, 0, 1), (Y, 1, 5, 0, 2), (CR, 1, 63, 0, 1), (CB, 1, 63, 0, 1), (Y, 6, 63, 0, 2),
               (Y, 1, 63, 2, 1), ([0, 1, 2], 0, 0, 1, 0), (CR, 1, 63, 1, 0), (CB, 1, 63, 1, 0), (Y, 1, 63, 1, 0)]
    files.append(("libjpeg_script_q95", synth.encode(pixels_for(synth, 6, 400, 304, "420"), "420", progressive=True,
                                                     quality=95, scans=libjpeg)[0]))
    files += [(n, f) for n, f in _pillow_files(synth) if "progressive" in n]
    rng = np.random.default_rng(12)
    for mode in (1, 2):
        host.set_option("selftest_refine", mode)
        try:
            before = host.counter("selftest_refine_scans")
            for name, jpg in files:
                fr, _, _ = oracle.decode(jpg)
                _, planes = host.decode_planes(jpg)          # raises constraint_error on a self-test mismatch
                for c in range(fr.ncomp):
                    assert np.array_equal(planes[c], fr.planes[c]), (name, c)
            assert host.counter("selftest_refine_scans") - before >= 2 * len(files)
            # damaged files: whatever the decoder proper does (planes or exception), the history decode did the
            # same up to the scan in which it happened -- a mismatch would surface as constraint_error
            jpg = files[-1][1]
            sos = jpg.index(b"\xff\xda")
            for t in range(60):
                bad = bytearray(jpg)
                pos = int(rng.integers(sos + 10, len(bad) - 2))
                if t % 3 == 0:
                    bad[pos] = int(rng.integers(0, 256))
                elif t % 3 == 1:
                    del bad[pos]
                else:
                    del bad[pos:pos + int(rng.integers(2, 30))]
                try:
                    host.decode_planes(bytes(bad))
                except host.GidHostError as e:
                    assert "self-test" not in str(e), (mode, t, str(e))
        finally:
            host.set_option("selftest_refine", 0)
