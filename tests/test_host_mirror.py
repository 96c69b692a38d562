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
