"""Properties of the oracle itself: the statements SURVEY.md makes about GID's
arithmetic, the agreement of its two decode routes, and its error behaviour."""
import numpy as np
import pytest

from tests.cases import CONFIG_SMALL, pixels_for
from tests.harness import OracleError


def test_column_shortcut_is_redundant_row_shortcut_is_not(oracle):
    rng = np.random.default_rng(1)
    diff_shift = diff_row = diff_col = total = 0
    for _ in range(400):
        blk = rng.integers(-256, 257, size=64)
        blk[rng.random(64) < 0.6] = 0
        ref = oracle.idct(blk)
        diff_shift += int((oracle.idct_variant(blk, shifts=True) != ref).sum())
        diff_row += int((oracle.idct_variant(blk, no_row=True) != ref).sum())
        diff_col += int((oracle.idct_variant(blk, no_col=True) != ref).sum())
        total += 64
    assert diff_col == 0                      # :858-862 can be dropped
    assert diff_row > 0                       # :810-812 cannot
    assert 0.2 < diff_shift / total < 0.7     # ">>" instead of "/" changes a large share of samples


@pytest.mark.parametrize("cfg", CONFIG_SMALL, ids=[c["name"] for c in CONFIG_SMALL])
def test_front_end_planes_and_both_routes_agree(oracle, synth, cfg):
    jpg, enc = synth.encode(pixels_for(synth, 5, cfg["w"] // 2, cfg["h"] // 2, cfg["samp"]), cfg["samp"],
                            progressive=cfg["prog"], restart_interval=cfg["ri"])
    fr, rgb, fb = oracle.decode(jpg)
    assert not fr.coef_overflow
    for c in range(fr.ncomp):
        assert np.array_equal(fr.planes[c], enc.planes[c])
        assert np.array_equal(fr.qt[c], enc.qt[c])
    assert np.array_equal(oracle.reconstruct(fr), rgb)
    assert fb == sorted(fb) and fb[-1] == 100


def test_client_callbacks_and_16bit_range(oracle, synth):
    jpg, _ = synth.encode(pixels_for(synth, 9, 40, 24, "420"), "420")
    _, rgb, _ = oracle.decode(jpg)
    _, img8, fb8 = oracle.decode_client(jpg, 256)
    assert np.array_equal(img8[:, :, :3].astype(np.uint8), rgb) and (img8[:, :, 3] == 255).all()
    _, img16, _ = oracle.decode_client(jpg, 65536)
    assert np.array_equal(img16[:, :, :3], rgb.astype(np.uint32) * 257) and (img16[:, :, 3] == 65535).all()
    with pytest.raises(OracleError) as e:
        oracle.decode_client(jpg, 1024)
    assert e.value.code == -5  # invalid_primary_color_range, gid-decoding_jpg.adb:934-936
    assert fb8 == sorted(fb8)


def test_error_behaviour(oracle, synth):
    jpg, _ = synth.encode(pixels_for(synth, 9, 64, 48, "420"), "420")
    with pytest.raises(OracleError) as e:
        oracle.decode(b"\x89PNG\r\n")
    assert e.value.code == -1
    with pytest.raises(OracleError) as e:          # header cut before SOF: End_Error propagates
        oracle.probe(jpg[:30])
    assert e.value.code == -6
    sof = jpg.index(b"\xff\xc0")
    bad = bytearray(jpg)
    bad[sof + 4] = 12                              # 12-bit precision
    with pytest.raises(OracleError) as e:
        oracle.probe(bytes(bad))
    assert e.value.code == -3
    bad = bytearray(jpg)
    bad[sof + 1] = 0xC9                            # arithmetic-coded SOF9
    with pytest.raises(OracleError) as e:
        oracle.probe(bytes(bad))
    assert e.value.code == -3
    # truncated entropy data: GID pads the bit buffer with FF and carries on or
    # reports a data error, never crashes (gid-decoding_jpg.adb:665-670)
    try:
        oracle.decode(jpg[:len(jpg) // 2])
    except OracleError as err:
        assert err.code in (-4, -6, -7)
    bad = bytearray(jpg)
    bad[jpg.index(b"\xff\xda") + 2 + 10:] = b"\xff\xe5" * 40   # invalid marker inside the scan
    with pytest.raises(OracleError) as e:
        oracle.decode(bytes(bad))
    assert e.value.code == -4


def test_restart_markers_reset_predictors(oracle, synth):
    pix = pixels_for(synth, 4, 160, 96, "420")
    a, _ = synth.encode(pix, "420", restart_interval=0)
    b, _ = synth.encode(pix, "420", restart_interval=4)
    fa, ra, _ = oracle.decode(a)
    fb_, rb, _ = oracle.decode(b)
    assert fb_.restart_interval == 4 and np.array_equal(ra, rb)
