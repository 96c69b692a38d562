"""The oracle (oracle/gidref.c) against a second restatement of the same Ada text (tests/pyref.py):
inverse DCT on random and extreme blocks, the colour transformation on every (Y, Cb, Cr), whole frames
with every sampling layout and partial MCUs."""
import numpy as np
import pytest

from tests import pyref
from tests.cases import pixels_for, random_planes


@pytest.mark.parametrize("kind", ["dense", "sparse", "dc_only", "rows", "extreme"])
def test_idct_two_restatements_agree(oracle, kind):
    rng = np.random.default_rng(hash(kind) & 0xFFFF)
    blocks = random_planes(rng, [40], [25], kind)[0].reshape(-1, 64).astype(np.int64)
    if kind != "extreme":
        blocks = blocks * rng.integers(1, 40, size=(1, 64))      # de-quantised magnitudes
    got = pyref.idct_blocks(blocks)
    for i in range(blocks.shape[0]):
        assert np.array_equal(got[i], np.asarray(oracle.idct(blocks[i])).reshape(64)), (kind, i)


def test_idct_single_coefficient_sweeps(oracle):
    """Every position alone, negative and positive, small and large: the row shortcut (:810-812), the
    column shortcut (:858-862) and the truncating divisions each see both signs."""
    blocks = []
    for pos in range(64):
        for v in (-2040, -1024, -300, -17, -16, -9, -1, 1, 7, 8, 100, 1016, 2000):
            b = np.zeros(64, dtype=np.int64)
            b[pos] = v
            blocks.append(b)
            b = b.copy()
            b[0] = -v
            blocks.append(b)
    blocks = np.array(blocks)
    got = pyref.idct_blocks(blocks)
    for i in range(blocks.shape[0]):
        assert np.array_equal(got[i], np.asarray(oracle.idct(blocks[i])).reshape(64)), i


def test_colour_every_triple(oracle):
    """All 2**24 (Y, Cb, Cr): the Python restatement against the documented survey vectors, and a
    sample of 30 000 triples against the oracle (a ctypes call each)."""
    y, cb, cr = np.meshgrid(np.arange(256), np.arange(256), np.arange(256), indexing="ij")
    r, g, b = pyref.ycbcr_to_rgb(y, cb, cr)
    for (yy, bb, rr), want in (((0, 0, 0), (0, 136, 0)), ((255, 255, 255), (255, 121, 255)), ((128, 128, 128), (128, 128, 128)),
                               ((76, 85, 255), (254, 0, 0)), ((150, 44, 21), (0, 255, 1)), ((29, 255, 107), (0, 0, 254)),
                               ((255, 0, 0), (76, 255, 28)), ((0, 255, 255), (178, 0, 225)), ((200, 100, 50), (91, 255, 150))):
        assert (r[yy, bb, rr], g[yy, bb, rr], b[yy, bb, rr]) == want
    rng = np.random.default_rng(4)
    for yy, bb, rr in rng.integers(0, 256, size=(30000, 3)):
        assert oracle.ycbcr_to_rgb(int(yy), int(bb), int(rr)) == (r[yy, bb, rr], g[yy, bb, rr], b[yy, bb, rr])


@pytest.mark.parametrize("w,h,samp", [(64, 48, "444"), (97, 61, "420"), (130, 35, "422"), (33, 70, "440"), (75, 50, "grey"),
                                      (100, 60, "411"), (17, 9, "420")])
def test_frames_two_restatements_agree(oracle, synth, w, h, samp):
    fr = synth.planes_only(pixels_for(synth, 12, w, h, samp), samp)
    got = pyref.reconstruct(w, h, fr.samp_h, fr.samp_v, np.asarray(fr.qt), fr.planes[:fr.ncomp])
    assert np.array_equal(got, oracle.reconstruct(fr))


def test_edge_of_the_domain_two_restatements_agree(oracle, synth):
    """Tables with entries of 0, products that are multiples of 2**21 (blk(4) * 2**11 wraps to 0 and the
    row shortcut is taken, :803-812), DC terms whose blk(0) * 2**11 wraps, full-range garbage: the two
    restatements of the Ada text give the same bytes."""
    from tests.cases import edge_frames
    for name, fr in edge_frames(synth):
        got = pyref.reconstruct(fr.width, fr.height, fr.samp_h, fr.samp_v, np.asarray(fr.qt), fr.planes[:fr.ncomp])
        assert np.array_equal(got, oracle.reconstruct(fr)), name
