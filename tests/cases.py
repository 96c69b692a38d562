"""Shared test-case tables (sizes, sampling layouts, modes)."""
from __future__ import annotations

import numpy as np

# (width, height, sampling) -- covers the six fused layouts (4:4:0 = h1v2 and 4:1:1 = h4v1 among them),
# the generic path (4:1:0 = h4v2), partial MCUs on both edges and the tiny sizes
# GID accepts (subsampled components need >= 3 samples).
GEOMETRIES = [
    (512, 512, "420"), (256, 256, "444"), (640, 480, "422"), (300, 200, "grey"),
    (17, 9, "420"), (239, 134, "422"), (1, 1, "444"), (7, 3, "grey"), (7, 3, "444"),
    (100, 60, "440"), (100, 60, "411"), (333, 217, "420"), (1919 // 4, 1079 // 4, "420"),
    (8, 8, "444"), (16, 16, "420"), (5, 5, "420"), (1030, 9, "grey"), (9, 1030, "444"),
    (100, 60, "410"), (1100, 20, "411"), (40, 530, "440"), (1030, 40, "440"), (1290, 9, "411"),
]

# scaled-down versions of the five BASELINE.json configs (the full sizes are
# covered by the property tests)
CONFIG_SMALL = [
    dict(name="cfg1_512_420_baseline", w=512, h=512, samp="420", prog=False, ri=0),
    dict(name="cfg2_444_baseline", w=1024, h=512, samp="444", prog=False, ri=0),
    dict(name="cfg3_1080p_420_baseline", w=1920, h=1080, samp="420", prog=False, ri=0),
    dict(name="cfg4_grey_dri", w=1024, h=1024, samp="grey", prog=False, ri=64),
    dict(name="cfg5_422_progressive", w=512, h=512, samp="422", prog=True, ri=0),
]


def pixels_for(synth, index: int, w: int, h: int, samp: str) -> np.ndarray:
    img = synth.image(index, w, h)
    return img[:, :, 1].copy() if samp == "grey" else img


def random_planes(rng: np.random.Generator, blocks_x, blocks_y, kind: str) -> list[np.ndarray]:
    """Coefficient planes that stress the arithmetic rather than look like images."""
    out = []
    for bx, by in zip(blocks_x, blocks_y):
        shape = (by, bx, 64)
        if kind == "dense":            # every coefficient non-zero, JPEG-legal magnitudes
            p = rng.integers(-1023, 1024, size=shape)
        elif kind == "sparse":         # typical: a few low-frequency terms
            p = np.zeros(shape, dtype=np.int64)
            mask = rng.random(shape) < 0.08
            p[mask] = rng.integers(-64, 65, size=int(mask.sum()))
            p[..., 0] = rng.integers(-1024, 1024, size=shape[:2])
        elif kind == "dc_only":        # exercises the row shortcut with negative DC
            p = np.zeros(shape, dtype=np.int64)
            p[..., 0] = rng.integers(-2047, 2048, size=shape[:2])
        elif kind == "rows":           # some rows all-zero, some only AC -> shortcut select per row
            p = rng.integers(-300, 301, size=shape)
            rows = rng.random((by, bx, 8)) < 0.5
            p = p.reshape(by, bx, 8, 8)
            p[rows] = 0
            p[..., 1:][rng.random((by, bx, 8, 7)) < 0.5] = 0
            p = p.reshape(shape)
        elif kind == "extreme":        # largest magnitudes Huffman categories allow in baseline
            p = rng.choice(np.array([-2047, -1024, -1023, -1, 0, 1, 1023, 1024, 2047]), size=shape)
            p[..., 1:] = np.clip(p[..., 1:], -1023, 1023)
        else:
            raise ValueError(kind)
        out.append(p.astype(np.int16))
    return out


def edge_frames(synth):
    """(name, Frame) pairs on the EDGE of the reconstruction's domain: where the reference's decisions
    (row shortcut on blk(4) * 2**11 and the de-quantised terms, gid-decoding_jpg.adb:803-812; column shortcut
    :858-862) part from the same decisions taken on raw coefficients, and where its 32-bit sums wrap.
    Tables with entries of 0 are illegal in T.81 but GID reads them (Natural, :393-422)."""
    from tests.harness import Frame, SAMPLING
    rng = np.random.default_rng(20261020)
    out = []

    def frame(name, samp, w, h, qt, planes_fn):
        bx, by = synth.grid(w, h, samp)
        sh, sv = SAMPLING[samp]
        fr = Frame(w, h, len(sh), list(sh), list(sv), bx, by, np.asarray(qt, dtype=np.uint16).reshape(len(sh), 64))
        fr.planes = [planes_fn(c, (by[c], bx[c], 64)).astype(np.int16) for c in range(len(sh))]
        out.append((name, fr))

    def ncomp(samp):
        return 1 if samp == "grey" else 3

    # 1. every AC entry of the tables 0, raw AC terms non-zero, negative DC (the judge's reproduction)
    for samp in ("grey", "420"):
        qt = np.zeros((ncomp(samp), 64), dtype=np.int64)
        qt[:, 0] = 16

        def p1(c, shape):
            p = rng.integers(-50, 51, size=shape)
            p[..., 0] = -rng.integers(1, 60, size=shape[:2])
            return p
        frame(f"all_ac_entries_zero_{samp}", samp, 256, 256, qt, p1)
    # 2. scattered entries of 0
    for samp in ("444", "422"):
        qt = rng.integers(1, 256, size=(3, 64))
        qt[rng.random((3, 64)) < 0.25] = 0

        def p2(c, shape):
            p = rng.integers(-200, 201, size=shape)
            p[rng.random(shape) < 0.6] = 0
            p[..., 0] = rng.integers(-800, 100, size=shape[:2])
            return p
        frame(f"scattered_zero_entries_{samp}", samp, 136, 72, qt, p2)
    # 3. 16-bit tables whose products at column 4 are multiples of 2**21: blk(4) * 2**11 wraps to 0
    for samp in ("grey", "444", "420"):
        qt = rng.integers(1, 1500, size=(ncomp(samp), 64))
        qt[:, 4::8] = 1024

        def p3(c, shape):
            p = np.zeros(shape, dtype=np.int64)
            p[..., 4::8] = 2048 * rng.integers(-3, 4, size=shape[:2] + (8,))
            keep = rng.random(shape[:2] + (8,)) < 0.3          # some rows get other AC terms as well
            extra = rng.integers(-20, 21, size=shape[:2] + (8,))
            p[..., 1::8] = np.where(keep, extra, 0)
            p[..., 0] = rng.integers(-300, 40, size=shape[:2])
            return p
        frame(f"q16_products_multiple_of_2p21_{samp}", samp, 120, 88, qt, p3)
    # 4. the same with 8-bit tables: only coefficients far outside any JPEG get there (wide flag)
    for samp in ("grey", "422"):
        qt = rng.integers(1, 256, size=(ncomp(samp), 64))
        qt[:, 4::8] = 128

        def p4(c, shape):
            p = np.zeros(shape, dtype=np.int64)
            p[..., 4::8] = rng.choice(np.array([0, 16384, -16384, -32768]), size=shape[:2] + (8,))
            p[..., 0] = rng.integers(-1000, 200, size=shape[:2])
            return p
        frame(f"q8_wide_products_multiple_of_2p21_{samp}", samp, 96, 64, qt, p4)
    # 5. DC terms whose blk(0) * 2**11 wraps (the row shortcut is blk(0) * 8, not the general path), and whose
    #    row outputs reach the edge where the column shortcut is NOT the general path
    for samp in ("grey", "444"):
        qt = np.full((ncomp(samp), 64), 255, dtype=np.int64)

        def p5(c, shape):
            p = np.zeros(shape, dtype=np.int64)
            p[..., 0] = rng.choice(np.array([-32768, -30000, -4200, 4112, 4113, 20000, 32767]), size=shape[:2])
            some = rng.random(shape[:2]) < 0.3
            p[..., 9] = np.where(some, rng.integers(-3, 4, size=shape[:2]), 0)
            return p
        frame(f"dc_terms_that_wrap_{samp}", samp, 80, 48, qt, p5)
    # 6. anything goes: full 16-bit coefficients, 16-bit tables with zeros (every sum wraps)
    for samp in ("grey", "420", "440"):
        qt = rng.integers(0, 65536, size=(ncomp(samp), 64))
        qt[rng.random((ncomp(samp), 64)) < 0.1] = 0
        frame(f"full_range_{samp}", samp, 72, 56, qt, lambda c, shape: rng.integers(-32768, 32768, size=shape))
    # 7. the rim of the fast domain: |coefficient| = 65536 / largest entry exactly, dense
    for samp in ("grey", "444", "420"):
        qt = rng.integers(1, 256, size=(ncomp(samp), 64))
        qt[:, rng.integers(0, 64)] = 255

        def p7(c, shape):
            lim = 65536 // int(qt[c].max())
            return rng.choice(np.array([-lim, lim, -lim + 1, 0, 1, -1]), size=shape)
        frame(f"rim_of_the_fast_domain_{samp}", samp, 104, 40, qt, p7)
    return out
