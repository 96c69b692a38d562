"""Shared test-case tables (sizes, sampling layouts, modes)."""
from __future__ import annotations

import numpy as np

# (width, height, sampling) -- covers the four fused layouts, the generic path
# (4:4:0 = h1v2, 4:1:1 = h4v1), partial MCUs on both edges and the tiny sizes
# GID accepts (subsampled components need >= 3 samples).
GEOMETRIES = [
    (512, 512, "420"), (256, 256, "444"), (640, 480, "422"), (300, 200, "grey"),
    (17, 9, "420"), (239, 134, "422"), (1, 1, "444"), (7, 3, "grey"), (7, 3, "444"),
    (100, 60, "440"), (100, 60, "411"), (333, 217, "420"), (1919 // 4, 1079 // 4, "420"),
    (8, 8, "444"), (16, 16, "420"), (5, 5, "420"), (1030, 9, "grey"), (9, 1030, "444"),
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
