"""A second, independent restatement of GID's reconstruction arithmetic -- vectorised numpy, written
from the Ada text (gid-decoding_jpg.adb) and sharing no code with oracle/gidref.c -- used only to pin
the oracle: two restatements that agree on random and extreme inputs are the strongest evidence there
is while GID itself cannot be compiled here (no GNAT).

    Inverse_DCT_8x8_Block   :790-907      (Ada "/" truncates toward zero)
    Color_Transformation    :1028-1038
    upsampling              :966-1012     (replication) and clipping to W x H, :1023-1026
"""
from __future__ import annotations

import numpy as np

# GID's Integer is 32 bits and the benchmark build (-gnatp) does not check for overflow: arithmetic wraps.
# numpy's int32 array arithmetic wraps the same way, silently.
I32 = np.int32

W1, W2, W3, W5, W6, W7 = 2841, 2676, 2408, 1609, 1108, 565


def ada_div(x: np.ndarray, k: int) -> np.ndarray:
    """x / 2**k with Ada semantics (toward zero) on int32 arrays."""
    x = np.asarray(x, dtype=I32)
    with np.errstate(over="ignore"):
        return np.where(x >= 0, x >> k, -((-x) >> k)).astype(I32)


def clip(x: np.ndarray) -> np.ndarray:  # :755-756
    return np.minimum(np.maximum(x, 0), 255)


def idct_blocks(coef: np.ndarray) -> np.ndarray:
    """coef: (n, 64) de-quantised coefficients, natural order.  Returns (n, 64) samples 0 .. 255."""
    b = np.asarray(coef).astype(I32).reshape(-1, 8, 8).copy()
    # ---- Row_IDCT on every row (:799-845)
    r = b.reshape(-1, 8)                       # one row per line
    K = lambda v: I32(v)  # (keep every constant 32 bits wide, or numpy would widen the product)
    x1 = r[:, 4] * K(2048)
    x2, x3, x4, x5, x6, x7 = r[:, 6], r[:, 2], r[:, 1], r[:, 7], r[:, 5], r[:, 3]
    shortcut = (x1 == 0) & (x2 == 0) & (x3 == 0) & (x4 == 0) & (x5 == 0) & (x6 == 0) & (x7 == 0)
    x0 = r[:, 0] * K(2048) + K(128)
    x8 = K(W7) * (x4 + x5)
    x4, x5 = x8 + K(W1 - W7) * x4, x8 - K(W1 + W7) * x5
    x8 = K(W3) * (x6 + x7)
    x6, x7 = x8 - K(W3 - W5) * x6, x8 - K(W3 + W5) * x7
    x8, x0 = x0 + x1, x0 - x1
    x1 = K(W6) * (x3 + x2)
    x2, x3 = x1 - K(W2 + W6) * x2, x1 + K(W2 - W6) * x3
    x1, x4 = x4 + x6, x4 - x6
    x6, x5 = x5 + x7, x5 - x7
    x7, x8 = x8 + x3, x8 - x3
    x3, x0 = x0 + x2, x0 - x2
    x2 = ada_div(K(181) * (x4 + x5) + K(128), 8)
    x4 = ada_div(K(181) * (x4 - x5) + K(128), 8)
    general = np.stack([ada_div(x7 + x1, 8), ada_div(x3 + x2, 8), ada_div(x0 + x4, 8), ada_div(x8 + x6, 8),
                        ada_div(x8 - x6, 8), ada_div(x0 - x4, 8), ada_div(x3 - x2, 8), ada_div(x7 - x1, 8)], axis=1)
    rows = np.where(shortcut[:, None], (r[:, 0] * K(8))[:, None], general).astype(I32)
    # ---- Col_IDCT on every column (:847-895)
    c = rows.reshape(-1, 8, 8).transpose(0, 2, 1).reshape(-1, 8)   # one column per line
    x1 = c[:, 4] * K(256)
    x2, x3, x4, x5, x6, x7 = c[:, 6], c[:, 2], c[:, 1], c[:, 7], c[:, 5], c[:, 3]
    shortcut = (x1 == 0) & (x2 == 0) & (x3 == 0) & (x4 == 0) & (x5 == 0) & (x6 == 0) & (x7 == 0)
    flat = clip(ada_div(c[:, 0] + K(32), 6) + K(128))
    x0 = c[:, 0] * K(256) + K(8192)
    x8 = K(W7) * (x4 + x5) + K(4)
    x4, x5 = ada_div(x8 + K(W1 - W7) * x4, 3), ada_div(x8 - K(W1 + W7) * x5, 3)
    x8 = K(W3) * (x6 + x7) + K(4)
    x6, x7 = ada_div(x8 - K(W3 - W5) * x6, 3), ada_div(x8 - K(W3 + W5) * x7, 3)
    x8, x0 = x0 + x1, x0 - x1
    x1 = K(W6) * (x3 + x2) + K(4)
    x2, x3 = ada_div(x1 - K(W2 + W6) * x2, 3), ada_div(x1 + K(W2 - W6) * x3, 3)
    x1, x4 = x4 + x6, x4 - x6
    x6, x5 = x5 + x7, x5 - x7
    x7, x8 = x8 + x3, x8 - x3
    x3, x0 = x0 + x2, x0 - x2
    x2 = ada_div(K(181) * (x4 + x5) + K(128), 8)
    x4 = ada_div(K(181) * (x4 - x5) + K(128), 8)
    general = np.stack([clip(ada_div(v, 14) + K(128)) for v in
                        (x7 + x1, x3 + x2, x0 + x4, x8 + x6, x8 - x6, x0 - x4, x3 - x2, x7 - x1)], axis=1)
    cols = np.where(shortcut[:, None], flat[:, None], general)
    return cols.reshape(-1, 8, 8).transpose(0, 2, 1).reshape(-1, 64)


def ycbcr_to_rgb(y, cb, cr):
    """:1028-1035 on arrays of samples 0 .. 255."""
    y = np.asarray(y).astype(I32) * I32(256)
    cb = np.asarray(cb).astype(I32) - I32(128)
    cr = np.asarray(cr).astype(I32) - I32(128)
    return (clip(ada_div(y + I32(359) * cr + I32(128), 8)), clip(ada_div(y - I32(88) * cb - I32(183) * cr + I32(128), 8)),
            clip(ada_div(y + I32(454) * cb + I32(128), 8)))


def reconstruct(width, height, samp_h, samp_v, qt, planes) -> np.ndarray:
    """Whole frame: planes[c] (by, bx, 64) un-dequantised int16, qt (ncomp, 64) natural order.
    Returns (H, W, 3) uint8, top-down (the buffer of test/benchmark.adb:67-81)."""
    ncomp = len(planes)
    hmax, vmax = max(samp_h[:ncomp]), max(samp_v[:ncomp])
    full = []
    for c in range(ncomp):
        p = np.asarray(planes[c]).astype(I32)
        by, bx = p.shape[:2]
        s = idct_blocks((p * np.asarray(qt[c]).astype(I32)).reshape(-1, 64)).reshape(by, bx, 8, 8)
        img = s.transpose(0, 2, 1, 3).reshape(by * 8, bx * 8)                     # sample plane
        ux, uy = hmax // samp_h[c], vmax // samp_v[c]
        full.append(np.repeat(np.repeat(img, uy, axis=0), ux, axis=1)[:height, :width])  # replication, then W x H
    if ncomp == 1:
        r = g = b = full[0]
    else:
        r, g, b = ycbcr_to_rgb(full[0], full[1], full[2])
    return np.stack([r, g, b], axis=2).astype(np.uint8)
