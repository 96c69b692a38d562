"""Generates the golden fixtures of tests/golden/ (run once, here, then commit).

    python tests/golden/make_golden.py

Inputs come from the harness encoder; expected outputs from the oracle, which
at generation time had been checked against the survey's KAT-1..KAT-10 (five
of them SHA-256 of GID's decode of reference/test/img files as restated
independently during the survey).  The fixtures pin the oracle from then on:
any later change of its arithmetic shows up as a hash mismatch.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from tests.harness import Oracle, Synth  # noqa: E402

CASES = [
    # name, w, h, sampling, progressive, restart, extra
    ("g01_420_baseline", 64, 48, "420", False, 0, {}),
    ("g02_444_baseline", 40, 24, "444", False, 0, {}),
    ("g03_422_baseline", 72, 40, "422", False, 0, {}),
    ("g04_grey_dri", 56, 40, "grey", False, 3, {}),
    ("g05_420_partial_mcu", 37, 29, "420", False, 5, {}),
    ("g06_422_progressive", 64, 64, "422", True, 0, {}),
    ("g07_420_progressive", 45, 35, "420", True, 0, {}),
    ("g08_444_progressive_dqt16", 32, 32, "444", True, 0, {"dqt16": True}),
    ("g09_grey_progressive", 40, 40, "grey", True, 0, {}),
    ("g10_440_baseline", 48, 48, "440", False, 0, {}),
    ("g11_411_baseline", 64, 24, "411", False, 0, {}),
    ("g12_1x1_444", 1, 1, "444", False, 0, {}),
    ("g13_7x3_grey", 7, 3, "grey", False, 0, {}),
    ("g14_420_q30_optimised", 96, 64, "420", False, 0, {"quality": 30, "optimize": True}),
    ("g15_420_libjpeg_script", 80, 56, "420", True, 0, {"quality": 90, "scans": [
        ([0, 1, 2], 0, 0, 0, 1), ([0], 1, 5, 0, 2), ([2], 1, 63, 0, 1), ([1], 1, 63, 0, 1),
        ([0], 6, 63, 0, 2), ([0], 1, 63, 2, 1), ([0, 1, 2], 0, 0, 1, 0), ([2], 1, 63, 1, 0),
        ([1], 1, 63, 1, 0), ([0], 1, 63, 1, 0)]}),
]


def main() -> None:
    S, O = Synth(), Oracle()
    files = []
    for i, (name, w, h, samp, prog, ri, extra) in enumerate(CASES):
        img = S.image(1000 + i, w, h)
        pix = img[:, :, 1].copy() if samp == "grey" else img
        jpg, enc = S.encode(pix, samp, progressive=prog, restart_interval=ri, **extra)
        fr, rgb, _ = O.decode(jpg)
        assert all(np.array_equal(a, b) for a, b in zip(fr.planes, enc.planes))
        open(os.path.join(HERE, name + ".jpg"), "wb").write(jpg)
        hp = hashlib.sha256()
        for p in fr.planes:
            hp.update(p.tobytes())
        ent = dict(jpg=name + ".jpg", width=w, height=h, sampling=samp, progressive=prog,
                   restart_interval=ri, rgb_sha256=hashlib.sha256(rgb.tobytes()).hexdigest(),
                   planes_sha256=hp.hexdigest())
        if w * h <= 64 * 48:
            rgb.reshape(-1).tofile(os.path.join(HERE, name + ".rgb"))
            ent["rgb_raw"] = name + ".rgb"
        files.append(ent)
    rng = np.random.default_rng(20240302)
    blocks = []
    for k in range(24):
        coef = np.zeros(64, dtype=int)
        n = int(rng.integers(1, 20))
        idx = rng.choice(64, size=n, replace=False)
        coef[idx] = rng.integers(-600, 601, size=n)
        if k % 3 == 0:
            coef[1:8] = 0           # first row: row shortcut with whatever DC
            coef[0] = -int(rng.integers(1, 900))
        blocks.append(dict(coef=coef.tolist(), samples=O.idct(coef).tolist()))
    json.dump(dict(files=files, blocks=blocks), open(os.path.join(HERE, "index.json"), "w"), indent=1)
    print(f"wrote {len(files)} files and {len(blocks)} blocks")


if __name__ == "__main__":
    main()
