"""The pin of the oracle to the reference ITSELF, for the day a GNAT box exists.

GID is Ada and this image has no Ada compiler, so the oracle (oracle/gidref.c) has never been held against
a GID binary: parity is unpinned by the reference.  oracle/ref_pin/ holds what closes
the gap with two commands on a box that has GNAT:

    oracle/ref_pin/build_pin.sh /path/to/gid        # gprbuild of gid_pin.adb over the UNMODIFIED GID,
                                                    # -> tests/golden/gid_pin_hashes.txt
    python -m pytest tests/test_reference_pin.py

gid_pin decodes every tests/golden/*.jpg and the five KAT files of the GID tree through
GID.Load_Image_Contents into the top-down RGB8 buffer of test/benchmark.adb:67-81 and prints the SHA-256.
When the hash file is present this test compares it with what the oracle says about the same files; when it
is not, the test SKIPS with the words "parity unpinned".
"""
import hashlib
import json
import os

import pytest

from tests.test_oracle_kat import KAT_FIXTURES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
HASHES = os.path.join(GOLDEN, "gid_pin_hashes.txt")


def _pin_lines():
    out = {}
    for line in open(HASHES):
        parts = line.split(None, 3)
        if len(parts) == 4:
            out[os.path.basename(parts[3].strip())] = (parts[0], int(parts[1]), int(parts[2]))
    return out


def test_pin_sources_are_in_place():
    """What a maintainer with GNAT needs is committed: the program, the project file (GID's own Fast_unchecked
    switches, and the -lgidcuda link line for the binding), the script."""
    pin = os.path.join(ROOT, "oracle", "ref_pin")
    for f in ("gid_pin.adb", "gid_pin.gpr", "build_pin.sh", "gid-cuda_check.ads", "gid-cuda_check.adb",
              "gid_cuda_check_main.adb"):
        assert os.path.exists(os.path.join(pin, f)), f
    gpr = open(os.path.join(pin, "gid_pin.gpr")).read()
    assert '"-O2", "-gnatpn"' in gpr and "-lgidcuda" in gpr and "gid-cuda.adb" in gpr
    adb = open(os.path.join(pin, "gid_pin.adb")).read()
    assert "GID.Load_Image_Contents" in adb and "image_height - 1 - y" in adb and "GNAT.SHA256" in adb


def test_oracle_agrees_with_the_gid_binary(oracle):
    if not os.path.exists(HASHES):
        pytest.skip("parity unpinned: tests/golden/gid_pin_hashes.txt has not been produced "
                    "(needs GNAT: oracle/ref_pin/build_pin.sh)")
    pin = _pin_lines()
    index = json.load(open(os.path.join(GOLDEN, "index.json")))
    checked = 0
    for ent in index["files"]:
        if ent["jpg"] not in pin:
            continue
        sha, w, h = pin[ent["jpg"]]
        assert (w, h) == (ent["width"], ent["height"]), ent["jpg"]
        assert sha == ent["rgb_sha256"], f"{ent['jpg']}: GID {sha} != oracle {ent['rgb_sha256']}"
        # ... and the oracle as it is built now still says what index.json recorded
        _, rgb, _ = oracle.decode(open(os.path.join(GOLDEN, ent["jpg"]), "rb").read())
        assert hashlib.sha256(rgb.tobytes()).hexdigest() == sha, ent["jpg"]
        checked += 1
    for name, (sha, _, _) in KAT_FIXTURES.items():
        if name in pin:
            assert pin[name][0] == sha, f"{name}: GID {pin[name][0]} != survey / oracle {sha}"
            checked += 1
    assert checked >= len(index["files"]), f"only {checked} files of the pin were compared"
