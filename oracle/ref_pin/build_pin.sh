#!/bin/bash
# Pins the oracle to the reference itself -- for a box that has GNAT (this repository's image has none).
#   oracle/ref_pin/build_pin.sh [/path/to/gid] [--with-cuda]
# 1. builds oracle/_ref/gid_pin from the UNMODIFIED GID sources (gprbuild, -O2 -gnatpn = GID's Fast_unchecked mode);
# 2. decodes tests/golden/*.jpg and the five KAT files of the GID tree with it and writes
#    tests/golden/gid_pin_hashes.txt, which tests/test_reference_pin.py holds against the oracle's hashes;
# 3. --with-cuda: also compiles the binding GID.CUDA, links libgidcuda.so and runs GID.CUDA_Check.
set -e
here="$(cd "$(dirname "$0")" && pwd)"; repo="$(cd "$here/../.." && pwd)"
gid="${1:-/root/reference}"; cuda=no; [ "$2" = "--with-cuda" ] && cuda=yes
command -v gprbuild >/dev/null || { echo "gprbuild (GNAT) not found: parity stays unpinned" >&2; exit 3; }
gprbuild -p -P "$here/gid_pin.gpr" -XGID_DIR="$gid" -XWITH_CUDA=$cuda
kat="exif_rotation_0.jpg jpeg_progressive_walensee.jpg jpeg_baseline_hifi_25_pct_quality.jpg jpeg_progressive_lyon_25_pct_quality.jpg jpeg_baseline_maschgenkamm.jpg"
files=$(ls "$repo"/tests/golden/*.jpg); for k in $kat; do files="$files $gid/test/img/$k"; done
"$repo/oracle/_ref/gid_pin" $files > "$repo/tests/golden/gid_pin_hashes.txt"
echo "wrote tests/golden/gid_pin_hashes.txt ($(wc -l < "$repo/tests/golden/gid_pin_hashes.txt") files); now: python -m pytest tests/test_reference_pin.py -q"
[ $cuda = yes ] && "$repo/oracle/_ref/gid_cuda_check_main"
exit 0
