#!/usr/bin/env python
"""One gidhost_batch_decode call over a few files of a workload (for ncu launch lists).
    python tools/files_once.py cfg3 [nfiles] [threads]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from gid_b200 import host_api
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
T = int(sys.argv[3]) if len(sys.argv) > 3 else 1
w, h, _, _, _ = bench.WORKLOADS[wl]
files = bench.make_files(wl, 2, 77)
files = [files[i % 2] for i in range(n)]
imgs, st, stats = host_api.decode_batch(files, threads=T)
print(wl, st, "%.1f MP/s" % (w * h * n / 1e6 / stats.seconds), "device scans", host_api.counter("device_scans"),
      "left to the host", host_api.counter("device_scan_fallbacks"))
