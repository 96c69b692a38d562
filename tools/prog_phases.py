#!/usr/bin/env python
"""Where a host thread's time goes on progressive files that take the device path (cfg5):
    python tools/prog_phases.py [threads ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gid_b200 import capi, host_api
w, h, _, _, _ = bench.WORKLOADS["cfg5"]
distinct = bench.make_files("cfg5", 2, 77)
if os.environ.get("BLOCKING_WAIT"):
    capi.set_option("blocking_wait", int(os.environ["BLOCKING_WAIT"]))
names = ("ns_pscan_queue", "ns_masks_wait", "ns_refine_host")
for T in [int(a) for a in sys.argv[1:]] or [1, 16]:
    n = 8 * T
    files = [distinct[i % 2] for i in range(n)]
    out = [capi.pinned_empty(3 * w * h)[0].reshape(h, w, 3) for _ in range(n)]
    with host_api.HostBatch(devices=(0,), threads=T) as hb:
        for _ in range(2):
            hb.decode(files, out)
        c0 = [host_api.counter(k) for k in names]
        t0 = time.perf_counter()
        _, st, stats = hb.decode(files, out)
        dt = time.perf_counter() - t0
        c1 = [host_api.counter(k) for k in names]
    per = 1e3 * stats.seconds * T / n
    print(f"threads {T}: {w*h*n/1e6/stats.seconds:.0f} MP/s, {per:.2f} ms per image per thread; of it (ms per image): " +
          ", ".join(f"{k[3:]} {(b - a) / 1e6 / n:.2f}" for k, a, b in zip(names, c0, c1)), flush=True)
    del out
