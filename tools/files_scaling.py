#!/usr/bin/env python
"""Host-thread scaling of the .jpg-files leg (gidhost_batch_decode): MP/s at 1, 2, 4 ... threads.

    python tools/files_scaling.py [workload ...]
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from gid_b200 import host_api
if os.environ.get("FILES_BLOCKING_WAIT"):
    from gid_b200 import capi as _capi
    _capi.set_option("blocking_wait", int(os.environ["FILES_BLOCKING_WAIT"]))
if os.environ.get("FILES_SPARSE_INPUT"):
    host_api.set_option("sparse_input", int(os.environ["FILES_SPARSE_INPUT"]))
for wl in (sys.argv[1:] or ["cfg3"]):
    w, h, _, _, _ = bench.WORKLOADS[wl]
    prog, ri, per_call = bench.FILE_FORM[wl]
    per_call = int(os.environ.get("FILES_PER_CALL", per_call))   # (length of a call: its ramp and drain are inside the time)
    distinct = bench.make_files(wl, 2, 77)
    files = [distinct[i % 2] for i in range(per_call)]
    from gid_b200 import capi
    pinned = os.environ.get("FILES_PINNED", "1") != "0"
    out = [capi.pinned_empty(3 * w * h)[0].reshape(h, w, 3) if pinned else np.empty((h, w, 3), dtype=np.uint8) for _ in range(per_call)]
    # entropy decode alone (no device), one thread
    t0 = time.perf_counter(); host_api.decode_planes(distinct[0]); t1 = time.perf_counter()
    print(f"{wl}: decode_planes (1 thread, incl. allocation + copy to numpy) {w*h/1e6/(t1-t0):.0f} MP/s", flush=True)
    T = int(os.environ.get("FILES_MIN_THREADS", 1))
    ncpu = os.cpu_count() or 1
    while True:
        with host_api.HostBatch(devices=(0,), threads=T) as hb:
            for _ in range(3):  # warm-up: contexts, three jobs per worker
                hb.decode(files, out)
            c0 = [host_api.counter(k) for k in ("device_scans", "device_scan_fallbacks")]
            _, st, stats = hb.decode(files, out)
            c1 = [host_api.counter(k) for k in ("device_scans", "device_scan_fallbacks")]
        print(f"{wl}: threads {stats.threads:3d} (x{stats.entropy_threads} inside an image)  {w*h*per_call/1e6/stats.seconds:8.0f} MP/s  {1e3*stats.seconds/per_call*stats.threads:7.2f} ms per image per thread"
              f"   device scans {c1[0]-c0[0]}, left to the host {c1[1]-c0[1]}", flush=True)
        tmax = int(os.environ.get("FILES_MAX_THREADS", ncpu))
        if T >= tmax:
            break
        T = min(tmax, T * 2 if T < ncpu else T + ncpu // 2)
