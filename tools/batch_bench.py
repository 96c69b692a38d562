#!/usr/bin/env python
"""The IN-PROCESS batch drivers on every GPU of the box (north star, subsystem 4: per-GPU work queues, no NCCL).

    python tools/batch_bench.py [cfg3 cfg2 ...]        # all visible devices, then 1, 2, 4 of them

Per workload and device count N, ONE process:
  records  gidcuda_batch_run_records: pinned coefficient records in (what an entropy decoder emits), pinned
           RGB8 out, one worker thread + three chunks in flight per device;
  planes   gidcuda_batch_run: the same with dense int16 planes in (pinned);
  files    gidhost_batch_decode: .jpg bytes in, Huffman decoding on the devices, pinned RGB8 out.
Prints MP/s, bytes over the link and images per device.  (bench.py measures the same path one RANK per GPU, as the
driver launches it; this is the one-process form.)
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np

import bench
from gid_b200 import capi, host_api
from tests.harness import frame_desc


def pinned_copy(a: np.ndarray, keep: list) -> int:
    buf, addr = capi.pinned_empty(max(a.nbytes, 16))
    keep.append(addr)
    buf[:a.nbytes] = np.frombuffer(a.tobytes(), dtype=np.uint8)
    return addr


def main() -> None:
    L = capi.load_library()
    ndev = L.gidcuda_device_count()
    counts = sorted({n for n in (1, 2, 4, 8, ndev) if n <= ndev})
    if os.environ.get("BATCH_DEVICES"):
        counts = [int(x) for x in os.environ["BATCH_DEVICES"].split(",")]
    lines = []
    for wl in (sys.argv[1:] or ["cfg3", "cfg2"]):
        w, h, samp, _, desc_txt = bench.WORKLOADS[wl]
        frames = bench.make_frames(wl, 2, 500)
        descs1 = [frame_desc(f) for f in frames]
        ctx = capi.GidCuda(0)
        stride, nbytes = ctx.output_geometry(descs1[0])
        ctx.close()
        keep: list = []
        src_planes = [[pinned_copy(f.planes[c], keep) for c in range(f.ncomp)] for f in frames]
        src_recs = []
        for f in frames:
            per = []
            for c in range(f.ncomp):
                first, rec = capi.records_from_plane(f.planes[c], f.samp_h[c], f.samp_v[c])
                per.append((pinned_copy(np.ascontiguousarray(first, dtype=np.uint32), keep),
                            pinned_copy(np.ascontiguousarray(rec, dtype=np.uint32), keep), len(rec)))
            src_recs.append(per)
        # enough images for ~0.5 s per device at the link's rate
        per_dev = max(6, int(9e9 / (w * h)))
        chunk = int(os.environ.get("BATCH_CHUNK", 8 if w * h < 4e6 else 1))   # images per launch
        for n_dev in counts:
            n = per_dev * n_dev
            outs = []
            for _ in range(min(n, 6 * chunk * n_dev)):  # reused round-robin, twice as many as can be in flight
                o, addr = capi.pinned_empty(nbytes)
                keep.append(addr)
                outs.append(addr)
            descs = [descs1[i % 2] for i in range(n)]
            pixels = [outs[i % len(outs)] for i in range(n)]
            ncomp = frames[0].ncomp
            planes, fa, ra, ca = [], [], [], []
            for i in range(n):
                for c in range(3):
                    planes.append(src_planes[i % 2][c] if c < ncomp else 0)
                    fa.append(src_recs[i % 2][c][0] if c < ncomp else 0)
                    ra.append(src_recs[i % 2][c][1] if c < ncomp else 0)
                    ca.append(src_recs[i % 2][c][2] if c < ncomp else 0)
            batch = capi.Batch(list(range(n_dev)), chunk_images=chunk)
            try:
                for kind in ("records", "planes"):
                    run = (lambda: batch.run_records(descs, fa, ra, ca, pixels)) if kind == "records" \
                        else (lambda: batch.run(descs, planes, pixels))
                    run()                                   # warm-up: buffers of the slots
                    st, stats = run()
                    assert not any(st), st
                    line = {"tool": "batch_bench", "driver": "gidcuda_batch_run" + ("_records" if kind == "records" else ""),
                            "workload": wl, "devices": n_dev, "images": n, "chunk_images": chunk,
                            "MP/s": w * h * n / 1e6 / stats.seconds, "seconds": stats.seconds,
                            "h2d_GBps": stats.h2d_bytes / stats.seconds / 1e9, "d2h_GBps": stats.d2h_bytes / stats.seconds / 1e9,
                            "images_per_device": list(stats.images_per_device[:n_dev])}
                    print(json.dumps(line), flush=True)
                    lines.append(line)
            finally:
                batch.close()
        # .jpg files through gidhost_batch_decode on every device
        prog, ri, per_call = bench.FILE_FORM[wl]
        distinct = bench.make_files(wl, 2, 77)
        for n_dev in counts:
            n = 2 * per_call * n_dev
            files = [distinct[i % 2] for i in range(n)]
            out = []
            for _ in range(n):
                a, addr = capi.pinned_empty(3 * w * h)
                keep.append(addr)
                out.append(a.reshape(h, w, 3))
            threads = min(os.cpu_count() or 1, 4 * n_dev)
            with host_api.HostBatch(devices=tuple(range(n_dev)), threads=threads) as hb:
                for _ in range(2):
                    hb.decode(files, out)
                _, st, stats = hb.decode(files, out)
            assert not any(st), st
            line = {"tool": "batch_bench", "driver": "gidhost_batch_decode", "workload": wl, "devices": n_dev, "images": n,
                    "threads": int(stats.threads), "MP/s": w * h * n / 1e6 / stats.seconds, "seconds": stats.seconds,
                    "d2h_GBps": 3 * w * h * n / stats.seconds / 1e9, "progressive": prog, "restart_interval": ri}
            print(json.dumps(line), flush=True)
            lines.append(line)
            del out
        for addr in keep:
            capi.pinned_free(addr)


if __name__ == "__main__":
    main()
