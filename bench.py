#!/usr/bin/env python
"""bench.py -- JPEG pixel-reconstruction throughput (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg4|cfg5]

One "step" = one pass of the hot path (dequant + IDCT + upsample + colour) over
one batch of synthetic coefficient planes.  Default workload = BASELINE.json
configs[1]: 4096x4096 baseline 4:4:4 frames (one frame per step per GPU).

Printed JSON (one line, rank 0):
  value        whole-job MP/s with the planes already resident in HBM (CUDA events,
               max over ranks)
  e2e          the same metric through the C-ABI job API with HOST buffers: pinned
               H2D + kernel + D2H inside the timed region
  roofline     algorithmic bytes / kernel time against the measured HBM peak
  cpu_baseline the oracle's reconstruction (GID restated in C; GNAT is not
               available, so the Ada benchmark cannot be built) on the host cores
  files        (beside the path) the same metric from .jpg bytes: Huffman decode
               on the host threads, reconstruction on the GPU, pixels back in host memory

`--impl reference` times that CPU restatement alone, with all host threads, on
the same workload definition, and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import os as _os
# read when the CUDA context is created (torch creates it before libgidcuda.so is loaded): as many hardware
# work queues as the device has, so that frames in flight on different streams really run side by side
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "jpeg_reconstruction_throughput"
UNIT = "MP/s"

WORKLOADS = {
    # name: (width, height, sampling, frames per step per GPU, description)
    "cfg1": (512, 512, "420", 64, "batch of 512x512 baseline 4:2:0 frames (configs[0], the reference's CPU-runnable case), 64 per step per GPU"),
    "cfg2": (4096, 4096, "444", 1, "4096x4096 baseline 4:4:4 frame (BASELINE.json configs[1], kernel roofline case)"),
    "cfg3": (1920, 1080, "420", 128, "batch of 1920x1080 baseline 4:2:0 frames (configs[2]), 128 frames per step per GPU"),
    "cfg4": (8192, 8192, "grey", 1, "8192x8192 grey baseline frame (configs[3])"),
    "cfg5": (2048, 2048, "422", 32, "batch of 2048x2048 4:2:2 frames, full coefficient planes after all scans (configs[4]), 32 per step per GPU"),
}


def env_int(name: str, default: int) -> int:
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def make_frames(workload: str, count: int, seed0: int):
    """`count` distinct synthetic frames (quantised coefficient planes) of the workload."""
    from tests.harness import Synth
    w, h, samp, _, _ = WORKLOADS[workload]
    S = Synth()
    frames = []
    for i in range(count):
        img = S.image(seed0 + i, w, h)
        pix = img[:, :, 1].copy() if samp == "grey" else img
        frames.append(S.planes_only(pix, samp))
    return frames


# JPEG-file form of each workload: (progressive, restart interval, files per call)
# Files per call: enough of them that the ramp of a call -- the first frame's way through header, Huffman kernels
# and reconstruction before the device-to-host link has anything to carry, about 1.3 ms for a 4096 x 4096 file --
# is a few per cent of it (cfg2 with 8 files per call: 16.5 GP/s, with 32: 18.1; profiles/r02_files_per_call.txt)
FILE_FORM = {"cfg1": (False, 0, 1024), "cfg2": (False, 0, 32), "cfg3": (False, 0, 256), "cfg4": (False, 64, 12),
             "cfg5": (True, 0, 128)}


def make_files(workload: str, count: int, seed0: int) -> list[bytes]:
    """`count` distinct synthetic .jpg files of the workload (in-run encoder of the test harness)."""
    from tests.harness import Synth
    w, h, samp, _, _ = WORKLOADS[workload]
    prog, ri, _ = FILE_FORM[workload]
    S = Synth()
    out = []
    for i in range(count):
        img = S.image(seed0 + i, w, h)
        pix = img[:, :, 1].copy() if samp == "grey" else img
        out.append(S.encode(pix, samp, progressive=prog, restart_interval=ri)[0])
    return out


def files_leg_ours(workload: str, device: int, threads: int, reps: int, pinned: bool = True, world: int = 1) -> dict:
    """.jpg bytes in host memory -> RGB8 in host memory: entropy decode on `threads` host threads
    (restart-interval-parallel inside a frame when there are fewer files than threads),
    reconstruction on the B200, through gidhost_batch_decode."""
    from gid_b200 import host_api
    w, h, _, _, _ = WORKLOADS[workload]
    prog, ri, per_call = FILE_FORM[workload]
    # many ranks on one host share its memory and its link: from four ranks on, fewer files per call and rank (the
    # pinned image buffers of all ranks together stay what two ranks' are; two ranks saturate the link anyway)
    per_call = max(2, per_call // max(1, min(world, 8) // 2))
    distinct = make_files(workload, 2, 77)
    files = [distinct[i % 2] for i in range(per_call)]
    # the client's image buffers: pinned host memory (the pixels arrive straight from the device)
    # or ordinary memory (one more copy per image on the host)
    from gid_b200 import capi
    addrs = []
    if pinned:
        out = []
        for _ in range(per_call):
            a, addr = capi.pinned_empty(3 * w * h)
            addrs.append(addr)
            out.append(a.reshape(h, w, 3))
    else:
        out = [np.empty((h, w, 3), dtype=np.uint8) for _ in range(per_call)]
    with host_api.HostBatch(devices=(device,), threads=threads) as hb:
        for _ in range(3):  # warm-up: every worker creates its contexts and its three jobs (pinned + device buffers)
            _, st, stats = hb.decode(files, out)
        assert not any(st), st
        secs = 0.0
        for _ in range(reps):
            _, st, stats = hb.decode(files, out)
            secs += stats.seconds
    del out
    for addr in addrs:
        capi.pinned_free(addr)
    return {"value": w * h * per_call * reps / 1e6 / secs, "unit": UNIT, "threads": int(stats.threads),
            "entropy_threads_per_image": int(stats.entropy_threads), "images_per_call": per_call, "calls": reps,
            "jpeg_bytes_per_image": len(distinct[0]), "progressive": prog, "restart_interval": ri,
            "output": "pinned client buffers (no host copy)" if pinned else "pageable client buffers (one host copy)",
            "api": "gidhost_batch_decode: .jpg bytes -> Huffman decode on host threads -> coefficient records -> "
                   "B200 reconstruction -> top-down RGB8 in host memory, three frames in flight per thread"}


def files_leg_reference(workload: str, threads: int, seconds_target: float) -> dict:
    """The oracle's whole decode of the same files (segments, Huffman, reconstruction, per-pixel
    Put_Pixel client: what test/benchmark.adb times) on `threads` host threads."""
    from tests.harness import Oracle
    w, h, _, _, _ = WORKLOADS[workload]
    prog, ri, _ = FILE_FORM[workload]
    files = make_files(workload, 2, 77)
    O = Oracle()
    t0 = time.perf_counter()
    O.decode(files[0], want_planes=False)
    one = time.perf_counter() - t0
    reps = max(1, int(round(seconds_target / max(one, 1e-3))))
    done = [0] * threads

    def work(k: int) -> None:
        for r in range(reps):
            O.decode(files[(k + r) % 2], want_planes=False)
            done[k] += 1

    ts = [threading.Thread(target=work, args=(k,)) for k in range(threads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.perf_counter() - t0
    return {"value": w * h * sum(done) / 1e6 / dt, "unit": UNIT, "threads": threads,
            "sample": f"{reps} file(s) per thread on {threads} thread(s), {dt:.1f} s wall",
            "jpeg_bytes_per_image": len(files[0]), "progressive": prog, "restart_interval": ri,
            "api": "oracle/gidref.c gidref_decode: whole GID decode restated in C (GNAT unavailable)"}


def workload_config(workload: str, world: int) -> dict:
    """The workload as BOTH arms state it (identical keys and values: the driver compares the two)."""
    w, h, samp, per_step, desc = WORKLOADS[workload]
    return {"workload": f"{workload}: {desc}", "sampling": samp, "width": w, "height": h,
            "frames_per_step_per_gpu": per_step,
            "parallelism": f"images sharded over {world} GPU(s), no collective"}


def link_peak(dev, world: int) -> dict:
    """What the host link of this box delivers: raw pinned cudaMemcpyAsync loops (torch copy_), 256 MiB per
    copy, one loop per GPU -- under torch.distributed every rank runs its loop at the same time and the rates are
    summed.  The ceiling the end-to-end leg is held against (3 bytes per pixel leave over this link)."""
    import torch
    import torch.distributed as dist
    from gid_b200 import sharding
    n = 256 << 20
    hbuf = torch.empty(n, dtype=torch.uint8).pin_memory()
    hbuf2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    dbuf = torch.empty(n, dtype=torch.uint8, device=dev)
    dbuf2 = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    out = {}
    for name in ("d2h", "h2d", "both"):
        reps = 8
        def go():
            for _ in range(reps):
                if name in ("d2h", "both"):
                    with torch.cuda.stream(s1):
                        hbuf.copy_(dbuf, non_blocking=True)
                if name in ("h2d", "both"):
                    with torch.cuda.stream(s2):
                        dbuf2.copy_(hbuf2, non_blocking=True)
        go()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        go()
        torch.cuda.synchronize()
        dt = sharding.max_over_ranks(time.perf_counter() - t0, dev)
        out[name + "_GBps"] = world * reps * n * (2 if name == "both" else 1) / dt / 1e9
    out["method"] = f"{world} concurrent loop(s) of 8 x 256 MiB pinned cudaMemcpyAsync per direction, wall clock, max over ranks"
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        self.rows: list[str] = []
        self.proc = None
        self.index = index

    def start(self) -> None:
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self) -> None:
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[4 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak() -> tuple[float, str]:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


def ncu_traffic(workload: str):
    """DRAM bytes per launch from the committed ncu capture, if there is one."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[workload]
    except Exception:
        return None


# --------------------------------------------------------------------------- CPU reference arm


def cpu_reconstruct_rate(frames, seconds_target: float, threads: int) -> tuple[float, str]:
    """Oracle reconstruction (dequant + IDCT + upsample + colour into a top-down RGB
    buffer, the client of test/benchmark.adb) on `threads` host threads; each thread
    runs its own copy of the work, as `one process per host core` would."""
    from tests.harness import Oracle
    O = Oracle()
    fr = frames[0]
    mp = fr.width * fr.height / 1e6
    t0 = time.perf_counter()
    O.reconstruct(fr)
    one = time.perf_counter() - t0
    reps = max(1, int(round(seconds_target / max(one, 1e-3))))
    done = [0] * threads

    def work(k: int) -> None:
        for _ in range(reps):
            O.reconstruct(frames[k % len(frames)])
            done[k] += 1

    ts = [threading.Thread(target=work, args=(k,)) for k in range(threads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.perf_counter() - t0
    total_mp = mp * sum(done)
    sample = f"{reps} x {fr.width}x{fr.height} frame(s) per thread on {threads} thread(s), {dt:.1f} s wall"
    return total_mp / dt, sample


def cpu_worker(core: int, path: str, reps: int) -> None:
    """One pinned process of cpu_reconstruct_rate_processes: `reps` reconstructions of the pickled frame."""
    import pickle
    try:
        os.sched_setaffinity(0, {core})
    except OSError:
        pass
    from tests.harness import Oracle
    O = Oracle()
    fr = pickle.load(open(path, "rb"))
    O.reconstruct(fr)  # warm
    print("ready", flush=True)
    sys.stdin.readline()  # start signal
    t0 = time.perf_counter()
    for _ in range(reps):
        O.reconstruct(fr)
    print(json.dumps({"reps": reps, "seconds": time.perf_counter() - t0}), flush=True)


def cpu_reconstruct_rate_processes(frames, seconds_target: float, cores: int) -> tuple[float, str]:
    """The same reconstruction as ONE PROCESS PER HOST CORE, each pinned to its core (sched_setaffinity),
    as SURVEY.md 8(d) plans the reference's CPU run (test/benchmark.adb, one process per core)."""
    import pickle
    import tempfile
    from tests.harness import Oracle
    fr = frames[0]
    O = Oracle()
    t0 = time.perf_counter()
    O.reconstruct(fr)
    one = time.perf_counter() - t0
    reps = max(1, int(round(seconds_target / max(one, 1e-3))))
    allowed = sorted(os.sched_getaffinity(0))[:cores]
    with tempfile.NamedTemporaryFile(suffix=".pkl", delete=False) as tf:
        pickle.dump(fr, tf)
        path = tf.name
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--cpu-worker", str(c), path, str(reps)],
                              stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
             for c in allowed]
    try:
        for pr in procs:
            assert pr.stdout.readline().strip() == "ready"
        t0 = time.perf_counter()
        for pr in procs:
            pr.stdin.write("go\n")
            pr.stdin.flush()
        done = [json.loads(pr.stdout.readline()) for pr in procs]
        dt = time.perf_counter() - t0
    finally:
        for pr in procs:
            try:
                pr.wait(timeout=10)
            except subprocess.TimeoutExpired:
                pr.kill()
                pr.wait()
        os.unlink(path)
    mp = fr.width * fr.height / 1e6
    return mp * sum(d["reps"] for d in done) / dt, \
        f"{reps} x {fr.width}x{fr.height} frame(s) per process, {len(procs)} pinned process(es), {dt:.1f} s wall"


def run_reference(args) -> None:
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    w, h, samp, per_step, desc = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    frames = make_frames(args.workload, 1, 0)
    # each "step" = a bounded sample of the workload; keep the whole run within a few minutes
    per_step_seconds = max(0.5, min(4.0, 60.0 / max(1, args.steps + args.warmup)))
    if os.environ.get("GIDBENCH_REFERENCE_SECONDS"):  # tests: a shorter sample
        per_step_seconds = float(os.environ["GIDBENCH_REFERENCE_SECONDS"])
    rate_fn = cpu_reconstruct_rate if os.environ.get("GIDBENCH_REFERENCE_THREADS") else cpu_reconstruct_rate_processes
    for _ in range(args.warmup):
        rate_fn(frames, per_step_seconds / 4, cores)
    rates, sample = [], ""
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r, sample = rate_fn(frames, per_step_seconds, cores)
        rates.append(r)
    dt = time.perf_counter() - t0
    value = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
        "data": "synthetic", "gpu_launches": 0,
        "config": workload_config(args.workload, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "per step: " + sample,
                         "note": "GID's reconstruction restated in C (oracle/gidref.c, gcc -O2), one pinned process per host core; "
                                 "GNAT is not installed so test/benchmark.adb cannot be built"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if not args.no_files:
        line["files"] = files_leg_reference(args.workload, cores, float(os.environ.get("GIDBENCH_REFERENCE_SECONDS", 3.0)))
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- our arm


def bind_to_gpu_numa_node(local: int) -> str:
    """One rank per GPU: run this rank's host threads, and so first-touch its pinned buffers, on the
    NUMA node the GPU hangs off (eight ranks streaming pixels into one host's memory are bound by the
    host side; memory on the far socket halves what a link delivers).  Best effort: returns what was done."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:       # nvml prints an 8-digit PCI domain, sysfs a 4-digit one
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return "no NUMA information for the GPU"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return f"NUMA node {node} has none of this process's CPUs"
        os.sched_setaffinity(0, cpus)
        return f"bound to NUMA node {node} ({len(cpus)} CPUs)"
    except Exception as e:  # no nvml, no sysfs entry, a container without the rights: carry on unbound
        return f"not bound ({type(e).__name__})"


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    from gid_b200 import GidCuda, build, capi, sharding
    from tests.harness import frame_desc

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    # stdout carries ONE line, the JSON: whatever libraries print there meanwhile (NCCL's version banner
    # goes to stdout) is sent to stderr, until the line itself is printed
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the reconstruction path has no CPU fallback)")
    if not os.path.exists(capi.library_path()):
        build.build_cuda()
    numa = bind_to_gpu_numa_node(local) if world > 1 else "single rank: not bound"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    w, h, samp, per_step, desc = WORKLOADS[args.workload]
    ctx = GidCuda(local)
    capi.set_option("plan_pdl", args.pdl)
    if args.turn_tile_rows:
        capi.set_option("turn_tile_rows", args.turn_tile_rows)   # (tuning: tile shape of frames turned by a quarter)

    # ---- synthetic inputs: a ring of distinct step-batches so that consecutive steps never
    # reuse L2-resident data (every step-batch is also larger than the 126 MB L2) ---------
    distinct = 2 if per_step > 1 else 4          # distinct source frames
    frames = make_frames(args.workload, distinct, 1000 * rank)
    alg_bytes_frame = frames[0].algorithmic_bytes() + (w * h if args.rgba else 0)  # (R,G,B,A: a byte more per pixel)
    ring = max(2, min(4, int(2.5e9 // max(1, alg_bytes_frame * per_step))))
    ring -= ring % 2  # a multiple of the number of launch streams
    rot_flag = {0: 0, 90: capi.FLAG_ROTATE_90, 180: capi.FLAG_ROTATE_180, 270: capi.FLAG_ROTATE_270}[args.rotate]
    if args.rgba:
        rot_flag |= capi.OUT_RGBA8
    descs_one = [frame_desc(f, rot_flag) for f in frames]
    stride, nbytes = ctx.output_geometry(descs_one[0])
    keep, plans = [], []
    for r in range(ring):
        descs, d_planes, d_pixels = [], [], []
        for i in range(per_step):
            fr = frames[(r + i) % distinct]
            descs.append(descs_one[(r + i) % distinct])
            for c in range(3):
                if c < fr.ncomp:
                    t = torch.from_numpy(fr.planes[c].reshape(-1)).to(dev)
                    keep.append(t)
                    d_planes.append(t.data_ptr())
                else:
                    d_planes.append(0)
            out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            keep.append(out)
            d_pixels.append(out.data_ptr())
        plans.append(ctx.plan_create(descs, d_planes, d_pixels))
    launches_per_step = ctx.plan_launch_count(plans[0])
    torch.cuda.synchronize()
    # an explicit stream: the events below and the kernels must be on the SAME stream
    # (torch.cuda.Event only sees torch's current stream; a NULL handle would mean the
    # context's own stream in the C ABI)
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0
    # Steps are independent frames, so like the job / batch paths of the library the
    # launches alternate between two streams: the tail of one launch (warps finishing
    # their last tile) overlaps the ramp of the next.  Step s always runs plan s % ring
    # on stream s % nstreams (ring is a multiple of nstreams), so a plan never runs
    # concurrently with itself.  Timing: events on the first stream, which forks to and
    # joins the second one inside the timed region.
    nstreams = max(1, args.streams)
    extra = [torch.cuda.Stream(device=dev) for _ in range(nstreams - 1)]
    streams = [tstream] + extra

    def launch_steps(n: int, use: int = 0) -> None:
        """n steps over the first `use` launch streams (0 = all of them)."""
        k = use or nstreams
        fork = torch.cuda.Event()
        fork.record(tstream)
        for st in extra[:k - 1]:
            st.wait_event(fork)
        for s_ in range(n):
            ctx.plan_launch(plans[s_ % ring], streams[s_ % k].cuda_stream)
        for st in extra[:k - 1]:
            join = torch.cuda.Event()
            join.record(st)
            tstream.wait_event(join)

    def barrier() -> None:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_regions(use: int, min_total_ms: float, max_regions: int = 400) -> list[float]:
        """Regions of EXACTLY args.steps steps each, every one bracketed by barrier + synchronize, CUDA events
        on the launching stream, max over ranks; repeated until the regions add up to min_total_ms of device
        time (a single region of a 40 us kernel at --steps 20 is under a millisecond: too short a sample)."""
        out: list[float] = []
        while True:
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(tstream)
            launch_steps(args.steps, use)
            b.record(tstream)
            barrier()
            out.append(sharding.max_over_ranks(a.elapsed_time(b), dev))
            if len(out) >= max_regions or (len(out) >= 3 and sum(out) >= min_total_ms):
                return out

    def summary(regions: list[float]) -> dict:
        per = sorted(r / args.steps for r in regions)
        return {"ms_per_launch_median": float(np.median(per)), "ms_per_launch_min": per[0], "ms_per_launch_max": per[-1],
                "regions": len(per), "timed_region_ms": float(np.median(regions)), "timed_total_ms": float(sum(regions))}

    # ---- device-resident timing --------------------------------------------------------
    # Primary figure: ONE stream, launches as the library issues them (programmatic dependent launches:
    # a launch may start filling SMs while its predecessor's last tiles drain; completion order is kept).
    launch_steps(args.warmup)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    regions = timed_regions(nstreams, 50.0)
    prim = summary(regions)
    ms_step = prim["ms_per_launch_median"]
    # Secondary: the same launches with the overlap switched off (every launch waits for its predecessor to
    # leave the device) -- the duration of one launch alone, ramp and drain included
    capi.set_option("plan_pdl", 0)
    launch_steps(min(args.warmup, 5), 1)
    serial = summary(timed_regions(1, 25.0, 20))
    capi.set_option("plan_pdl", args.pdl)
    # keep the clocks under observation for at least ~0.5 s of the same work
    t_end = time.perf_counter() + 0.5
    while time.perf_counter() < t_end:
        launch_steps(64)
        torch.cuda.synchronize()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    mp_step = w * h * per_step / 1e6
    value = sharding.whole_job_rate(mp_step, world, ms_step * 1e-3)

    # ---- end to end through the job API: host planes -> host pixels -----------------------
    L = ctx.lib
    depth = max(2, args.e2e_depth)
    jobs = []
    h2d_frame = 0  # bytes that cross the link per frame
    for k in range(depth):
        job = ctypes.c_void_p()
        fr = frames[k % distinct]
        d = descs_one[k % distinct]
        ctx._check(L.gidcuda_job_begin(ctx.ctx, ctypes.byref(d), ctypes.byref(job)), ctx.ctx)
        moved = 0
        for c in range(fr.ncomp):
            if args.e2e_input == "records":
                # what the entropy decoder emits: one record per non-zero coefficient
                try:
                    moved += ctx.job_fill_records(job, d, c, fr.planes[c])
                    continue
                except OverflowError:
                    pass  # too dense for the record buffer: this component travels as a plane
            p = L.gidcuda_job_plane(job, c, None, None)
            ctypes.memmove(p, fr.planes[c].ctypes.data, fr.planes[c].nbytes)
            # what the entropy decoder knows for free: the last block row it stored into
            ctx._check(L.gidcuda_job_hint_rows(job, c, capi.rows_used(fr.planes[c])), ctx.ctx)
            # hinted planes move the upper half of every block
            pl = fr.planes[c]
            moved += pl.nbytes // 2 if capi.rows_used(pl) <= 4 and pl.nbytes >= (8 << 20) else pl.nbytes
        if k == 0:
            h2d_frame = moved
        jobs.append(job)
    e2e_frames = max(1, min(per_step, 8))  # frames per e2e step (bounded: host-link limited)
    pix, st = ctypes.c_void_p(), ctypes.c_size_t()
    checksum = 0

    def e2e_steps(n: int) -> None:
        nonlocal checksum
        total = n * e2e_frames
        for i in range(total + depth - 1):
            if i < total:
                ctx._check(L.gidcuda_job_submit(jobs[i % depth]), ctx.ctx)
            if i >= depth - 1:
                ctx._check(L.gidcuda_job_wait(jobs[(i - (depth - 1)) % depth], ctypes.byref(pix), ctypes.byref(st)), ctx.ctx)
                # device->host read of the step's result: first and last byte of the host pixel buffer
                row = (ctypes.c_uint8 * 1).from_address(pix.value)
                checksum += row[0]

    e2e_steps(max(1, min(args.warmup, 3)))
    e2e_regions: list[float] = []
    while True:
        barrier()
        t0 = time.perf_counter()
        e2e_steps(args.steps)
        torch.cuda.synchronize()
        e2e_regions.append(sharding.max_over_ranks(time.perf_counter() - t0, dev))
        if len(e2e_regions) >= 400 or (len(e2e_regions) >= 3 and sum(e2e_regions) >= 1.0):
            break
    e2e_s = float(np.median(e2e_regions))
    e2e_value = sharding.whole_job_rate((w * h * e2e_frames / 1e6) * args.steps, world, e2e_s)
    link = link_peak(dev, world)
    coef_bytes = h2d_frame
    for job in jobs:
        L.gidcuda_job_release(job)

    # ---- from .jpg files: host entropy decode on this rank's share of the cores + the path above ----
    files = None
    if not args.no_files:
        # this rank's share of the cores (where the device decodes the scans, 2-4 workers already saturate
        # the device-to-host link: profiles/r01f_files_thread_scaling.txt)
        share = max(1, (os.cpu_count() or 1) // world)
        files = files_leg_ours(args.workload, local, share, 3, world=world)
        secs = w * h * files["images_per_call"] * files["calls"] / 1e6 / files["value"]
        files["value"] = sharding.whole_job_rate(w * h * files["images_per_call"] * files["calls"] / 1e6, world,
                                                 sharding.max_over_ranks(secs, dev))

    # ---- CPU baseline (rank 0, N = 1 only) ---------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        v, sample = cpu_reconstruct_rate_processes(frames[:1], 3.0, cores)
        vt, sample_t = cpu_reconstruct_rate(frames[:1], 3.0, cores)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
               "threads_of_one_process": {"value": vt, "sample": sample_t},
               "note": "oracle/gidref.c reconstruction only (GID restated in C, gcc -O2), one pinned process per "
                       "host core; GNAT unavailable"}

    if rank == 0:
        peak, peak_src = measured_peak()
        alg = alg_bytes_frame * per_step
        achieved = alg / (ms_step * 1e-3) / 1e9
        ach_serial = alg / (serial["ms_per_launch_median"] * 1e-3) / 1e9
        e2e_bytes_s = e2e_value * 1e6 * (nbytes / (w * h))   # pixel bytes leaving over the link per second
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": workload_config(args.workload, world),
            "method": {"rotate": args.rotate, "rgba": bool(args.rgba), "launch_streams": nstreams, "programmatic_dependent_launch": bool(args.pdl), "host_numa": numa,
                       "l2": f"inputs larger than L2: ring of {ring} step-batches of {alg / 1e6:.0f} MB each",
                       "timing": "CUDA events on the launching stream around EXACTLY --steps launches, barrier + synchronize "
                                 "on both sides, max over ranks; regions repeated until >= 50 ms of device time; the median "
                                 "region is reported", **prim},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": coef_bytes * e2e_frames,
                    "d2h_bytes_per_step": nbytes * e2e_frames, "frames_per_step": e2e_frames,
                    "input": args.e2e_input, "regions": len(e2e_regions), "timed_total_s": float(sum(e2e_regions)),
                    "region_s_min": float(min(e2e_regions)), "region_s_max": float(max(e2e_regions)),
                    "link_peak": link, "d2h_GBps": e2e_bytes_s / 1e9,
                    "frac_of_link_d2h_peak": e2e_bytes_s / 1e9 / link["d2h_GBps"],
                    "api": ("gidcuda_job_submit/wait, pinned host coefficient records (one per non-zero coefficient, "
                            "gidcuda_job_records) -> expand + reconstruction kernels -> pinned host pixels, %d jobs in flight" % depth
                            if args.e2e_input == "records" else
                            "gidcuda_job_submit/wait, pinned host planes -> pinned host pixels, %d jobs in flight; " % depth +
                            "planes with empty block rows 4..7 (gidcuda_job_hint_rows) travel as pitched copies")},
            "gpu_launches": launches_per_step * args.steps,
            # the kernel's own figure: algorithmic bytes of ONE launch / the average duration of a launch in ONE
            # stream (launched as the library launches: programmatic dependent launches).  `serialized` = the
            # same with the overlap of consecutive launches switched off.
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(args.workload), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg,
                         "ms_per_launch": ms_step, "ms_per_launch_min": prim["ms_per_launch_min"],
                         "ms_per_launch_max": prim["ms_per_launch_max"],
                         "serialized": {"ms_per_launch": serial["ms_per_launch_median"], "achieved": ach_serial,
                                        "frac": ach_serial / peak, "regions": serial["regions"],
                                        "note": "plan_pdl = 0: every launch waits for its predecessor to leave the device"}},
            "cpu_baseline": cpu,
        }
        if files is not None:
            line["files"] = files
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    for p in plans:
        ctx.plan_destroy(p)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    if len(sys.argv) >= 5 and sys.argv[1] == "--cpu-worker":
        cpu_worker(int(sys.argv[2]), sys.argv[3], int(sys.argv[4]))
        return
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="cfg2")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-files", action="store_true", help="skip the .jpg-files leg (host entropy decode + the path)")
    ap.add_argument("--e2e-input", choices=["records", "planes"], default="records",
                    help="host-side form of the coefficients in the end-to-end leg: sparse records (what an entropy "
                         "decoder emits) or dense int16 planes")
    ap.add_argument("--e2e-depth", type=int, default=3, help="jobs in flight in the end-to-end leg")
    ap.add_argument("--streams", type=int, default=1, choices=[1, 2, 4],
                    help="launch streams for the device-resident leg (default 1: the kernel's own figure)")
    ap.add_argument("--rotate", type=int, default=0, choices=[0, 90, 180, 270],
                    help="apply a display orientation while writing the pixels (GIDCUDA_FLAG_ROTATE_*; not the default workload)")
    ap.add_argument("--turn-tile-rows", type=int, default=0, choices=[0, 4, 8, 16, 32],
                    help="tuning: MCU rows per tile of frames turned by a quarter (0 = the library chooses)")
    ap.add_argument("--rgba", action="store_true", help="R,G,B,A pixels (GIDCUDA_OUT_RGBA8) instead of the benchmark's R,G,B")
    ap.add_argument("--pdl", type=int, default=1, choices=[0, 1],
                    help="plan launches as programmatic dependent launches (library default 1); 0 = every launch "
                         "waits for its predecessor to leave the device")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
