"""bench.py prints ONE JSON line with the keys the driver reads (our arm, on the GPU)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_bench_line_contract():
    env = dict(os.environ)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "cfg1", "--steps", "20",
                        "--warmup", "3"], capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = lines[0]
    assert d["metric"] == "jpeg_reconstruction_throughput" and d["unit"] == "MP/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 20 and d["warmup"] == 3 and d["scaling"] == "weak"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["dtype"] == "int32" and "workload" in d["config"] and "model" not in d["config"]
    assert d["gpu_launches"] == 20
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0 < r["frac"] < 1 and r["algorithmic_bytes_per_launch"] == 64 * 1572864
    assert r["serialized"]["frac"] <= r["frac"] * 1.05 and r["serialized"]["regions"] >= 3
    assert abs(r["ms_per_launch"] - d["ms_per_step"]) < 1e-12 and r["ms_per_launch_min"] <= r["ms_per_launch"] <= r["ms_per_launch_max"]
    m = d["method"]
    assert m["launch_streams"] == 1 and m["regions"] >= 3 and m["timed_total_ms"] >= 50.0
    assert set(d["config"]) == {"workload", "sampling", "width", "height", "frames_per_step_per_gpu", "parallelism"}
    e = d["e2e"]
    assert e["value"] > 0 and e["unit"] == "MP/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]                      # the host link is inside this one
    assert e["timed_total_s"] >= 1.0 and e["link_peak"]["d2h_GBps"] > 1 and 0 < e["frac_of_link_d2h_peak"] < 1.2
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and c["unit"] == "MP/s" and c["sample"]
    k = d["clocks"]
    assert k["sm_max_mhz"] and not set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    f = d["files"]
    assert f["value"] > 0 and f["threads"] >= 1 and f["images_per_call"] == 1024


def test_reference_arm_line_has_the_same_config_keys():
    """`--impl reference` (CPU, no GPU needed): same metric, unit and config keys as our arm."""
    env = dict(os.environ, GIDBENCH_REFERENCE_SECONDS="0.2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1",
                        "--steps", "1", "--warmup", "0", "--no-files"], capture_output=True, text=True, timeout=600,
                       cwd=ROOT, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = lines[0]
    assert d["impl"] == "reference" and d["metric"] == "jpeg_reconstruction_throughput" and d["unit"] == "MP/s"
    assert set(d["config"]) == {"workload", "sampling", "width", "height", "frames_per_step_per_gpu", "parallelism"}
    assert d["cpu_baseline"]["kind"] == "port" and "pinned process" in d["cpu_baseline"]["sample"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
