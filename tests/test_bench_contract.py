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
    assert r["single_stream"]["frac"] <= r["frac"] * 1.05
    e = d["e2e"]
    assert e["value"] > 0 and e["unit"] == "MP/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]                      # the host link is inside this one
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and c["unit"] == "MP/s" and c["sample"]
    k = d["clocks"]
    assert k["sm_max_mhz"] and not set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    f = d["files"]
    assert f["value"] > 0 and f["threads"] >= 1 and f["images_per_call"] == 256
