#!/bin/bash
# Usual GPU-box sequence: parity tests, then one bench line per workload.
# usage: tools/gpu_check.sh [tag] [steps]   (outputs under gpurun_out/)
tag=${1:-run}; steps=${2:-50}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${tag}_pytest.log
cat gpurun_out/${tag}_pytest.log
: > gpurun_out/${tag}_bench.jsonl
for w in cfg2 cfg3 cfg4 cfg5; do
  python bench.py --workload $w --steps $steps --warmup 5 --no-cpu >> gpurun_out/${tag}_bench.jsonl 2> gpurun_out/${tag}_bench_$w.err || tail -n 5 gpurun_out/${tag}_bench_$w.err
done
python - <<PY
import json
for l in open("gpurun_out/${tag}_bench.jsonl"):
    d=json.loads(l)
    print(d["config"]["workload"][:4], "value %.0f MP/s" % d["value"], "ms/step %.4f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "e2e %.0f" % d["e2e"]["value"], d["clocks"])
PY
