#!/bin/bash
# bench lines only (no tests): tools/bench_quick.sh tag steps "extra bench args" [workloads...]
tag=${1:-q}; steps=${2:-50}; extra=${3:-}; shift; shift; shift
wl=${@:-cfg2 cfg3 cfg4 cfg5}
mkdir -p gpurun_out; : > gpurun_out/${tag}_bench.jsonl
for w in $wl; do
  python bench.py --workload $w --steps $steps --warmup 5 --no-cpu $extra >> gpurun_out/${tag}_bench.jsonl 2> gpurun_out/${tag}_bench_$w.err || tail -n 5 gpurun_out/${tag}_bench_$w.err
done
python - <<PY
import json
for l in open("gpurun_out/${tag}_bench.jsonl"):
    d=json.loads(l)
    print("${tag}", d["config"]["workload"][:4], "value %.0f MP/s" % d["value"], "ms/step %.4f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "e2e %.0f" % d["e2e"]["value"], "files %.0f (%d thr)" % (d["files"]["value"], d["files"]["threads"]) if d.get("files") else "", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
