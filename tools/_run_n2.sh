#!/bin/bash
out=gpurun_out
for w in cfg2 cfg5; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 20 --warmup 3 --workload $w > $out/r02c_n2_bench_$w.json 2> $out/r02c_n2_bench_$w.err || tail -5 $out/r02c_n2_bench_$w.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $out/r02c_n2_reference.json 2> $out/r02c_n2_reference.err
python - <<PY
import json
for w in ("cfg2","cfg5"):
    d=json.load(open("$out/r02c_n2_bench_%s.json"%w))
    print(w, "n", d["n_gpus"], "value %.0f"%d["value"], "frac %.3f"%d["roofline"]["frac"], "e2e %.0f"%d["e2e"]["value"], "e2e/link %.2f"%d["e2e"]["frac_of_link_d2h_peak"], "files %.0f"%d["files"]["value"])
r=json.load(open("$out/r02c_n2_reference.json")); print("reference", r["value"], r["config"]==json.load(open("$out/r02c_n2_bench_cfg2.json"))["config"])
PY
