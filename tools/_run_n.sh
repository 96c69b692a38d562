#!/bin/bash
# multi-GPU evidence: tools/_run_n.sh N tag
N=$1; tag=$2; out=gpurun_out
nvidia-smi topo -m > $out/${tag}_topo.txt 2>&1
for w in cfg2 cfg3; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 20 --warmup 3 --workload $w > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.err || tail -5 $out/${tag}_bench_$w.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --impl reference --gpus $N --steps 3 --warmup 1 --no-files > $out/${tag}_reference.json 2> $out/${tag}_reference.err
python -m pytest tests/test_gpu_parity.py -q -k "batch_driver" 2>&1 | tail -3 > $out/${tag}_pytest_batch.log; cat $out/${tag}_pytest_batch.log
python tools/batch_bench.py cfg3 cfg2 > $out/${tag}_batch_bench.jsonl 2> $out/${tag}_batch_bench.err || tail -5 $out/${tag}_batch_bench.err
python - <<PY
import json
for w in ("cfg2","cfg3"):
    try:
        d=json.load(open("$out/${tag}_bench_%s.json"%w))
        print(w, "n", d["n_gpus"], "value %.0f"%d["value"], "frac %.3f"%d["roofline"]["frac"], "e2e %.0f"%d["e2e"]["value"], "link d2h %.1f GB/s"%d["e2e"]["link_peak"]["d2h_GBps"], "e2e/link %.2f"%d["e2e"]["frac_of_link_d2h_peak"], "files %.0f"%d["files"]["value"], d["method"]["host_numa"])
    except Exception as e: print(w, "failed", e)
for l in open("$out/${tag}_batch_bench.jsonl"):
    d=json.loads(l); print(d["driver"], d["workload"], d["devices"], "%.0f MP/s"%d["MP/s"], "d2h %.1f GB/s"%d["d2h_GBps"], d.get("images_per_device"))
PY
