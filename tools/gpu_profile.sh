#!/bin/bash
# Full evidence run for profiles/: parity tests, the default bench line (with cpu_baseline),
# the reference arm, one bench line per workload, the ncu launch list of the default command
# and one `ncu --set full` capture of the dominant kernel per workload.
# usage: tools/gpu_profile.sh tag [workloads for the full captures]     (outputs under gpurun_out/)
tag=${1:-prof}; shift
wl=${@:-cfg2 cfg3 cfg4 cfg5}
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $out/${tag}_pytest.log; cat $out/${tag}_pytest.log
python bench.py > $out/${tag}_default.json 2> $out/${tag}_default.err || tail -5 $out/${tag}_default.err
python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_reference.json 2> $out/${tag}_reference.err || tail -5 $out/${tag}_reference.err
tools/bench_quick.sh ${tag} 100 "" cfg1 cfg2 cfg3 cfg4 cfg5
# launch list of the default command (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu > $out/${tag}_launches.log 2>&1
for w in $wl; do
  ncu --set full --clock-control none --import-source on -k regex:recon -s 6 -c 1 -f -o $out/${tag}_full_$w \
      python bench.py --workload $w --steps 4 --warmup 3 --no-cpu --streams 1 > $out/${tag}_full_$w.log 2>&1 || tail -3 $out/${tag}_full_$w.log
done
# the .jpg-files leg: host-thread scaling, and one full capture of each device Huffman kernel
python tools/files_scaling.py cfg3 cfg2 cfg4 cfg1 cfg5 2>&1 | grep -v "^$" > $out/${tag}_files_thread_scaling.txt; cat $out/${tag}_files_thread_scaling.txt
ncu --set full --clock-control none --import-source on -k regex:huffman -c 4 -f -o $out/${tag}_full_huffman \
    python tools/files_once.py cfg3 1 1 > $out/${tag}_full_huffman.log 2>&1 || tail -3 $out/${tag}_full_huffman.log
ls -la $out | grep ${tag}_ | head -40
