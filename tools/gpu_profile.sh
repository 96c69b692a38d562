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
# (gpurun brings back at most 64 MiB: the reports -- 20 MB each with the source -- are condensed on the box and removed)
condense() {  # condense <report without extension>
  python tools/ncu_summary.py $1.ncu-rep >> $out/${tag}_ncu_summary.md 2>> $out/${tag}_ncu_summary.err
  ncu -i $1.ncu-rep --page source --csv 2> /dev/null > $1_source.csv && python tools/ncu_opmix.py $1_source.csv > $1_opmix.txt 2>&1
  python tools/ncu_hotspots.py $1_source.csv 25 > $1_hotspots.txt 2>&1
  gzip -f $1_source.csv
  rm -f $1.ncu-rep
}
: > $out/${tag}_ncu_summary.md
for w in $wl; do
  ncu --set full --clock-control none --import-source on -k regex:recon -s 6 -c 1 -f -o $out/${tag}_full_$w \
      python bench.py --workload $w --steps 4 --warmup 3 --no-cpu --no-files --pdl 0 > $out/${tag}_full_$w.log 2>&1 || tail -3 $out/${tag}_full_$w.log
  condense $out/${tag}_full_$w
done
# the .jpg-files leg: host-thread scaling, and one full capture of each device Huffman kernel
python tools/files_scaling.py cfg3 cfg2 cfg4 cfg1 cfg5 2>&1 | grep -v "^$" > $out/${tag}_files_thread_scaling.txt; cat $out/${tag}_files_thread_scaling.txt
ncu --set full --clock-control none --import-source on -k regex:huffman -c 4 -f -o $out/${tag}_full_huffman \
    python tools/files_once.py cfg3 1 1 > $out/${tag}_full_huffman.log 2>&1 || tail -3 $out/${tag}_full_huffman.log
condense $out/${tag}_full_huffman
ncu --set full --clock-control none --import-source on -k regex:huffman_intervals -c 1 -f -o $out/${tag}_full_huffman_intervals \
    python tools/files_once.py cfg4 1 1 > $out/${tag}_full_huffman_intervals.log 2>&1 || tail -3 $out/${tag}_full_huffman_intervals.log
condense $out/${tag}_full_huffman_intervals
ls -la $out | grep ${tag}_ | head -40
