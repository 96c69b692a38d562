python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/t3_pytest.log; cat gpurun_out/t3_pytest.log
tools/bench_quick.sh t3 100 ""
tools/bench_quick.sh t3d3 30 "--e2e-depth 3" cfg2 cfg3
tools/bench_quick.sh t3pl 30 "--e2e-input planes" cfg2 cfg3
