python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/t5_pytest.log; cat gpurun_out/t5_pytest.log
tools/bench_quick.sh t5 100 ""
