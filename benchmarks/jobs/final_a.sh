#!/bin/bash
# full GPU test suite, the bench line, the other-config sweeps, smoke
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_t8_full_tests.log
tail -3 gpurun_out/r2_t8_full_tests.log
timeout 900 python bench.py > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
tail -c 500 gpurun_out/r2_bench_b.err
timeout 600 python benchmarks/bench_configs.py > gpurun_out/r2_other_configs.jsonl 2> gpurun_out/r2_other_configs.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
timeout 300 python benchmarks/n_sweep.py > gpurun_out/r2_n_sweep.txt 2> gpurun_out/r2_n_sweep.err
