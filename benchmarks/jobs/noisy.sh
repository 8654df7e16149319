#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_pipe_kernel.py tests/test_gpu_parity_serial.py tests/test_gpu_parity_div.py tests/test_gpu_step_et.py tests/test_gpu_canary.py tests/test_gpu_dfo_div.py -x -q 2>&1 | tail -25 > gpurun_out/r2_t10_tests.log
tail -3 gpurun_out/r2_t10_tests.log
# rewards-only replay of a stored plan with the env-per-thread period forced on the serial chain
for et in 0 1; do IMX_STEP_ET=$et timeout 200 python benchmarks/bench_configs.py --only "rewards only" > gpurun_out/r2_noobs_et$et.jsonl 2>&1; done
cat gpurun_out/r2_noobs_et0.jsonl gpurun_out/r2_noobs_et1.jsonl | cut -c1-220
