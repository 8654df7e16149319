#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_t12_tests.log
tail -3 gpurun_out/r2_t12_tests.log
for p in serial2 serial4; do timeout 300 python benchmarks/cc_sweep.py --decompose --preset $p; done > gpurun_out/r2_cc_decompose.jsonl 2> gpurun_out/r2_cc_decompose.err
cat gpurun_out/r2_cc_decompose.jsonl; tail -3 gpurun_out/r2_cc_decompose.err
