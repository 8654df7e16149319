#!/bin/bash
mkdir -p gpurun_out
for p in serial2 serial4; do timeout 300 python benchmarks/cc_sweep.py --decompose --preset $p; done > gpurun_out/r2_cc_decompose.jsonl 2> gpurun_out/r2_cc_decompose.err
cat gpurun_out/r2_cc_decompose.jsonl; tail -3 gpurun_out/r2_cc_decompose.err
