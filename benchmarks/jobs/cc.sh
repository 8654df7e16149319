#!/bin/bash
mkdir -p gpurun_out
timeout 900 python benchmarks/cc_sweep.py > gpurun_out/r2_cc_sweep.jsonl 2> gpurun_out/r2_cc_sweep.err
cut -c1-200 gpurun_out/r2_cc_sweep.jsonl; tail -3 gpurun_out/r2_cc_sweep.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel_pipe -s 3 -c 1 -f -o gpurun_out/r2_cc_pipe_serial2 python benchmarks/cc_sweep.py --one > gpurun_out/r2_ncu_cc.log 2>&1
tail -2 gpurun_out/r2_ncu_cc.log
