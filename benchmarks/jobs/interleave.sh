#!/bin/bash
mkdir -p gpurun_out
timeout 900 python benchmarks/interleave_sweep.py --check --configs serial4,div1,div2 --envs 32768,65536,262144 --groups 1,2,4,8 > gpurun_out/r2_interleave_sweep.jsonl 2> gpurun_out/r2_interleave_sweep.err
cut -c1-260 gpurun_out/r2_interleave_sweep.jsonl
tail -5 gpurun_out/r2_interleave_sweep.err
