#!/bin/bash
mkdir -p gpurun_out
( echo "# IMX_ACT_PREFETCH=1 (actions + demand rows + table into L1)"
  timeout 300 python benchmarks/n_sweep.py 2>/dev/null | head -10
  timeout 300 python benchmarks/interleave_sweep.py --configs div1,div2,serial8 --envs 32768,262144 --groups 1 2>/dev/null | cut -c1-130 ) > gpurun_out/r2_act_prefetch_ab3.txt 2>&1
cat gpurun_out/r2_act_prefetch_ab3.txt
