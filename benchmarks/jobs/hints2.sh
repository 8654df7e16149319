#!/bin/bash
mkdir -p gpurun_out
for h in 0 2 1; do
  echo "# IMX_L2_HINTS=$h"
  for n in 32768 65536 131072; do IMX_L2_HINTS=$h timeout 200 python benchmarks/bookkeeping.py --config serial4 --envs $n 2>/dev/null | cut -c1-400; done
  IMX_L2_HINTS=$h timeout 200 python benchmarks/bookkeeping.py --config div1 --envs 32768 2>/dev/null | cut -c1-400
done > gpurun_out/r2_l2_hints_small_ab.txt 2>&1
cat gpurun_out/r2_l2_hints_small_ab.txt
