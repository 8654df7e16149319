#!/bin/bash
mkdir -p gpurun_out
for h in 0 1; do
  echo "# IMX_L2_HINTS=$h"
  IMX_L2_HINTS=$h timeout 300 python benchmarks/n_sweep.py 2>/dev/null | head -10
  IMX_L2_HINTS=$h IMX_PIPE=1 timeout 300 python benchmarks/interleave_sweep.py --configs serial4,div1,div2,serial8 --envs 262144,524288,1048576 --groups 1 2>/dev/null | cut -c1-130
  IMX_L2_HINTS=$h timeout 200 python benchmarks/bookkeeping.py --config serial4 --envs 65536 2>/dev/null
  IMX_L2_HINTS=$h timeout 200 python benchmarks/bookkeeping.py --config div2 --envs 262144 2>/dev/null
done > gpurun_out/r2_l2_hints_ab.txt 2>&1
cat gpurun_out/r2_l2_hints_ab.txt
IMX_L2_HINTS=1 timeout 600 python -m pytest tests/test_gpu_pipe_kernel.py tests/test_gpu_full_size.py -x -q 2>&1 | tail -2
