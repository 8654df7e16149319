#!/bin/bash
mkdir -p gpurun_out
for pf in 0 1; do
  echo "# IMX_ACT_PREFETCH=$pf"
  IMX_ACT_PREFETCH=$pf timeout 300 python benchmarks/n_sweep.py 2>/dev/null | head -12
  IMX_ACT_PREFETCH=$pf timeout 300 python benchmarks/interleave_sweep.py --configs div1,div2,serial8 --envs 32768,262144 --groups 1 2>/dev/null | cut -c1-130
done > gpurun_out/r2_act_prefetch_ab.txt 2>&1
cat gpurun_out/r2_act_prefetch_ab.txt
timeout 600 python -m pytest tests/test_gpu_pipe_kernel.py tests/test_gpu_step_many.py tests/test_gpu_canary.py -x -q 2>&1 | tail -2
