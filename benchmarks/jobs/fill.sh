#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipe_kernel.py tests/test_gpu_step_et.py tests/test_gpu_full_size.py -x -q 2>&1 | tail -2
( echo "# first ring fill dealt over all warps"
  timeout 300 python benchmarks/n_sweep.py 2>/dev/null | head -10
  timeout 300 python benchmarks/interleave_sweep.py --configs div1,div2,serial8,serial2 --envs 32768,65536,262144 --groups 1 2>/dev/null | cut -c1-130 ) > gpurun_out/r2_fill_all_warps.txt 2>&1
cat gpurun_out/r2_fill_all_warps.txt
