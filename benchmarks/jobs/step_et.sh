#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_step_et.py -x -q 2>&1 | tail -25 > gpurun_out/r2_t6_tests.log
tail -3 gpurun_out/r2_t6_tests.log
for thr in 32 64 128; do
  IMX_STEP_ET=1 IMX_STEP_ET_THREADS=$thr timeout 400 python benchmarks/pipe_sweep.py --configs div1,div2 --envs 32768,65536,262144,1048576 --threads 128 --stages 2,3,4 --ctas 2,3,4,6,8 --reps 10 > gpurun_out/r2_step_et_sweep_$thr.jsonl 2> gpurun_out/r2_step_et_sweep_$thr.err
done
tail -2 gpurun_out/r2_step_et_sweep_64.err
