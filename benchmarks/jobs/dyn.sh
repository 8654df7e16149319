#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_pipe_kernel.py tests/test_gpu_step_et.py tests/test_gpu_canary.py tests/test_cc_observer.py tests/test_gpu_cells.py -x -q 2>&1 | tail -25 > gpurun_out/r2_t7_tests.log
tail -3 gpurun_out/r2_t7_tests.log
timeout 600 python benchmarks/pipe_sweep.py --configs serial4 --envs 16384,32768,65536,131072,262144 --threads 64,128 --stages 3,4 --ctas 2,4,6,8 --reps 10 > gpurun_out/r2_dyn_sweep_serial4.jsonl 2> gpurun_out/r2_dyn_sweep_serial4.err
timeout 600 python benchmarks/pipe_sweep.py --configs div1,div2,serial8,serial2 --envs 32768,65536,262144 --threads 128 --stages 3,4 --ctas 2,4,6,8 --reps 10 > gpurun_out/r2_dyn_sweep_others.jsonl 2> gpurun_out/r2_dyn_sweep_others.err
IMX_STEP_ET=1 IMX_STEP_ET_THREADS=32 timeout 400 python benchmarks/pipe_sweep.py --configs div2,div1 --envs 32768,65536,262144,1048576 --threads 128 --stages 2,3,4 --ctas 2,4,6,8 --reps 10 > gpurun_out/r2_dyn_sweep_et32.jsonl 2> gpurun_out/r2_dyn_sweep_et32.err
IMX_STEP_ET=1 IMX_STEP_ET_THREADS=64 timeout 400 python benchmarks/pipe_sweep.py --configs div2 --envs 65536,262144,1048576 --threads 128 --stages 2,3 --ctas 2,4,6 --reps 10 > gpurun_out/r2_dyn_sweep_et64.jsonl 2> gpurun_out/r2_dyn_sweep_et64.err
tail -2 gpurun_out/r2_dyn_sweep_et32.err
