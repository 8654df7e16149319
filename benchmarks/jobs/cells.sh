#!/bin/bash
# GPU job: parity of the cell-mapped / pipelined kernels, the canaries, then cells on/off sweeps of the divergent step kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_cells.py tests/test_gpu_rollout.py tests/test_gpu_dfo_div.py tests/test_gpu_pipe_kernel.py tests/test_gpu_parity_div.py tests/test_gpu_step_many.py tests/test_gpu_full_size.py tests/test_cc_observer.py tests/test_gpu_canary.py -x -q 2>&1 | tail -25 > gpurun_out/r2_t3_cells_tests.log
tail -4 gpurun_out/r2_t3_cells_tests.log
for cells in 0 1; do
  IMX_CELLS=$cells timeout 600 python benchmarks/pipe_sweep.py --configs div1,div2 --envs 32768,65536,262144,1048576 --threads 128 --stages 4 --ctas 2,3,4,6 > gpurun_out/r2_cells_sweep_$cells.jsonl 2> gpurun_out/r2_cells_sweep_$cells.err
done
tail -2 gpurun_out/r2_cells_sweep_1.err
