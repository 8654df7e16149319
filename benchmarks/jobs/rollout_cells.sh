#!/bin/bash
# parity of the cell-mapped kernels + critic rows, then the fused rollouts with the cell mapping off / on, then the bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cells.py tests/test_cc_observer.py tests/test_gpu_canary.py -x -q 2>&1 | tail -25 > gpurun_out/r2_t4_tests.log
tail -3 gpurun_out/r2_t4_tests.log
for cells in 0 1; do
  IMX_ROLLOUT_CELLS=$cells timeout 300 python benchmarks/bench_configs.py --only rollout > gpurun_out/r2_rollouts_cells$cells.jsonl 2> gpurun_out/r2_rollouts_cells$cells.err
done
timeout 900 python bench.py --steps 20 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
tail -c 600 gpurun_out/r2_bench_a.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
