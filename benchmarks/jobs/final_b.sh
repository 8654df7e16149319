#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_t9_full_tests.log
tail -3 gpurun_out/r2_t9_full_tests.log
# lane utilisation of the env-per-thread div2 step kernel without the producer warp's spin loop (one tile per CTA)
IMX_PIPE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel_tma -s 2 -c 1 -f -o gpurun_out/r2_et_onetile_div2 python benchmarks/one_step.py --config div2 --envs 1048576 --steps 5 > gpurun_out/r2_ncu_et_onetile_div2.log 2>&1
timeout 900 python bench.py > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err
tail -c 300 gpurun_out/r2_bench_c.err
python benchmarks/bench_python_loop.py > gpurun_out/r2_python_loop.jsonl 2> gpurun_out/r2_python_loop.err
