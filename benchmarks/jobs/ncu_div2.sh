#!/bin/bash
# ncu --set full of the div2 step kernel at 262144 envs: lane mapping vs cell mapping (one tile per CTA, IMX_PIPE=0)
mkdir -p gpurun_out
for cells in 0 1; do
  IMX_CELLS=$cells IMX_PIPE=0 python benchmarks/one_step.py --config div2 --envs 262144 > gpurun_out/one_step_$cells.log 2>&1 || exit 1
  IMX_CELLS=$cells IMX_PIPE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel_tma -s 1 -c 1 -f -o gpurun_out/r2_div2_cells$cells \
      python benchmarks/one_step.py --config div2 --envs 262144 > gpurun_out/ncu_div2_$cells.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
