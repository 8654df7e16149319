#!/bin/bash
# ncu evidence for the final kernels: launch list of the bench command, then --set full captures of the dominant kernels
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu --skip-large --skip-configs > gpurun_out/r2_bench_short.json 2> gpurun_out/r2_bench_short.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 3 --warmup 3 --skip-cpu --skip-large --skip-configs > gpurun_out/r2_ncu_launches.log 2>&1
cap() {  # name, kernel regex, skip, extra env..., -- command
  name=$1; regex=$2; skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o gpurun_out/r2_$name "$@" > gpurun_out/r2_ncu_$name.log 2>&1
}
cap pipe_serial4_65536 step_kernel_pipe 2 python benchmarks/one_step.py --config serial4 --envs 65536 --steps 6
cap tma_serial4_4Mi step_kernel_tma 2 python benchmarks/one_step.py --config serial4 --envs 4194304 --steps 4
cap pipe_div2_262144 step_kernel_pipe 2 python benchmarks/one_step.py --config div2 --envs 262144 --steps 6
cap pipe_div1_262144 step_kernel_pipe 2 python benchmarks/one_step.py --config div1 --envs 262144 --steps 6
cap rollout_et_serial8 rollout_kernel_et 1 python benchmarks/one_step.py --config serial8 --envs 131072 --rollout --steps 3
cap rollout_et_div2 rollout_kernel_et 1 python benchmarks/one_step.py --config div2 --envs 262144 --rollout --steps 3
cap many_div2 step_kernel_tma_many 0 python benchmarks/one_step.py --config div2 --envs 262144 --many --steps 30
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel_pipe -s 3 -c 1 -f -o gpurun_out/r2_cc_pipe_serial2 python benchmarks/cc_sweep.py --one > gpurun_out/r2_ncu_cc.log 2>&1
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
