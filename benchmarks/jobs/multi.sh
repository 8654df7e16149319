#!/bin/bash
# usage: multi.sh <N>  — N-rank PCIe probe, then bench.py on N GPUs (torchrun), outputs in gpurun_out/
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_${N}gpu.txt 2>&1
timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 benchmarks/pcie_probe_nrank.py > gpurun_out/r2_pcie_probe_${N}rank.json 2> gpurun_out/r2_pcie_probe_${N}rank.err
tail -c 300 gpurun_out/r2_pcie_probe_${N}rank.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
tail -c 600 gpurun_out/r2_bench_${N}gpu.err
timeout 120 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>> gpurun_out/r2_bench_${N}gpu.err
cut -c1-300 gpurun_out/r2_bench_${N}gpu.json
