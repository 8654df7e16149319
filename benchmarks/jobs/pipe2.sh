#!/bin/bash
# ring depth x resident CTAs again, now that the cold inputs are prefetched ahead of the dependency wait
mkdir -p gpurun_out
timeout 900 python benchmarks/pipe_sweep.py --no-plain --configs serial4 --envs 32768,65536 --threads 64 --stages 2,3,4 --ctas 6,8,10,12 > gpurun_out/r2_pipe_sweep2.jsonl 2> gpurun_out/r2_pipe_sweep2.err
timeout 900 python benchmarks/pipe_sweep.py --no-plain --configs div1 --envs 32768,262144 --threads 64,128 --stages 2,3,4 --ctas 4,6,8 >> gpurun_out/r2_pipe_sweep2.jsonl 2>> gpurun_out/r2_pipe_sweep2.err
python - <<'PY'
import json
for l in open("gpurun_out/r2_pipe_sweep2.jsonl"):
    d = json.loads(l)
    print(d["config"], d["envs"], d["tma_threads"], d["stages"], d["ctas"], round(d.get("us_per_launch", -1), 2))
PY
tail -2 gpurun_out/r2_pipe_sweep2.err
