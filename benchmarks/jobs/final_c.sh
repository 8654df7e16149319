#!/bin/bash
# round-2 final: full GPU test suite, the bench line, the other-config sweeps, smoke, batch-size sweep, bookkeeping, reference arm
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_t14_full_tests.log
tail -3 gpurun_out/r2_t14_full_tests.log
timeout 900 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
tail -c 300 gpurun_out/r2_bench_1gpu.err
timeout 600 python benchmarks/bench_configs.py > gpurun_out/r2_other_configs_1gpu.jsonl 2> gpurun_out/r2_other_configs.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
timeout 300 python benchmarks/n_sweep.py > gpurun_out/r2_step_kernel_n_sweep.txt 2> gpurun_out/r2_n_sweep.err
for c in serial4 div2; do timeout 300 python benchmarks/bookkeeping.py --config $c --envs $([ $c = serial4 ] && echo 65536 || echo 262144); done > gpurun_out/r2_bookkeeping.jsonl 2> gpurun_out/r2_bookkeeping.err
for n in 131072 1048576; do timeout 200 python benchmarks/rollout_stats.py --envs $n; done > gpurun_out/r2_rollout_stats.jsonl 2> gpurun_out/r2_rollout_stats.err
for p in serial2 serial4; do timeout 300 python benchmarks/cc_sweep.py --decompose --preset $p; done > gpurun_out/r2_cc_decompose.jsonl 2> gpurun_out/r2_cc_decompose.err
timeout 200 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_1gpu.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches") if k in d}, d.get("roofline", {}).get("frac"), d.get("e2e", {}).get("value"), d.get("e2e_f32_obs", {}).get("value"))
for k, v in d["configs"].items():
    print(k, v.get("agent_steps_per_sec", v.get("samples_per_sec")), v.get("ms_per_batch", v.get("ms_per_episode", v.get("us_per_period"))), v.get("frac", v.get("step_kernel", {}).get("frac")), (v.get("replay_fused") or {}).get("agent_steps_per_sec"))
PY
cat gpurun_out/r2_bookkeeping.jsonl gpurun_out/r2_rollout_stats.jsonl | cut -c1-400
