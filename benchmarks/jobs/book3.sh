#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_t13_tests.log
tail -3 gpurun_out/r2_t13_tests.log
timeout 300 python benchmarks/cc_sweep.py --decompose --preset serial2 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/r2_bench_1gpu_b.json 2> gpurun_out/r2_bench_1gpu_b.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_1gpu_b.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches") if k in d}, d.get("roofline", {}).get("frac"), d.get("e2e", {}).get("value"), d.get("e2e_f32_obs", {}).get("value"))
for k, v in d["configs"].items():
    print(k, v.get("agent_steps_per_sec", v.get("samples_per_sec")), v.get("ms_per_batch", v.get("ms_per_episode", v.get("us_per_period"))), v.get("frac", v.get("step_kernel", {}).get("frac")))
PY
tail -3 gpurun_out/r2_bench_1gpu_b.err
