#!/bin/bash
# bookkeeping kernels (reset + transpose fused, fused return + statistics, PDL): full GPU tests, then the default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_t11_tests.log
tail -3 gpurun_out/r2_t11_tests.log
timeout 600 python bench.py > gpurun_out/r2_bench_book.json 2> gpurun_out/r2_bench_book.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_book.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches") if k in d}, d.get("roofline", {}).get("frac"), d.get("e2e", {}).get("value"))
PY
tail -3 gpurun_out/r2_bench_book.err
