"""Prints the headline numbers of bench.py JSON lines read from the files given as arguments, or from stdin."""
import json
import os
import sys

import fileinput

for line in fileinput.input():
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    r, rl = d["roofline"], d.get("roofline_large_n") or {}
    tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("IMX_"))
    print(f"[{tag}] value={d['value'] / 1e9:.2f}G ms/step={d['ms_per_step']:.4f} step_us={r['us_per_launch']:.3f} frac={r['frac']:.3f} "
          f"large_us={rl.get('us_per_launch', 0):.1f} large_frac={rl.get('frac', 0):.3f} e2e={d['e2e']['value'] / 1e6:.0f}M")
