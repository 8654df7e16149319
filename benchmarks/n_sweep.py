#!/usr/bin/env python
"""Launch time of the config-2 step kernel against the batch size (graph of 30 dependent launches / 30): separates the fixed
cost of a dependent launch from the streaming slope.  Prints a small table (profiles/r1_step_kernel_n_sweep.txt)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "benchmarks"))

from bench_configs import PEAK, time_steps  # noqa: E402
from marl_for_im_b200 import presets  # noqa: E402


def main():
    print("# MAIM 4-stage MA_6 step kernel (476 algorithmic bytes per env-step); measured HBM copy peak %.0f GB/s" % PEAK)
    print("#      envs   us/launch   GB/s(algorithmic)   frac   G agent-steps/s   kernel")
    prev = None
    for n in (1024, 4096, 16384, 32768, 65536, 131072, 262144, 524288, 1048576, 2097152, 4194304):
        r = time_steps("MAIM", presets.serial4(), n, 20 if n <= 262144 else 6)
        slope = ""
        if prev is not None:
            slope = "   marginal %.0f GB/s" % (476.0 * (n - prev[0]) / ((r["us_per_launch"] - prev[1]) * 1e-6) / 1e9)
        print(f"{n:11d} {r['us_per_launch']:10.2f} {r['achieved_gbs']:14.0f} {r['frac_of_measured_hbm_peak']:12.3f} "
              f"{r['agent_steps_per_sec'] / 1e9:12.1f}   {r['kernel_variant']}{slope}")
        prev = (n, r["us_per_launch"])


if __name__ == "__main__":
    main()
