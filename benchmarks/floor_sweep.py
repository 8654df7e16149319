#!/usr/bin/env python
"""Launch floor of the per-step API: step-kernel time per launch (graph of 30 dependent launches / 30) against the batch
size, the CTA size and the network, next to the floor of an EMPTY dependent launch chain on the same GPU.

    python benchmarks/floor_sweep.py [--threads 64,128,256] [--envs 4096,16384,...] [--configs serial4,div1,div2]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "benchmarks"))

import torch  # noqa: E402

from bench_configs import time_steps  # noqa: E402
from marl_for_im_b200 import presets  # noqa: E402

CONFIGS = {"serial4": ("MAIM", presets.serial4), "serial8": ("MAIM", presets.serial8), "div1": ("MAIM_div", presets.div1),
           "div2": ("MAIM_div", presets.div2), "serial2": ("MAIM", presets.serial2)}


def empty_chain_floor(n=30, reps=50):
    """us per launch of a graph of n tiny dependent torch kernels (no PDL): the floor any kernel boundary pays."""
    x = torch.zeros(32, device="cuda")
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        x.add_(1.0)
        s.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n):
                x.add_(1.0)
    torch.cuda.current_stream().wait_stream(s)
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * n)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--threads", default="64,128,256")
    ap.add_argument("--envs", default="4096,16384,32768,65536,131072,262144")
    ap.add_argument("--configs", default="serial4,div1,div2")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--env", action="append", default=[], help="extra NAME=VALUE environment settings for every run")
    args = ap.parse_args()
    for kv in args.env:
        k, v = kv.split("=", 1)
        os.environ[k] = v
    print(json.dumps({"empty_dependent_launch_us": empty_chain_floor()}), flush=True)
    for name in args.configs.split(","):
        kind, preset = CONFIGS[name]
        for thr in args.threads.split(","):
            os.environ["IMX_TMA_THREADS"] = thr
            for n in (int(x) for x in args.envs.split(",")):
                try:
                    r = time_steps(kind, preset(), n, args.reps)
                except Exception as exc:
                    r = {"error": str(exc)[:200]}
                r.update(config=name, tma_threads=int(thr), envs=n)
                print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
