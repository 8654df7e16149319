#!/usr/bin/env python
"""Episode bookkeeping around the 30 step launches of bench.py's episode graph: us per imx_reset (state fill + initial
observation + demand transpose), per imx_episode_stats (return + statistics), and the whole episode, each as a CUDA graph.

    python benchmarks/bookkeeping.py [--config serial4] [--envs 65536]
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from marl_for_im_b200 import _lib, presets  # noqa: E402
from marl_for_im_b200.envs import ENV_CLASSES  # noqa: E402

CONFIGS = {"serial4": ("MAIM", presets.serial4), "serial8": ("MAIM", presets.serial8), "div1": ("MAIM_div", presets.div1),
           "div2": ("MAIM_div", presets.div2)}


def graph_us(fn, reps=30, inner=1):
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn(s.cuda_stream)
        s.synchronize()
        with torch.cuda.graph(g, stream=s):
            fn(torch.cuda.current_stream().cuda_stream)
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
    return e0.elapsed_time(e1) * 1e3 / (reps * inner)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="serial4")
    ap.add_argument("--envs", type=int, default=65536)
    args = ap.parse_args()
    kind, preset = CONFIGS[args.config]
    dev = torch.device("cuda:0")
    N = args.envs
    env = ENV_CLASSES[kind](dict(preset(), num_envs=N))
    m, T, O, R = env.num_nodes, env.num_periods, env.obs_len, len(env._retailers)
    demand = torch.poisson(torch.full((N, R, T), 5.0, device=dev)).to(torch.int32)
    actions = torch.rand((T, N, m), dtype=torch.float64, device=dev) * 2 - 1
    obs = torch.empty((T, N, m, O), dtype=torch.float64, device=dev)
    rew = torch.empty((T, N, m), dtype=torch.float64, device=dev)
    stats = torch.zeros(3 + 2 * m, dtype=torch.float64, device=dev)
    lib, h = env._lib, env._handle

    def reset(s):
        _lib.check(lib.imx_reset(h, C.c_void_p(demand.data_ptr()), None, 0, 1, C.c_void_p(obs[0].data_ptr()), C.c_void_p(s)))

    def steps(s):
        lib.imx_set_period(h, 0)
        for t in range(T):
            _lib.check(lib.imx_step(h, C.c_void_p(actions[t].data_ptr()), C.c_void_p(obs[t].data_ptr()), C.c_void_p(rew[t].data_ptr()), None, C.c_void_p(s)))

    def stat(s):
        _lib.check(lib.imx_episode_stats(h, C.c_void_p(rew.data_ptr()), T, None, C.c_void_p(stats.data_ptr()), 1, C.c_void_p(s)))

    def episode(s):
        reset(s)
        steps(s)
        stat(s)

    def x10(fn):
        def run(s):
            for _ in range(10):
                fn(s)
        return run

    out = {"config": args.config, "envs": N,
           "reset_us": graph_us(x10(reset), inner=10), "reset_single_graph_us": graph_us(reset),
           "stats_us": graph_us(x10(stat), inner=10), "stats_single_graph_us": graph_us(stat),
           "steps30_us": graph_us(steps), "episode_us": graph_us(episode)}
    out["bookkeeping_in_episode_us"] = out["episode_us"] - out["steps30_us"]
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
