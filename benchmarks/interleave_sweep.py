#!/usr/bin/env python
"""Sub-batch interleaving of the per-step API: the batch of N envs is cut into G independent handles (contiguous env ranges,
`env_offset` keeps the global env index), each stepped by its own chain of 30 `imx_step` launches on its own stream; the G
chains are forked from / joined into one stream and captured as ONE CUDA graph.  Envs are independent, so chain g's launch
floor (dependent-launch latency, first-tile load latency, store drain) can hide behind the transfers of the other chains.

    python benchmarks/interleave_sweep.py [--configs serial4,div1,div2] [--envs 32768,65536] [--groups 1,2,4,8]

Prints us per period (all G launches of one period) and the algorithmic rate against the measured HBM peak.
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "benchmarks"))

import torch  # noqa: E402

from bench_configs import PEAK, bytes_per_env_step  # noqa: E402
from marl_for_im_b200 import _lib, presets  # noqa: E402
from marl_for_im_b200.envs import ENV_CLASSES  # noqa: E402

CONFIGS = {"serial4": ("MAIM", presets.serial4), "serial8": ("MAIM", presets.serial8), "div1": ("MAIM_div", presets.div1),
           "div2": ("MAIM_div", presets.div2), "serial2": ("MAIM", presets.serial2)}


def time_interleaved(kind, cfg, N, G, reps, check=False):
    dev = torch.device("cuda:0")
    n = N // G
    envs = [ENV_CLASSES[kind](dict(cfg, num_envs=n, env_offset=g * n)) for g in range(G)]
    e0_ = envs[0]
    m, T, O, R = e0_.num_nodes, e0_.num_periods, e0_.obs_len, len(e0_._retailers)
    gen = torch.Generator(device=dev)
    gen.manual_seed(0)
    demand = torch.poisson(torch.full((N, R, T), 5.0, device=dev), generator=gen).to(torch.int32)
    actions = (torch.randn((T, N, m), dtype=torch.float64, device=dev, generator=gen) * 0.5 - 0.6).clamp(-1, 1)
    obs = torch.empty((T, N, m, O), dtype=torch.float64, device=dev)
    rew = torch.empty((T, N, m), dtype=torch.float64, device=dev)
    lib = e0_._lib
    streams = [torch.cuda.Stream() for _ in range(G)]
    s0 = torch.cuda.current_stream().cuda_stream
    for g, env in enumerate(envs):
        _lib.check(lib.imx_reset(env._handle, C.c_void_p(demand[g * n:(g + 1) * n].contiguous().data_ptr()), None, 0, 1, None, C.c_void_p(s0)))
    torch.cuda.synchronize()

    def chain(g, stream):
        h = envs[g]._handle
        lib.imx_set_period(h, 0)
        for t in range(T):
            _lib.check(lib.imx_step(h, C.c_void_p(actions[t, g * n:(g + 1) * n].data_ptr()), C.c_void_p(obs[t, g * n:(g + 1) * n].data_ptr()),
                                    C.c_void_p(rew[t, g * n:(g + 1) * n].data_ptr()), None, C.c_void_p(stream)))

    def all_chains(main):
        # fork: every group stream waits for the capturing stream; join: the capturing stream waits for every group stream
        for g in range(G):
            if G == 1:
                chain(0, main.cuda_stream)
            else:
                streams[g].wait_stream(main)
                with torch.cuda.stream(streams[g]):
                    chain(g, streams[g].cuda_stream)
        if G > 1:
            for g in range(G):
                main.wait_stream(streams[g])

    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        all_chains(side)
        side.synchronize()
        torch.cuda.synchronize()
        with torch.cuda.graph(graph, stream=side):
            all_chains(torch.cuda.current_stream())
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(3):
        graph.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) * 1e-3 / (reps * T)
    B = bytes_per_env_step(e0_)
    out = {"envs": N, "groups": G, "envs_per_launch": n, "us_per_period": dt * 1e6, "agent_steps_per_sec": N * m / dt,
           "achieved_gbs": B * N / dt / 1e9, "frac_of_measured_hbm_peak": B * N / dt / 1e9 / PEAK,
           "kernel_variant": lib.imx_kernel_variant(e0_._handle), "watchdog_flags": sum(int(e.error_flags.abs().sum()) for e in envs)}
    if check:
        out["reward_checksum"] = float(rew.sum().item())
        out["obs_checksum"] = float(obs[-1].sum().item())
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="serial4,div1,div2")
    ap.add_argument("--envs", default="32768,65536,262144")
    ap.add_argument("--groups", default="1,2,4,8")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--check", action="store_true", help="also print reward / last-observation checksums (must not depend on G)")
    args = ap.parse_args()
    for name in args.configs.split(","):
        kind, preset = CONFIGS[name]
        for N in (int(x) for x in args.envs.split(",")):
            for G in (int(x) for x in args.groups.split(",")):
                try:
                    r = time_interleaved(kind, preset(), N, G, args.reps, args.check)
                except Exception as exc:
                    r = {"error": str(exc)[:300], "envs": N, "groups": G}
                r["config"] = name
                print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
