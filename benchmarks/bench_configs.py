#!/usr/bin/env python
"""Secondary measurements: the other BASELINE.json configs (parity-test cases in bench.py's contract)
timed on one GPU, one JSON line each — step kernels of the 8-stage chain and the divergent
networks (config 4), and the fused 30-period base-stock rollout (config 3, one GPU's share).

    python benchmarks/bench_configs.py [--reps 20]
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from marl_for_im_b200 import _lib, presets  # noqa: E402
from marl_for_im_b200.envs import ENV_CLASSES  # noqa: E402

PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
VARIANT = {0: "aot-direct", 1: "aot-tma", 2: "nvrtc-specialised-tma", 3: "nvrtc-specialised-tma-pipeline"}


def bytes_per_env_step(env):
    S, R, m, O = env.state_words, len(env._retailers), env.num_nodes, env.obs_len
    return 2 * 4 * S + 4 * R + 8 * m * (2 + O)


def time_steps(kind, cfg, N, reps, action_mode="uniform"):
    dev = torch.device("cuda:0")
    env = ENV_CLASSES[kind](dict(cfg, num_envs=N))
    m, T, O, R = env.num_nodes, env.num_periods, env.obs_len, len(env._retailers)
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    demand = torch.poisson(torch.full((N, R, T), 5.0, device=dev), generator=g).to(torch.int32)
    if action_mode == "uniform":
        actions = torch.rand((T, N, m), dtype=torch.float64, device=dev, generator=g) * 2 - 1
    else:
        actions = (torch.randn((T, N, m), dtype=torch.float64, device=dev, generator=g) * 0.5 - 0.6).clamp(-1, 1)
    nbuf = 4 if N * m * O * 8 > (64 << 20) else T
    obs = [torch.empty((N, m, O), dtype=torch.float64, device=dev) for _ in range(nbuf)]
    rew = [torch.empty((N, m) if env.MULTI else (N,), dtype=torch.float64, device=dev) for _ in range(nbuf)]
    lib, h = env._lib, env._handle

    def episode(stream):
        _lib.check(lib.imx_reset(h, C.c_void_p(demand.data_ptr()), None, 0, 1, C.c_void_p(obs[0].data_ptr()), C.c_void_p(stream)))
        for t in range(T):
            _lib.check(lib.imx_step(h, C.c_void_p(actions[t].data_ptr()), C.c_void_p(obs[t % nbuf].data_ptr()),
                                    C.c_void_p(rew[t % nbuf].data_ptr()), None, C.c_void_p(stream)))

    def steps_only(stream):
        lib.imx_set_period(h, 0)
        for t in range(T):
            _lib.check(lib.imx_step(h, C.c_void_p(actions[t].data_ptr()), C.c_void_p(obs[t % nbuf].data_ptr()),
                                    C.c_void_p(rew[t % nbuf].data_ptr()), None, C.c_void_p(stream)))

    s = torch.cuda.current_stream().cuda_stream
    episode(s)
    torch.cuda.synchronize()
    variant = lib.imx_kernel_variant(h)
    gs = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        steps_only(side.cuda_stream)
        side.synchronize()
        with torch.cuda.graph(gs, stream=side):
            steps_only(torch.cuda.current_stream().cuda_stream)
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(3):
        gs.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        gs.replay()
    e1.record()
    torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) * 1e-3 / (reps * T)
    B = bytes_per_env_step(env)
    errs = int(env.error_flags.abs().sum())
    return {"us_per_launch": dt * 1e6, "agent_steps_per_sec": N * m / dt, "algorithmic_bytes_per_env_step": B,
            "achieved_gbs": B * N / dt / 1e9, "frac_of_measured_hbm_peak": B * N / dt / 1e9 / PEAK, "kernel_variant": VARIANT[variant],
            "watchdog_flags": errs}


def time_steps_many(kind, cfg, N, reps, action_mode="uniform", want_obs=True):
    """The same T periods as one imx_step_many launch (stored action plan, state resident on chip)."""
    dev = torch.device("cuda:0")
    env = ENV_CLASSES[kind](dict(cfg, num_envs=N))
    m, T, O, R = env.num_nodes, env.num_periods, env.obs_len, len(env._retailers)
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    demand = torch.poisson(torch.full((N, R, T), 5.0, device=dev), generator=g).to(torch.int32)
    if action_mode == "uniform":
        actions = torch.rand((T, N, m), dtype=torch.float64, device=dev, generator=g) * 2 - 1
    else:
        actions = (torch.randn((T, N, m), dtype=torch.float64, device=dev, generator=g) * 0.5 - 0.6).clamp(-1, 1)
    obs = torch.empty((T, N, m, O), dtype=torch.float64, device=dev) if want_obs else None
    rew = torch.empty((T, N, m) if env.MULTI else (T, N), dtype=torch.float64, device=dev)
    lib, h = env._lib, env._handle
    s = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.imx_reset(h, C.c_void_p(demand.data_ptr()), None, 0, 1, None, C.c_void_p(s)))

    def run():
        lib.imx_set_period(h, 0)
        _lib.check(lib.imx_step_many(h, C.c_void_p(actions.data_ptr()), T, C.c_void_p(obs.data_ptr()) if want_obs else None,
                                     C.c_void_p(rew.data_ptr()), None, C.c_void_p(s)))

    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) * 1e-3 / reps
    S = env.state_words
    B = 4 * R + 8 * m * (2 + (O if want_obs else 0)) + 2 * 4 * S / T
    return {"ms_per_launch": dt * 1e3, "us_per_period": dt * 1e6 / T, "agent_steps_per_sec": N * m * T / dt,
            "algorithmic_bytes_per_env_step": B, "achieved_gbs": B * N * T / dt / 1e9, "frac_of_measured_hbm_peak": B * N * T / dt / 1e9 / PEAK,
            "kernel_variant": VARIANT[lib.imx_kernel_variant(h)], "watchdog_flags": int(env.error_flags.abs().sum())}


def time_rollout(kind, cfg, N, reps, replay):
    dev = torch.device("cuda:0")
    env = ENV_CLASSES[kind](dict(cfg, num_envs=N))
    m, T, R = env.num_nodes, env.num_periods, len(env._retailers)
    z = torch.full((m,), 25.0, dtype=torch.float64, device=dev)
    demand = None
    if replay:
        demand = torch.poisson(torch.full((N, R, T), 5.0, device=dev)).to(torch.int32)
    for _ in range(3):
        out = env.rollout_basestock(z, customer_demand=demand)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        out = env.rollout_basestock(z, customer_demand=demand)
    e1.record()
    torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) * 1e-3 / reps
    return {"ms_per_launch": dt * 1e3, "agent_steps_per_sec": N * m * T / dt, "episodes_per_sec": N / dt,
            "mean_return": float(out["returns"].sum(dim=-1).mean().item()) if env.MULTI else float(out["returns"].mean().item()),
            "kernel_variant": VARIANT[env._lib.imx_kernel_variant(env._handle)], "demand": "replayed" if replay else "philox"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--only", default="", help="run only the cases whose name contains this text (ncu captures)")
    args = ap.parse_args()
    jobs = [
        ("config2 MAIM 4-stage step, 65536 envs", lambda: time_steps("MAIM", presets.serial4(), 65536, args.reps)),
        ("MAIM 8-stage step (MA_6 mode), 65536 envs", lambda: time_steps("MAIM", presets.serial8(), 65536, args.reps)),
        ("MAIM 8-stage step, 1048576 envs", lambda: time_steps("MAIM", presets.serial8(), 1 << 20, max(3, args.reps // 4))),
        ("MAIM 2-stage CC_5 step, 65536 envs", lambda: time_steps("MAIM", presets.serial2(), 65536, args.reps)),
        ("IM 4-stage step (FTT), 65536 envs", lambda: time_steps("IM", presets.serial4(time_dependency=False, prev_actions=True), 65536, args.reps)),
        ("config4 MAIM_div div1 step, 262144 envs, uniform actions", lambda: time_steps("MAIM_div", presets.div1(), 262144, args.reps)),
        ("config4 MAIM_div div1 step, 262144 envs, near-equilibrium actions", lambda: time_steps("MAIM_div", presets.div1(), 262144, args.reps, "near_eq")),
        ("config4 MAIM_div div2 step, 262144 envs, uniform actions", lambda: time_steps("MAIM_div", presets.div2(), 262144, args.reps)),
        ("config4 MAIM_div div2 step, 262144 envs, near-equilibrium actions", lambda: time_steps("MAIM_div", presets.div2(), 262144, args.reps, "near_eq")),
        ("config4 MAIM_div div2 step, 2097152 envs, near-equilibrium actions", lambda: time_steps("MAIM_div", presets.div2(), 1 << 21, max(3, args.reps // 4), "near_eq")),
        ("step_many config2 MAIM 4-stage, 65536 envs x 30 periods", lambda: time_steps_many("MAIM", presets.serial4(), 65536, args.reps)),
        ("step_many rewards only (obs = NULL) config2 MAIM 4-stage, 65536 envs x 30 periods", lambda: time_steps_many("MAIM", presets.serial4(), 65536, args.reps, want_obs=False)),
        ("step_many rewards only (obs = NULL) MAIM 4-stage, 1048576 envs x 30 periods", lambda: time_steps_many("MAIM", presets.serial4(), 1 << 20, max(3, args.reps // 4), want_obs=False)),
        ("step_many MAIM 4-stage, 1048576 envs x 30 periods", lambda: time_steps_many("MAIM", presets.serial4(), 1 << 20, max(3, args.reps // 4))),
        ("step_many MAIM 8-stage, 262144 envs x 30 periods", lambda: time_steps_many("MAIM", presets.serial8(), 262144, max(3, args.reps // 2))),
        ("step_many MAIM_div div1, 262144 envs x 30 periods, near-equilibrium", lambda: time_steps_many("MAIM_div", presets.div1(), 262144, max(3, args.reps // 2), "near_eq")),
        ("step_many MAIM_div div2, 262144 envs x 30 periods, near-equilibrium", lambda: time_steps_many("MAIM_div", presets.div2(), 262144, max(3, args.reps // 2), "near_eq")),
        ("config3 MAIM 8-stage fused rollout, 131072 envs, philox", lambda: time_rollout("MAIM", presets.serial8(standardise_actions=False), 131072, args.reps, False)),
        ("config3 MAIM 8-stage fused rollout, 131072 envs, replayed", lambda: time_rollout("MAIM", presets.serial8(standardise_actions=False), 131072, args.reps, True)),
        ("config3 MAIM 8-stage fused rollout, 1048576 envs, philox", lambda: time_rollout("MAIM", presets.serial8(standardise_actions=False), 1 << 20, args.reps, False)),
        ("config1 IM 4-stage DFO rollout, 1048576 envs, philox", lambda: time_rollout("IM", presets.serial4_dfo(), 1 << 20, args.reps, False)),
        ("MAIM_div div2 fused rollout, 262144 envs, philox", lambda: time_rollout("MAIM_div", presets.div2(), 262144, args.reps, False)),
        ("MAIM_div div2 fused rollout, 262144 envs, replayed", lambda: time_rollout("MAIM_div", presets.div2(), 262144, args.reps, True)),
        ("MAIM_div div1 fused rollout, 262144 envs, philox", lambda: time_rollout("MAIM_div", presets.div1(), 262144, args.reps, False)),
        ("IM_div div2 fused rollout, 262144 envs, philox", lambda: time_rollout("IM_div", presets.div2(), 262144, args.reps, False)),
    ]
    for name, fn in jobs:
        if args.only and args.only not in name:
            continue
        try:
            res = fn()
        except Exception as exc:
            res = {"error": str(exc)[:300]}
        print(json.dumps({"case": name, **res}), flush=True)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
