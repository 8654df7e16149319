#!/usr/bin/env python
"""Config 3's evaluation batch taken apart: us per fused rollout launch, per return-statistics call, and per batch (both),
as bench.py's config3_fused_rollout issues them (stream launches, no graph).

    python benchmarks/rollout_stats.py [--envs 1048576]
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
from marl_for_im_b200.envs import MultiAgentInvManagement  # noqa: E402


def timed(fn, reps):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 20)
    ap.add_argument("--reps", type=int, default=30)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    cfg = presets.serial8(time_dependency=False, prev_demand=False)
    cfg["standardise_actions"] = False
    cfg.update(demand_dist="poisson", mu=5)
    N = args.envs
    env = MultiAgentInvManagement(dict(cfg, num_envs=N))
    m, T = env.num_nodes, env.num_periods
    z = torch.full((m,), 25.0, dtype=torch.float64, device=dev)
    ret = torch.empty((N, m), dtype=torch.float64, device=dev)
    stats = torch.zeros(3 + 2 * m, dtype=torch.float64, device=dev)
    lib, h = env._lib, env._handle
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ep = [0]

    def rollout():
        ep[0] += 1
        _lib.check(lib.imx_rollout_basestock(h, C.c_void_p(z.data_ptr()), 0, None, None, 0, ep[0], None, C.c_void_p(ret.data_ptr()), None, None, 0, s))

    def stat():
        _lib.check(lib.imx_episode_stats(h, C.c_void_p(ret.data_ptr()), 1, None, C.c_void_p(stats.data_ptr()), 1, s))

    def stat_plain():
        _lib.check(lib.imx_return_stats(h, C.c_void_p(ret.data_ptr()), C.c_void_p(stats.data_ptr()), s))

    def both():
        rollout()
        stat()

    out = {"envs": N, "rollout_us": timed(rollout, args.reps), "episode_stats_T1_us": timed(stat, args.reps),
           "return_stats_us": timed(stat_plain, args.reps), "batch_us": timed(both, args.reps)}
    out["agent_steps_per_sec_rollout"] = N * m * T / (out["rollout_us"] * 1e-6)
    out["agent_steps_per_sec_batch"] = N * m * T / (out["batch_us"] * 1e-6)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
