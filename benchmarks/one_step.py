#!/usr/bin/env python
"""A few step() launches of one configuration (the target of `ncu --set full` captures).

    python benchmarks/one_step.py --config div2 --envs 262144 [--steps 4] [--rollout] [--many]
"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "benchmarks"))

import torch  # noqa: E402

from floor_sweep import CONFIGS  # noqa: E402
from marl_for_im_b200 import _lib  # noqa: E402
from marl_for_im_b200.envs import ENV_CLASSES  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="serial4")
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--rollout", action="store_true")
    ap.add_argument("--many", action="store_true")
    args = ap.parse_args()
    kind, preset = CONFIGS[args.config]
    cfg = preset()
    if args.rollout:
        cfg.update(time_dependency=False, prev_demand=False, prev_actions=False, demand_dist="poisson", mu=5)
    env = ENV_CLASSES[kind](dict(cfg, num_envs=args.envs))
    N, m, T, O, R = args.envs, env.num_nodes, env.num_periods, env.obs_len, len(env._retailers)
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    if args.rollout:
        z = torch.full((m,), 25.0 if kind in ("IM", "MAIM") and not cfg.get("standardise_actions", True) else 0.2, dtype=torch.float64, device=dev)
        for _ in range(args.steps):
            env.rollout_basestock(z)
        torch.cuda.synchronize()
        print("rollout variant", env._lib.imx_kernel_variant(env._handle))
        return
    demand = torch.poisson(torch.full((N, R, T), 5.0, device=dev), generator=g).to(torch.int32)
    actions = (torch.randn((T, N, m), dtype=torch.float64, device=dev, generator=g) * 0.5 - 0.6).clamp(-1, 1)
    obs = torch.empty((args.steps, N, m, O), dtype=torch.float64, device=dev)
    rew = torch.empty((args.steps, N, m) if env.MULTI else (args.steps, N), dtype=torch.float64, device=dev)
    lib, h = env._lib, env._handle
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.imx_reset(h, C.c_void_p(demand.data_ptr()), None, 0, 1, None, s))
    if args.many:
        _lib.check(lib.imx_step_many(h, C.c_void_p(actions.data_ptr()), args.steps, C.c_void_p(obs.data_ptr()), C.c_void_p(rew.data_ptr()), None, s))
    else:
        for t in range(args.steps):
            _lib.check(lib.imx_step(h, C.c_void_p(actions[t].data_ptr()), C.c_void_p(obs[t].data_ptr()), C.c_void_p(rew[t].data_ptr()), None, s))
    torch.cuda.synchronize()
    print("step variant", lib.imx_kernel_variant(h), "flags", int(env.error_flags.abs().sum()))


if __name__ == "__main__":
    main()
