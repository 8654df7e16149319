#!/usr/bin/env python
"""Shape sweep of the step + centralised-critic-rows kernel (imx_step_cc, config 5's env part: MAIM 2-stage, CC_5 mode, float32
observations): us per period against the CTA size, the ring depth and the resident CTAs per SM of the pipelined kernel.

    python benchmarks/cc_sweep.py [--envs 65536] [--one]      (--one: a few launches of the default shape, for ncu)
"""
import argparse
import ctypes as C
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from marl_for_im_b200 import _lib, presets  # noqa: E402
from marl_for_im_b200.envs import MultiAgentInvManagement  # noqa: E402


def run(N, reps, f32=True, one=False, preset="serial2", plain=False):
    dev = torch.device("cuda:0")
    env = MultiAgentInvManagement(dict(presets.PRESETS[preset](), num_envs=N, **({"obs_dtype": "float32"} if f32 else {})))
    m, T, O = env.num_nodes, env.num_periods, env.obs_len
    W = (m - 1) * (1 + O) + O
    dt_obs = torch.float32 if f32 else torch.float64
    demand = torch.poisson(torch.full((N, 1, T), 5.0, device=dev)).to(torch.int32)
    actions = torch.rand((T, N, m), dtype=torch.float64, device=dev) * 2 - 1
    obs = torch.empty((T, N, m, O), dtype=dt_obs, device=dev)
    cc = torch.empty((T, N, m, W), dtype=dt_obs, device=dev)
    rew = torch.empty((T, N, m), dtype=torch.float64, device=dev)
    lib, h = env._lib, env._handle
    s0 = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.imx_prepare(h, 0))
    _lib.check(lib.imx_reset(h, C.c_void_p(demand.data_ptr()), None, 0, 1, None, C.c_void_p(s0)))

    def steps(stream, n=T):
        lib.imx_set_period(h, 0)
        for t in range(n):
            if plain:
                _lib.check(lib.imx_step(h, C.c_void_p(actions[t].data_ptr()), C.c_void_p(obs[t].data_ptr()), C.c_void_p(rew[t].data_ptr()), None,
                                        C.c_void_p(stream)))
                continue
            _lib.check(lib.imx_step_cc(h, C.c_void_p(actions[t].data_ptr()), C.c_void_p(obs[t].data_ptr()), C.c_void_p(cc[t].data_ptr()), 1,
                                       -1.0, 1.0, C.c_void_p(rew[t].data_ptr()), C.c_void_p(stream)))

    if one:
        steps(s0, 6)
        torch.cuda.synchronize()
        return {"variant": lib.imx_kernel_variant(h)}
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        steps(side.cuda_stream)
        side.synchronize()
        with torch.cuda.graph(g, stream=side):
            steps(torch.cuda.current_stream().cuda_stream)
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * T)
    es = 4 if f32 else 8
    B = 2 * 4 * env.state_words + 4 + 8 * m * 2 + es * m * O + (0 if plain else es * m * W)
    return {"us_per_period": us, "frac": B * N / (us * 1e-6) / 1e9 / 6542.1, "variant": lib.imx_kernel_variant(h), "bytes_per_env_step": B}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--one", action="store_true")
    ap.add_argument("--f64", action="store_true")
    ap.add_argument("--preset", default="serial2")
    ap.add_argument("--decompose", action="store_true", help="plain step vs step + critic rows, float64 vs float32 observations, default shapes")
    args = ap.parse_args()
    if args.decompose:
        for plain in (True, False):
            for f32 in (False, True):
                r = run(args.envs, args.reps, f32, preset=args.preset, plain=plain)
                r.update(critic_rows=not plain, obs="float32" if f32 else "float64")
                print(json.dumps(r), flush=True)
        return
    if args.one:
        print(json.dumps(run(args.envs, 1, not args.f64, one=True, preset=args.preset)))
        return
    print(json.dumps(dict(run(args.envs, args.reps, not args.f64, preset=args.preset), shape="default")), flush=True)
    for thr, st, ct in itertools.product(("64", "128", "256"), ("2", "3", "4"), ("2", "3", "4", "6", "8")):
        os.environ.update(IMX_TMA_THREADS=thr, IMX_PIPE_STAGES=st, IMX_PIPE_CTAS=ct, IMX_PIPE="1")
        try:
            r = run(args.envs, args.reps, not args.f64, preset=args.preset)
        except Exception as exc:
            r = {"error": str(exc)[:200]}
        r.update(threads=int(thr), stages=int(st), ctas=int(ct))
        print(json.dumps(r), flush=True)
    for thr in ("64", "128", "256"):
        os.environ.update(IMX_TMA_THREADS=thr, IMX_PIPE="0")
        os.environ.pop("IMX_PIPE_STAGES", None); os.environ.pop("IMX_PIPE_CTAS", None)
        r = run(args.envs, args.reps, not args.f64, preset=args.preset)
        r.update(threads=int(thr), pipe=0)
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
