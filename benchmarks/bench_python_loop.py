#!/usr/bin/env python
"""Host-side cost of the per-step Python API: a plain Python loop of env.step(actions) on GPU-resident tensors (what an RL
loop with a GPU policy executes), against the CUDA-graph replay of the same launches.  One JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from marl_for_im_b200 import presets  # noqa: E402
from marl_for_im_b200.envs import MultiAgentInvManagement  # noqa: E402


def main():
    N, T, m = 65536, 30, 4
    dev = torch.device("cuda:0")
    out = {"case": "python loop of env.step() on device tensors, config 2, 65536 envs"}
    actions = torch.rand((T, N, m), dtype=torch.float64, device=dev) * 2 - 1
    for reuse in (False, True):
        env = MultiAgentInvManagement(dict(presets.serial4(), num_envs=N, reuse_buffers=reuse))
        for _ in range(3):
            env.reset()
            for t in range(T):
                env.step(actions[t])
        torch.cuda.synchronize()
        reps = 20
        t0 = time.perf_counter()
        for _ in range(reps):
            env.reset()
            for t in range(T):
                env.step(actions[t])
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        out[f"reuse_buffers={reuse}"] = {"us_per_step_call": dt / (T + 1) * 1e6, "agent_steps_per_sec": N * m * T / dt}
        if reuse:
            t0 = time.perf_counter()
            for _ in range(reps):
                env.reset()
                for t in range(T):
                    env.step_packed(actions[t])
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / reps
            out["step_packed, reuse_buffers=True"] = {"us_per_step_call": dt / (T + 1) * 1e6, "agent_steps_per_sec": N * m * T / dt}
    # drop-in mode (no num_envs key): one env, numpy in / numpy out, info dicts and history arrays like the reference
    import numpy as np
    env = MultiAgentInvManagement(presets.serial4())
    rng = np.random.default_rng(0)
    acts = rng.uniform(-1, 1, size=(T, m))
    demand = rng.poisson(5, size=T)
    for _ in range(2):
        env.reset(customer_demand=demand)
        for t in range(T):
            env.step({f"stage_{i}": np.array([acts[t, i]]) for i in range(m)})
    reps = 10
    t0 = time.perf_counter()
    for _ in range(reps):
        env.reset(customer_demand=demand)
        for t in range(T):
            env.step({f"stage_{i}": np.array([acts[t, i]]) for i in range(m)})
    dt = (time.perf_counter() - t0) / reps
    out["drop-in N=1 (numpy dicts, info, histories)"] = {"us_per_step_call": dt / (T + 1) * 1e6, "agent_steps_per_sec": m * T / dt,
                                                         "reference_cpu_us_per_step": 241}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
