#!/usr/bin/env python
"""BASELINE.json config 5 stand-in: PPO-style rollout with a centralised critic on GPU-resident
2-stage ``MultiAgentInvManagement`` observations (parameters Oracle_2.py:26-37, CC_5 obs mode, O = 8).

RLlib is not installable here, so the sampler is a torch-only stand-in with the shapes of
``CentralizedCriticModel`` (models/CC_Model.py:27-73): per agent an action MLP on own_obs
(8 -> 256 -> 256 -> 2: mean and log-std of the 1-d action) and a value MLP on the flat
centralised-critic observation (17 -> 256 -> 256 -> 1).  Per period: policy forward -> Gaussian
sample -> env.step() (CUDA kernel) -> cc_observe() (CUDA kernel, float32) -> value forward; actions,
log-probs, values, rewards are stored in [T, N, m] trajectory tensors.  Everything stays on the GPU.
The MLPs are library GEMMs (cuBLAS through torch) — plumbing around the hot path, not part of it.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from marl_for_im_b200 import presets  # noqa: E402
from marl_for_im_b200.cc import cc_observe  # noqa: E402
from marl_for_im_b200.envs import MultiAgentInvManagement  # noqa: E402


def mlp(i, o):
    return torch.nn.Sequential(torch.nn.Linear(i, 256), torch.nn.ReLU(), torch.nn.Linear(256, 256), torch.nn.ReLU(), torch.nn.Linear(256, o))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--episodes", type=int, default=5)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    N, T = args.envs, 30
    torch.backends.cuda.matmul.allow_tf32 = True         # the policy / critic GEMMs are plumbing here: tensor-core TF32
    torch.backends.cudnn.allow_tf32 = True
    env = MultiAgentInvManagement(dict(presets.serial2(), num_envs=N, reuse_buffers=True, obs_dtype="float32"))
    m, O = env.num_nodes, env.obs_len
    W = (m - 1) * (1 + O) + O
    torch.manual_seed(0)
    actors = [mlp(O, 2).to(dev) for _ in range(m)]
    critics = [mlp(W, 1).to(dev) for _ in range(m)]
    traj = {k: torch.empty((T, N, m), device=dev) for k in ("action", "logp", "value", "reward")}

    @torch.no_grad()
    def episode():
        env.reset()                                          # Philox Poisson(5) demand drawn on the device
        full = env.last_obs                                  # packed [N, m, O] float32 tensor behind the per-agent views
        for t in range(T):
            own32 = full
            acts = []
            for i in range(m):
                out = actors[i](own32[:, i])
                mean, log_std = out[:, 0], out[:, 1].clamp(-5, 2)
                a = mean + log_std.exp() * torch.randn_like(mean)
                traj["logp"][t, :, i] = -0.5 * ((a - mean) / log_std.exp()) ** 2 - log_std
                acts.append(a)
            action = torch.stack(acts, dim=1)
            traj["action"][t] = action
            cc = cc_observe(env, full, actions=action, dtype=torch.float32)     # critic input of the CURRENT obs + actions
            for i in range(m):
                traj["value"][t, :, i] = critics[i](cc[:, i]).squeeze(-1)
            _, _, done, _ = env.step(action.double())
            full = env.last_obs
            traj["reward"][t] = env.last_reward
        return done

    episode()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.episodes):
        episode()
    e1.record()
    torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) * 1e-3
    print(json.dumps({"case": "config5 stand-in: 2-stage MAIM_env CC_5 + centralised-critic rollout (torch MLPs + imx kernels)",
                      "envs": N, "episodes": args.episodes, "samples_per_sec": N * m * T * args.episodes / dt,
                      "env_steps_per_sec": N * T * args.episodes / dt, "ms_per_period": dt / (args.episodes * T) * 1e3,
                      "mean_episode_return": float(traj["reward"].sum(dim=0).mean().item()),
                      "note": "torch-only stand-in for RLlib's sampler; MLPs are cuBLAS GEMMs"}))


if __name__ == "__main__":
    main()
