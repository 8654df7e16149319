"""Drop-in for the reference's base_restock_policy.py: the order-up-to heuristic, the DFO
objective and the Powell search, with the episode loop replaced by the fused rollout kernel.

    base_stock_policy(policy, env)                      base_restock_policy.py:4-21
    dfo_func(policy, env, demand=None)                  base_restock_policy.py:24-45
    optimize_inventory_policy(env, fun, init_policy...) base_restock_policy.py:48-63
"""
from __future__ import annotations

import numpy as np
import torch


def base_stock_policy(policy, env):
    """Order-up-to levels ``policy`` → actions: z - (inv + order_u - backlog), clipped to
    [0, order_max].  Works on drop-in envs (numpy, shape [m]) and batched envs (torch, [N, m])."""
    t = env.period
    inv, order_u, backlog = env.inv[t, :], env.order_u[t, :], env.backlog[t, :]
    if isinstance(inv, torch.Tensor):
        z = torch.as_tensor(np.asarray(policy, dtype=np.float64), device=inv.device) if not isinstance(policy, torch.Tensor) else policy
        inv_ech = inv.double() + order_u.double() - backlog.double()
        om = torch.as_tensor(np.asarray(env.order_max, dtype=np.float64), device=inv.device)
        return torch.minimum(om, torch.clamp(z - inv_ech, min=0.0))
    inv_ech = inv + order_u - backlog
    unc_actions = policy - inv_ech
    return np.minimum(env.order_max, np.maximum(unc_actions, np.zeros(env.num_nodes)))


def dfo_func(policy, env, demand=None, *args):
    """Negative pmf-weighted mean reward of one base-stock episode (base_restock_policy.py:24-45).  One kernel launch:
    the ``while not done`` loop of the reference runs inside imx_rollout_basestock with state on chip, a second small
    kernel forms ``-1 / T * np.sum(prob * rewards)`` in numpy's summation order.  Serial envs take a ``[T]`` trace,
    divergent envs ``[R, T]`` (inv_management_div.py:228: ``demand=test_demand[0, :]``) — ``prob`` then broadcasts against
    the per-period rewards exactly as in the reference.  A sticky ``noisy_delay`` (MAIM_env.py:192-194) carries into the
    rollout like it does into the reference's ``env.reset(customer_demand=demand)``.
    Drop-in env -> float; batched env -> [N] tensor (one objective value per env / demand trace)."""
    if demand is None:
        env.reset()                       # same side effect as the reference: draws a fresh trace
        demand = env.customer_demand if env.customer_demand is not None else env.customer_demand_device().permute(2, 1, 0)
    d_host = demand.cpu().numpy() if isinstance(demand, torch.Tensor) else np.asarray(demand)
    prob = env.dist.pmf(d_host, **env.dist_param)
    out = env.rollout_basestock(np.asarray(policy, dtype=np.float64), customer_demand=demand, pmf=prob)
    if not env.batched:
        env.customer_demand = demand
        return float(out["dfo"][0].item())
    return out["dfo"]


def optimize_inventory_policy(env, fun, init_policy=None, method="Powell", demand=None):
    """scipy Powell search over the base-stock levels (host optimiser, calls ``fun`` per candidate)."""
    from scipy.optimize import minimize
    if init_policy is None:
        init_policy = np.ones(env.num_stages) * env.mu
    if demand is None:
        out = minimize(fun=fun, x0=init_policy, args=env, method=method)
    else:
        out = minimize(fun=fun, x0=init_policy, args=(env, demand), method=method)
    policy = out.x.copy()
    policy = np.round(np.maximum(policy, 0), 0).astype(int)
    return policy, out


# --------------------------------------------------------------------------------------
# population evaluation: many candidate policies x many demand traces per launch
# --------------------------------------------------------------------------------------
def dfo_func_batch(policies, env, demands):
    """Objective of ``dfo_func`` for every (candidate, trace) pair in ONE rollout launch.

    policies [K, m] base-stock levels, demands [D, T] (serial envs) or [D, R, T] (divergent envs) integer traces; ``env`` must be a
    batched env with ``num_envs == K * D``.  Returns a [K, D] float64 tensor whose entry (k, d) equals
    ``dfo_func(policies[k], env1, demands[d])`` of a single-env instance bit for bit."""
    policies = np.asarray(policies, dtype=np.float64)
    demands = np.asarray(demands)
    K, D = policies.shape[0], demands.shape[0]
    if env.num_envs != K * D:
        raise ValueError(f"env.num_envs = {env.num_envs}, need K * D = {K * D}")
    prob = env.dist.pmf(demands, **env.dist_param)                      # [D, T] or [D, R, T]
    z = np.repeat(policies, D, axis=0)                                   # env index = k * D + d
    reps = (K,) + (1,) * (demands.ndim - 1)
    dem = np.tile(demands, reps)
    pmf = np.tile(prob, reps)
    out = env.rollout_basestock(z, customer_demand=dem, pmf=pmf)
    return out["dfo"].reshape(K, D)


def population_search_inventory_policy(env_cls, config, demands, init_policy, sweeps=3, radius=4, device="cuda:0"):
    """Coordinate search over integer base-stock levels using ``dfo_func_batch``: per sweep and per
    stage, all levels within ``radius`` of the incumbent are evaluated on all traces in one launch and
    the level with the lowest mean objective is kept.  A GPU-friendly replacement for the serial Powell
    loop of ``optimize_inventory_policy`` when many traces should shape the policy at once."""
    demands = np.asarray(demands)
    policy = np.round(np.asarray(init_policy, dtype=np.float64))
    m = policy.size
    K = 2 * radius + 1
    env = env_cls(dict(config, num_envs=K * demands.shape[0], device=device))
    best = None
    for _ in range(sweeps):
        for i in range(m):
            cand = np.repeat(policy[None], K, axis=0)
            cand[:, i] = np.maximum(policy[i] + np.arange(-radius, radius + 1), 0)
            score = dfo_func_batch(cand, env, demands).mean(dim=1)       # [K]
            k = int(torch.argmin(score).item())
            policy[i] = cand[k, i]
            best = float(score[k].item())
    return policy.astype(int), best
