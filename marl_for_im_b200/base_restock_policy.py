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
    """Negative pmf-weighted mean reward of one base-stock episode.  One kernel launch: the 30-step
    ``while not done`` loop of the reference runs inside imx_rollout_basestock with state on chip.
    Drop-in env → float; batched env → [N] tensor (one objective value per env / demand trace)."""
    if demand is None:
        env.reset()                       # same side effect as the reference: draws a fresh trace
        demand = env.customer_demand if env.customer_demand is not None else env.customer_demand_device().permute(2, 1, 0)
    d_host = demand.cpu().numpy() if isinstance(demand, torch.Tensor) else np.asarray(demand)
    prob = env.dist.pmf(d_host, **env.dist_param)
    if prob.ndim > 1 and getattr(env, "DIV", False):
        raise NotImplementedError("dfo_func is defined for the serial envs (one demand trace per episode)")
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
