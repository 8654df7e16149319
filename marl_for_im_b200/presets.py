"""Canonical environment parameter sets used by the reference's experiment scripts.

The reference has no config files: every script builds an ``env_config`` dict from
module-level constants.  These functions return fresh dicts with the same keys and
values (citations are ``path:line`` in the reference tree), so the drop-in envs,
the parity tests and ``bench.py`` all use the workloads BASELINE.json names.

All functions take the observation-mode flags as keyword arguments because the
scripts toggle them per experiment (``hyperparams.py`` holds 36 such combinations).
"""
from __future__ import annotations

import numpy as np


def _obs_flags(cfg, time_dependency, prev_demand, prev_actions, prev_length):
    cfg["time_dependency"] = time_dependency
    cfg["prev_demand"] = prev_demand
    cfg["prev_actions"] = prev_actions
    cfg["prev_length"] = prev_length
    return cfg


def serial4(time_dependency=True, prev_demand=True, prev_actions=False, prev_length=1, independent=False,
            standardise_state=True, standardise_actions=True, mu=5):
    """4-stage serial chain — inv_management.py:36-56; defaults = preset ``MA_6``
    (hyperparams.py:482-484: td=T, pd=T, pa=F, P=1, shared reward)."""
    cfg = {
        "num_stages": 4, "num_periods": 30,
        "init_inv": np.ones(4) * 10, "inv_target": np.zeros(4), "inv_max": np.ones(4) * 30,
        "price": np.array([5, 4, 3, 2, 1]),
        "stock_cost": np.array([0.35, 0.3, 0.4, 0.2]),
        "backlog_cost": np.array([0.5, 0.7, 0.6, 0.9]),
        "delay": np.array([1, 2, 3, 1], dtype=np.int8),
        "demand_dist": "poisson", "mu": mu, "seed": 52,
        "standardise_state": standardise_state, "standardise_actions": standardise_actions,
        "a": -1, "b": 1, "independent": independent,
    }
    return _obs_flags(cfg, time_dependency, prev_demand, prev_actions, prev_length)


def serial4_dfo(mu=5):
    """``DFO_CONFIG`` — inv_management.py:121-126: raw state/actions, no history (base-stock rollouts)."""
    return serial4(time_dependency=False, prev_demand=False, prev_actions=False,
                   standardise_state=False, standardise_actions=False, mu=mu)


def serial8(time_dependency=True, prev_demand=True, prev_actions=False, prev_length=1, independent=False,
            standardise_state=True, standardise_actions=True, mu=5):
    """8-stage serial chain — MA_inv_management.py:40-61."""
    cfg = {
        "num_stages": 8, "num_periods": 30,
        "init_inv": np.ones(8) * 10, "inv_target": np.zeros(8), "inv_max": np.ones(8) * 30,
        "price": np.array([9, 8, 7, 6, 5, 4, 3, 2, 1]),
        "stock_cost": np.array([0.35, 0.3, 0.4, 0.2, 0.35, 0.3, 0.4, 0.2]),
        "backlog_cost": np.array([0.5, 0.7, 0.6, 0.9, 0.5, 0.7, 0.6, 0.9]),
        "delay": np.array([1, 2, 3, 1, 4, 2, 3, 1], dtype=np.int8),
        "demand_dist": "poisson", "mu": mu, "seed": 52,
        "standardise_state": standardise_state, "standardise_actions": standardise_actions,
        "a": -1, "b": 1, "independent": independent,
    }
    return _obs_flags(cfg, time_dependency, prev_demand, prev_actions, prev_length)


def serial2(time_dependency=True, prev_demand=True, prev_actions=True, prev_length=1, independent=False, mu=5):
    """2-stage serial chain — Oracle_2.py:26-37; default flags = ``CC_5`` mode (O = 8)."""
    cfg = {
        "num_stages": 2, "num_periods": 30,
        "init_inv": np.ones(2) * 10, "inv_target": np.zeros(2), "inv_max": np.ones(2) * 30,
        "price": np.array([3, 2, 1]),
        "stock_cost": np.array([0.5, 0.2]), "backlog_cost": np.array([0.6, 0.9]),
        "delay": np.array([3, 1], dtype=np.int8),
        "demand_dist": "poisson", "mu": mu, "seed": 52,
        "standardise_state": True, "standardise_actions": True,
        "a": -1, "b": 1, "independent": independent,
    }
    return _obs_flags(cfg, time_dependency, prev_demand, prev_actions, prev_length)


def div1(time_dependency=True, prev_demand=True, prev_actions=False, prev_length=1, independent=False,
         share_network=False, mu=5):
    """Divergent network ``div1`` (1 factory → 1 distributor → 2 retailers) —
    MA_inv_management_div.py:46-69."""
    cfg = {
        "num_nodes": 4, "num_periods": 30,
        "connections": {0: [1], 1: [2, 3], 2: [], 3: []},
        "init_inv": np.ones(4) * 10, "inv_target": np.zeros(4), "inv_max": np.ones(4) * 30,
        "stock_cost": np.array([0.35, 0.3, 0.4, 0.4]),
        "backlog_cost": np.array([0.5, 0.7, 0.6, 0.6]),
        "delay": np.array([1, 2, 1, 1], dtype=np.int8),
        "demand_dist": "poisson", "mu": mu, "seed": 52,
        "standardise_state": True, "standardise_actions": True,
        "a": -1, "b": 1, "independent": independent, "share_network": share_network,
    }
    return _obs_flags(cfg, time_dependency, prev_demand, prev_actions, prev_length)


def div2(time_dependency=True, prev_demand=True, prev_actions=False, prev_length=1, independent=False,
         share_network=False, mu=5):
    """Divergent network ``div2`` (6 nodes, two split nodes) — SHLP_div2.py:27-47."""
    cfg = {
        "num_nodes": 6, "num_periods": 30,
        "connections": {0: [1], 1: [2, 3], 2: [4, 5], 3: [], 4: [], 5: []},
        "init_inv": np.ones(6) * 10, "inv_target": np.zeros(6), "inv_max": np.ones(6) * 30,
        "stock_cost": np.ones(6) * 0.4, "backlog_cost": np.ones(6) * 0.6,
        "delay": np.array([1, 2, 1, 1, 2, 1], dtype=np.int8),
        "demand_dist": "poisson", "mu": mu, "seed": 52,
        "standardise_state": True, "standardise_actions": True,
        "a": -1, "b": 1, "independent": independent, "share_network": share_network,
    }
    return _obs_flags(cfg, time_dependency, prev_demand, prev_actions, prev_length)


PRESETS = {"serial4": serial4, "serial4_dfo": serial4_dfo, "serial8": serial8, "serial2": serial2,
           "div1": div1, "div2": div2}


# hyperparams.py:3-1010 — the 36 named experiment configurations ("S_k" single-agent PPO, "MA_k" independent / shared
# multi-agent PPO, "CC_k" centralised critic; k = 1..12).  Their env_config is always the 4-stage chain above; k selects
# the observation mode (k and k + 6 differ only in trainer hyper-parameters, which are not part of this path).
_NAMED_OBS_MODES = {1: (False, False, False), 2: (True, False, False), 3: (False, True, True),
                    4: (False, True, False), 5: (True, True, True), 0: (True, True, False)}
NAMED_ENV_CLASS = {"S": "InvManagement", "MA": "MultiAgentInvManagement", "CC": "MultiAgentInvManagement"}
NAMED_CONFIGS = [f"{p}_{k}" for p in ("S", "MA", "CC") for k in range(1, 13)]


def named_env_config(name):
    """``hyperparams.get_hyperparams(name)['env_config']`` for the reference's named configurations, e.g. ``"MA_6"``
    (BASELINE config 2) or ``"CC_5"``.  Returns ``(env class name, env_config dict)``."""
    prefix, _, k = name.partition("_")
    if prefix not in NAMED_ENV_CLASS or not k.isdigit() or not 1 <= int(k) <= 12:
        raise KeyError(f"unknown configuration {name!r}")
    td, pd, pa = _NAMED_OBS_MODES[int(k) % 6]
    cfg = serial4(time_dependency=td, prev_demand=pd, prev_actions=pa, prev_length=1)
    if prefix == "S":
        del cfg["independent"]                  # the single-agent configurations carry no such key (hyperparams.py:6)
    return NAMED_ENV_CLASS[prefix], cfg
