"""Generates tests/golden/dfo/dfo_cases.npz by running the UNMODIFIED reference (build container only):

    python tests/golden/make_golden_dfo.py

Each case is one call of the reference's ``dfo_func(policy, env, demand)`` (base_restock_policy.py:24-45) —
on the serial ``InvManagement`` and on ``InvManagementDiv`` exactly as inv_management_div.py:228 calls it
(``demand`` of shape [R, T]) — with and without a sticky noisy delay (the env was reset once with
``noisy_delay=True``; the uniforms of the rollout are replayed from a stored mask).  Also long episodes
(T = 130, 300) where numpy's pairwise summation splits recursively.
"""
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from harness import _uniform_replayer, copy_config, make_delay_mask  # noqa: E402
from make_golden import cfg_to_json  # noqa: E402
from marl_for_im_b200 import presets  # noqa: E402
from oracle.ref_import import load_reference  # noqa: E402


def chain_as_div(T=30):
    """the 4-node serial chain expressed in the divergent env: inv_management_div.py:37-55"""
    return {"num_nodes": 4, "connections": {0: [1], 1: [2], 2: [3], 3: []}, "num_periods": T, "init_inv": np.ones(4) * 10,
            "stock_cost": np.array([0.35, 0.3, 0.4, 0.2]), "backlog_cost": np.array([0.5, 0.7, 0.6, 0.9]), "demand_dist": "poisson",
            "inv_target": np.ones(4) * 0, "inv_max": np.ones(4) * 30, "seed": 52, "delay": np.array([1, 2, 3, 1], dtype=np.int8), "mu": 5,
            "time_dependency": False, "prev_demand": False, "prev_actions": False, "prev_length": 1,
            "standardise_state": False, "standardise_actions": False}


def raw(cfg, T=None):
    cfg = dict(cfg)
    cfg.update(time_dependency=False, prev_demand=False, prev_actions=False, standardise_state=False, standardise_actions=False,
               demand_dist="poisson", mu=5)
    if T is not None:
        cfg["num_periods"] = T
        cfg["delay"] = np.asarray(cfg["delay"], dtype=np.int64)   # the shipped int8 lead times overflow at t = 128 under numpy 2
    return cfg


def cases():
    yield "im4", "IM", raw(presets.serial4_dfo()), False
    yield "im4_T130", "IM", raw(presets.serial4_dfo(), 130), False
    yield "im8_T300_noisy", "IM", raw(presets.serial8(), 300), True
    yield "imdiv_chain", "IM_div", chain_as_div(), False
    yield "imdiv_chain_noisy", "IM_div", chain_as_div(), True
    yield "imdiv1", "IM_div", raw(presets.div1()), False
    yield "imdiv2", "IM_div", raw(presets.div2()), False
    yield "imdiv2_noisy", "IM_div", raw(presets.div2()), True
    yield "imdiv2_T130_std", "IM_div", dict(raw(presets.div2(), 130), standardise_state=True, standardise_actions=True), False


def main():
    R = load_reference()
    rng = np.random.default_rng(20261019)
    out = {}
    names = []
    for name, kind, cfg, noisy in cases():
        cls = R.InvManagement if kind == "IM" else R.InvManagementDiv
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            env = cls(copy_config(cfg))
            if kind == "IM":
                env.reset()                                   # sets env.dist / dist_param (the ctor popped mu: quirk 4 -> mu = 5)
            m, T = env.num_nodes, env.num_periods
            nret = len(env.retailers) if kind == "IM_div" else 1
            zs, dems, masks, wants = [], [], [], []
            for rep in range(4):
                z = rng.integers(8, 38, size=m).astype(float) + (0.37 if rep % 2 else 0.0)
                if kind == "IM_div" and cfg.get("standardise_actions", False):
                    z = rng.uniform(-1, 1, size=m)            # with standardised actions the policy's output is read as a scaled action
                demand = rng.poisson(5, size=(nret, T)) if kind == "IM_div" else rng.poisson(5, size=T)
                mask = make_delay_mask(kind, [int(d) for d in env.delay], T, 0.3, rng) if noisy else np.zeros((T, m), dtype=bool)
                saved = np.random.uniform
                try:
                    if noisy:
                        env.reset(customer_demand=demand, noisy_delay=True, noisy_delay_threshold=0.5)   # sticky from here on
                        seq = iter(_uniform_replayer(kind, [int(d) for d in env.delay], T, mask))
                        np.random.uniform = lambda *a, **k: next(seq)
                    want = R.dfo_func(z, env, demand)
                finally:
                    np.random.uniform = saved
                zs.append(z); dems.append(np.asarray(demand).reshape(nret, T)); masks.append(mask); wants.append(want)
        out[name + "__z"] = np.stack(zs)
        out[name + "__demand"] = np.stack(dems).astype(np.int64)
        out[name + "__mask"] = np.stack(masks)
        out[name + "__dfo"] = np.array(wants)
        out[name + "__meta"] = json.dumps({"kind": kind, "noisy": noisy, "config": cfg_to_json(cfg)})
        names.append(name)
        print("wrote", name, wants)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "dfo", "dfo_cases.npz"), **out)


if __name__ == "__main__":
    main()
