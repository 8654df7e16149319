"""Generates tests/golden/*.npz by running the UNMODIFIED reference envs (build container only).

    python tests/golden/make_golden.py

Each fixture holds one episode: the env_config (JSON), the replayed demand trace, the action
trace, the optional noisy-delay mask, and everything the reference produced (obs after reset and
after every step, rewards, info fields, and the inv/backlog/order_u history arrays).  The GPU
box has no reference tree, so these files are what pins the oracle and the CUDA path there.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from harness import make_delay_mask, random_case, run_reference  # noqa: E402
from marl_for_im_b200 import presets  # noqa: E402


def cfg_to_json(cfg):
    out = {}
    for k, v in cfg.items():
        if isinstance(v, np.ndarray):
            out[k] = {"__nd__": v.tolist(), "dtype": str(v.dtype)}
        elif isinstance(v, dict):
            out[k] = {"__dict__": {str(a): list(b) for a, b in v.items()}}
        else:
            out[k] = v
    return json.dumps(out)


CASES = []


def add(name, kind, preset, mu=5, amode="uniform", noisy=False, hetero=False, nonfinite=False, **kw):
    CASES.append(dict(name=name, kind=kind, preset=preset, mu=mu, amode=amode, noisy=noisy, hetero=hetero, nonfinite=nonfinite, kw=kw))


def inject_nonfinite(actions, kind, rng):
    """Overwrites ~15 % of the action trace with +-inf, huge finite values either side of 2^63 and (MAIM kinds only) NaN.
    The MAIM kinds round and .astype(int) before clipping (MAIM_env.py:344-347): on x86-64 everything outside [-2^63, 2^63)
    and NaN converts to INT64_MIN and clips to order 0.  The IM kinds clip first (IM_env.py:300-302); a NaN there turns the
    reference's state into INT64_MIN garbage, so it is left out."""
    specials = [np.inf, -np.inf, 1e19, -1e19, 3e9, -3e9, 1e300, 9.3e18, 9.2e18, 2.0 ** 63, np.nextafter(2.0 ** 63, 0), 6.2e17, 1e18]
    if kind.startswith("MAIM"):
        specials += [np.nan, -np.nan]
    a = np.array(actions, dtype=np.float64, copy=True)
    hit = rng.uniform(size=a.shape) < 0.15
    a[hit] = rng.choice(np.array(specials), size=int(hit.sum()))
    return a


# config 2 of BASELINE.json (MA_6 mode) and its siblings
add("maim4_ma6", "MAIM", "serial4")
add("maim4_ma6_mu15", "MAIM", "serial4", mu=15)
add("maim4_ttt_p3_indep", "MAIM", "serial4", prev_actions=True, prev_length=3, independent=True, hetero=True)
add("maim4_fff", "MAIM", "serial4", time_dependency=False, prev_demand=False)
add("maim4_quirk2_ftf", "MAIM", "serial4", time_dependency=False, prev_demand=True, prev_actions=False)
add("maim4_raw_quirk13", "MAIM", "serial4", prev_actions=True, standardise_state=False, standardise_actions=False)
add("maim4_noisy", "MAIM", "serial4", noisy=True, amode="near_eq")
add("maim8_ma6", "MAIM", "serial8")
add("maim8_ttt_p2", "MAIM", "serial8", prev_actions=True, prev_length=2, hetero=True, mu=12)
add("maim2_cc5", "MAIM", "serial2")
add("im4_ftt", "IM", "serial4", time_dependency=False, prev_demand=True, prev_actions=True)
add("im4_ttt_p2", "IM", "serial4", prev_actions=True, prev_length=2, hetero=True)
add("im4_dfo_raw", "IM", "serial4_dfo")
add("im8_tft", "IM", "serial8", prev_demand=False, prev_actions=True)
add("im8_noisy", "IM", "serial8", noisy=True)
add("maimdiv1_ma6", "MAIM_div", "div1")
add("maimdiv1_neareq", "MAIM_div", "div1", amode="near_eq")
add("maimdiv1_share_ttt", "MAIM_div", "div1", prev_actions=True, prev_length=2, share_network=True, independent=True)
add("maimdiv2_ma6", "MAIM_div", "div2")
add("maimdiv2_neareq_mu8", "MAIM_div", "div2", amode="near_eq", mu=8, hetero=True)
add("maimdiv2_noisy", "MAIM_div", "div2", noisy=True, amode="near_eq")
add("imdiv1_ttt", "IM_div", "div1", prev_actions=True)
add("imdiv2_fff", "IM_div", "div2", time_dependency=False, prev_demand=False, amode="near_eq")
add("imdiv2_tft_hetero", "IM_div", "div2", prev_demand=False, prev_actions=True, hetero=True, mu=9)
# non-finite and out-of-range actions (appended last: the generator stream of the cases above is unchanged)
add("maim4_nonfinite", "MAIM", "serial4", nonfinite=True)
add("maim4_nonfinite_raw", "MAIM", "serial4", nonfinite=True, standardise_actions=False, prev_actions=True)
add("maimdiv2_nonfinite", "MAIM_div", "div2", nonfinite=True, amode="near_eq")
add("im4_nonfinite", "IM", "serial4", nonfinite=True, prev_actions=True)
add("imdiv1_nonfinite", "IM_div", "div1", nonfinite=True)


def main():
    rng = np.random.default_rng(20261018)
    for c in CASES:
        cfg = presets.PRESETS[c["preset"]](**c["kw"])
        m = cfg.get("num_nodes", cfg.get("num_stages"))
        if c["hetero"]:
            cfg["inv_max"] = np.array([30, 25, 40, 35, 30, 45, 20, 30][:m], dtype=float)
            cfg["inv_target"] = np.array([0, 3, 5.5, 1, 0, 2, 4, 0][:m], dtype=float)
        demand, actions = random_case(c["kind"], cfg, rng, mu=c["mu"], action_mode=c["amode"])
        if c["nonfinite"]:
            actions = inject_nonfinite(actions, c["kind"], rng)
        mask = make_delay_mask(c["kind"], cfg["delay"], cfg["num_periods"], 0.3, rng) if c["noisy"] else None
        out = run_reference(c["kind"], cfg, demand, actions, mask)
        np.savez_compressed(
            os.path.join(HERE, c["name"] + ".npz"),
            kind=c["kind"], config=cfg_to_json(cfg), demand_trace=np.asarray(demand, dtype=np.int64),
            actions=actions, delay_mask=(mask if mask is not None else np.zeros((0, 0), dtype=bool)),
            **out)
        print("wrote", c["name"], out["obs"].shape)


if __name__ == "__main__":
    main()
