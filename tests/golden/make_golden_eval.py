"""Generates tests/golden/evalloop/eval_loop.npz FROM THE REFERENCE (build container only): per-episode accumulators
of the scripts' evaluation loops (MA_inv_management.py:538-587, inv_management.py:570-606, DSHLP_4.py:896-928) on
replayed demand / action traces.  Run:  python tests/golden/make_golden_eval.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from harness import random_case  # noqa: E402
from marl_for_im_b200 import presets  # noqa: E402
from test_eval_loop_reference import reference_eval_loop  # noqa: E402

CASES = [("MAIM", "serial4", {}, True), ("MAIM", "serial8", dict(independent=True), True), ("MAIM", "serial2", {}, True),
         ("IM", "serial4", {}, True), ("IM", "serial8", {}, True), ("IM", "serial4_dfo", {}, False),
         ("MAIM_div", "div1", {}, True), ("MAIM_div", "div2", dict(independent=True), True), ("IM_div", "div2", {}, True),
         ("IM_div", "div1", {}, True)]
EPISODES = 6


def case_config(preset, kw):
    cfg = presets.PRESETS[preset](**kw)
    m = cfg.get("num_nodes", cfg.get("num_stages"))
    if preset != "serial4_dfo":
        cfg["inv_max"] = np.array([30, 25, 40, 35, 30, 45, 20, 30][:m], dtype=float)
    return cfg


def main():
    rng = np.random.default_rng(2026)
    out = {"cases": json.dumps([[k, p, kw, r] for k, p, kw, r in CASES])}
    for ci, (kind, preset, kw, rescaled) in enumerate(CASES):
        cfg = case_config(preset, kw)
        dem, act, want = [], [], []
        for _ in range(EPISODES):
            d, a = random_case(kind, cfg, rng, mu=7, action_mode="near_eq" if kind.endswith("div") else "uniform")
            dem.append(d)
            act.append(a)
            want.append(reference_eval_loop(kind, cfg, d, a, rescaled))
        out[f"demand_{ci}"] = np.stack(dem)
        out[f"actions_{ci}"] = np.stack(act)
        out[f"want_{ci}"] = np.stack(want)
    os.makedirs(os.path.join(HERE, "evalloop"), exist_ok=True)
    np.savez_compressed(os.path.join(HERE, "evalloop", "eval_loop.npz"), **out)
    print("wrote", os.path.join(HERE, "evalloop", "eval_loop.npz"))


if __name__ == "__main__":
    main()
