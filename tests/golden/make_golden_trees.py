"""Generates tests/golden/*div_tree*.npz FROM THE REFERENCE (build container only): one episode each on random
divergent networks (5-12 nodes, up to 5 children per node, heterogeneous lead times and capacities) — the cases that
exercise the round-robin split with more than two children.  Same file format as make_golden.py, so the golden tests
pick them up by name.  Run:  python tests/golden/make_golden_trees.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from harness import random_case, random_tree_config, run_reference  # noqa: E402
from make_golden import cfg_to_json  # noqa: E402

CASES = [("maimdiv_tree0", "MAIM_div", 5, 3, dict()), ("maimdiv_tree1", "MAIM_div", 9, 4, dict(prev_actions=True, prev_length=2, share_network=True)),
         ("maimdiv_tree2", "MAIM_div", 12, 5, dict(independent=True)), ("imdiv_tree0", "IM_div", 6, 3, dict()),
         ("imdiv_tree1", "IM_div", 10, 5, dict(prev_actions=True, prev_length=3)), ("imdiv_tree2", "IM_div", 12, 2, dict(time_dependency=False))]


def main():
    rng = np.random.default_rng(777)
    for name, kind, m, maxc, flags in CASES:
        cfg = random_tree_config(rng, m, maxc, periods=24, **flags)
        demand, actions = random_case(kind, cfg, rng, mu=4, action_mode="near_eq")
        out = run_reference(kind, cfg, demand, actions, None)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), kind=kind, config=cfg_to_json(cfg),
                            demand_trace=np.asarray(demand, dtype=np.int64), actions=actions,
                            delay_mask=np.zeros((0, 0), dtype=bool), **out)
        print("wrote", name, {k: v for k, v in cfg["connections"].items() if v})


if __name__ == "__main__":
    main()
