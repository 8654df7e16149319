#!/usr/bin/env python
"""Times the UNMODIFIED reference envs (imported from /root/reference through oracle/ref_import.py's stub gym / ray modules)
on this machine's host cores, beside the oracle port (oracle/im_oracle.py) on the same inputs, and writes
profiles/r2_reference_cpu_timing.json.  The reference tree does not exist on the GPU box, so bench.py ships these figures
as `cpu_baseline_reference` next to the port it times live there (VERDICT r1 task 7d).

    python benchmarks/time_reference.py [--episodes 200]
"""
import argparse
import json
import multiprocessing as mp
import os
import platform
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

from marl_for_im_b200 import presets  # noqa: E402

CONFIGS = {
    "config2_maim4_ma6": ("MAIM", "serial4", 4, 1),
    "maim8_ma6": ("MAIM", "serial8", 8, 1),
    "config4_div1": ("MAIM_div", "div1", 4, 2),
    "config4_div2": ("MAIM_div", "div2", 6, 3),
    "config5_maim2_cc5": ("MAIM", "serial2", 2, 1),
}
T = 30


def _run(args):
    """one worker: `episodes` episodes of one config through the reference (impl 'reference') or the port (impl 'port')"""
    name, impl, episodes, seed = args
    kind, preset, m, R = CONFIGS[name]
    cfg = presets.PRESETS[preset]()
    rng = np.random.default_rng(seed)
    demand = rng.poisson(5, size=(episodes, R, T)) if kind.endswith("_div") else rng.poisson(5, size=(episodes, T))
    actions = rng.uniform(-1, 1, size=(episodes, T, m))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if impl == "reference":
            from harness import KIND_TO_CLASS, agent_names, copy_config
            from oracle.ref_import import load_reference
            env = getattr(load_reference(), KIND_TO_CLASS[kind])(copy_config(cfg))
            names = agent_names(kind, m)
            t0 = time.perf_counter()
            for e in range(episodes):
                env.reset(customer_demand=demand[e])
                for t in range(T):
                    env.step({names[i]: np.array([actions[e, t, i]]) for i in range(m)})
            return time.perf_counter() - t0
        from oracle import im_oracle
        env = im_oracle.OracleEnv(kind, cfg)
        t0 = time.perf_counter()
        for e in range(episodes):
            env.reset(demand[e])
            for t in range(T):
                env.step(actions[e, t])
        return time.perf_counter() - t0


def _run_basestock(args):
    """config 1: IM_env 4-stage + base_stock_policy rollout (inv_management.py:217-232), z = 25"""
    impl, episodes, seed = args
    cfg = presets.serial4_dfo()
    rng = np.random.default_rng(seed)
    demand = rng.poisson(5, size=(episodes, T))
    z = np.full(4, 25.0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if impl == "reference":
            from harness import copy_config
            from oracle.ref_import import load_reference
            R = load_reference()
            env = R.InvManagement(copy_config(cfg))
            t0 = time.perf_counter()
            for e in range(episodes):
                env.reset(customer_demand=demand[e])
                done = False
                while not done:
                    _, _, done, _ = env.step(R.base_stock_policy(z, env))
            return time.perf_counter() - t0
        from oracle import im_oracle
        env = im_oracle.OracleEnv("IM", cfg)
        t0 = time.perf_counter()
        for e in range(episodes):
            im_oracle.base_stock_rollout(env, z, demand[e])
        return time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--episodes", type=int, default=200)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_reference_cpu_timing.json"))
    args = ap.parse_args()
    cores = len(os.sched_getaffinity(0))
    out = {"what": "unmodified reference envs (environments/*.py via stub gym/ray) vs the oracle port, same inputs, same machine",
           "machine": {"cpu": platform.processor() or platform.machine(), "cores_available": cores, "python": platform.python_version(),
                       "numpy": np.__version__, "where": "build container (the GPU box has no /root/reference)"},
           "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), "episodes_per_measurement": args.episodes, "periods": T, "configs": {}}
    pool = mp.get_context("fork").Pool(cores)
    try:
        pool.map(_run, [("config2_maim4_ma6", impl, 2, i) for i in range(cores) for impl in ("reference", "port")])   # imports, warm caches
        for name, (kind, preset, m, R) in CONFIGS.items():
            row = {"kind": kind, "preset": preset, "agents": m}
            for impl in ("reference", "port"):
                dt1 = _run((name, impl, args.episodes, 1))
                t0 = time.perf_counter()
                pool.map(_run, [(name, impl, args.episodes, 100 + i) for i in range(cores)])
                wall = time.perf_counter() - t0
                row[impl] = {"agent_steps_per_sec_1core": args.episodes * T * m / dt1,
                             "agent_steps_per_sec_all_cores": cores * args.episodes * T * m / wall, "cores": cores}
            row["port_over_reference_1core"] = row["port"]["agent_steps_per_sec_1core"] / row["reference"]["agent_steps_per_sec_1core"]
            out["configs"][name] = row
            print(name, json.dumps(row), flush=True)
        row = {"kind": "IM", "preset": "serial4_dfo", "agents": 4}
        for impl in ("reference", "port"):
            dt1 = _run_basestock((impl, args.episodes, 1))
            t0 = time.perf_counter()
            pool.map(_run_basestock, [(impl, args.episodes, 100 + i) for i in range(cores)])
            wall = time.perf_counter() - t0
            row[impl] = {"agent_steps_per_sec_1core": args.episodes * T * 4 / dt1,
                         "agent_steps_per_sec_all_cores": cores * args.episodes * T * 4 / wall, "cores": cores}
        row["port_over_reference_1core"] = row["port"]["agent_steps_per_sec_1core"] / row["reference"]["agent_steps_per_sec_1core"]
        out["configs"]["config1_im4_basestock_rollout"] = row
        print("config1", json.dumps(row), flush=True)
    finally:
        pool.close()
        pool.join()
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
