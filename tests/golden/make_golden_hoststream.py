"""Fixtures for the drop-in envs' HOST random stream (build container only).

The reference seeds the global numpy generator in its constructor (MAIM_env.py:50) and then draws
(a) the demand trace inside reset() when none is passed, (b) the noisy-demand mutations of the
divergent envs (MAIM_div_env.py:287-295) and (c) one uniform per eligible stage per period for noisy
delays (MAIM_env.py:449-452).  These fixtures record complete episodes driven ONLY by the seed, so a
drop-in env must consume the same stream in the same order to reproduce them.

    python tests/golden/make_golden_hoststream.py      ->  tests/golden/hoststream/*.npz
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from harness import KIND_TO_CLASS, agent_names, copy_config  # noqa: E402
from make_golden import cfg_to_json  # noqa: E402
from marl_for_im_b200 import presets  # noqa: E402
from oracle.ref_import import load_reference  # noqa: E402

OUT = os.path.join(HERE, "hoststream")


def run(kind, cfg, actions, episodes, noisy_delay_thr):
    R = load_reference()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = getattr(R, KIND_TO_CLASS[kind])(copy_config(cfg))
        m, T = env.num_nodes, env.num_periods
        names = agent_names(kind, m)
        multi = kind.startswith("MAIM")
        demands, all_obs, all_rew = [], [], []
        for e in range(episodes):
            if noisy_delay_thr is not None:
                o = env.reset(noisy_delay=True, noisy_delay_threshold=noisy_delay_thr)
            else:
                o = env.reset()
            demands.append(np.array(env.customer_demand, dtype=np.int64))
            obs = [np.stack([o[n] for n in names]) if multi else np.array(o)]
            rew = np.zeros((T, m))
            for t in range(T):
                act = {names[i]: np.array([actions[e, t, i]]) for i in range(m)} if multi else np.array(actions[e, t])
                o, r, done, info = env.step(act)
                obs.append(np.stack([o[n] for n in names]) if multi else np.array(o))
                if multi:
                    rew[t] = [r[n] for n in names]
                else:
                    rew[t, 0] = r
            all_obs.append(np.stack(obs))
            all_rew.append(rew)
    return np.stack(demands), np.stack(all_obs), np.stack(all_rew)


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(99)
    cases = [
        ("maim4_draw", "MAIM", presets.serial4(mu=7), None),
        ("im4_draw_uniform", "IM", dict(presets.serial4(prev_actions=True), demand_dist="uniform", lower_upper=(2, 9)), None),
        ("maim8_noisy_delay", "MAIM", presets.serial8(), 0.3),
        ("maimdiv2_noisy_demand_and_delay", "MAIM_div", dict(presets.div2(), noisy_demand=True, noisy_demand_threshold=0.2), 0.25),
        ("imdiv1_draw", "IM_div", presets.div1(prev_actions=True), None),
    ]
    for name, kind, cfg, thr in cases:
        m = cfg.get("num_nodes", cfg.get("num_stages"))
        episodes = 3
        actions = np.clip(rng.normal(-0.4, 0.5, size=(episodes, 30, m)), -1, 1)
        demands, obs, rew = run(kind, cfg, actions, episodes, thr)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), kind=kind, config=cfg_to_json(cfg), actions=actions,
                            noisy_delay_threshold=(-1.0 if thr is None else thr), demands=demands, obs=obs, reward=rew)
        print("wrote", name, demands.shape, obs.shape)


if __name__ == "__main__":
    main()
