"""Pins oracle.im_oracle.eval_loop_accumulators to the reference's evaluation loops, run here with the
UNMODIFIED reference envs (build container only).  The loops are script code, not functions, so the loop
bodies below are the statements of MA_inv_management.py:568-581 / inv_management.py:585-600 / DSHLP_4.py:908-913
driven by a replayed action trace instead of ``agent.compute_single_action``."""
import numpy as np
import pytest

from harness import KIND_TO_CLASS, agent_names, copy_config, random_case, reference_available
from marl_for_im_b200 import presets
from oracle import im_oracle

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present")


def reference_eval_loop(kind, cfg, demand, actions, rescaled=True):
    from oracle.ref_import import load_reference
    R = load_reference()
    test_env = getattr(R, KIND_TO_CLASS[kind])(copy_config(cfg))
    multi = kind.startswith("MAIM")
    num_stages = test_env.num_nodes if kind.endswith("_div") else test_env.num_stages
    agent_ids = agent_names(kind, num_stages)
    obs = test_env.reset(customer_demand=np.array(demand))
    episode_reward = 0
    total_inventory = 0
    total_backlog = 0
    customer_backlog = 0
    stage_rewards = np.zeros(num_stages)
    done = False
    t = 0
    while not done:
        if multi:
            obs, reward, dones, info = test_env.step({agent_ids[m]: np.array([actions[t][m]]) for m in range(num_stages)})
            done = dones['__all__']
            total_step_inv = 0
            total_step_bl = 0
            for m in range(num_stages):
                episode_reward += reward[agent_ids[m]]
                stage_rewards[m] += info[agent_ids[m]]['profit']
                total_step_inv += test_env.rev_scale(obs[agent_ids[m]][0], 0, test_env.inv_max[m], test_env.a, test_env.b)
                total_step_bl += test_env.rev_scale(obs[agent_ids[m]][1], 0, test_env.inv_max[m], test_env.a, test_env.b)
            total_inventory += total_step_inv
            total_backlog += total_step_bl
            customer_backlog += test_env.rev_scale(obs[agent_ids[0]][1], 0, test_env.inv_max[0], test_env.a, test_env.b)
        else:
            obs, reward, done, info = test_env.step(actions[t])
            if rescaled:
                inv = test_env.rev_scale(obs[:, 0], np.zeros(num_stages), test_env.inv_max, test_env.a, test_env.b)
                bl = test_env.rev_scale(obs[:, 1], np.zeros(num_stages), test_env.inv_max, test_env.a, test_env.b)
                total_inventory += sum(inv)
                total_backlog += sum(bl)
                customer_backlog += bl[0]
            else:                                          # DSHLP_4.py:911-913
                total_inventory += sum(obs[:, 0])
                total_backlog += sum(obs[:, 1])
                customer_backlog += obs[0, 1]
            for m in range(num_stages):
                stage_rewards[m] += info["profit"][m]
            episode_reward += reward
        t += 1
    return np.array([episode_reward, total_inventory, total_backlog, customer_backlog] + list(stage_rewards), dtype=np.float64)


@pytest.mark.parametrize("kind,preset,kw,rescaled", [
    ("MAIM", "serial4", {}, True), ("MAIM", "serial8", dict(independent=True), True), ("MAIM", "serial2", {}, True),
    ("IM", "serial4", {}, True), ("IM", "serial8", {}, True), ("IM", "serial4_dfo", {}, False),
    ("MAIM_div", "div1", {}, True), ("MAIM_div", "div2", dict(independent=True), True), ("IM_div", "div2", {}, True)])
def test_eval_loop_restatement_matches_reference_loop(kind, preset, kw, rescaled):
    rng = np.random.default_rng(123)
    cfg = presets.PRESETS[preset](**kw)
    m = cfg.get("num_nodes", cfg.get("num_stages"))
    if preset != "serial4_dfo":
        cfg["inv_max"] = np.array([30, 25, 40, 35, 30, 45, 20, 30][:m], dtype=float)
    for _ in range(4):
        demand, actions = random_case(kind, cfg, rng, mu=7, action_mode="near_eq" if kind.endswith("div") else "uniform")
        want = reference_eval_loop(kind, cfg, demand, actions, rescaled)
        got = im_oracle.eval_loop_accumulators(im_oracle.OracleEnv(kind, copy_config(cfg)), demand, actions, rescaled)
        np.testing.assert_array_equal(got, want)
